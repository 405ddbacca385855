"""BASELINE.json's full recall size (configs[1]: 10 M items x 64-d f32, top-1000, batch 64) through size-independent
properties — the CPU oracle cannot scan 10 M x 64 x 64 in test time, so exactness is established from:
  1. sortedness: every result list is strictly descending in the order key (score desc, row asc), no duplicates;
  2. exact scores: the returned rows, re-scored by the ORACLE (fmaf chain) from the same matrix rows, give bit-identical
     scores and the same order;
  3. completeness: no row of the catalog outside the list can beat the k-th entry — checked against an independent
     fp32 matmul of the whole catalog (different summation order, so with a 2e-5 margin);
  4. determinism: a second call returns the same bits;
  5. shard invariance: the 2-way row-sharded global-threshold protocol merges to the same bits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, D, B, K = 10_000_000, 64, 64, 1000


def test_recall_full_size_properties(oracle_lib):
    import torch
    from pairec_b200 import Engine
    from pairec_b200.binding import MEM_DEVICE
    from tests.test_shard_gpu import _global_protocol
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    E = torch.randn(N, D, device=dev, generator=g) / D ** 0.5
    g.manual_seed(3)
    Q = torch.randn(B, D, device=dev, generator=g) / D ** 0.5
    eng = Engine(0)
    halves = [Engine(0), Engine(0)]
    try:
        eng.set_item_matrix(E.data_ptr(), rows=N, dim=D, mem=MEM_DEVICE)
        rows = torch.empty(B, K, dtype=torch.int32, device=dev)
        sc = torch.empty(B, K, dtype=torch.float32, device=dev)
        n = torch.empty(B, dtype=torch.int32, device=dev)
        eng.recall_topk_dev(Q.data_ptr(), B, K, rows.data_ptr(), sc.data_ptr(), n.data_ptr())
        eng.sync()
        assert eng.recall_stats()["fallback_queries"] == 0
        assert (n == K).all()
        rows_l = rows.long()

        # 1. strictly descending order keys (score desc, row asc)
        s0, s1, r0, r1 = sc[:, :-1], sc[:, 1:], rows_l[:, :-1], rows_l[:, 1:]
        assert bool(((s0 > s1) | ((s0 == s1) & (r0 < r1))).all()), "a result list is not in order-key order"

        # 2. oracle re-score of the returned rows (bit-exact), for a few queries
        Qh = Q.cpu().numpy()
        for qi in (0, 17, 63):
            order = torch.argsort(rows_l[qi])                      # ascending global row: the oracle breaks ties by row
            sub_rows = rows_l[qi][order]
            E_sub = E[sub_rows].cpu().numpy()
            keys = oracle_lib.recall_topk(E_sub, Qh[qi:qi + 1], K)
            orows, oscores, on = oracle_lib.keys_split(keys)
            assert on[0] == K
            got_rows = rows_l[qi].cpu().numpy()
            assert (sub_rows.cpu().numpy()[orows[0]] == got_rows).all(), "order differs from the oracle's on the same rows"
            assert (oscores[0].view(np.uint32) == sc[qi].cpu().numpy().view(np.uint32)).all(), "scores are not bit-identical"

        # 3. completeness against an independent fp32 matmul of the whole catalog
        kth = sc[:, -1]
        thr = kth + kth.abs() * 2e-5 + 1e-6
        sorted_rows, _ = torch.sort(rows_l, dim=1)
        torch.backends.cuda.matmul.allow_tf32 = False
        chunk = 1_000_000
        for c0 in range(0, N, chunk):
            S = Q @ E[c0:c0 + chunk].T                               # [B, chunk]
            qi, ri = (S > thr[:, None]).nonzero(as_tuple=True)
            ri = ri + c0
            pos = torch.searchsorted(sorted_rows[qi], ri[:, None]).squeeze(1).clamp_(max=K - 1)
            found = sorted_rows[qi, pos] == ri
            assert bool(found.all()), f"{int((~found).sum())} rows beat the k-th score but are not in the list"
            del S

        # 4. determinism
        rows2, sc2 = torch.empty_like(rows), torch.empty_like(sc)
        eng.recall_topk_dev(Q.data_ptr(), B, K, rows2.data_ptr(), sc2.data_ptr(), n.data_ptr())
        eng.sync()
        assert torch.equal(rows, rows2) and torch.equal(sc.view(torch.int32), sc2.view(torch.int32))

        # 4b. the threshold taken from sample KEYS (recall_tilemax=0, round 1's rule) instead of tile maxima: same bits
        eng_keys = Engine(0, recall_tilemax=0)
        try:
            eng_keys.set_item_matrix(E.data_ptr(), rows=N, dim=D, mem=MEM_DEVICE)
            eng_keys.recall_topk_dev(Q.data_ptr(), B, K, rows2.data_ptr(), sc2.data_ptr(), n.data_ptr())
            eng_keys.sync()
            assert eng_keys.recall_stats()["fallback_queries"] == 0
            assert torch.equal(rows, rows2) and torch.equal(sc.view(torch.int32), sc2.view(torch.int32))
        finally:
            eng_keys.close()

        # 5. two row shards, global-threshold protocol, merged: same bits
        half = N // 2
        halves[0].set_item_matrix(E.data_ptr(), rows=half, dim=D, row_base=0, mem=MEM_DEVICE)
        halves[1].set_item_matrix(E.data_ptr() + half * D * 4, rows=N - half, dim=D, row_base=half, mem=MEM_DEVICE)
        gathered, retry = _global_protocol(halves, Qh, K, dev)
        assert retry[0] == 0
        lists = gathered[:, :B * K].reshape(2, B, K).contiguous()
        mrows, mscores, mcnt = halves[0].merge_keys(lists.data_ptr(), 2, B, K)
        assert (mcnt == K).all()
        assert (mrows.astype(np.int64) == rows_l.cpu().numpy()).all(), "sharded result differs from the unsharded one"
        assert (mscores.view(np.uint32) == sc.cpu().numpy().view(np.uint32)).all()
    finally:
        eng.close()
        for h in halves:
            h.close()


def test_fused_path_full_size_equals_staged_calls_and_oracle_spot_checks(oracle_lib):
    """configs[3] at full size (10 M items, 32 tables x 1 M x 16, MLP 512-512-256-128-1, DPP top-50 on 128-d, batch 64):
    the fused prg_recommend equals the four staged C-ABI calls bit for bit (composition), its output is a set of
    distinct recalled rows, and for two requests the oracle reproduces the rank scores (1e-5) from the same tables and
    the DPP selection sequence (exactly) from the GPU's scores."""
    import torch
    import bench
    from pairec_b200 import DppParams, Engine
    from pairec_b200.binding import MEM_DEVICE, MODEL_FM_MLP
    w = bench.WORKLOADS["c4"]
    dev = torch.device("cuda:0")
    T = bench.make_tables_torch(w, dev, 0, 1)
    eng = Engine(0)
    try:
        bench.load_engine(eng, w, T, MEM_DEVICE)     # item matrix, 32 item + 8 user tables, tower 640-512-256-128-1, diversity
        W, b = bench.mlp_weights_np(w["mlp"])
        U = w["user_fields"]
        Bq, k, Tn = w["batch"], w["k"], w["top_n"]
        p = DppParams(top_n=Tn, alpha=1.0, window_size=w["window"])
        g = torch.Generator(device=dev)
        g.manual_seed(3)
        Q = (torch.randn(Bq, w["dim"], device=dev, generator=g) / w["dim"] ** 0.5).cpu().numpy()

        uids = torch.randint(0, w["user_rows"], (Bq, U), device=dev, generator=g, dtype=torch.int32).cpu().numpy().astype(np.uint32)

        rows_f, scores_f, n_f = eng.recommend(Q, k, MODEL_FM_MLP, p, user_ids=uids)   # fused
        rrows, rscores, rn = eng.recall_topk(Q, k)                               # staged: four calls
        rank = eng.rank(MODEL_FM_MLP, rrows, user_ids=uids)
        perm = eng.sort_desc(rank)
        srows = np.take_along_axis(rrows, perm, axis=1)
        sscores = np.take_along_axis(rank, perm, axis=1)
        idx, cnt, st = eng.dpp(srows, sscores, p)
        assert (st == 0).all() and (cnt == Tn).all() and (n_f == Tn).all()
        assert (np.take_along_axis(srows, idx, axis=1) == rows_f).all(), "fused rows differ from the staged calls"
        assert (np.take_along_axis(sscores, idx, axis=1).view(np.uint64) == scores_f.view(np.uint64)).all()
        for bi in range(Bq):                                                     # distinct rows, all from the recall list
            assert len(set(rows_f[bi].tolist())) == Tn
            assert set(rows_f[bi].tolist()) <= set(rrows[bi].tolist())

        # oracle spot checks (two requests): rank scores from the same tables, DPP sequence from the GPU's scores
        factors = [T["factors"][t].cpu().numpy() for t in range(w["n_fields"])] + [T["ufactors"][u].cpu().numpy() for u in range(U)]
        linear = [T["linear"][t].cpu().numpy() for t in range(w["n_fields"])] + [T["ulinear"][u].cpu().numpy() for u in range(U)]
        for bi in (0, Bq - 1):
            r = torch.from_numpy(rrows[bi].astype(np.int64)).to(dev)
            fields_sub = T["fields"][r].cpu().numpy().astype(np.uint32)
            fm, x = oracle_lib.gather_fm(fields_sub, factors, linear, 0.05, np.arange(k, dtype=np.uint32), want_x=True,
                                         user_ids=uids[bi])
            ml = oracle_lib.mlp_forward(x, w["mlp"], W, b)
            want = oracle_lib.sigmoid((fm + ml).astype(np.float32)).astype(np.float64)
            rel = np.abs(rank[bi] - want) / np.maximum(np.abs(want), 1e-30)
            assert rel.max() <= 1e-5, f"rank score error {rel.max():.2e}"
            rs = torch.from_numpy(srows[bi].astype(np.int64)).to(dev)
            emb = T["D"][rs].cpu().numpy().astype(np.float64)
            oidx, ost = oracle_lib.dpp_request(emb, sscores[bi], Tn, alpha=1.0, window_size=w["window"])
            assert ost == 0 and (oidx == idx[bi]).all(), "DPP selection sequence differs from the oracle"
    finally:
        eng.close()
