"""CPU checks of device-side index arithmetic that a compile cannot catch (the kernels themselves need a B200):
the three-store feature layout of csrc/dpp_pair.cu and the compare-exchange network of csrc/bitonic.cuh, restated with
the kernels' own formulas."""
import numpy as np


def test_dpp_pair_feature_layout_covers_every_feature_once():
    # constants of csrc/dpp_pair.cu: 512 threads, groups of 8 lanes x 8 candidates, chains of 16 positions:
    # [0,4) registers, [4,8) shared memory, [8,16) tensor memory
    R, CL, TR, TS, THREADS = 8, 16, 4, 8, 512
    seen, smem, tmem = set(), set(), set()
    for tid in range(THREADS):
        grp, lam = tid // 8, tid % 8
        q, bb = lam & 3, lam >> 2
        warp, lane = tid >> 5, tid & 31
        tcol0, tlane = (warp >> 2) * 128, (warp & 3) * 32 + lane      # tbase in the kernel
        for r in range(R):
            cand = grp * R + r
            for t in range(CL):
                key = (cand, 64 * bb + 4 * t + q)
                assert key not in seen
                seen.add(key)
                if TR <= t < TS:                                        # f_idx(r, t)
                    idx = ((r * ((TS - TR) // 2) + ((t - TR) >> 1)) * THREADS + tid) * 2 + ((t - TR) & 1)
                    assert idx not in smem
                    smem.add(idx)
                elif t >= TS:                                           # granule (rh, tp), 16 columns
                    tp, u, rh, r4 = (t - TS) // 2, (t - TS) & 1, r >> 2, r & 3
                    col = tcol0 + (rh * 4 + tp) * 16 + 4 * r4 + 2 * u
                    for c in (col, col + 1):
                        assert (tlane, c) not in tmem
                        tmem.add((tlane, c))
    assert len(seen) == 512 * 128                                       # 512 candidates x 128 features
    assert len(smem) == 16384 and max(smem) == 16383                    # kPrFDoubles: 128 KB
    assert len(tmem) == 128 * 512                                       # all of tensor memory, nothing twice
    assert max(c for _, c in tmem) == 511 and max(l for l, _ in tmem) == 127
    # staging swizzle: chunk c of row `cl` is written at 4 * (c ^ (cl & 31)); the chain read must find it
    for cl in (0, 1, 31, 32, 77, 255):
        for feat in range(128):
            bb, t, q = feat // 64, (feat % 64) // 4, feat % 4
            assert cl * 128 + q + 4 * ((16 * bb + t) ^ (cl & 31)) == cl * 128 + 4 * ((feat // 4) ^ (cl & 31)) + feat % 4


def test_bitonic_1024_network_sorts_like_the_shared_memory_version():
    # csrc/bitonic.cuh, one element per thread: e = (keep_first == before(e, o)) ? e : o with
    # keep_first = ((t & stride) == 0) == ((t & size) == 0); order = key desc, index asc, padding last
    rng = np.random.default_rng(0)
    t = np.arange(1024)
    for n in (1, 2, 31, 32, 33, 500, 1000, 1023, 1024):
        sc = np.round(rng.random(n), 2)                                 # ties
        k = np.zeros(1024, dtype=np.int64)
        i = np.full(1024, 0x7FFFFFFF, dtype=np.int64)
        k[:n] = (sc * 1000).astype(np.int64) + 1
        i[:n] = np.arange(n)
        size = 2
        while size <= 1024:
            fwd = (t & size) == 0
            stride = size >> 1
            while stride > 0:
                ok, oi = k[t ^ stride], i[t ^ stride]
                keep_first = ((t & stride) == 0) == fwd
                before = (k > ok) | ((k == ok) & (i < oi))
                take_e = keep_first == before
                k, i = np.where(take_e, k, ok), np.where(take_e, i, oi)
                stride >>= 1
            size <<= 1
        assert (i[:n] == np.lexsort((np.arange(n), -sc))).all()
        assert (i[n:] == 0x7FFFFFFF).all()


def test_tile_maximum_threshold_estimates_the_same_row_share_as_the_key_rule():
    # csrc/recall.cu tilemax_tau_kernel: tau = r_t-th largest tile maximum with r_t = T (1 - exp(-256 target / N));
    # the rows reaching it number about `target`, like with the r-th largest sample key (r = target * sample share)
    rng = np.random.default_rng(5)
    N, target = 8_000_000, 4096
    n_tiles = N // 256
    T = max(64, n_tiles // 128)
    stride = n_tiles // T
    x = 256 * target / N
    assert x <= 0.35
    assert T * (1 - np.exp(-x)) >= 24                                  # the rule's own applicability test
    r_t = int(T * (1 - np.exp(-x)) + 0.999)
    r_key = int(target * (T * 256 / N) + 0.999)
    got_t, got_k = [], []
    for _ in range(6):
        s = rng.standard_normal(N).astype(np.float32)
        samp = s[: n_tiles * 256].reshape(n_tiles, 256)[::stride][:T]
        tau_t = np.sort(samp.max(axis=1))[::-1][r_t - 1]
        tau_k = np.sort(samp.ravel())[::-1][r_key - 1]
        got_t.append(int((s >= tau_t).sum()))
        got_k.append(int((s >= tau_k).sum()))
    # both rules are +-1/sqrt(r) estimates of the same quantity: means within 25 % of each other and of the target scale
    assert 0.5 * target < np.mean(got_t) < 2.5 * target, got_t
    assert abs(np.mean(got_t) - np.mean(got_k)) < 0.35 * np.mean(got_k), (got_t, got_k)
