#!/usr/bin/env python
"""bench.py — requests/s and p99 ms of the recall -> gather+rank -> sort -> DPP hot path on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic requests.

  python bench.py --gpus 1 --steps K --warmup W            the CUDA path (libpairec_gpu.so through its C ABI)
  python bench.py --impl reference --gpus N ...            the reference's CPU path (oracle port) on the host cores
  python bench.py --workload c2|c3|c4|c5|small             BASELINE.json configs[1..4] (default c4 = the metric's config)

Workloads (BASELINE.json `configs`):
  c2  configs[1]: 10 M items x 64-d f32, vector recall top-1000 + ItemRankScore sort, batch 64        (recall + sort)
  c3  configs[2]: 32 categorical tables (1 M rows x 16-d) gather + FM rank of 1000 candidates, batch 256 (gather + FM)
  c4  configs[3]: recall -> 32 item tables + 8 user tables gather + FM + MLP (item 512 + user 128)-512-256-128-1 ->
      score sort -> DPP top-50 (window 10) on 128-d diversity embeddings, batch 64     (the metric's config; default)
  c5  configs[4]: 100 M items x 128-d row-sharded across the ranks, 128 requests per GPU (1024 at 8 GPUs), rest as c4
N>1 (torchrun, one rank per GPU): the item matrix is row-sharded, every rank scans its shard for the global batch,
the shards' candidates are exchanged over NCCL (global-threshold protocol) and each rank merges, ranks and
re-ranks its own requests with replicated feature/diversity tables: per-GPU work is constant -> "weak".  After the
timed region rank 0's results are compared, untimed, with the unsharded path (c4) and with the exact-local protocol.

Timing: W >= 3 warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the library's
stream, max over ranks.  The item matrix (2.56 GB / its 1.28 GB bf16 index) is far larger than the 126 MB L2, so every
step streams it from HBM; the timed loop rotates through 8 distinct request batches (queries, users, candidates).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_ROT = 8   # distinct request batches the timed loop rotates through

WORKLOADS = {
    "c2": dict(kind="recall_sort", items=10_000_000, dim=64, k=1000, batch=64, n_fields=0, table_rows=0, user_fields=0,
               mlp=None, div_dim=0, top_n=0, window=0),
    "c3": dict(kind="fm", items=10_000_000, dim=64, k=1000, batch=256, n_fields=32, table_rows=1_000_000, user_fields=0,
               mlp=None, div_dim=0, top_n=0, window=0),
    "c4": dict(kind="full", items=10_000_000, dim=64, k=1000, batch=64, n_fields=32, table_rows=1_000_000, user_fields=8,
               user_rows=100_000, mlp=[640, 512, 256, 128, 1], div_dim=128, top_n=50, window=10),
    # BASELINE.json configs[4]: 100 M x 128-d row-sharded, 128 requests per GPU = 1024 at N = 8
    "c5": dict(kind="full", items=100_000_000, dim=128, k=1000, batch=128, n_fields=32, table_rows=1_000_000, user_fields=8,
               user_rows=100_000, mlp=[640, 512, 256, 128, 1], div_dim=128, top_n=50, window=10),
    "small": dict(kind="full", items=400_000, dim=64, k=1000, batch=64, n_fields=32, table_rows=50_000, user_fields=8,
                  user_rows=10_000, mlp=[640, 512, 256, 128, 1], div_dim=128, top_n=50, window=10),
}
METRIC = "requests/sec (10M-item recall->rank->DPP)"
UNIT = "requests/s"


def describe(name, w):
    if w["kind"] == "recall_sort":
        return (f"{name}: {w['items']} items x {w['dim']}-d f32 vector recall top-{w['k']} -> ItemRankScore sort")
    if w["kind"] == "fm":
        return (f"{name}: {w['n_fields']}-table ({w['table_rows']} rows x 16-d f32) gather + FM rank of {w['k']} candidates "
                f"per request out of {w['items']} items")
    return (f"{name}: {w['items']} items x {w['dim']}-d f32 recall top-{w['k']} -> {w['n_fields']} item + {w['user_fields']} "
            f"user tables gather + FM + MLP {'-'.join(map(str, w['mlp']))} (bf16 input, bf16x2 hidden act, bf16 W) rank -> "
            f"score sort -> DPP top-{w['top_n']} (window {w['window']}, {w['div_dim']}-d f32 table, fp64 arithmetic)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


_NVML_POLLER = r"""
import sys, time
import pynvml as n
n.nvmlInit()
d = n.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
print("max", n.nvmlDeviceGetMaxClockInfo(d, n.NVML_CLOCK_SM), flush=True)
while True:
    try:
        r = n.nvmlDeviceGetCurrentClocksEventReasons(d)
    except Exception:
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(d)
    print(time.time(), n.nvmlDeviceGetClockInfo(d, n.NVML_CLOCK_SM), int(r), flush=True)
    time.sleep(0.002)
"""


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  A separate PROCESS polls NVML every ~2 ms for the whole
    run (a thread in this process starves behind the GIL while the timed loop issues ctypes calls back to back);
    begin()/end() bracket the timed region and only the samples stamped inside it are reported.  Falls back to
    `nvidia-smi -lms 50` when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.lines = []
        self.sm_max = None
        self.proc = None
        self.kind = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _NVML_POLLER, str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            first = self.proc.stdout.readline().split()
            if len(first) == 2 and first[0] == "max":
                self.sm_max = float(first[1])
                self.kind = "nvml"
                self.t = threading.Thread(target=self._read_nvml, daemon=True)
                self.t.start()
                return
            self.proc.kill()
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.kind = "nvidia-smi"
            self.t = threading.Thread(target=self._read_smi, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read_nvml(self):
        for line in self.proc.stdout:
            f = line.split()
            if len(f) == 3:
                self.lines.append((float(f[0]), float(f[1]), int(f[2])))

    def _read_smi(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source"], "samples": 0}
        time.sleep(0.01)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 or 0.0
        t1 = self.t1 or time.time()
        if self.kind == "nvml":
            inside = [(c, r) for (t, c, r) in self.lines if t0 <= t <= t1]
            if not inside:  # region shorter than one poll: take the nearest samples around it
                inside = [(c, r) for (t, c, r) in self.lines if t0 - 0.01 <= t <= t1 + 0.01]
            bits = 0
            for _, r in inside:
                bits |= r
            return {"sm_mhz": float(np.median([c for c, _ in inside])) if inside else None, "sm_max_mhz": self.sm_max,
                    "reasons": sorted(k for k, b in self.BITS.items() if bits & b), "samples": len(inside),
                    "source": "nvml poller process, 2 ms"}
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.05 and len(r) >= 8]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 50"}


# ---------------------------------------------------------------------------------------------- synthetic tables
def make_tables_torch(w, dev, rank, world):
    """Synthetic tables generated ON the device with torch (plumbing), in row chunks so that the 100 M-item workload
    never holds more than one chunk of temporaries.  Seeds follow SURVEY §8d."""
    import torch
    g = torch.Generator(device=dev)
    items, dim = w["items"], w["dim"]
    shard = items // world
    row_base = rank * shard
    if rank == world - 1:
        shard = items - row_base
    CH = 5_000_000
    E = torch.empty(shard, dim, device=dev)
    g.manual_seed(2 + 1000 * rank)
    for c0 in range(0, shard, CH):
        c1 = min(shard, c0 + CH)
        E[c0:c1] = torch.randn(c1 - c0, dim, device=dev, generator=g) / dim ** 0.5
    T = dict(E=E, row_base=row_base)
    F, R = w["n_fields"], w["table_rows"]
    if F:
        fields = torch.empty(items, F, dtype=torch.int32, device=dev)
        g.manual_seed(5)
        for c0 in range(0, items, CH):
            c1 = min(items, c0 + CH)
            u = torch.rand(c1 - c0, F, device=dev, generator=g)
            fields[c0:c1] = (u * u * u * R).to(torch.int32).clamp_(0, R - 1)  # skewed ids (hot head), replicated on every rank
            del u
        g.manual_seed(4)
        T["fields"] = fields
        T["factors"] = torch.randn(F, R, 16, device=dev, generator=g) * 0.4
        T["linear"] = torch.randn(F, R, device=dev, generator=g) * 0.05
    U = w.get("user_fields", 0)
    if U:
        g.manual_seed(9)
        T["ufactors"] = torch.randn(U, w["user_rows"], 16, device=dev, generator=g) * 0.4
        T["ulinear"] = torch.randn(U, w["user_rows"], device=dev, generator=g) * 0.05
    if w["div_dim"]:
        D = torch.empty(items, w["div_dim"], device=dev)
        g.manual_seed(7)
        for c0 in range(0, items, CH):
            c1 = min(items, c0 + CH)
            d = torch.randn(c1 - c0, w["div_dim"], device=dev, generator=g)
            D[c0:c1] = d / d.norm(dim=1, keepdim=True)
            del d
        T["D"] = D
    return T


def bf16_bits(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def mlp_weights_np(dims, seed=6):
    rng = np.random.default_rng(seed)
    W, b = [], []
    for l in range(len(dims) - 1):
        lim = np.sqrt(6.0 / (dims[l] + dims[l + 1]))
        W.append(bf16_bits(rng.uniform(-lim, lim, size=(dims[l + 1], dims[l])).astype(np.float32)))
        b.append((rng.standard_normal(dims[l + 1]) * 0.01).astype(np.float32))
    return W, b


def load_engine(eng, w, T, mem):
    eng.set_item_matrix(T["E"].data_ptr(), rows=T["E"].shape[0], dim=w["dim"], row_base=T["row_base"], mem=mem)
    F, U = w["n_fields"], w.get("user_fields", 0)
    if F:
        eng.set_item_fields(T["fields"].data_ptr(), rows=w["items"], n_fields=F, mem=mem)
        for t in range(F):
            eng.set_feature_table(t, T["factors"][t].data_ptr(), T["linear"][t].data_ptr(), rows=w["table_rows"], fdim=16, mem=mem)
        for u in range(U):
            eng.set_feature_table(F + u, T["ufactors"][u].data_ptr(), T["ulinear"][u].data_ptr(), rows=w["user_rows"], fdim=16,
                                  mem=mem)
        eng.set_fm_bias(0.05)
        if U:
            eng.set_user_fields(U, 0)
    if w["mlp"]:
        W, b = mlp_weights_np(w["mlp"])
        eng.set_mlp(w["mlp"], W, b)
    if w["div_dim"]:
        eng.set_diversity_matrix(T["D"].data_ptr(), rows=w["items"], dim=w["div_dim"], dtype=0, mem=mem)


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_tables(w):
    """Host tables for the CPU arm at the workload's FULL sizes (numpy, same distributions as the device tables)."""
    rng = np.random.default_rng(2)
    n, F, R, U = w["items"], w["n_fields"], w["table_rows"], w.get("user_fields", 0)
    T = {}
    if w["kind"] != "fm":
        T["E"] = (rng.standard_normal((n, w["dim"]), dtype=np.float32) / np.float32(w["dim"] ** 0.5))
    if F:
        fields = np.empty((n, F), dtype=np.uint32)
        for c0 in range(0, n, 2_000_000):
            u = rng.random((min(n, c0 + 2_000_000) - c0, F), dtype=np.float32)
            fields[c0:c0 + u.shape[0]] = np.minimum((u * u * u * R).astype(np.uint32), R - 1)
        T["fields"] = fields
        T["factors"] = [(rng.standard_normal((R, 16), dtype=np.float32) * np.float32(0.4)) for _ in range(F)]
        T["linear"] = [(rng.standard_normal(R, dtype=np.float32) * np.float32(0.05)) for _ in range(F)]
        for _ in range(U):
            T["factors"].append(rng.standard_normal((w["user_rows"], 16), dtype=np.float32) * np.float32(0.4))
            T["linear"].append(rng.standard_normal(w["user_rows"], dtype=np.float32) * np.float32(0.05))
    if w["div_dim"]:
        D = rng.standard_normal((n, w["div_dim"]), dtype=np.float32)
        D /= np.linalg.norm(D, axis=1, keepdims=True)
        T["D"] = D
    if w["mlp"]:
        T["W"], T["b"] = mlp_weights_np(w["mlp"])
    return T


def cpu_step(w, T, Q, rows_in, uids, threads):
    """The oracle port of the workload's path on the host cores for the requests given; seconds per stage."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    k = w["k"]
    st = {}
    if w["kind"] != "fm":
        B = Q.shape[0]
        t0 = time.perf_counter()
        keys = oracle.recall_topk(T["E"], Q, k, n_threads=threads)
        st["recall"] = time.perf_counter() - t0
        rows, scores, _ = oracle.keys_split(keys)
    else:
        rows, B = rows_in, rows_in.shape[0]
    if w["kind"] == "recall_sort":
        t0 = time.perf_counter()
        for b in range(B):
            oracle.go_sort(scores[b].astype(np.float64))   # Item.Score = float64(score) (vector_recall.go:99), ItemRankScore
        st["sort"] = time.perf_counter() - t0
        return st
    U = w.get("user_fields", 0)
    t0 = time.perf_counter()
    fm, xs = [], []
    for b in range(B):   # per request: its user's features are merged into every candidate's feature map
        f, x = oracle.gather_fm(T["fields"], T["factors"], T["linear"], 0.05, rows[b], want_x=w["mlp"] is not None,
                                user_ids=uids[b] if U else None)
        fm.append(f)
        xs.append(x)
    st["gather_fm"] = time.perf_counter() - t0
    if w["kind"] == "fm":
        oracle.sigmoid(np.concatenate(fm))
        return st
    t0 = time.perf_counter()
    ml = oracle.mlp_forward(np.concatenate(xs), w["mlp"], T["W"], T["b"])
    sc = oracle.sigmoid((np.concatenate(fm) + ml).astype(np.float32)).astype(np.float64).reshape(B, k)
    st["mlp"] = time.perf_counter() - t0

    def one(b):
        perm = oracle.stable_sort_desc(sc[b])
        r = rows[b][perm]
        idx, _ = oracle.dpp_request(T["D"][r].astype(np.float64), sc[b][perm], w["top_n"], alpha=1.0, window_size=w["window"])
        return r[idx]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=max(1, min(threads, B))) as ex:
        list(ex.map(one, range(B)))
    st["sort_dpp"] = time.perf_counter() - t0
    return st


def cpu_measure(w, steps, warmup, reqs_per_step):
    """`steps` timed steps of `reqs_per_step` requests each (a bounded sample of the workload's batch) at FULL table
    sizes, distinct requests every step."""
    import oracle
    oracle.build()
    threads = oracle.num_threads()
    scale = 1.0
    if w["items"] > 20_000_000:   # c5: 115 GB of host tables do not fit every box: recall over 10 M rows, time scaled
        scale = w["items"] / 10_000_000
        w = dict(w, items=10_000_000)
    T = cpu_tables(w)
    rng = np.random.default_rng(3)
    U = w.get("user_fields", 0)
    per = []
    for i in range(warmup + steps):
        Q = (rng.standard_normal((reqs_per_step, w["dim"]), dtype=np.float32) / np.float32(w["dim"] ** 0.5))
        rows_in = rng.integers(0, w["items"], size=(reqs_per_step, w["k"])).astype(np.uint32) if w["kind"] == "fm" else None
        uids = rng.integers(0, w["user_rows"], size=(reqs_per_step, U)).astype(np.uint32) if U else None
        t0 = time.perf_counter()
        st = cpu_step(w, T, Q, rows_in, uids, threads)
        st["total"] = time.perf_counter() - t0
        if scale != 1.0:
            st["total"] += st["recall"] * (scale - 1.0)
            st["recall_scaled_to_full_catalog"] = st["recall"] * scale
        if i >= warmup:
            per.append(st)
    tot = [p["total"] for p in per]
    mean = float(np.mean(tot))
    stages = {k: float(np.median([p[k] for p in per])) for k in per[0]}
    return dict(value=reqs_per_step / mean, ms_per_step=mean * 1e3, cores=threads, stages_s=stages,
                sample=f"{steps} steps of {reqs_per_step} requests (of the workload's batch of {w['batch']}) at full table "
                       f"sizes ({w['items']} items); oracle port (C, AVX2 FMA, pthreads) on all {threads} host threads; "
                       f"distinct requests per step" + (f"; catalog capped at 10 M rows, recall time scaled x{scale:.0f}" if scale != 1.0 else ""))


# ---------------------------------------------------------------------------------------------- prg_group (one process)
def main_group(args, w, config):
    """The row-sharded path of ONE process over G GPUs (prg_group_*: P2P exchanges inside the library, no NCCL): the call a
    Go host makes.  Host buffers in and out, so every number here is end to end.  Not the driver's N>1 contract (that is
    torchrun, one rank per GPU); an extra line for profiles/."""
    import torch
    from pairec_b200 import DppParams, Engine, Group
    from pairec_b200.binding import MEM_DEVICE, MODEL_FM_MLP
    G = args.group
    assert w["kind"] == "full"
    engs, tabs = [], []
    for g in range(G):
        torch.cuda.set_device(g)
        dev = torch.device("cuda", g)
        T = make_tables_torch(w, dev, g, G)
        e = Engine(g)
        load_engine(e, w, T, MEM_DEVICE)
        engs.append(e)
        tabs.append(T)
    grp = Group(engs)
    B, k, Tn, U = w["batch"], w["k"], w["top_n"], w.get("user_fields", 0)
    Bg = B * G
    rng = np.random.default_rng(3)
    Qs = [(rng.standard_normal((Bg, w["dim"])) / w["dim"] ** 0.5).astype(np.float32) for _ in range(N_ROT)]
    Us = [rng.integers(0, w["user_rows"], size=(Bg, U)).astype(np.uint32) for _ in range(N_ROT)] if U else [None] * N_ROT
    p = DppParams(top_n=Tn, alpha=1.0, window_size=w["window"])
    warm = max(args.warmup, N_ROT)
    redone = 0
    for i in range(warm):
        grp.recommend(Qs[i % N_ROT], k, MODEL_FM_MLP, p, user_ids=Us[i % N_ROT])
    l0 = sum(e.launches for e in engs)
    lat = []
    t0 = time.perf_counter()
    for i in range(args.steps):
        t1 = time.perf_counter()
        out = grp.recommend(Qs[i % N_ROT], k, MODEL_FM_MLP, p, user_ids=Us[i % N_ROT])
        lat.append((time.perf_counter() - t1) * 1e3)
        redone += int(out[3])
    wall = time.perf_counter() - t0
    launches = sum(e.launches for e in engs) - l0
    parity = None
    if w["items"] * w["dim"] * 4 <= 8e9 and not args.no_verify:   # the unsharded path over the gathered matrix on GPU 0
        torch.cuda.set_device(0)
        full = torch.cat([tabs[g]["E"].to("cuda:0") for g in range(G)])
        T1 = dict(tabs[0])
        T1["E"], T1["row_base"] = full, 0
        e1 = Engine(0)
        load_engine(e1, w, T1, MEM_DEVICE)
        want = e1.recommend(Qs[0][:B * G], k, MODEL_FM_MLP, p, user_ids=Us[0])
        got = grp.recommend(Qs[0], k, MODEL_FM_MLP, p, user_ids=Us[0])
        parity = bool((got[0] == want[0]).all() and (got[1].view(np.uint64) == want[1].view(np.uint64)).all() and
                      (got[2] == want[2]).all())
        e1.close()
    line = {"impl": "b200-group", "metric": METRIC, "value": Bg * args.steps / wall, "unit": UNIT, "n_gpus": G,
            "steps": args.steps, "warmup": warm, "ms_per_step": wall / args.steps * 1e3,
            "p50_ms_per_step": float(np.percentile(lat, 50)), "p99_ms_per_step": float(np.percentile(lat, 99)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (recall, FM), bf16 / bf16x2 -> f32 (MLP), f64 (sort, DPP)", "data": "synthetic",
            "config": dict(config, global_batch=Bg,
                           sharding="item matrix row-sharded over the GPUs of ONE process; prg_group_recommend: sample keys "
                                    "and candidate prefixes travel as peer stores over NVLink, no NCCL"),
            "gpu_launches": int(launches), "batches_redone_exactly": redone,
            "e2e": {"value": Bg * args.steps / wall, "unit": UNIT, "h2d_bytes_per_step": G * (Bg * w["dim"] * 4 + Bg * U * 4),
                    "d2h_bytes_per_step": Bg * Tn * 12 + Bg * 4,
                    "note": "host buffers in and out: every GPU copies the whole request batch in, returns its own results"},
            "equals_unsharded_path": parity}
    print(json.dumps(line))
    sys.stdout.flush()
    os._exit(0)


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PRG_WORKLOAD", "c4"), choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-requests", type=int, default=0,
                    help="requests per step of the CPU arm (0 = the workload's whole batch; a smaller bounded sample "
                         "under-uses the cache reuse of the CPU recall across queries)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batcher", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the untimed N>1 parity checks")
    ap.add_argument("--group", type=int, default=0, help="ONE process driving this many GPUs through prg_group_recommend "
                                                         "(the in-library multi-GPU path a Go host would call)")
    ap.add_argument("--items", type=int, default=0, help="override the catalog size (e.g. c5's per-GPU shard shape on fewer GPUs)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.items:
        w["items"] = args.items
    if args.cpu_requests <= 0:
        args.cpu_requests = min(w["batch"], 64)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)
    kind = w["kind"]
    if world > 1 and kind != "full":
        raise SystemExit("c2 / c3 are single-GPU workloads (parity-test configs)")
    config = {"workload": describe(args.workload, w),
              "batch_per_gpu": w["batch"], "global_batch": w["batch"] * world,
              "sharding": ("item matrix row-sharded; " + {"local": "all-gather of per-shard top-k keys",
                           "global": "global threshold: all-gather of per-shard sample keys, then all-gather of the "
                                     "candidates that reach it",
                           "a2a": "global threshold: all-gather of per-shard sample keys, then ONE all-to-all of the candidates "
                                  "that reach it (a rank receives its own requests' lists only)"}[
                               os.environ.get("PRG_SHARD_PROTOCOL", "a2a")] if world > 1 else "none"),
              "l2": (f"inputs larger than L2 (item matrix: its {w['items'] // world * w['dim'] * 2 / 1e9:.2f} GB bf16 or "
                     f"{w['items'] // world * (w['dim'] + 8) / 1e9:.2f} GB int8 filter index streamed per step)" if kind != "fm" else
                     "inputs larger than L2 (2.2 GB of feature tables + 1.28 GB of item fields, random rows)"),
              "request_batches_rotated": N_ROT}

    if args.group:
        return main_group(args, w, config)
    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_measure(w, max(1, args.steps), max(0, args.warmup), args.cpu_requests)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 recall/FM, f64 MLP-acc/DPP",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                 "sample": r["sample"], "stages_s": r["stages_s"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    from pairec_b200 import DppParams, Engine
    from pairec_b200.binding import MEM_DEVICE, MODEL_FM, MODEL_FM_MLP, UserFeatures
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local_rank)
    T = make_tables_torch(w, dev, rank, world)
    load_engine(eng, w, T, MEM_DEVICE)
    p = DppParams(top_n=max(1, w["top_n"]), alpha=1.0, window_size=max(1, w["window"]))

    B, k, Tn, U = w["batch"], w["k"], w["top_n"], w.get("user_fields", 0)
    Bg = B * world
    gq = torch.Generator(device=dev)
    gq.manual_seed(3)
    # N_ROT distinct global request batches (the same on every rank): queries, user ids
    Qs = [torch.randn(Bg, w["dim"], device=dev, generator=gq) / w["dim"] ** 0.5 for _ in range(N_ROT)]
    Us = [torch.randint(0, w["user_rows"], (Bg, U), device=dev, generator=gq, dtype=torch.int32) for _ in range(N_ROT)] if U else None
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    lib, h = eng._lib, eng._h

    def uptr(i, local=True):   # this rank's requests' user ids of batch i
        if not U:
            return None
        return Us[i].data_ptr() + (rank * B * U * 4 if local else 0)

    launches_extra = 0
    if kind == "recall_sort":
        r_rows = torch.empty(B, k, dtype=torch.int32, device=dev)
        r_sc = torch.empty(B, k, dtype=torch.float32, device=dev)
        r_n = torch.empty(B, dtype=torch.int32, device=dev)
        r_perm = torch.empty(B, k, dtype=torch.int32, device=dev)
        out_rows = r_rows

        def step_device(i):
            eng.recall_topk_dev(Qs[i % N_ROT].data_ptr(), B, k, r_rows.data_ptr(), r_sc.data_ptr(), r_n.data_ptr())
            with torch.cuda.stream(stream):
                sc64 = r_sc.to(torch.float64)   # Item.Score = float64(score), vector_recall.go:99 (a torch cast: plumbing)
            eng.sort_desc_dev(sc64.data_ptr(), B, k, r_perm.data_ptr())
    elif kind == "fm":
        # candidates: N_ROT sets of B x k item rows (what C2's recall hands over), drawn once, untimed
        cand = [torch.randint(0, w["items"], (B, k), device=dev, generator=gq, dtype=torch.int32) for _ in range(N_ROT)]
        fm_out = torch.empty(B, k, dtype=torch.float64, device=dev)
        out_rows = fm_out

        def step_device(i):
            eng.rank_dev(MODEL_FM, cand[i % N_ROT].data_ptr(), B, k, fm_out.data_ptr())
    else:
        out_rows = torch.empty(B, Tn, dtype=torch.int32, device=dev)
        out_scores = torch.empty(B, Tn, dtype=torch.float64, device=dev)
        out_n = torch.empty(B, dtype=torch.int32, device=dev)
        # "a2a" (default): one global threshold per query, candidates exchanged with ONE all-to-all (each rank receives only
        # its own requests' lists); "global": the same with an all-gather of everything; "local": exact per-shard top-k lists
        protocol = os.environ.get("PRG_SHARD_PROTOCOL", "a2a")
        if world > 1 and protocol == "local":
            keys_local = torch.empty(Bg, k, dtype=torch.int64, device=dev)
            keys_all = torch.empty(world, Bg, k, dtype=torch.int64, device=dev)
        elif world > 1:
            r_s = eng.shard_sample_len(k)
            samp_local = torch.empty(Bg, r_s, dtype=torch.int64, device=dev)
            samp_all = torch.empty(world, Bg, r_s, dtype=torch.int64, device=dev)
            blk = Bg * k + Bg                                        # per rank: Bg x k keys + Bg status words
            cand_local = torch.empty(blk, dtype=torch.int64, device=dev)
            retry = torch.zeros(2, dtype=torch.int32, device=dev)
            if protocol == "a2a":
                chunk = B * k + B                                    # what one rank needs of one shard: its B lists + B status words
                cand_packed = torch.empty(world, chunk, dtype=torch.int64, device=dev)
                cand_recv = torch.empty(world, chunk, dtype=torch.int64, device=dev)
            else:
                cand_all = torch.empty(world, blk, dtype=torch.int64, device=dev)

        def step_local_protocol(i, rows_t, scores_t, n_t):
            Qg = Qs[i % N_ROT]
            eng.recall_local_keys_dev(Qg.data_ptr(), Bg, k, keys_local.data_ptr())
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(keys_all, keys_local)
            eng.recommend_from_keys_dev(keys_all.data_ptr() + rank * B * k * 8, world, Bg * k, B, k, MODEL_FM_MLP, p,
                                        rows_t.data_ptr(), scores_t.data_ptr(), n_t.data_ptr(), user_ids_ptr=uptr(i % N_ROT))

        def step_device(i):
            Qg = Qs[i % N_ROT]
            if world == 1:
                eng.recommend_dev(Qg.data_ptr(), B, k, MODEL_FM_MLP, p, out_rows.data_ptr(), out_scores.data_ptr(),
                                  out_n.data_ptr(), user_ids_ptr=uptr(i % N_ROT))
            elif protocol == "local":
                step_local_protocol(i, out_rows, out_scores, out_n)
            else:
                eng.shard_sample_dev(Qg.data_ptr(), Bg, k, world, samp_local.data_ptr())
                with torch.cuda.stream(stream):
                    dist.all_gather_into_tensor(samp_all, samp_local)          # exchange 1: G x Bg x r sample keys
                eng.shard_candidates_dev(Qg.data_ptr(), Bg, k, world, samp_all.data_ptr(), cand_local.data_ptr())
                if protocol == "a2a":
                    eng.shard_pack_owner_dev(cand_local.data_ptr(), Bg, B, k, cand_packed.data_ptr())
                    with torch.cuda.stream(stream):
                        dist.all_to_all_single(cand_recv, cand_packed)         # exchange 2: each rank gets ITS requests' lists
                    eng.shard_check_owner_dev(cand_recv.data_ptr(), world, B, k, rank * B, retry.data_ptr())
                    eng.recommend_from_keys_dev(cand_recv.data_ptr(), world, chunk, B, k, MODEL_FM_MLP, p,
                                                out_rows.data_ptr(), out_scores.data_ptr(), out_n.data_ptr(),
                                                user_ids_ptr=uptr(i % N_ROT))
                else:
                    with torch.cuda.stream(stream):
                        dist.all_gather_into_tensor(cand_all, cand_local)      # exchange 2: candidates that reach tau
                    eng.shard_check_dev(cand_all.data_ptr(), world, Bg, k, retry.data_ptr())
                    eng.recommend_from_keys_dev(cand_all.data_ptr() + rank * B * k * 8, world, blk, B, k, MODEL_FM_MLP, p,
                                                out_rows.data_ptr(), out_scores.data_ptr(), out_n.data_ptr(),
                                                user_ids_ptr=uptr(i % N_ROT))

    def sync_all():
        eng.sync()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # ---- warm-up, then K timed steps (device-resident inputs): `value`
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # the poller process needs a moment to come up: start it before the warm-up
    for i in range(max(warmup, N_ROT)):   # every rotated batch is touched once before timing
        step_device(i)
    sync_all()
    roof_stage = "gather_fm" if kind == "fm" else "scan"
    eng.timing(1 if kind == "fm" else 2)   # CUDA-event spans around the roofline kernel only while the step is timed
    eng.timing(1 if kind == "fm" else 2, read=True)  # reset accumulators
    launches0 = eng.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    sync_all()
    if rank == 0:
        sampler.begin()
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(args.steps):
            step_device(i)
            ev[i + 1].record(stream)
    sync_all()
    if rank == 0:
        sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    stage = eng.timing(1, read=True)   # spans of the timed region; now switch to all stages
    launches = eng.launches - launches0
    for i in range(20):                # untimed: per-stage breakdown (the extra events perturb the step slightly)
        step_device(i)
    sync_all()
    stage_all = eng.timing(0, read=True)
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    value = Bg * args.steps / (total_ms * 1e-3)

    # ---- e2e: the same step through the C ABI with HOST buffers: pinned host inputs, results out, copies inside
    def pinned(t):
        return t.cpu().pin_memory()
    q_host = [pinned(Qs[i] if world > 1 else Qs[i][:B]) for i in range(N_ROT)]
    u_host = [pinned(Us[i]) for i in range(N_ROT)] if U else None
    if kind == "recall_sort":
        rows_h = torch.empty(B, k, dtype=torch.int32).pin_memory()
        sc_h = torch.empty(B, k, dtype=torch.float32).pin_memory()
        n_h = torch.empty(B, dtype=torch.int32).pin_memory()
        sc64_h = torch.empty(B, k, dtype=torch.float64).pin_memory()
        perm_h = torch.empty(B, k, dtype=torch.int32).pin_memory()

        def step_host(i):   # prg_recall_topk + prg_sort_desc, PRG_MEM_HOST
            rc = lib.prg_recall_topk(h, C.c_void_p(q_host[i % N_ROT].data_ptr()), B, k, C.c_void_p(rows_h.data_ptr()),
                                     C.c_void_p(sc_h.data_ptr()), C.c_void_p(n_h.data_ptr()), 0)
            assert rc == 0, lib.prg_last_error()
            sc64_h.copy_(sc_h)          # Item.Score = float64(score): what the Go glue does per item
            rc = lib.prg_sort_desc(h, C.c_void_p(sc64_h.data_ptr()), B, k, C.c_void_p(perm_h.data_ptr()), 0)
            assert rc == 0, lib.prg_last_error()
        h2d, d2h = B * w["dim"] * 4 + B * k * 8, B * k * 8 + B * 4 + B * k * 4
        note = "prg_recall_topk + prg_sort_desc with host buffers: H2D queries / scores, D2H rows, scores, counts, permutation"
    elif kind == "fm":
        cand_h = [pinned(c) for c in cand]
        fm_h = torch.empty(B, k, dtype=torch.float64).pin_memory()

        def step_host(i):   # prg_rank(PRG_MEM_HOST)
            rc = lib.prg_rank(h, MODEL_FM, C.c_void_p(cand_h[i % N_ROT].data_ptr()), B, k, C.c_void_p(fm_h.data_ptr()), 0)
            assert rc == 0, lib.prg_last_error()
        h2d, d2h = B * k * 4, B * k * 8
        note = "prg_rank(PRG_MODEL_FM) with host buffers: H2D candidate rows, gather + FM, D2H f64 scores, per step"
    else:
        rows_h = torch.empty(B, Tn, dtype=torch.int32).pin_memory()
        sc_h = torch.empty(B, Tn, dtype=torch.float64).pin_memory()
        n_h = torch.empty(B, dtype=torch.int32).pin_memory()
        if world == 1:
            def step_host(i):   # prg_recommend_ex(PRG_MEM_HOST): H2D, all stages, D2H inside the C ABI call
                uf = UserFeatures(u_host[i % N_ROT].data_ptr(), None) if U else None
                rc = lib.prg_recommend_ex(h, C.c_void_p(q_host[i % N_ROT].data_ptr()), B, k, MODEL_FM_MLP, C.byref(p),
                                          C.byref(uf) if U else None, C.c_void_p(rows_h.data_ptr()),
                                          C.c_void_p(sc_h.data_ptr()), C.c_void_p(n_h.data_ptr()), 0)
                assert rc == 0, lib.prg_last_error()
            h2d = B * w["dim"] * 4 + B * U * 4
            note = ("prg_recommend_ex with host buffers: H2D of the queries and user ids, all stages, D2H of "
                    "rows/scores/counts, per step")
        else:
            def step_host(i):   # sharded path: every rank receives the global request batch from its host, returns its results
                j = i % N_ROT
                with torch.cuda.stream(stream):
                    Qs[j].copy_(q_host[j], non_blocking=True)
                    if U:
                        Us[j].copy_(u_host[j], non_blocking=True)
                step_device(i)
                with torch.cuda.stream(stream):
                    rows_h.copy_(out_rows, non_blocking=True)
                    sc_h.copy_(out_scores, non_blocking=True)
                    n_h.copy_(out_n, non_blocking=True)
                eng.sync()
            h2d = Bg * w["dim"] * 4 + Bg * U * 4
            note = ("pinned host queries + user ids -> device, sharded recall + exchanges + rank/sort/DPP, results -> pinned "
                    "host, per step and per rank")
        d2h = B * Tn * 12 + B * 4
    for i in range(3):
        step_host(i)
    sync_all()
    lat = []
    t0 = time.perf_counter()
    for i in range(args.steps):
        t1 = time.perf_counter()
        step_host(i)
        lat.append((time.perf_counter() - t1) * 1e3)
    sync_all()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e = {"value": Bg * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "note": note}

    # ---- the same path entered ONE REQUEST PER HOST THREAD (what the reference's request goroutines do) through the
    #      cross-call batcher: closed loop of 3*B native client threads, each a blocking prg_batcher_recommend call
    batcher = None
    if world == 1 and kind == "full" and not args.no_batcher:
        from pairec_b200.binding import Batcher
        if U:
            eng.set_user_fields(0, 0)          # the native load generator drives query-only requests
            W0, b0 = mlp_weights_np([w["n_fields"] * 16] + w["mlp"][1:])
            eng.set_mlp([w["n_fields"] * 16] + w["mlp"][1:], W0, b0)
        bat = Batcher(eng, k, MODEL_FM_MLP, p, max_batch=B, max_wait_us=0)
        q_pool = q_host[0].numpy().copy()
        n_thr = 3 * B
        per = max(4, (args.steps * B) // n_thr)
        bat.drive(q_pool, n_thr, 2)                      # warm-up
        st0 = bat.stats()
        lat_us, wall, rows_b, n_b = bat.drive(q_pool, n_thr, per)
        st1 = bat.stats()
        rows_d, _, n_d = eng.recommend(q_pool, k, MODEL_FM_MLP, p)   # the direct batch call on the same queries
        same = bool((rows_b == rows_d).all() and (n_b == n_d).all())
        lat1, _, _, _ = bat.drive(q_pool, 1, 50)          # one caller at a time: unloaded single-request latency
        bat.close()
        nb = st1["batches"] - st0["batches"]
        batcher = {"value": n_thr * per / wall, "unit": UNIT, "client_threads": n_thr, "requests": n_thr * per,
                   "p50_ms": float(np.percentile(lat_us, 50)) / 1e3, "p99_ms": float(np.percentile(lat_us, 99)) / 1e3,
                   "mean_batch": (st1["requests"] - st0["requests"]) / max(1, nb),
                   "unloaded_request_p50_ms": float(np.percentile(lat1, 50)) / 1e3,
                   "unloaded_request_p99_ms": float(np.percentile(lat1, 99)) / 1e3,
                   "answers_equal_direct_batch_call": same,
                   "pipelined": os.environ.get("PRG_BATCHER_PIPELINE", "1") != "0",
                   "note": "prg_batcher_recommend: one request per host thread, coalesced into batches of the fused path "
                           "(host buffers, copies inside; a full batch is enqueued behind the running one when "
                           "pipelined); item-side features only (the load generator sends no user block); latency = "
                           "per request, queueing included"}

    # ---- N>1, untimed: what the sharded NCCL path produced against (a) the exact-local protocol on the same ranks and
    #      (b) the unsharded path over the gathered matrix (c4: it fits beside the shard)
    shard_retries = None
    verify = None
    if world > 1 and kind == "full":
        if protocol in ("global", "a2a"):
            eng.sync()
            rt = retry.clone()
            if protocol == "a2a":   # the check ran per owner: the ranks' counts add up
                dist.all_reduce(rt, op=dist.ReduceOp.SUM)
            shard_retries = int(rt.cpu()[1].item())   # queries that asked for the exact protocol (expected: 0)
        if not args.no_verify:
            step_device(0)
            eng.sync()
            got = (out_rows.clone(), out_scores.clone(), out_n.clone())
            verify = {}
            if protocol in ("global", "a2a"):
                keys_local = torch.empty(Bg, k, dtype=torch.int64, device=dev)
                keys_all = torch.empty(world, Bg, k, dtype=torch.int64, device=dev)
                a_rows, a_sc, a_n = torch.empty_like(got[0]), torch.empty_like(got[1]), torch.empty_like(got[2])
                step_local_protocol(0, a_rows, a_sc, a_n)
                eng.sync()
                ok = bool(torch.equal(a_rows, got[0]) and torch.equal(a_sc, got[1]) and torch.equal(a_n, got[2]))
                verify["equals_exact_local_protocol"] = ok
            full_bytes = w["items"] * w["dim"] * 4
            if full_bytes <= 8e9:
                shard_rows = w["items"] // world
                parts = [torch.empty(shard_rows, w["dim"], device=dev) for _ in range(world)]
                if w["items"] % world == 0:
                    dist.all_gather(parts, T["E"])
                    if rank == 0:
                        full = torch.cat(parts)
                        del parts
                        eng1 = Engine(local_rank)
                        T1 = dict(T)
                        T1["E"], T1["row_base"] = full, 0
                        load_engine(eng1, w, T1, MEM_DEVICE)
                        b_rows, b_sc, b_n = torch.empty_like(got[0]), torch.empty_like(got[1]), torch.empty_like(got[2])
                        eng1.recommend_dev(Qs[0].data_ptr(), B, k, MODEL_FM_MLP, p, b_rows.data_ptr(), b_sc.data_ptr(),
                                           b_n.data_ptr(), user_ids_ptr=uptr(0))
                        eng1.sync()
                        verify["rank0_equals_unsharded_path"] = bool(torch.equal(b_rows, got[0]) and
                                                                     torch.equal(b_sc, got[1]) and torch.equal(b_n, got[2]))
                        eng1.close()
            vt = torch.tensor([1 if all(verify.values()) else 0], device=dev)
            dist.all_reduce(vt, op=dist.ReduceOp.MIN)
            verify["all_ranks_ok"] = bool(vt.item())
    def finish():
        # Tensors that NCCL used on the library's stream are recorded against it by torch's allocator: freeing them after
        # the Engine (and its stream) has gone aborts the interpreter at exit.  Leave in a defined order instead.
        sys.stdout.flush()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(0)

    if rank != 0:
        finish()

    pk, pk_kind = peaks()
    rs = stage[roof_stage]
    roof_ms = rs["ms"] / max(1, rs["spans"])
    cfg_env = json.loads(os.environ.get("PRG_CFG", "{}"))
    if kind == "fm":
        M = B * k
        F = w["n_fields"]
        alg_bytes = M * (F * 4 + F * 64 + F * 4 + 8)      # ids + 64-B factor rows + linear weights + f64 score
        kernel = "gather_fm_kernel<8,4>"
        note = ("algorithmic bytes per candidate = F*4 (field ids) + F*64 (factor rows) + F*4 (linear weights) + 8 (score); "
                "SURVEY §8(d) C3")
        extra = {}
    else:
        rows_local = T["E"].shape[0]
        filt = "ffma2" if cfg_env.get("scan_ffma2") else ("tf32" if cfg_env.get("scan_tf32") else "bf16")
        try:   # what the last full pass actually ran (prg_recall_filter): int8 index at dim 64 and <= 64 queries per pass
            used = eng.recall_stats()["filter"]
            if used in ("ffma2", "tf32", "bf16", "int8"):
                filt = used
        except Exception:
            pass
        q_pass = min(Bg, 256 if filt != "ffma2" else 64)
        elem = {"int8": 1, "bf16": 2}.get(filt, 4)
        row_side = {"int8": 8, "ffma2": 0}.get(filt, 4)          # per-row filter parameters: {a_r, hl_r} / norm bound
        alg_bytes = rows_local * w["dim"] * elem + rows_local * row_side + q_pass * w["dim"] * 4
        grp16 = int(os.environ.get("PRG_SCAN_GRP16", cfg_env.get("scan_grp16", 1))) != 0
        kernel = {"int8": (f"recall_scan_i8_kernel (dim {w['dim']}, int8 shadow index)" if q_pass <= 64 else
                           f"recall_scan_grp_kernel<int8> (dim {w['dim']}, int8 shadow index, GROUP mode)"),
                  "bf16": (f"recall_scan_grp_kernel<bf16> (dim {w['dim']}, bf16 shadow index, GROUP mode)"
                           if (w["dim"] == 64 and q_pass > 128 and grp16) else
                           f"recall_scan_tc_kernel<{w['dim']},NQB,bf16 shadow index>"),
                  "tf32": f"recall_scan_tc_kernel<{w['dim']},NQB,tf32 on fp32 rows>",
                  "ffma2": f"recall_scan_kernel<{w['dim']},THRESH>"}[filt]
        note = ("algorithmic bytes = what one pass must stream: rows*dim (int8 filter index) + rows*8 (row scale and bound), "
                "rows*dim*2 (bf16 filter index) or rows*dim*4 (tf32 / ffma2 over the fp32 rows) + rows*4 row norms, + queries; "
                "the fp32 matrix is only touched for the ~5-7 k survivors per query (exact re-score); see DESIGN.md 3.1")
        extra = {"queries_per_pass": q_pass, "passes_per_step": -(-Bg // q_pass),
                 "fp32_matrix_bytes_per_ms": rows_local * w["dim"] * 4 / roof_ms if roof_ms > 0 else 0.0}
    achieved = alg_bytes / (roof_ms * 1e-3) / 1e9 if roof_ms > 0 else 0.0
    roof = {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / pk["hbm_gbs"]}
    if kind != "fm" and roof_ms > 0:
        # the same pass as tensor-core work: 2 * rows * dim * queries flops per launch against the sustained bf16 rate.
        # Whichever fraction is larger names the bound: C4 (64-d, 64 queries per pass) streams, C5 (128-d, 256 queries per
        # pass: 819 GFLOP per 3.2 GB) computes.
        flops = 2.0 * rows_local * w["dim"] * q_pass
        tf = flops / (roof_ms * 1e-3) / 1e12
        # (kind::i8 covers 32 K-elements per instruction where kind::f16 covers 16, in the same cycles: the int8 pass is
        # held against twice the measured bf16 rate — not measured separately)
        tpeak = pk["bf16_tflops_sustained"] * (2.0 if filt == "int8" else 1.0)
        tfrac = tf / tpeak
        extra.update({"hbm": {"achieved_gbs": achieved, "peak_gbs": pk["hbm_gbs"], "frac": achieved / pk["hbm_gbs"]},
                      "tensor": {"achieved_tflops": tf, "peak_tflops": tpeak, "frac": tfrac, "flops_per_launch": flops,
                                 "peak_note": "2 x bf16_tflops_sustained (int8 operands)" if filt == "int8" else "bf16_tflops_sustained"}})
        if tfrac > roof["frac"]:
            roof.update({"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tfrac})
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            traffic = json.load(open(tp)).get(args.workload if kind == "fm" else
                                              (f"scan_dim{w['dim']}_int8" if filt == "int8" else f"scan_dim{w['dim']}"))
        except Exception:
            traffic = None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": total_ms / args.steps, "p50_ms_per_step": float(np.percentile(step_ms, 50)),
            "p99_ms_per_step": float(np.percentile(step_ms, 99)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": {"recall_sort": "f32 (recall), f64 (sort)", "fm": "f32 (FM)",
                      "full": "f32 (recall, FM), bf16 / bf16x2 -> f32 (MLP), f64 (sort, DPP)"}[kind],
            "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
            "stage_ms_per_step": {s: stage_all[s]["ms"] / 20 for s in stage_all},
            "roofline": dict(roof, **{"traffic": traffic, "peak_source": f"MEASURED_PEAKS.json ({pk_kind})",
                                      "launch_ms": roof_ms, "algorithmic_bytes_per_launch": alg_bytes, "note": note}, **extra),
            "e2e": e2e}
    if kind == "full":
        # DPP is not HBM- or tensor-bound: per selection step a CTA of the pair kernel re-reads the fp64 features of
        # its 512 candidates from shared memory (positions 4-7 of 16) and tensor memory (positions 8-15)
        steps_dpp = w["top_n"]
        waves = -(-B * 2 // 148)                      # 2 CTAs per request
        feat = 512 * 129 * 8
        smem_b, tmem_b = feat * 4 // 16, feat * 8 // 16
        dops = 512 * 129 * 2                          # one DMUL + one DADD per feature element and step
        clk = (clocks or {}).get("sm_mhz") or 1900.0
        cyc = smem_b / 128.0 + tmem_b / 870.0 + dops / 64.0
        t_bound = waves * steps_dpp * cyc / (clk * 1e6) * 1e3
        line["dpp_bound"] = {"kernel": "dpp_pair_kernel", "bound": "on-chip feature re-read + fp64 issue, per selection step",
                             "per_step_per_cta": {"shared_bytes": smem_b, "tensor_memory_bytes": tmem_b, "fp64_ops": dops},
                             "rates_per_clk_per_sm": {"shared_bytes": 128, "tensor_memory_read_bytes": 870, "fp64_ops": 64},
                             "cycles_per_step": cyc, "selection_steps": steps_dpp, "waves": waves,
                             "bound_ms": t_bound, "measured_ms": stage_all["dpp"]["ms"] / 20,
                             "frac": t_bound / max(1e-9, stage_all["dpp"]["ms"] / 20),
                             "note": "sum of the three on-chip phases of a step (every warp is in the same phase at the same "
                                     "time); tensor-memory read rate MEASURED on this part (tools/ubench_tmem_ld.cu, 16 warps, "
                                     "profiles/r02_ubench_tmem_ld.txt: 870 B/clk/SM, not the 64 B/clk of the microarchitecture "
                                     "notes used in earlier lines).  The rest of a step is latency: 4 warps per scheduler and a "
                                     "serial arg-max / record exchange between the two CTAs of a request; DESIGN.md 3.6"}
    knobs = {kk: v for kk, v in sorted(os.environ.items()) if kk.startswith("PRG_")}
    if knobs:
        line["knobs"] = knobs   # experiment switches read by the library (A/B lines describe themselves)
    if batcher is not None:
        line["e2e_batcher"] = batcher
    if shard_retries is not None:
        line["shard_retry_queries"] = shard_retries
    if verify is not None:
        line["multi_gpu_parity"] = verify
    if not args.no_cpu_baseline and world == 1:
        try:
            r = cpu_measure(w, 1, 0, args.cpu_requests)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"], "stages_s": r["stages_s"]}
        except Exception as ex:  # the baseline must not take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
