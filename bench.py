#!/usr/bin/env python
"""bench.py — requests/s and p99 ms of the recall -> gather+rank -> sort -> DPP hot path on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic requests.

  python bench.py --gpus 1 --steps K --warmup W            the CUDA path (libpairec_gpu.so through its C ABI)
  python bench.py --impl reference --gpus N ...            the reference's CPU path (oracle port) on the host cores

N=1 workload ("c4"): BASELINE.json configs[3] — 10 M items x 64-d f32 recall top-1000, 32 categorical tables
(1 M rows x 16-d) gather + FM + MLP 512-512-256-128-1 (bf16 tensor cores) rank, score sort, DPP top-50 (window 10) on
128-d diversity embeddings, batch 64 requests.  The recall stage alone is configs[1]; it is the dominant kernel and
the one the roofline object describes.

N>1 (torchrun, one rank per GPU): the item matrix is row-sharded (10 M / N rows per rank), every rank scans its shard
for the global batch of 64*N queries, ONE all-gather (NCCL) exchanges the per-shard top-k keys, and each rank merges,
ranks and re-ranks its own 64 requests with replicated feature/diversity tables: per-GPU work is constant -> "weak".

Timing: W >= 3 warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the library's
stream, max over ranks.  The 2.56 GB item matrix is far larger than the 126 MB L2, so every step streams it from
HBM (config.l2 = "inputs larger than L2").
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

WORKLOADS = {
    # name: (items, dim, k, batch, n_fields, table_rows, mlp dims, div_dim, top_n, window)
    "c4": dict(items=10_000_000, dim=64, k=1000, batch=64, n_fields=32, table_rows=1_000_000,
               mlp=[512, 512, 256, 128, 1], div_dim=128, top_n=50, window=10),
    # BASELINE.json config 5 (not yet run: needs 8 GPUs): 100 M x 128-d row-sharded, 128 requests per GPU = 1024 at N = 8;
    # ~90 GB per GPU while the tables are generated (replicated fields / diversity tables)
    "c5": dict(items=100_000_000, dim=128, k=1000, batch=128, n_fields=32, table_rows=1_000_000,
               mlp=[512, 512, 256, 128, 1], div_dim=128, top_n=50, window=10),
    "small": dict(items=400_000, dim=64, k=1000, batch=64, n_fields=32, table_rows=50_000,
                  mlp=[512, 512, 256, 128, 1], div_dim=128, top_n=50, window=10),
}
METRIC = "requests/sec (10M-item recall->rank->DPP)"
UNIT = "requests/s"


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
    """Synthetic tables generated ON the device with torch (plumbing).  Seeds follow SURVEY §8d."""
    import torch
    g = torch.Generator(device=dev)
    items, dim = w["items"], w["dim"]
    shard = items // world
    row_base = rank * shard
    if rank == world - 1:
        shard = items - row_base
    g.manual_seed(2 + 1000 * rank)
    E = torch.randn(shard, dim, device=dev, generator=g) / dim ** 0.5
    g.manual_seed(5)
    F, R = w["n_fields"], w["table_rows"]
    u = torch.rand(items, F, device=dev, generator=g)
    fields = (u * u * u * R).to(torch.int32).clamp_(0, R - 1)  # skewed ids (hot head), replicated on every rank
    del u
    g.manual_seed(4)
    factors = torch.randn(F, R, 16, device=dev, generator=g) * 0.4
    linear = torch.randn(F, R, device=dev, generator=g) * 0.05
    g.manual_seed(7)
    D = torch.randn(items, w["div_dim"], device=dev, generator=g)
    D /= D.norm(dim=1, keepdim=True)
    return dict(E=E, row_base=row_base, fields=fields, factors=factors, linear=linear, D=D)


def mlp_weights(dims, seed=6):
    import oracle
    rng = np.random.default_rng(seed)
    W, b = [], []
    for l in range(len(dims) - 1):
        lim = np.sqrt(6.0 / (dims[l] + dims[l + 1]))
        W.append(oracle.f32_to_bf16(rng.uniform(-lim, lim, size=(dims[l + 1], dims[l])).astype(np.float32)))
        b.append((rng.standard_normal(dims[l + 1]) * 0.01).astype(np.float32))
    return W, b


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


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_pipeline(w, T, Q, sample_rows, threads):
    """The oracle port of the path on the host cores, over the first `sample_rows` catalog rows; returns seconds per
    stage for one batch.  Recall time scales linearly with rows, the other stages do not depend on the catalog size."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    B, k = Q.shape[0], w["k"]
    E = T["E"][:sample_rows]
    t0 = time.perf_counter()
    keys = oracle.recall_topk(E, Q, k, n_threads=threads)
    t_recall = time.perf_counter() - t0
    rows, _, _ = oracle.keys_split(keys)
    t0 = time.perf_counter()
    fm, x = oracle.gather_fm(T["fields"], T["factors"], T["linear"], 0.05, rows.reshape(-1), want_x=True)
    t_gather = time.perf_counter() - t0
    t0 = time.perf_counter()
    ml = oracle.mlp_forward(x, w["mlp"], T["W"], T["b"])
    sc = oracle.sigmoid((fm + ml).astype(np.float32)).astype(np.float64).reshape(B, k)
    t_mlp = time.perf_counter() - t0

    def one(b):
        perm = oracle.stable_sort_desc(sc[b])
        r = rows[b][perm]
        idx, st = oracle.dpp_request(T["D"][r].astype(np.float64), sc[b][perm], w["top_n"], alpha=1.0,
                                     window_size=w["window"])
        return r[idx]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=max(1, min(threads, B))) as ex:
        out = list(ex.map(one, range(B)))
    t_dpp = time.perf_counter() - t0
    return dict(recall=t_recall, gather_fm=t_gather, mlp=t_mlp, sort_dpp=t_dpp), out


def cpu_tables(w, sample_rows):
    """Host tables for the CPU arm (numpy, same distributions as the device tables; the CPU arm is a timing arm)."""
    rng = np.random.default_rng(2)
    F, R = w["n_fields"], min(w["table_rows"], 200_000)
    E = (rng.standard_normal((sample_rows, w["dim"]), dtype=np.float32) / np.float32(w["dim"] ** 0.5))
    u = rng.random((sample_rows, F), dtype=np.float32)
    fields = np.minimum((u * u * u * R).astype(np.uint32), R - 1)
    factors = [(rng.standard_normal((R, 16), dtype=np.float32) * np.float32(0.4)) for _ in range(F)]
    linear = [(rng.standard_normal(R, dtype=np.float32) * np.float32(0.05)) for _ in range(F)]
    D = rng.standard_normal((sample_rows, w["div_dim"]), dtype=np.float32)
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    W, b = mlp_weights_np(w["mlp"])
    return dict(E=E, fields=fields, factors=factors, linear=linear, D=D, W=W, b=b)


def cpu_measure(w, steps, warmup, sample_rows):
    import oracle
    oracle.build()
    threads = oracle.num_threads()
    T = cpu_tables(w, sample_rows)
    rng = np.random.default_rng(3)
    Q = (rng.standard_normal((w["batch"], w["dim"]), dtype=np.float32) / np.float32(w["dim"] ** 0.5))
    per = []
    for i in range(warmup + steps):
        st, _ = cpu_pipeline(w, T, Q, sample_rows, threads)
        if i >= warmup:
            per.append(st)
    scale = w["items"] / sample_rows
    tot = [p["recall"] * scale + p["gather_fm"] + p["mlp"] + p["sort_dpp"] for p in per]
    med = float(np.median(tot))
    stages = {k: float(np.median([p[k] for p in per])) for k in per[0]}
    stages["recall_scaled_to_full_catalog"] = stages["recall"] * scale
    return dict(value=w["batch"] / med, ms_per_step=med * 1e3, cores=threads, stages_s=stages,
                sample=f"batch {w['batch']} requests; recall over the first {sample_rows} of {w['items']} rows "
                       f"(time scaled x{scale:.1f}), gather+FM / MLP / sort+DPP at full size on tables of "
                       f"{min(w['table_rows'], 200_000)} rows; oracle port (C, AVX2 FMA), all host threads")


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PRG_WORKLOAD", "c4"), choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batcher", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)
    config = {"workload": f"{args.workload}: {w['items']} items x {w['dim']}-d f32 recall top-{w['k']} -> "
                          f"{w['n_fields']}-table gather + FM + MLP {'-'.join(map(str, w['mlp']))} (bf16x2 act, bf16 W) "
                          f"rank -> score sort -> DPP top-{w['top_n']} (window {w['window']}, {w['div_dim']}-d f32 table, "
                          f"fp64 arithmetic)",
              "batch_per_gpu": w["batch"], "global_batch": w["batch"] * world,
              "sharding": ("item matrix row-sharded; " + ("all-gather of per-shard top-k keys"
                           if os.environ.get("PRG_SHARD_PROTOCOL", "global") == "local" else
                           "global threshold: all-gather of per-shard sample keys, then all-gather of the candidates "
                           "that reach it") if world > 1 else "none"),
              "l2": f"inputs larger than L2 (item matrix / its {w['items'] // world * w['dim'] * 2 / 1e9:.2f} GB bf16 filter "
                    f"index streamed per step)"}

    if args.impl == "reference":
        if rank != 0:
            return
        sample_rows = min(args.cpu_sample_rows, w["items"])
        r = cpu_measure(w, max(1, min(args.steps, 3)), 1, sample_rows)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": max(1, min(args.steps, 3)), "warmup": 1, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 recall/FM, f64 MLP-acc/DPP",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                 "sample": r["sample"], "stages_s": r["stages_s"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from pairec_b200 import DppParams, Engine
    from pairec_b200.binding import MEM_DEVICE, MODEL_FM_MLP
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local_rank)
    T = make_tables_torch(w, dev, rank, world)
    eng.set_item_matrix(T["E"].data_ptr(), rows=T["E"].shape[0], dim=w["dim"], row_base=T["row_base"], mem=MEM_DEVICE)
    eng.set_item_fields(T["fields"].data_ptr(), rows=w["items"], n_fields=w["n_fields"], mem=MEM_DEVICE)
    for t in range(w["n_fields"]):
        eng.set_feature_table(t, T["factors"][t].data_ptr(), T["linear"][t].data_ptr(), rows=w["table_rows"], fdim=16,
                              mem=MEM_DEVICE)
    eng.set_fm_bias(0.05)
    W, b = mlp_weights_np(w["mlp"])
    eng.set_mlp(w["mlp"], W, b)
    eng.set_diversity_matrix(T["D"].data_ptr(), rows=w["items"], dim=w["div_dim"], dtype=0, mem=MEM_DEVICE)
    p = DppParams(top_n=w["top_n"], alpha=1.0, window_size=w["window"])

    B, k, Tn = w["batch"], w["k"], w["top_n"]
    Bg = B * world
    gq = torch.Generator(device=dev)
    gq.manual_seed(3)
    Qg = torch.randn(Bg, w["dim"], device=dev, generator=gq) / w["dim"] ** 0.5   # same global batch on every rank
    out_rows = torch.empty(B, Tn, dtype=torch.int32, device=dev)
    out_scores = torch.empty(B, Tn, dtype=torch.float64, device=dev)
    out_n = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    protocol = os.environ.get("PRG_SHARD_PROTOCOL", "global")   # "global": one threshold per query across shards
    if world > 1 and protocol == "local":
        keys_local = torch.empty(Bg, k, dtype=torch.int64, device=dev)
        keys_all = torch.empty(world, Bg, k, dtype=torch.int64, device=dev)
    elif world > 1:
        r_s = eng.shard_sample_len(k)
        samp_local = torch.empty(Bg, r_s, dtype=torch.int64, device=dev)
        samp_all = torch.empty(world, Bg, r_s, dtype=torch.int64, device=dev)
        blk = Bg * k + Bg                                        # per rank: Bg x k keys + Bg status words
        cand_local = torch.empty(blk, dtype=torch.int64, device=dev)
        cand_all = torch.empty(world, blk, dtype=torch.int64, device=dev)
        retry = torch.zeros(2, dtype=torch.int32, device=dev)

    def step_device():
        if world == 1:
            eng.recommend_dev(Qg.data_ptr(), B, k, MODEL_FM_MLP, p, out_rows.data_ptr(), out_scores.data_ptr(),
                              out_n.data_ptr())
        elif protocol == "local":
            eng.recall_local_keys_dev(Qg.data_ptr(), Bg, k, keys_local.data_ptr())
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(keys_all, keys_local)
            eng.recommend_from_keys_dev(keys_all.data_ptr() + rank * B * k * 8, world, Bg * k, B, k, MODEL_FM_MLP, p,
                                        out_rows.data_ptr(), out_scores.data_ptr(), out_n.data_ptr())
        else:
            eng.shard_sample_dev(Qg.data_ptr(), Bg, k, world, samp_local.data_ptr())
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(samp_all, samp_local)          # exchange 1: G x Bg x r sample keys
            eng.shard_candidates_dev(Qg.data_ptr(), Bg, k, world, samp_all.data_ptr(), cand_local.data_ptr())
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(cand_all, cand_local)          # exchange 2: candidates that reach tau
            eng.shard_check_dev(cand_all.data_ptr(), world, Bg, k, retry.data_ptr())
            eng.recommend_from_keys_dev(cand_all.data_ptr() + rank * B * k * 8, world, blk, B, k, MODEL_FM_MLP, p,
                                        out_rows.data_ptr(), out_scores.data_ptr(), out_n.data_ptr())

    def sync_all():
        eng.sync()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # ---- warm-up, then K timed steps (device-resident inputs): `value`
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # the poller process needs a moment to come up: start it before the warm-up
    for _ in range(warmup):
        step_device()
    sync_all()
    eng.timing(2)             # CUDA-event spans around the recall scan only while the step is timed
    eng.timing(2, read=True)  # reset accumulators
    launches0 = eng.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    sync_all()
    if rank == 0:
        sampler.begin()
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(args.steps):
            step_device()
            ev[i + 1].record(stream)
    sync_all()
    if rank == 0:
        sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    stage = eng.timing(1, read=True)   # scan spans of the timed region; now switch to all stages
    launches = eng.launches - launches0
    for _ in range(20):                # untimed: per-stage breakdown (the extra events perturb the step slightly)
        step_device()
    sync_all()
    stage_all = eng.timing(0, read=True)
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    value = Bg * args.steps / (total_ms * 1e-3)

    # ---- e2e: the same step with HOST buffers: pinned host queries in, results out, copies inside the timed region
    e2e = None
    rows_h = torch.empty(B, Tn, dtype=torch.int32).pin_memory()
    sc_h = torch.empty(B, Tn, dtype=torch.float64).pin_memory()
    n_h = torch.empty(B, dtype=torch.int32).pin_memory()
    if world == 1:
        q_host = torch.empty(B, w["dim"], dtype=torch.float32).pin_memory()
        q_host.copy_(Qg.cpu())
        import ctypes as C
        lib, h = eng._lib, eng._h

        def step_host():   # prg_recommend(PRG_MEM_HOST): H2D, all stages, D2H inside the C ABI call
            rc = lib.prg_recommend(h, C.c_void_p(q_host.data_ptr()), B, k, MODEL_FM_MLP, C.byref(p),
                                   C.c_void_p(rows_h.data_ptr()), C.c_void_p(sc_h.data_ptr()),
                                   C.c_void_p(n_h.data_ptr()), 0)
            assert rc == 0, lib.prg_last_error()
        h2d = B * w["dim"] * 4
        note = "prg_recommend with host buffers: H2D of the queries, all stages, D2H of rows/scores/counts, per step"
    else:
        q_host = torch.empty(Bg, w["dim"], dtype=torch.float32).pin_memory()
        q_host.copy_(Qg.cpu())

        def step_host():   # sharded path: every rank receives the global query batch from its host, returns its 64 results
            with torch.cuda.stream(stream):
                Qg.copy_(q_host, non_blocking=True)
            step_device()
            with torch.cuda.stream(stream):
                rows_h.copy_(out_rows, non_blocking=True)
                sc_h.copy_(out_scores, non_blocking=True)
                n_h.copy_(out_n, non_blocking=True)
            eng.sync()
        h2d = Bg * w["dim"] * 4
        note = ("pinned host queries -> device, sharded recall + all-gather + rank/sort/DPP, results -> pinned host, "
                "per step and per rank")
    for _ in range(3):
        step_host()
    sync_all()
    lat = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t1 = time.perf_counter()
        step_host()
        lat.append((time.perf_counter() - t1) * 1e3)
    sync_all()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e = {"value": Bg * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": B * Tn * 12 + B * 4, "p50_ms": float(np.percentile(lat, 50)),
           "p99_ms": float(np.percentile(lat, 99)), "note": note}

    # ---- the same path entered ONE REQUEST PER HOST THREAD (what the reference's request goroutines do) through the
    #      cross-call batcher: closed loop of 3*B native client threads, each a blocking prg_batcher_recommend call
    batcher = None
    if world == 1 and not args.no_batcher:
        from pairec_b200.binding import Batcher
        bat = Batcher(eng, k, MODEL_FM_MLP, p, max_batch=B, max_wait_us=0)
        q_pool = q_host.numpy().copy()
        n_thr = 3 * B
        per = max(4, (args.steps * B) // n_thr)
        bat.drive(q_pool, n_thr, 2)                      # warm-up
        st0 = bat.stats()
        lat_us, wall, rows_b, n_b = bat.drive(q_pool, n_thr, per)
        st1 = bat.stats()
        step_host()                                       # the direct batch call on the same queries, for the check
        same = bool((rows_b.view(np.int32)[:, :] == rows_h.numpy()).all() and (n_b == n_h.numpy()).all())
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
                           "pipelined); latency = per request, queueing included"}

    shard_retries = None
    if world > 1 and protocol == "global":
        eng.sync()
        shard_retries = int(retry.cpu()[1].item())   # queries that asked for the exact protocol (expected: 0)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_kind = peaks()
    scan = stage["scan"]
    scan_ms = scan["ms"] / max(1, scan["spans"])
    rows_local = T["E"].shape[0]
    # the full-matrix pass: tensor-core filter over the bf16 shadow index (default), the fp32 rows (scan_tf32) or the
    # exact FFMA2 scan (scan_ffma2); algorithmic bytes per launch = the operand the pass must stream + the row norms
    cfg_env = json.loads(os.environ.get("PRG_CFG", "{}"))
    filt = "ffma2" if cfg_env.get("scan_ffma2") else ("tf32" if cfg_env.get("scan_tf32") else "bf16")
    q_pass = min(Bg, 256 if filt != "ffma2" else 64)
    elem = 2 if filt == "bf16" else 4
    alg_bytes = rows_local * w["dim"] * elem + (rows_local * 4 if filt != "ffma2" else 0) + q_pass * w["dim"] * 4
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    kernel = {"bf16": f"recall_scan_tc_kernel<{w['dim']},NQB,bf16 shadow index>",
              "tf32": f"recall_scan_tc_kernel<{w['dim']},NQB,tf32 on fp32 rows>",
              "ffma2": f"recall_scan_kernel<{w['dim']},THRESH>"}[filt]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch") if tj.get("filter", "tf32") == filt else None
        except Exception:
            traffic = None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": total_ms / args.steps, "p50_ms_per_step": float(np.percentile(step_ms, 50)),
            "p99_ms_per_step": float(np.percentile(step_ms, 99)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (recall, FM), bf16x2/bf16->f32 (MLP), f64 (sort, DPP)",
            "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
            "stage_ms_per_step": {s: stage_all[s]["ms"] / 20 for s in stage_all},
            "roofline": {"kernel": kernel, "bound": "hbm", "achieved": achieved,
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
                         "peak_source": f"MEASURED_PEAKS.json ({pk_kind})", "launch_ms": scan_ms,
                         "algorithmic_bytes_per_launch": alg_bytes, "queries_per_pass": q_pass,
                         "fp32_matrix_bytes_per_ms": rows_local * w["dim"] * 4 / scan_ms if scan_ms > 0 else 0.0,
                         "note": "algorithmic bytes = what this pass must stream: rows*dim*2 (bf16 filter index) or "
                                 "rows*dim*4 (tf32 / ffma2 over the fp32 rows) + rows*4 row norms + queries; the fp32 "
                                 "matrix is only touched for the ~5 k survivors per query (exact re-score); see DESIGN.md 3.1"},
            "e2e": e2e}
    knobs = {k: v for k, v in sorted(os.environ.items()) if k.startswith("PRG_")}
    if knobs:
        line["knobs"] = knobs   # experiment switches read by the library (A/B lines describe themselves)
    if batcher is not None:
        line["e2e_batcher"] = batcher
    if shard_retries is not None:
        line["shard_retry_queries"] = shard_retries
    if not args.no_cpu_baseline and world == 1:
        try:
            r = cpu_measure(w, 1, 0, min(args.cpu_sample_rows, w["items"]))
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"], "stages_s": r["stages_s"]}
        except Exception as ex:  # the baseline must not take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
