"""DPP-only timing probe: prg_dpp (device buffers) at several batch sizes -> per-request latency vs throughput."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pairec_b200 import DppParams, Engine
from pairec_b200.binding import MEM_DEVICE
import ctypes as C

dev = torch.device("cuda:0")
N, D, n, T = 2_000_000, int(os.environ.get("D", 128)), 1000, 50
torch.manual_seed(0)
Dm = torch.randn(N, D, device=dev); Dm /= Dm.norm(dim=1, keepdim=True)
eng = Engine(0)
eng.set_diversity_matrix(Dm.data_ptr(), rows=N, dim=D, dtype=0, mem=MEM_DEVICE)
p = DppParams(top_n=T, alpha=1.0, window_size=10)
for B in (1, 8, 32, 37, 64, 128):
    rows = torch.randint(0, N, (B, n), device=dev, dtype=torch.int32)
    score = torch.rand(B, n, device=dev, dtype=torch.float64)
    idx = torch.empty(B, T, dtype=torch.int32, device=dev); cnt = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
    def run():
        rc = eng._lib.prg_dpp(eng._h, C.c_void_p(rows.data_ptr()), C.c_void_p(score.data_ptr()), B, n, C.byref(p),
                              C.c_void_p(idx.data_ptr()), C.c_void_p(cnt.data_ptr()), C.c_void_p(st.data_ptr()), MEM_DEVICE)
        assert rc == 0
    for _ in range(3): run()
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(20): run()
    eng.sync()
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"B={B:4d}: {ms:.3f} ms per launch, {ms/B*1e3:.1f} us per request amortised")
