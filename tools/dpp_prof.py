"""clock64 breakdown of the cluster DPP kernel (needs libpairec_gpu_prof.so built with -DPRG_DPP_PROFILE)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PRG_LIB"] = os.path.join(ROOT, "pairec_b200", "libpairec_gpu_prof.so")
import torch, ctypes as C
from pairec_b200 import DppParams, Engine
from pairec_b200.binding import MEM_DEVICE
dev = torch.device("cuda:0")
N, D, n, T = 2_000_000, 128, 1000, 50
B = int(os.environ.get("B", 1))
torch.manual_seed(0)
Dm = torch.randn(N, D, device=dev); Dm /= Dm.norm(dim=1, keepdim=True)
eng = Engine(0)
eng.set_diversity_matrix(Dm.data_ptr(), rows=N, dim=D, dtype=0, mem=MEM_DEVICE)
p = DppParams(top_n=T, alpha=1.0, window_size=10)
rows = torch.randint(0, N, (B, n), device=dev, dtype=torch.int32)
score = torch.rand(B, n, device=dev, dtype=torch.float64)
idx = torch.empty(B, T, dtype=torch.int32, device=dev); cnt = torch.empty(B, dtype=torch.int32, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
out = (C.c_longlong * 16)()
names = ["presort/setup", "2a norm", "2b features", "diag gram", "gram", "update", "argmax shfl", "cta barrier", "publish", "cluster.sync", "pick head", "tail"]
for rep in range(3):
    rc = eng._lib.prg_dpp(eng._h, C.c_void_p(rows.data_ptr()), C.c_void_p(score.data_ptr()), B, n, C.byref(p),
                          C.c_void_p(idx.data_ptr()), C.c_void_p(cnt.data_ptr()), C.c_void_p(st.data_ptr()), MEM_DEVICE)
    assert rc == 0
    eng._lib.prg_debug_dpp_clocks(out, 1)
    tot = sum(out)
    print(f"rep {rep}: total {tot} clk = {tot/1.965e3:.1f} us :: " + ", ".join(f"{nm}={out[i]}" for i, nm in enumerate(names)))
