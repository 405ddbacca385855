"""Single-GPU emulation of rank 0 of a G-way row-sharded recall (global-threshold protocol): G engines hold the shards
of one 10 M-row catalog; the two all-gathers are emulated by writing every shard's output into its slot.  Times the
recall part of rank 0's step (sample -> tau -> filter -> refine -> check; the merge of the own 64 queries is not included), the part that grows with G.
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel list (NCU=1 runs one repetition)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pairec_b200 import Engine
from pairec_b200.binding import MEM_DEVICE

G = int(os.environ.get("G", 8)); N = int(os.environ.get("N", 10_000_000)); D = int(os.environ.get("D", 64))
B = int(os.environ.get("B", 64)); K = 1000
Bg = B * G
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
shard = N // G
engs, Es = [], []
for r in range(G):
    g.manual_seed(2 + 1000 * r)
    E = torch.randn(shard, D, device=dev, generator=g) / D ** 0.5
    e = Engine(0)
    e.set_item_matrix(E.data_ptr(), rows=shard, dim=D, row_base=r * shard, mem=MEM_DEVICE)
    engs.append(e); Es.append(E)
g.manual_seed(3)
Q = torch.randn(Bg, D, device=dev, generator=g) / D ** 0.5
e0 = engs[0]
rs = e0.shard_sample_len(K)
samp = torch.zeros(G, Bg, rs, dtype=torch.int64, device=dev)
blk = Bg * K + Bg
cand = torch.zeros(G, blk, dtype=torch.int64, device=dev)
retry = torch.zeros(2, dtype=torch.int32, device=dev)
keys = torch.zeros(B, K, dtype=torch.int64, device=dev)
for r, e in enumerate(engs):            # the other ranks' contributions (computed once)
    e.shard_sample_dev(Q.data_ptr(), Bg, K, G, samp[r].data_ptr()); e.sync()
for r, e in enumerate(engs):
    e.shard_candidates_dev(Q.data_ptr(), Bg, K, G, samp.data_ptr(), cand[r].data_ptr()); e.sync()


def rank0_step():
    e0.shard_sample_dev(Q.data_ptr(), Bg, K, G, samp[0].data_ptr())
    e0.shard_candidates_dev(Q.data_ptr(), Bg, K, G, samp.data_ptr(), cand[0].data_ptr())
    e0.shard_check_dev(cand.data_ptr(), G, Bg, K, retry.data_ptr())


reps = 1 if os.environ.get("NCU") else 20
for _ in range(1 if os.environ.get("NCU") else 3):
    rank0_step()
e0.sync()
e0.timing(1); e0.timing(1, read=True)
t0 = time.perf_counter()
for _ in range(reps):
    rank0_step()
e0.sync()
ms = (time.perf_counter() - t0) / reps * 1e3
st = e0.timing(0, read=True)
print(f"G={G} Bg={Bg} shard={shard}: rank-0 recall part {ms:.3f} ms/step; retry={retry.cpu().tolist()}")
print({k: round(v["ms"] / reps, 4) for k, v in st.items()})
per = (cand[:, :Bg * K].reshape(G, Bg, K) != 0).sum(dim=2).float().mean().item()
print("candidates per query and shard:", per)
