"""Developer tool: globaltimer stamps of the chain steps of one triangular solve (option SAB_TRSM_TRACE_PTR).

Per block of job 0 (a finishing row): 1 consumers have read the converted x of block c-1 from the low-latency buffer,
2 W tile landed, 3 W product done, 4 after the block-max barrier, 6 low-latency stores of x_c issued (the chain),
5 plain (bulk-copy) publication issued, 7 item start.
    python tools/trace_trsm.py [M] [dp] [transpose]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 50048
dp = int(sys.argv[2]) if len(sys.argv) > 2 else 20
tr = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
Z = torch.randn(M, 64, device=dev, generator=g)
A = torch.cdist(Z, Z).mul_(-1.0 / 8.0).exp_()
A.diagonal().add_(1e-5 * M)
invd = torch.empty(M // 128, 128, 128, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
nat.potrf(A, invd, info)
F = nat.trsm_pack(A, invd)
del A
B = torch.randn(M, dp, device=dev)
nat.trsm(F, B.clone(), tr)
torch.cuda.synchronize()
nblk = M // 128
trace = torch.zeros(nblk * 8, dtype=torch.int64, device=dev)
nat.set_option("SAB_TRSM_TRACE_PTR", str(trace.data_ptr()))
nat.trsm(F, B.clone(), tr)
torch.cuda.synchronize()
nat.set_option("SAB_TRSM_TRACE_PTR", None)
t = trace.cpu().view(nblk, 8).double()
order = list(range(nblk - 1, -1, -1)) if tr else list(range(nblk))
t = t[order]
pub = t[:, 6]
step = (pub[1:] - pub[:-1]) * 1e-3
print(f"M={M} dp={dp} transpose={tr}: {nblk} blocks, total {(pub[-1] - pub[0]) * 1e-6:.3f} ms; chain step us: "
      f"mean {step.mean():.2f} median {step.median():.2f} p90 {step.quantile(0.9):.2f} max {step.max():.2f}")
names = ["LL stores of c-1 issued -> x of c-1 read (LL poll)", "x read -> W tile landed", "W landed -> product done",
         "product -> block max", "block max -> LL stores issued", "item start -> x read"]
rows = t[1:]
prev_pub = pub[:-1]
segs = [rows[:, 1] - prev_pub, rows[:, 2] - rows[:, 1], rows[:, 3] - rows[:, 2], rows[:, 4] - rows[:, 3],
        rows[:, 6] - rows[:, 4], rows[:, 1] - rows[:, 7]]
for nm, sg in zip(names, segs):
    sg = sg * 1e-3
    print(f"  {nm:52s} mean {sg.mean():7.2f} us  median {sg.median():7.2f}  p90 {sg.quantile(0.9):7.2f}")
for lo in range(0, nblk - 1, max(1, nblk // 8)):
    hi = min(nblk - 1, lo + max(1, nblk // 8))
    print(f"  blocks {lo:4d}-{hi:4d}: {(pub[hi] - pub[lo]) * 1e-3 / (hi - lo):.2f} us per step")
