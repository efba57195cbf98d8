"""Developer tool: globaltimer stamps of the last chain step of every block in one chained triangular solve."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device=dev).manual_seed(1)
Z = torch.randn(M, 64, device=dev, generator=g)
A = (torch.exp(-torch.cdist(Z, Z) / 8.0) + 1e-2 * torch.eye(M, device=dev)).contiguous()
invd = torch.empty(M // 128, 128, 128, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
nat.potrf(A, invd, info)
B = torch.randn(M, 20, device=dev)
nat.trsm(A, invd, B.clone(), False)
torch.cuda.synchronize()
nblk = M // 128
trace = torch.zeros(8 * nblk, dtype=torch.int64, device=dev)
os.environ["SAB_TRSM_TRACE_PTR"] = str(trace.data_ptr())
nat.trsm(A, invd, B.clone(), False)
torch.cuda.synchronize()
t = trace.cpu().view(nblk, 8).double()
names = ["wait start", "flag seen", "x + tile landed", "update product done", "rhs staged", "diag product done",
         "stores + barrier", "published"]
mid = t[nblk // 4: 3 * nblk // 4]
print("mean ns between consecutive stamps (middle half of the chain):")
for i in range(1, 8):
    print(f"  {names[i - 1]:>22s} -> {names[i]:<22s} {float((mid[:, i] - mid[:, i - 1]).mean()):8.0f}")
pub = t[:, 7]
print(f"publish-to-publish interval: {float((pub[1:] - pub[:-1])[nblk // 4: 3 * nblk // 4].mean()):.0f} ns")
seen_after_pub = t[1:, 1] - t[:-1, 7]
print(f"previous block published -> this block sees the flag: {float(seen_after_pub[nblk // 4: 3 * nblk // 4].mean()):.0f} ns")
