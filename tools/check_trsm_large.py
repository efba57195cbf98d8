"""Developer tool: residual check of the packed triangular solves at M = 100 096 (two chunk columns with the default chunk)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 100096
dp = 20
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
Z = torch.randn(M, 64, device=dev, generator=g)
A = torch.cdist(Z, Z).mul_(-1.0 / 8.0).exp_()
A.diagonal().add_(1e-5 * M)
del Z
inv = torch.empty(M // 128, 128, 128, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
nat.potrf(A, inv, info)
assert int(info) == 0
F = nat.trsm_pack(A, inv)
A.tril_()
B = torch.randn(M, dp, device=dev, generator=g)
V = torch.randn(M, dp, device=dev, generator=g)


def residual(X, rhs, transpose):
    """|op(L) X - rhs| / |rhs| in row blocks (fp32 matmul, fp64 accumulation of the norms)."""
    num = torch.zeros((), dtype=torch.float64, device=dev)
    step = 8192
    for r0 in range(0, M, step):
        r1 = min(M, r0 + step)
        blk = (A[:, r0:r1].T if transpose else A[r0:r1]) @ X          # rows r0:r1 of op(L) X
        num += (blk.double() - rhs[r0:r1].double()).pow(2).sum()
    return float((num / rhs.double().pow(2).sum()).sqrt())


for tr in (False, True):
    X = B.clone()
    nat.trsm(F, X, tr)
    print(f"M={M} single transpose={tr}: relative residual {residual(X, B, tr):.2e}")
    xa, xb = B.clone(), torch.empty_like(B)
    nat.trsm_pair(F, F, xa, xb, V, 0.25, tr)
    print(f"M={M} pair   transpose={tr}: job 0 equals single: {bool(torch.equal(xa, X))}, "
          f"job 1 residual {residual(xb, xa + 0.25 * V, tr):.2e}")
