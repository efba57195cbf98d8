"""potrf + forward/backward triangular solves at M=16384 (for ncu)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device=dev).manual_seed(1)
Z = torch.randn(M, 64, device=dev, generator=g)
A = (torch.exp(-torch.cdist(Z, Z) / 8.0) + 1e-2 * torch.eye(M, device=dev)).contiguous()
invd = torch.empty(M // 128, 128, 128, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
nat.potrf(A, invd, info)
B = torch.randn(M, 20, device=dev)
torch.cuda.synchronize()
for tr in (False, True, False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); nat.trsm(A, invd, B, tr); e1.record(); torch.cuda.synchronize()
    print(f"trsm trans={tr}: {e0.elapsed_time(e1):.3f} ms ({M // 128} steps, {e0.elapsed_time(e1) * 1e3 / (M // 128):.1f} us/step)")

# paired solves (two pipelined chains in one launch)
G = (torch.exp(-torch.cdist(Z, Z) / 5.0) + 1e-2 * torch.eye(M, device=dev)).contiguous()
invg = torch.empty(M // 128, 128, 128, device=dev)
nat.potrf(G, invg, info)
Xa, Xb = torch.randn(M, 20, device=dev), torch.empty(M, 20, device=dev)
torch.cuda.synchronize()
for tr in (False, True, False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); nat.trsm_pair(A, invd, G, invg, Xa, Xb, None, 0.0, tr); e1.record(); torch.cuda.synchronize()
    print(f"trsm_pair trans={tr}: {e0.elapsed_time(e1):.3f} ms")
wa, wg = nat.trsm_premul(A, invd), nat.trsm_premul(G, invg)
torch.cuda.synchronize()
for tr in (False, True, False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); nat.trsm_pair(A, invd, G, invg, Xa, Xb, None, 0.0, tr, wa, wg); e1.record(); torch.cuda.synchronize()
    print(f"trsm_pair with pre-multiplied tiles trans={tr}: {e0.elapsed_time(e1):.3f} ms")
