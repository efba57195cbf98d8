"""Developer tool: fused-kernel time at the C3 launch shapes against the column-chunk knobs (SAB_CHUNK_TILES, SAB_L2_BUDGET_MB)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
n, M, p, d = 1_000_000, 50_000, 3000, 20
g = torch.Generator(device=dev).manual_seed(0)
PX = nat.alloc_prepared(n, p, dev)
for s in range(0, n, 100_000):
    blk = torch.randn(100_000, p, device=dev, generator=g) * 0.2
    nat.prep(blk, None, None, 1.0, out=PX, row_offset=s)
C = torch.randn(M, p, device=dev, generator=g) * 0.2
PC = nat.prep(C, None, None, 1.0)
del C, blk
Z = torch.randn(M, d, device=dev, generator=g)
T = torch.randn(n, d, device=dev, generator=g)
t_out = torch.zeros(n, d, device=dev)
w_out = torch.zeros(M, d, device=dev)
fl = 2.0 * n * M * (p + d)


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


modes = sys.argv[1:] or ["tr", "fwd"]
for mode in modes:
    fn = (lambda: nat.kmv(1, 0.1, PC, PX, T, w_out)) if mode == "tr" else (lambda: nat.kmv(1, 0.1, PX, PC, Z, t_out))
    for chunk in (32, 64, 96, 128, 192):
        for l2 in (16, 32, 48, 64):
            nat.set_option("SAB_CHUNK_TILES", chunk)
            nat.set_option("SAB_L2_BUDGET_MB", l2)
            ms = timed(fn)
            print(f"{mode} chunk_tiles={chunk:3d} l2_budget={l2:2d} MB: {ms:7.2f} ms  {fl / ms / 1e9:6.0f} TFLOP/s", flush=True)
