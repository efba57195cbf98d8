"""One warm-up + two timed launches of the fused kernel at a C3-shaped slice (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
rows = 148 * 128 * 2
a, b = (rows, 50000) if mode == "fwd" else (50000, rows)
p, d = 3000, 20
A = torch.randn(a, p, device=dev) * 0.2
B = torch.randn(b, p, device=dev) * 0.2
PA, PB = nat.prep(A, None, None, 1.0), nat.prep(B, None, None, 1.0)
del A, B
V = torch.randn(b, d, device=dev)
out = torch.zeros(a, d, device=dev)
for i in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nat.kmv(1, 0.1, PA, PB, V, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{mode} launch {i}: {ms:.2f} ms {2.0 * a * b * (p + d) / ms / 1e9:.1f} TFLOP/s", flush=True)
