"""One warm-up + two timed launches of the fused kernel at a C3-shaped slice (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
rows = int(os.environ.get("ROWS", 148 * 128 * 2))   # ROWS=1000000: the C3 launch shapes
a, b = (rows, 50000) if mode == "fwd" else (50000, rows)
p, d = 3000, 20


def prepared(n):
    out = nat.alloc_prepared(n, p, dev)
    for s0 in range(0, n, 65536):
        e0 = min(n, s0 + 65536)
        nat.prep(torch.randn(e0 - s0, p, device=dev) * 0.2, None, None, 1.0, out=out, row_offset=s0)
    return out


PA, PB = prepared(a), prepared(b)
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
