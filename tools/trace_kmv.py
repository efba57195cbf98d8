"""Developer tool: clock64 event log of CTA 0's MMA-issuing thread for one launch of the fused kernel.

events: 1 before / 2 after the wait for a pipeline stage, 3 stage issued (+commit), 4 before / 5 after the
wait for the epilogue (P ready), 6 P V issued.   python tools/trace_kmv.py [SAB_DEBUG]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
if len(sys.argv) > 1:
    nat.set_option("SAB_DEBUG", sys.argv[1])
a, b, p, d = 148 * 128 * 2, 50000, 3000, 20
A = torch.randn(a, p, device=dev) * 0.2
B = torch.randn(b, p, device=dev) * 0.2
PA, PB = nat.prep(A, None, None, 1.0), nat.prep(B, None, None, 1.0)
del A, B
V = torch.randn(b, d, device=dev)
out = torch.zeros(a, d, device=dev)
nat.kmv(1, 0.1, PA, PB, V, out)
torch.cuda.synchronize()
trace = torch.zeros(4000, dtype=torch.int64, device=dev)
nat.set_option("SAB_TRACE_PTR", str(trace.data_ptr()))
nat.kmv(1, 0.1, PA, PB, V, out)
torch.cuda.synchronize()
ev = trace.cpu().view(-1, 2).tolist()
ev = [(int(t), int(c)) for t, c in ev if t != 0]
t0 = ev[0][1]
# per-stage: wait time (2 - 1), issue time (3 - 2); boundaries: 4..6
waits, issues, bound = [], [], []
last = {}
rows = []
for typ, c in ev:
    rows.append((typ, c - t0))
    last[typ] = c
for i in range(len(rows) - 1):
    pass
print("first 120 events (type, cycles since start, delta):")
prev = 0
for typ, c in rows[:120]:
    print(f"  {typ} {c:9d} +{c - prev}")
    prev = c
# summary over the steady state (skip the first 2 tiles)
import collections
acc = collections.defaultdict(list)
for (t1, c1), (t2, c2) in zip(rows, rows[1:]):
    acc[(t1, t2)].append(c2 - c1)
print("transition: count mean max")
for k, v in sorted(acc.items()):
    v2 = v[10:] if len(v) > 20 else v
    print(f"  {k}: {len(v)} {sum(v2) / len(v2):.0f} {max(v2)}")
print("total cycles", rows[-1][1], "events", len(rows))
