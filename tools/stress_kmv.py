"""Developer tool: repeat one small fused product many times and report runs that deviate (race hunting)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402
from sobolev_alignment_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
shapes = [(3000, 1000, 333, 32), (2000, 500, 500, 12), (129, 257, 65, 20), (5000, 3000, 1000, 20)]
for (a, b, p, d) in shapes:
    A = torch.from_numpy(syn.log_counts_numpy(a, p, seed=a + b)).to(dev)
    B = torch.from_numpy(syn.log_counts_numpy(b, p, seed=a + b + 1)).to(dev)
    shift = B.mean(0).contiguous()
    PA, PB = nat.prep(A, shift, None, 1.0), nat.prep(B, shift, None, 1.0)
    dp = nat.round_up(d, 4)
    V = torch.randn(b, dp, device=dev)
    sigma = max(1.0, 0.3 * math.sqrt(p))
    for kid in (1, 3):
        ref = None
        bad = 0
        worst = 0.0
        for i in range(reps):
            out = torch.zeros(a, dp, device=dev)
            nat.kmv(kid, 1.0 / sigma, PA, PB, V, out)
            if i % 7 == 0:   # perturb timing / allocator state between launches
                torch.cuda.synchronize()
            if ref is None:
                ref = out.clone()
                continue
            err = float((out - ref).abs().max() / ref.abs().max())
            worst = max(worst, err)
            if err > 1e-5:
                bad += 1
                if bad <= 3:
                    rows = ((out - ref).abs().max(1).values > 1e-5 * ref.abs().max()).nonzero().flatten()
                    cols = ((out - ref).abs().max(0).values > 1e-5 * ref.abs().max()).nonzero().flatten()
                    print(f"  run {i}: err {err:.3e}; bad rows {rows.numel()} [{rows[:6].tolist()} ...] bad cols {cols.tolist()}")
        print(f"a={a} b={b} p={p} d={d} kid={kid} CG={os.environ.get('SAB_CTA_GROUP', '2')}: {bad}/{reps - 1} deviating runs, worst {worst:.2e}", flush=True)
