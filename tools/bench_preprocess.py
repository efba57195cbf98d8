"""Timing of the device preprocessing chain (sa_colstats + sa_transform_rows) next to the numpy / sklearn chain
the reference runs (sobolev_alignment.py:851-862), on Poisson counts n x 3000."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200.preprocessing import InputPreprocessor  # noqa: E402

n, p = int(sys.argv[1]) if len(sys.argv) > 1 else 400000, 3000
X = torch.poisson(torch.rand(p, device="cuda")[None, :].expand(n, p) * 20).float()
pre = InputPreprocessor(log_input=True, mean_center=True, unit_std=True, frob_norm_source=True)
for name in ("source", "target", "source"):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = pre.fit_transform(name, X)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    # traffic: staging copy (r+w), column statistics (r), transform (r+w) [+ scalar multiply (r+w) for the target]
    passes = 5 + (2 if name == "target" else 0)
    print(f"device-resident counts {n}x{p}: {name:6s} {dt * 1e3:7.1f} ms  ({n * p * 4 * passes / dt / 1e9:.0f} GB/s of HBM traffic)")
    del out
nh = min(n, 100000)
Xh = X[:nh].cpu().numpy()
t = time.perf_counter()
out = pre.fit_transform("source", Xh)
torch.cuda.synchronize()
print(f"host (pageable numpy) counts {nh}x{p} incl. H2D staging: {(time.perf_counter() - t) * 1e3:.0f} ms")
from sklearn.preprocessing import StandardScaler  # noqa: E402

t = time.perf_counter()
Y = np.log10(Xh + 1)
Y = StandardScaler().fit_transform(Y)
f = np.mean(np.linalg.norm(Y, axis=1))
print(f"numpy / sklearn chain of the reference on the same {nh}x{p}: {(time.perf_counter() - t) * 1e3:.0f} ms")
