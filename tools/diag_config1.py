import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
from conftest import load_golden, relerr
from oracle import falkon_oracle as fo
from sobolev_alignment_b200 import synthetic as syn
from sobolev_alignment_b200.engine import CudaOps, KernelSpec
ops = CudaOps()
g = load_golden("config1")
Xall = syn.log_counts_numpy(2050, 500, seed=0)
X = torch.from_numpy(Xall[:2000]); C = X[g["idx"]]
spec = KernelSpec("laplacian", 10.0)
frame = ops.make_frame(C, spec)
print("gscale", frame.gscale)
Xp, Cp = ops.prepare(X, frame), ops.prepare(C, frame)
K = ops.empty(2000, 500)
ops.kdense(spec, frame, Xp, Cp, K)
Kt = fo.kernel_matrix(X, C, "laplacian", 10.0)
Xr, Cr = Xp.data[:, :500].double().cpu(), Cp.data[:, :500].double().cpu()
d2r = fo.sq_dists(Xr, Cr, direct=True)
Kr = torch.exp(-d2r.sqrt() / (10.0 * frame.gscale))
E = (K.cpu().double() - Kr).abs()
i = int(E.argmax()); r, c = divmod(i, 500)
print("GPU vs rounded-exact: max", float(E.max()), "at", r, c, "d2r", float(d2r[r, c]), "Kr", float(Kr[r, c]), "K", float(K[r, c]))
print("rounded-exact vs true:", float((Kr - Kt).abs().max()))
print("count |E|>1e-5:", int((E > 1e-5).sum()), " of which d2r<1:", int(((E > 1e-5) & (d2r < 1)).sum()))
srt = torch.sort(d2r.flatten())[0]
print("smallest nonzero d2r:", [float(x) for x in srt[srt > 0][:5]], "zeros:", int((d2r == 0).sum()))
gen = torch.Generator().manual_seed(5)
V = torch.normal(0, 1, size=(500, 10), generator=gen)
Vp = ops.zeros(500, 12); Vp[:, :10] = V.cuda()
out = ops.zeros(2000, 12)
ops.kmv(spec, frame, Xp, Cp, Vp, out)
print("kmv vs golden", relerr(out[:, :10], g["lap_kmv"]), "kmv vs rounded-exact", relerr(out[:, :10], Kr @ V.double()),
      "dense@V vs golden", relerr(K.cpu().double() @ V.double(), g["lap_kmv"]))
