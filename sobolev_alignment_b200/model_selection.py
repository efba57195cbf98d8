"""
Hyper-parameter grid of KRR fits sharing one dataset.

Reference: `krr_model_selection.py:68-274` -- `model_selection_nu` / `model_alignment_penalization` refit
`KRRApprox(method="falkon")` for every point of `_make_hyperparameters_grid(sigma, M)`
(4 nu x 11 penalizations, :260-266) on the SAME artificial samples (`sample_artificial=False`, :141-147), each
fit starting from scratch: staging and preparing X, K_MM, both Cholesky factorisations, the CG.

Sharing what does not change:

  over the whole grid     the prepared rows X~ (fp16 + norms), the prepared centres, Y on the device
                          -- one H2D pass and one `sa_prep` pass for 44 fits;
  over the penalties of   K_MM, T = chol(K_MM + eps M I) (packed once), G0 = T T^T / M, and
  one kernel (nu)         T^-T K_nM^T (Y / n); per penalty only A = chol(G0 + lambda I) is new;
  inside the CG           K_nM^T K_nM z does not depend on lambda: 32 // dp penalties ride through the two
                          fused-kernel launches of an iteration side by side (multi-lambda CG).

One deviation, deliberate: Falkon draws fresh random centres for every fit (unseeded host RNG); the grid
draws them ONCE (`center_seed`) -- which is also what makes the grid points comparable.
"""

from __future__ import annotations

import time

import numpy as np
import torch

from .engine import CudaOps
from .krr_approx import KRRApprox
from .solver import FalkonSolver


class KRRGrid:
    """Fits `KRRApprox(method="falkon", kernel=..., kernel_params={"sigma", "nu"}, penalization=...)` over a grid."""

    def __init__(self, X: torch.Tensor, Y: torch.Tensor, M: int, sigma: float, kernel: str = "matern",
                 maxiter: int = 20, center_seed: int = 0, falkon_options: dict | None = None, ops=None):
        self.ops = ops or CudaOps()
        self.kernel, self.sigma, self.M, self.maxiter = kernel, float(sigma), int(M), int(maxiter)
        self.falkon_options = dict(falkon_options or {})
        X, Y = torch.as_tensor(X), torch.as_tensor(Y)
        if Y.dim() == 1:
            Y = Y.reshape(-1, 1)
        self.n, self.dtype, self.host = X.shape[0], Y.dtype, not X.is_cuda
        m = min(self.M, self.n)
        idx = np.arange(self.n) if m == self.n else np.sort(
            np.random.default_rng(center_seed).choice(self.n, size=m, replace=False))
        self.centre_idx = idx
        self.centres = X[torch.from_numpy(idx).to(X.device)]
        t0 = time.perf_counter()
        spec = self._kernel_object(1.5).spec()
        C = self.ops.to_device(self.centres)
        self.frame = self.ops.make_frame(C, spec)
        self.Cprep = self.ops.prepare(C, self.frame)
        self.Xprep = self.ops.prepare(X, self.frame)
        self.Yd = self.ops.to_device(Y)
        self.ops.synchronize()
        self.times = {"prepare": time.perf_counter() - t0, "kernels": {}}

    def _kernel_object(self, nu):
        params = {"sigma": self.sigma}
        if self.kernel == "matern":
            params["nu"] = nu
        return KRRApprox.falkon_kernel[self.kernel](**params)

    def fit_kernel(self, nu, penalizations) -> list:
        """All penalties of one kernel; returns the fitted KRRApprox objects in the order of `penalizations`."""
        ops = self.ops
        pens = [float(x) for x in penalizations]
        t0 = time.perf_counter()
        kobj = self._kernel_object(nu)
        from .falkon import FalkonOptions

        opt = FalkonOptions(**self.falkon_options)
        solver = FalkonSolver(ops, kobj.spec(), pens, self.maxiter, pc_eps=opt.pc_eps(),
                              cg_eps=float(opt.cg_epsilon_32), cg_tol=float(opt.cg_tolerance),
                              full_grad_every=int(opt.cg_full_gradient_every))
        solver.n_total, solver.M = self.n, self.Cprep.n
        solver.frame, solver.Cprep, solver.Xprep = self.frame, self.Cprep, self.Xprep
        from .solver import Preconditioner

        solver.pre = Preconditioner(ops, solver.spec, self.frame, self.Cprep, pens, solver.pc_eps)
        solver.pre.check()
        alphas = solver.solve_prepared(self.Yd)
        ops.synchronize()
        out = []
        for lam, alpha, res in zip(pens, alphas, solver.residuals_per_penalty):
            params = {"sigma": self.sigma}
            if self.kernel == "matern":
                params["nu"] = nu
            clf = KRRApprox(method="falkon", kernel=self.kernel, kernel_params=params, M=self.M, penalization=lam,
                            maxiter=self.maxiter, falkon_options=self.falkon_options)
            clf._setup_clf()
            clf.ridge_clf_.alpha_ = alpha.to(self.dtype).cpu() if self.host else alpha.to(self.dtype)
            clf.ridge_clf_.ny_points_ = self.centres
            clf.ridge_clf_.solver_stats_ = {"residuals": res, "n_iter": len(res)}
            clf.sample_weights_ = clf.ridge_clf_.alpha_
            clf.training_data_ = self.centres
            out.append(clf)
        self.times["kernels"][nu] = {"seconds": time.perf_counter() - t0, "n_kmv": solver.n_kmv_}
        solver.release()
        return out

    def fit(self, nus=(0.5, 1.5, 2.5, np.inf), penalizations=tuple(np.logspace(-8, 2, 11))) -> dict:
        """{(nu, penalization): fitted KRRApprox} over the reference's default grid (krr_model_selection.py:24-52,260-266)."""
        grid = {}
        for nu in (nus if self.kernel == "matern" else (None,)):
            for lam, clf in zip(penalizations, self.fit_kernel(nu, penalizations)):
                grid[(nu, float(lam))] = clf
        return grid
