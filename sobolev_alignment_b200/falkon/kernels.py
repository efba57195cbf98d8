"""
`falkon.kernels` look-alike: GaussianKernel, LaplacianKernel, MaternKernel.

The reference constructs them at /root/reference/sobolev_alignment/krr_approx.py:67-70,182, CALLS them
as functions (`kernel_(A, B)` -> dense tensor) at sobolev_alignment.py:950,955-958,1790-1801 and
tests/test_krr_approx.py:255, and pickles the whole object (krr_approx.py:340).  Objects therefore hold
plain Python state only; device operands are built per call.

    Gaussian   exp(-|a-b|^2 / (2 sigma^2))          Laplacian  exp(-|a-b|_2 / sigma)
    Matern     nu=0.5 -> Laplacian, nu=inf -> Gaussian,
               nu=1.5 (1+sqrt3 r) exp(-sqrt3 r), nu=2.5 (1+sqrt5 r+5r^2/3) exp(-sqrt5 r),  r=|a-b|/sigma
"""

from __future__ import annotations

import math

import torch

from ..engine import CudaOps, KernelSpec, pad_cols


def _as_float_or_tensor(sigma):
    if torch.is_tensor(sigma):
        return sigma.detach().clone().to(torch.float32).cpu() if sigma.numel() > 1 else float(sigma.item())
    if isinstance(sigma, (list, tuple)):
        return torch.tensor(sigma, dtype=torch.float32) if len(sigma) > 1 else float(sigma[0])
    return float(sigma)


class Kernel:
    kernel_name = "abstract"

    def __init__(self, sigma, opt=None):
        self.sigma = _as_float_or_tensor(sigma)
        self.params = opt

    # -- what the CUDA epilogue evaluates
    def _family(self) -> str:
        raise NotImplementedError

    def spec(self) -> KernelSpec:
        return KernelSpec(self._family(), self.sigma)

    # -- falkon.kernels.Kernel API
    def __call__(self, X1, X2, diag: bool = False, out=None, opt=None):
        """Dense k(X1, X2) (or its diagonal); result lives where X1 lives."""
        X1, X2 = torch.as_tensor(X1), torch.as_tensor(X2)
        if X1.shape[1] != X2.shape[1]:
            raise ValueError("feature dimensions differ")
        ops = CudaOps()
        spec = self.spec()
        frame = ops.make_frame(X2, spec)
        B = ops.prepare(X2, frame)
        same = X1 is X2 or (X1.data_ptr() == X2.data_ptr() and X1.shape == X2.shape and X1.device == X2.device)
        A = B if same else ops.prepare(X1, frame)
        K = ops.empty(X1.shape[0], X2.shape[0])
        ops.kdense(spec, frame, A, B, K, symmetric=same)
        if diag:
            K = K.diagonal().clone()
        K = K.to(X1.dtype) if X1.dtype in (torch.float64, torch.float32) else K
        res = K if X1.is_cuda else K.cpu()
        if out is not None:
            out.copy_(res)
            return out
        return res

    def mmv(self, X1, X2, v, out=None, opt=None):
        """k(X1, X2) @ v without materialising the kernel matrix."""
        X1, X2, v = torch.as_tensor(X1), torch.as_tensor(X2), torch.as_tensor(v)
        ops = CudaOps()
        spec = self.spec()
        frame = ops.make_frame(X2, spec)
        B = ops.prepare(X2, frame)
        A = ops.prepare(X1, frame)
        res = _kmv_cols(ops, spec, frame, A, B, ops.to_device(v.reshape(v.shape[0], -1)))
        res = res.to(v.dtype)
        res = res if X1.is_cuda else res.cpu()
        if out is not None:
            out.copy_(res)
            return out
        return res

    def dmmv(self, X1, X2, v, w, out=None, opt=None):
        """k(X1, X2)^T (k(X1, X2) v + w)  -- the product Falkon's CG is built on."""
        X1, X2 = torch.as_tensor(X1), torch.as_tensor(X2)
        if v is None and w is None:
            raise ValueError("one of v, w must be given")
        ops = CudaOps()
        spec = self.spec()
        frame = ops.make_frame(X2, spec)
        B = ops.prepare(X2, frame)
        A = ops.prepare(X1, frame)
        ref = v if v is not None else w
        t = None
        if v is not None:
            v = torch.as_tensor(v)
            t = _kmv_cols(ops, spec, frame, A, B, ops.to_device(v.reshape(v.shape[0], -1)))
        if w is not None:
            wd = ops.to_device(torch.as_tensor(w).reshape(X1.shape[0], -1))
            t = wd if t is None else t + wd
        res = _kmv_cols(ops, spec, frame, B, A, t.contiguous()).to(torch.as_tensor(ref).dtype)
        res = res if X1.is_cuda else res.cpu()
        if out is not None:
            out.copy_(res)
            return out
        return res

    def __repr__(self):
        return f"{type(self).__name__}(sigma={self.sigma})"


def _kmv_cols(ops, spec, frame, A, B, V):
    """k(A, B) @ V for any number of columns (32 padded columns per launch)."""
    d = V.shape[1]
    out = ops.empty(A.n, d)
    for c in range(0, d, 32):
        w = min(32, d - c)
        dp = pad_cols(w)
        Vp = ops.zeros(B.n, dp)
        Vp[:, :w] = V[:, c : c + w]
        o = ops.zeros(A.n, dp)
        ops.kmv(spec, frame, A, B, Vp, o)
        out[:, c : c + w] = o[:, :w]
    return out


class GaussianKernel(Kernel):
    kernel_name = "gaussian"

    def _family(self):
        return "gaussian"


class LaplacianKernel(Kernel):
    kernel_name = "laplacian"

    def _family(self):
        return "laplacian"


class MaternKernel(Kernel):
    kernel_name = "matern"
    _valid_nu = (0.5, 1.5, 2.5, math.inf)

    def __init__(self, sigma, nu, opt=None):
        super().__init__(sigma, opt)
        nu = float(nu.item() if torch.is_tensor(nu) else nu)
        if nu not in self._valid_nu:
            raise ValueError(f"The given value of nu = {nu} can only take values {self._valid_nu}.")
        self.nu = nu

    def _family(self):
        return {0.5: "laplacian", 1.5: "matern15", 2.5: "matern25", math.inf: "gaussian"}[self.nu]

    def __repr__(self):
        return f"MaternKernel(sigma={self.sigma}, nu={self.nu})"
