"""
TEST-ONLY stand-in for `sobolev_alignment_b200.engine.CudaOps`: same interface, torch-CPU float64
arithmetic taken from the oracle.  It exists so the HOST logic (FALKON driver, row sharding, gloo
all-reduce, API plumbing) can be exercised without a GPU.  It lives under tests/ and is never
importable from the product package.
"""

import math

import torch

from oracle import falkon_oracle as fo

_FAMILY = {"gaussian": ("gaussian", None), "laplacian": ("laplacian", None),
           "matern15": ("matern", 1.5), "matern25": ("matern", 2.5)}


class _Rows:
    def __init__(self, data):
        self.data, self.n, self.p = data, data.shape[0], data.shape[1]

    def rows(self, r0, r1):
        return _Rows(self.data[r0:r1])


class _CgScratch:
    """Same contract as _native.CgScratch: residual history, stop flag, late polls."""

    def __init__(self, maxiter, dp):
        self.hist = torch.zeros(max(1, maxiter), dp, dtype=torch.float64)
        self.stop = False
        self._polls = {}

    def record(self, row, sums, tol):
        self.hist[row] = sums
        if float(sums.max().clamp_min(0).sqrt()) < tol:
            self.stop = True

    def poll(self, slot):
        self._polls[slot] = self.stop

    def stopped_at(self, slot):
        return self._polls.get(slot, False)

    def history(self):
        return self.hist


class CpuOps:
    name = "cpu-oracle"

    def __init__(self, device=None, stage_rows=65536):
        self.device = torch.device("cpu")

    def zeros(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, dtype=torch.float64 if dtype == torch.float32 else dtype)

    def empty(self, *shape, dtype=torch.float32):
        return self.zeros(*shape, dtype=dtype)

    def to_device(self, t):
        return t.detach().to(torch.float64).contiguous()

    def to_host(self, t):
        return t

    def make_frame(self, ref_rows, spec):
        return None

    def prepare(self, X, frame):
        return _Rows(X.detach().to(torch.float64))

    def _k(self, spec, A, B):
        fam, nu = _FAMILY[spec.family]
        return fo.kernel_matrix(A.data, B.data, fam, spec.sigma, nu, direct=False)

    def kmv(self, spec, frame, A, B, V, out, symmetric=False, kstore=None):
        K = self._k(spec, A, B)
        out += K @ V
        if kstore is not None:
            self._stored = K            # the CUDA engine leaves packed fp16 hi/lo tiles in `kstore`
        return out

    def ktv(self, kstore, a, b, t, out):
        assert self._stored.shape == (a, b)
        out += self._stored.T @ t
        return out

    def kstore_bytes(self, a, b):
        return ((a + 127) // 128) * ((b + 127) // 128) * 65536

    def free_bytes(self):
        return 1 << 40

    def kdense(self, spec, frame, A, B, out, symmetric=False):
        out[: A.n, : B.n] = self._k(spec, A, B)
        return out

    def potrf(self, A, invdiag, info):
        L = torch.linalg.cholesky(torch.tril(A) + torch.tril(A, -1).T)
        A.copy_(L)

    # the factorisation in steps (outer panels dealt out over the ranks); one 128-block per outer panel here
    def potrf_outer_blocks(self):
        return 1

    def potrf_begin(self, A, info):
        info.zero_()

    def potrf_panel(self, A, invdiag, info, k0, kend):
        a, b = k0 * 128, kend * 128
        D = torch.tril(A[a:b, a:b])
        L = torch.linalg.cholesky(D + torch.tril(D, -1).T)
        A[a:b, a:b] = L
        if b < A.shape[0]:
            A[b:, a:b] = torch.linalg.solve_triangular(L, A[b:, a:b].T.contiguous(), upper=False).T

    def potrf_split(self, m, panel, W):
        pass

    def potrf_update(self, A, kend, c0, c1, panel, W):
        off, ncols = (c0 - kend) * 128, (c1 - c0) * 128
        P = panel[off:]
        A[c0 * 128:, c0 * 128: c1 * 128] -= P @ P[:ncols].T

    def potrf_update_cyclic(self, A, kend, nb, first, stride, skip, panel, W):
        nblk = A.shape[0] // 128
        for j in range(first, (nblk + nb - 1) // nb, stride):
            if j != skip:
                self.potrf_update(A, kend, j * nb, min(nblk, j * nb + nb), panel, W)

    def ltl(self, L, G, alpha, beta):
        Lt = torch.tril(L)
        G.copy_(alpha * (Lt.T @ Lt) + beta * torch.eye(L.shape[0], dtype=L.dtype))

    def add_diag(self, A, m, value):
        A.diagonal()[:m] += value

    def trsm_pack(self, L, invdiag):
        return torch.tril(L).clone()      # the CUDA engine returns the packed fp16 hi/lo tiles

    def trsm(self, F, B, transpose):
        B.copy_(torch.linalg.solve_triangular(F.T if transpose else F, B, upper=bool(transpose)))
        return B

    def trsm_pair(self, Fa, Fb, Xa, Xb, V, lam, transpose):
        """Xa <- op(La)^-1 Xa ; Xb <- op(Lb)^-1 (Xa + lam V)  (same contract as _native.trsm_pair)."""
        self.trsm(Fa, Xa, transpose)
        Xb.copy_(Xa if V is None else Xa + lam * V)
        self.trsm(Fb, Xb, transpose)
        return Xb

    def cg_scratch(self, maxiter, dp):
        return _CgScratch(maxiter, dp)

    def coldot(self, P, Q, out, cgs, hist_row=None, tol=0.0):
        out.copy_((P * Q).sum(0))
        if hist_row is not None:
            cgs.record(hist_row, out, tol)
        return out

    def cg_update(self, P, AP, X, R, rs_old, pAp, eps, rs_new, cgs, hist_row=None, tol=0.0):
        if cgs.stop:
            return
        alpha = rs_old / (pAp + eps)
        X += P * alpha[None, :]
        if AP is not None:
            R -= AP * alpha[None, :]
            rs_new.copy_((R * R).sum(0))
            if hist_row is not None:
                cgs.record(hist_row, rs_new, tol)

    def cg_direction(self, R, P, rs_new, rs_old, eps, cgs):
        if cgs.stop:
            return
        beta = torch.where(rs_old > 0, rs_new / (rs_old + eps), torch.zeros_like(rs_old))
        P.copy_(R + P * beta[None, :])

    def axpby(self, a, X, b, Y):
        Y.copy_(a * X + b * Y)
        return Y

    def synchronize(self):
        pass
