"""
Null model of principal-vector similarity.

Reference: `SobolevAlignment.null_model_similarity`, `compute_random_direction_`, `sample_random_vector_`
(/root/reference/sobolev_alignment/sobolev_alignment.py:1731-1810).  The reference materialises K_X, K_Y and
K_XY on the host (3 x 10 GB at M = 50k, :1788-1799) and, for each of the n_iter random draws, runs a
Gram-Schmidt orthogonalisation in the K inner product with one `c_i^T K c_j` per pair of factors (:1748-1757).

Here no M x M matrix is formed.  Gram-Schmidt only ever needs `K c` of vectors that are already final:
w_i = K c_i is computed once per factor with the fused `kmv` kernel (symmetric product) and every similarity
is the dot product w_i . c_j.  The draws are independent, so the j-th factors of up to 32 draws ride in the
columns of ONE launch; the cross term C_s^T K_XY C_t packs (draw, factor) pairs into 32-column launches
the same way.  The d x d inverse square roots and SVDs stay on the host, as in the reference.

Random numbers are drawn with the same torch calls in the same order as the reference running with
n_jobs = 1 (randn(M, n_factors), then uniform_ for the norms; source before target; draw after draw), so a
fixed `torch.manual_seed` reproduces the reference's draws.

Note on `n_factors`: the reference reads it from `anchors().shape[1]` (:1734) -- the number of GENES, not of
latent factors.  That is kept as the default for parity; pass `n_factors=krr.sample_weights_.shape[1]` for
the number of latent factors.
"""

from __future__ import annotations

import numpy as np
import torch

from .engine import CudaOps
from .kernel_operations import mat_inv_sqrt

MAX_COLS = 32


class _Side:
    """One data source: prepared anchors, M matrix (for the norm range), and the K-products it needs."""

    def __init__(self, ops, spec, frame, anchors, M_mat):
        self.ops, self.spec, self.frame = ops, spec, frame
        self.prep = ops.prepare(torch.as_tensor(anchors), frame)
        self.m = self.prep.n
        sv = torch.linalg.svd(torch.as_tensor(np.asarray(M_mat), dtype=torch.float32))[1]
        self.norm_lo, self.norm_hi = float(torch.sqrt(sv.min())), float(torch.sqrt(sv.max()))

    def kmul(self, other_prep, cols: torch.Tensor, symmetric: bool) -> torch.Tensor:
        """k(self, other) @ cols for cols [m_other x c] (any c): launches of <= 32 columns."""
        ops = self.ops
        out = ops.zeros(self.m, cols.shape[1])
        for c0 in range(0, cols.shape[1], MAX_COLS):
            c1 = min(cols.shape[1], c0 + MAX_COLS)
            dp = max(4, (c1 - c0 + 3) // 4 * 4)
            V = ops.zeros(cols.shape[0], dp)
            V[:, : c1 - c0] = cols[:, c0:c1]
            part = ops.zeros(self.m, dp)
            ops.kmv(self.spec, self.frame, self.prep, other_prep, V, part, symmetric=symmetric)
            out[:, c0:c1] = part[:, : c1 - c0]
        return out


def _orthonormalise(side: _Side, coef: torch.Tensor, norms: torch.Tensor):
    """
    sobolev_alignment.py:1748-1765 for a group of draws at once.  coef [G x m x f] (device), norms [G x f].
    Returns (coefficients [G x m x f], K @ coefficients [G x m x f]).
    """
    G, m, f = coef.shape
    C = coef.clone()
    W = torch.zeros_like(C)                   # W[:, :, i] = K c_i for the finished factors
    for j in range(f):
        cj = C[:, :, j]
        for i in range(j):                    # modified Gram-Schmidt in the K inner product
            sim = (W[:, :, i] * cj).sum(1, keepdim=True)
            cj -= sim * C[:, :, i]
        u = side.kmul(side.prep, cj.T.contiguous(), symmetric=True).T          # [G x m] = K c_j per draw
        nrm = torch.sqrt((cj * u).sum(1, keepdim=True))
        cj /= nrm
        W[:, :, j] = u / nrm
    # "correct for norm": unit K-norm (up to rounding) -> the sampled factor norms
    kn = torch.sqrt((C * W).sum(1))           # [G x f]
    scale = (norms / kn)[:, None, :]
    return C * scale, W * scale


def null_model_similarity(krr_source, krr_target, M_X, M_Y, n_iter: int = 100, quantile: float = 0.95,
                          return_all: bool = False, n_factors: int | None = None, n_jobs: int = 1, ops=None):
    """
    Null distribution of the principal-vector cosines between random encoders living in the spans of the two
    anchor sets (sobolev_alignment.py:1785-1810): [n_iter x n_factors] singular values if `return_all`, else
    the `quantile` of the top one.  `n_jobs` is accepted for signature compatibility (the draws are batched
    on the device instead of threaded).
    """
    ops = ops or CudaOps()
    spec = krr_target.kernel_.spec()          # the reference evaluates all three kernels with the TARGET kernel object
    a_s, a_t = torch.as_tensor(krr_source.anchors()), torch.as_tensor(krr_target.anchors())
    frame = ops.make_frame(torch.cat([ops.to_device(a_s), ops.to_device(a_t)]), spec)
    src = _Side(ops, spec, frame, a_s, M_X)
    tgt = _Side(ops, spec, frame, a_t, M_Y)
    f_s = a_s.shape[1] if n_factors is None else int(n_factors)
    f_t = a_t.shape[1] if n_factors is None else int(n_factors)
    # random draws in the reference's order: per draw, source (randn, uniform) then target (randn, uniform)
    draws = []
    for _ in range(n_iter):
        cs = torch.randn(src.m, f_s)
        ns = torch.FloatTensor(f_s).uniform_(src.norm_lo, src.norm_hi)
        ct = torch.randn(tgt.m, f_t)
        nt = torch.FloatTensor(f_t).uniform_(tgt.norm_lo, tgt.norm_hi)
        draws.append((cs, ns, ct, nt))
    out = []
    for g0 in range(0, n_iter, MAX_COLS):
        grp = draws[g0: g0 + MAX_COLS]
        Cs, KCs = _orthonormalise(src, ops.to_device(torch.stack([d[0] for d in grp])),
                                  ops.to_device(torch.stack([d[1] for d in grp])))
        Ct, KCt = _orthonormalise(tgt, ops.to_device(torch.stack([d[2] for d in grp])),
                                  ops.to_device(torch.stack([d[3] for d in grp])))
        G = len(grp)
        # K_XY C_t for every (draw, factor) column
        cols = Ct.permute(1, 0, 2).reshape(tgt.m, G * f_t).contiguous()
        KxyCt = src.kmul(tgt.prep, cols, symmetric=False).reshape(src.m, G, f_t).permute(1, 0, 2)
        M_xy = torch.einsum("gmi,gmj->gij", Cs.double(), KxyCt.double()).cpu().numpy()
        M_x = torch.einsum("gmi,gmj->gij", Cs.double(), KCs.double()).cpu().numpy()
        M_y = torch.einsum("gmi,gmj->gij", Ct.double(), KCt.double()).cpu().numpy()
        for k in range(G):
            cos = mat_inv_sqrt(M_x[k]).dot(M_xy[k]).dot(mat_inv_sqrt(M_y[k]))
            out.append(np.linalg.svd(cos)[1])
    out = np.array(out)
    if return_all:
        return out
    return np.quantile(out[:, 0], quantile)
