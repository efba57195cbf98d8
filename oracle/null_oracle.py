"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY (nothing in the product package may import this file).

float64 restatement, with dense kernel matrices, of the null model of principal-vector similarity:
`sample_random_vector_`, `compute_random_direction_`, `null_model_similarity`
(/root/reference/sobolev_alignment/sobolev_alignment.py:1731-1810).  Random numbers are drawn with the
reference's torch calls in the reference's order (float32 `randn` / `uniform_`), everything after that runs
in float64.

PARITY STATUS: parity unpinned against running reference code -- `SobolevAlignment` cannot be imported here
(its module imports scvi / anndata at import time) and the reference has no test or fixture for these
functions.  The restatement follows the cited lines statement by statement.
"""

from __future__ import annotations

import numpy as np
import torch

from oracle import falkon_oracle as fo


def sample_random_vector(K, M_mat, n_factors):
    """sobolev_alignment.py:1731-1765.  K: dense [m x m] float64; M_mat: M_X or M_Y."""
    m = K.shape[0]
    coef = torch.randn(m, n_factors)
    sv = torch.linalg.svd(torch.as_tensor(np.asarray(M_mat), dtype=torch.float32))[1]
    norms = torch.FloatTensor(n_factors).uniform_(torch.sqrt(torch.min(sv)), torch.sqrt(torch.max(sv)))
    c = coef.double()
    for j in range(n_factors):
        for i in range(j):
            c[:, j] = c[:, j] - (c[:, i] @ K @ c[:, j]) * c[:, i]
        c[:, j] = c[:, j] / torch.sqrt(c[:, j] @ K @ c[:, j])
    kn = torch.sqrt(torch.diag(c.T @ K @ c))
    return c / kn * norms.double()


def null_model(anchors_s, anchors_t, M_X, M_Y, kernel, sigma, nu=None, n_iter=5, n_factors=None):
    """sobolev_alignment.py:1767-1810 with return_all=True: [n_iter x n_factors] singular values."""
    a_s, a_t = torch.as_tensor(anchors_s), torch.as_tensor(anchors_t)
    K_X = fo.kernel_matrix(a_s, a_s, kernel, sigma, nu)
    K_Y = fo.kernel_matrix(a_t, a_t, kernel, sigma, nu)
    K_XY = fo.kernel_matrix(a_s, a_t, kernel, sigma, nu)
    f_s = a_s.shape[1] if n_factors is None else n_factors      # (sic) :1734 reads the number of genes
    f_t = a_t.shape[1] if n_factors is None else n_factors
    out = []
    for _ in range(n_iter):
        cs = sample_random_vector(K_X, M_X, f_s)
        ct = sample_random_vector(K_Y, M_Y, f_t)
        m_xy = (cs.T @ K_XY @ ct).numpy()
        m_x = (cs.T @ K_X @ cs).numpy()
        m_y = (ct.T @ K_Y @ ct).numpy()
        out.append(np.linalg.svd(fo.mat_inv_sqrt(m_x).dot(m_xy).dot(fo.mat_inv_sqrt(m_y)))[1])
    return np.array(out)
