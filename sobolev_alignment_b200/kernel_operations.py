"""
Small dense helpers of the alignment step (/root/reference/sobolev_alignment/kernel_operations.py:12-16).
d x d matrices, float64 on the host.
"""

from __future__ import annotations

import numpy as np


def mat_inv_sqrt(M, threshold: float = 1e-6):
    """M^{-1/2} of a symmetric PSD matrix through its SVD; singular values <= threshold are dropped."""
    M = np.asarray(M.detach().cpu().numpy() if hasattr(M, "detach") else M, dtype=np.float64)
    u, s, vt = np.linalg.svd(M)
    inv_root = np.where(s > threshold, 1.0 / np.sqrt(np.where(s > threshold, s, 1.0)), 0.0)
    return (vt.T * inv_root[None, :]) @ u.T
