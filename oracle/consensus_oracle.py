"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY (nothing in the product package may import this file).

float64 restatement of the consensus-feature step the reference builds on top of the KRR path:
`SobolevAlignment.compute_consensus_features` (/root/reference/sobolev_alignment/sobolev_alignment.py:1027-1088)
and `interpolated_features.py:12-48` (Grassmann interpolation between matched principal vectors, interpolation
time chosen by the smallest two-sample Kolmogorov-Smirnov statistic over 101 steps).

PARITY STATUS: the two interpolation functions are pinned against the reference's own
`interpolated_features.py`, imported by file path in oracle/make_golden.py (fixture tests/golden/consensus.npz);
the projection step composes the already-pinned oracle pieces (kernel matrices, principal vectors).
"""

from __future__ import annotations

import numpy as np
import scipy.stats
import torch

from oracle import falkon_oracle as fo


def pv_projections(alpha, centres, X, pv_coef, kernel, sigma, nu=None, center=True):
    """
    sobolev_alignment.py:1027-1064.  alpha / centres / pv_coef: dicts over {"source", "target"} (pv_coef[m] is the
    [d x M] matrix rotation^T sqrt_inv alpha^T of :980-987, unused here on purpose: the reference rotates the
    transform output, and so does this restatement); X: dict of row matrices.  Returns the nested dict of
    [n x d] float64 arrays.
    """
    out = {}
    for m in ("source", "target"):
        out[m] = {}
        for data in ("source", "target"):
            K = fo.kernel_matrix(torch.as_tensor(X[data]), torch.as_tensor(centres[m]), kernel, sigma, nu)
            P = (K @ torch.as_tensor(pv_coef[m], dtype=torch.float64).T).numpy()
            if center:
                P = P - P.mean(axis=0, keepdims=True)
            out[m][data] = P
    return out


def interpolate(angle, k, tau, proj):
    """interpolated_features.py:29-48."""
    w0, w1, nrm = np.sin((1.0 - tau) * angle), np.sin(tau * angle), np.sin(angle)
    src = (w0 * proj["source"]["source"][:, k] + w1 * proj["target"]["source"][:, k]) / nrm
    tgt = (w0 * proj["source"]["target"][:, k] + w1 * proj["target"]["target"][:, k]) / nrm
    return src, tgt


def ks_curve(k, proj, angles, n_interpolation=100):
    taus = np.linspace(0, 1, n_interpolation + 1)
    return taus, np.array([scipy.stats.ks_2samp(*interpolate(angles[k], k, t, proj)).statistic for t in taus])


def optimal_tau(k, proj, angles, n_interpolation=100):
    """
    interpolated_features.py:12-26: smallest KS statistic.  The statistic is discrete, ties are common; the
    reference resolves them through `DataFrame.sort_values("ks").iloc[0]`, i.e. numpy's (unstable but
    deterministic) quicksort argsort of the column -- restated literally.
    """
    taus, ks = ks_curve(k, proj, angles, n_interpolation)
    return taus[int(np.argsort(ks, kind="quicksort")[0])]
