"""
Consensus features: projections of both datasets on the source and target
principal vectors, and their Grassmann interpolation.

Reference: `SobolevAlignment.compute_consensus_features`
(/root/reference/sobolev_alignment/sobolev_alignment.py:990-1102) and `interpolated_features.py:12-48`.
The reference calls `transform` four times (2 models x 2 datasets, :1027-1036) -- each time staging the
whole dataset -- then rotates every [n x d] projection on the host (:1041-1052) and mean-centres it with a
StandardScaler (:1056-1064).  Here

* each dataset is staged and prepared (fp16 operand + norms) ONCE, in a frame shared by both centre sets,
  one row chunk ahead of the tensor cores (engine.stream_rows);
* the rotation is folded into the coefficient block: U^T S (K alpha)^T = (K (alpha S^T U))^T, so each
  chunk needs one fused `kmv` per model (V = principal_vectors_coef_^T) and no host matmul;
* centring happens on the device; only the four [n x d] results travel back.

`compute_optimal_tau` / `project_on_interpolate_PV` keep the reference's names, arguments and return values;
the 101 two-sample Kolmogorov-Smirnov statistics per principal vector are evaluated together with torch
sort / searchsorted (on the device when one is present).
"""

from __future__ import annotations

import numpy as np
import torch

from .engine import CudaOps, pad_cols
from .solver import stream_rows

SOURCES = ("source", "target")


def pv_projections(krr: dict, X: dict, principal_vectors_coef: dict, center: bool = True,
                   row_chunk: int = 131072, ops=None) -> dict:
    """
    pv_projections[pv_data_source][proj_data_source] = (mean-centred) projection of dataset `proj_data_source`
    on the principal vectors of model `pv_data_source` -- numpy [n x d_pv], the dict layout of
    sobolev_alignment.py:1041-1064.

    krr: {"source": KRRApprox, "target": KRRApprox} (fitted, method="falkon");
    X:   {"source": tensor [n_s x p], "target": tensor [n_t x p]} after the reference's log / Frobenius step;
    principal_vectors_coef: {"source": [d x M_s], "target": [d x M_t]} = untransformed_rotations^T sqrt_inv alpha^T.
    """
    ops = ops or CudaOps()
    specs = {m: krr[m].kernel_.spec() for m in SOURCES}
    centres = {m: torch.as_tensor(krr[m].anchors()) for m in SOURCES}
    # one frame for both centre sets and both datasets: distances do not depend on the shift, and a common
    # gain lets every prepared row be used against either centre set
    ref_rows = torch.cat([ops.to_device(centres[m]) for m in SOURCES])
    frame = ops.make_frame(ref_rows, specs["source"])
    del ref_rows
    if specs["source"].sigma_vec is not None or specs["target"].sigma_vec is not None:
        raise NotImplementedError("per-feature length scales are not supported here")
    Cprep = {m: ops.prepare(centres[m], frame) for m in SOURCES}
    V = {}
    for m in SOURCES:
        coef = torch.as_tensor(np.asarray(principal_vectors_coef[m], dtype=np.float32))   # [d x M]
        d = coef.shape[0]
        dp = pad_cols(d)
        if dp > 32:
            raise NotImplementedError("more than 32 principal vectors")
        blk = ops.zeros(coef.shape[1], dp)
        blk[:, :d] = ops.to_device(coef.T.contiguous())
        V[m] = (d, blk)
    out = {m: {} for m in SOURCES}
    for data in SOURCES:
        Xd = torch.as_tensor(X[data])
        n = Xd.shape[0]
        proj = {m: ops.zeros(n, V[m][1].shape[1]) for m in SOURCES}

        def one_chunk(Xp, s, proj=proj):
            for m in SOURCES:
                ops.kmv(specs[m], frame, Xp, Cprep[m], V[m][1], proj[m][s: s + Xp.n])

        stream_rows(ops, Xd, frame, one_chunk, row_chunk)
        for m in SOURCES:
            P = proj[m][:, : V[m][0]]
            if center:
                P = P - P.mean(0, keepdim=True)
            out[m][data] = ops.to_host(P).numpy()
    return out


def project_on_interpolate_PV(angle, PV_number, tau_step, pv_projections):
    """interpolated_features.py:29-48: source / target samples projected on the PV interpolated at `tau_step`."""
    a, b = np.sin((1 - tau_step) * angle), np.sin(tau_step * angle)
    s = (a * pv_projections["source"]["source"][:, PV_number]
         + b * pv_projections["target"]["source"][:, PV_number]) / np.sin(angle)
    t = (a * pv_projections["source"]["target"][:, PV_number]
         + b * pv_projections["target"]["target"][:, PV_number]) / np.sin(angle)
    return s, t


def ks_statistics(PV_number, pv_projections, principal_angles, n_interpolation=100, device=None) -> np.ndarray:
    """Two-sample KS statistic between the interpolated source and target projections at every tau step."""
    if device is None:
        device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    angle = float(principal_angles[PV_number])
    cols = [torch.as_tensor(np.ascontiguousarray(pv_projections[m][dsrc][:, PV_number]), dtype=torch.float64,
                            device=device) for m in SOURCES for dsrc in SOURCES]
    ss, st, ts, tt = cols       # [pv model][projected data]
    taus = torch.linspace(0, 1, n_interpolation + 1, dtype=torch.float64, device=device)
    a = torch.sin((1 - taus) * angle)[:, None]
    b = torch.sin(taus * angle)[:, None]
    src = (a * ss[None, :] + b * ts[None, :]) / np.sin(angle)      # [steps x n_s]
    tgt = (a * st[None, :] + b * tt[None, :]) / np.sin(angle)      # [steps x n_t]
    src, tgt = src.sort(dim=1).values, tgt.sort(dim=1).values
    n1, n2 = src.shape[1], tgt.shape[1]
    allv = torch.cat([src, tgt], dim=1)
    # empirical CDFs of both samples at every pooled data point (scipy.stats.ks_2samp, two-sided statistic)
    cdf1 = torch.searchsorted(src, allv, right=True).double() / n1
    cdf2 = torch.searchsorted(tgt, allv, right=True).double() / n2
    return (cdf1 - cdf2).abs().max(dim=1).values.cpu().numpy()


def compute_optimal_tau(PV_number, pv_projections, principal_angles, n_interpolation=100):
    """
    interpolated_features.py:12-26: the interpolation step with the smallest source/target KS statistic.
    Ties (the statistic is discrete) fall the way the reference's `sort_values("ks").iloc[0]` makes them
    fall: first element of numpy's quicksort argsort.
    """
    ks = ks_statistics(PV_number, pv_projections, principal_angles, n_interpolation)
    taus = np.linspace(0, 1, n_interpolation + 1)
    return taus[int(np.argsort(ks, kind="quicksort")[0])]


def consensus_features(krr: dict, X: dict, principal_vectors_coef: dict, principal_angles, n_similar_pv: int,
                       optimal_interpolation_step: dict | None = None, row_chunk: int = 131072):
    """
    The numeric core of `compute_consensus_features` (sobolev_alignment.py:1027-1088): returns
    (interpolated [n_s + n_t x n_similar_pv] numpy array, source rows first; optimal step per PV).
    `principal_angles` are the cosines `SobolevAlignment.principal_angles` holds (the reference passes
    np.arccos of them on, :1073,1082).
    """
    proj = pv_projections(krr, X, principal_vectors_coef, center=True, row_chunk=row_chunk)
    angles = np.arccos(np.clip(np.asarray(principal_angles, dtype=np.float64), -1.0, 1.0))
    if optimal_interpolation_step is None:
        optimal_interpolation_step = {k: compute_optimal_tau(k, proj, angles, n_interpolation=100)
                                      for k in range(n_similar_pv)}
    cols = [np.concatenate(list(project_on_interpolate_PV(angles[k], k, tau, proj)))
            for k, tau in optimal_interpolation_step.items()]
    return np.stack(cols, axis=1), optimal_interpolation_step
