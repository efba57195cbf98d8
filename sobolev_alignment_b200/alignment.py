"""
Comparison of two approximated encoders: M_X, M_Y, M_XY, cosine-similarity matrix and principal
vectors (/root/reference/sobolev_alignment/sobolev_alignment.py:929-988).

The reference materialises K(anchors, anchors) on the host (10 GB at M = 50k) only to form
alpha^T K beta.  Here K beta comes out of the fused `kmv` kernel (K never leaves the SM) and the
remaining d x M by M x d product is a tiny matmul.
"""

from __future__ import annotations

import numpy as np
import torch

from .engine import CudaOps, pad_cols
from .kernel_operations import mat_inv_sqrt


def gram(kernel, anchors_a, weights_a, anchors_b, weights_b) -> torch.Tensor:
    """weights_a^T k(anchors_a, anchors_b) weights_b  ->  [d_a x d_b] float64 CPU tensor."""
    ops = CudaOps()
    spec = kernel.spec()
    frame = ops.make_frame(torch.as_tensor(anchors_b), spec)
    same = anchors_a is anchors_b
    B = ops.prepare(torch.as_tensor(anchors_b), frame)
    A = B if same else ops.prepare(torch.as_tensor(anchors_a), frame)
    wa = ops.to_device(torch.as_tensor(weights_a))
    wb = torch.as_tensor(weights_b)
    d_b = wb.shape[1]
    dp = pad_cols(d_b)
    if dp > 32:
        raise NotImplementedError("more than 32 latent dimensions")
    V = ops.zeros(B.n, dp)
    V[:, :d_b] = ops.to_device(wb)
    KV = ops.zeros(A.n, dp)
    ops.kmv(spec, frame, A, B, V, KV, symmetric=same)
    return (wa.double().T @ KV[:, :d_b].double()).cpu()


class EncoderComparison:
    """Holds what `SobolevAlignment._compare_approximated_encoders` / `_compute_principal_vectors` set."""

    def __init__(self, source_krr, target_krr):
        self.krr = {"source": source_krr, "target": target_krr}

    def compare(self):
        s, t = self.krr["source"], self.krr["target"]
        self.M_X = gram(s.kernel_, s.anchors(), s.sample_weights_, s.anchors(), s.sample_weights_)
        self.M_Y = gram(t.kernel_, t.anchors(), t.sample_weights_, t.anchors(), t.sample_weights_)
        # the reference evaluates the cross term with the TARGET kernel object (sobolev_alignment.py:955)
        self.M_XY = gram(t.kernel_, s.anchors(), s.sample_weights_, t.anchors(), t.sample_weights_)
        self.sqrt_inv_M_X_ = mat_inv_sqrt(self.M_X)
        self.sqrt_inv_M_Y_ = mat_inv_sqrt(self.M_Y)
        self.sqrt_inv_matrices_ = {"source": self.sqrt_inv_M_X_, "target": self.sqrt_inv_M_Y_}
        self.cosine_sim = self.sqrt_inv_M_X_.dot(self.M_XY.numpy()).dot(self.sqrt_inv_M_Y_)
        return self

    def principal_vectors(self, all_PVs: bool = False):
        u, s, vt = np.linalg.svd(self.cosine_sim, full_matrices=all_PVs)
        self.principal_angles = s
        self.untransformed_rotations_ = {"source": u, "target": vt.T}
        self.principal_vectors_coef_ = {
            x: self.untransformed_rotations_[x].T.dot(self.sqrt_inv_matrices_[x]).dot(
                self.krr[x].sample_weights_.T.detach().cpu().numpy().astype(np.float64))
            for x in ("source", "target")
        }
        return self


def permutation_principal_angles(X: dict, Y: dict, krr_params: dict, n_permutations: int = 10, seed=None):
    """
    `SobolevAlignment.permutation_test_number_similar_pvs` (sobolev_alignment.py:1698-1729), numeric core: for every
    permutation the GENES (columns) of the source and of the target artificial samples are shuffled independently,
    both KRRs are refitted on the shuffled samples against the unchanged embeddings, and the principal angles
    (cosines) of the resulting pair are collected.  X / Y: {"source": ..., "target": ...} artificial samples [n x p]
    and their latent embeddings [n x d]; krr_params: {"source": kwargs, "target": kwargs} of `KRRApprox`.
    Returns [n_permutations x d] (the reference's `random_principal_angles`); `number_similar_pvs` is then
    `np.sum(principal_angles > np.percentile(out[:, 0], quantile))`.
    """
    from .krr_approx import KRRApprox

    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_permutations):
        krr = {}
        for side in ("source", "target"):
            Xs = torch.as_tensor(X[side])
            perm = torch.from_numpy(rng.permutation(Xs.shape[1])).to(Xs.device)
            krr[side] = KRRApprox(**krr_params[side]).fit(Xs[:, perm], torch.as_tensor(Y[side]))
        out.append(EncoderComparison(krr["source"], krr["target"]).compare().principal_vectors().principal_angles)
    return np.array(out)
