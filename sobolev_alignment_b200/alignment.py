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
