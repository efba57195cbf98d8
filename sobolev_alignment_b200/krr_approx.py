"""
`KRRApprox`: the kernel-ridge-regression plugin boundary of Sobolev Alignment, API-compatible with
/root/reference/sobolev_alignment/krr_approx.py:48-406 (constructor :83-94, fit :192, transform :290,
anchors :286, save :320, load :369; attributes `kernel_`, `sample_weights_`, `training_data_`,
`ridge_clf_`).

method="falkon" runs the B200-native Nystrom solver of this package (no `falkon` install needed, no
CPU path).  method="sklearn" is the reference's exact-KRR alternative on scikit-learn: it is kept
because callers select it by name, and it is never used as a substitute for the falkon method.
"""

from __future__ import annotations

import gc
import os
import pickle

import numpy as np
import torch

from .falkon import Falkon
from .falkon.kernels import GaussianKernel, LaplacianKernel, MaternKernel
from .falkon.options import FalkonOptions

FALKON_IMPORTED = True  # the reference's module-level flag (krr_approx.py:34-36)

_PERSISTED_PARAMS = ("method", "M", "penalization", "mean_center", "unit_std")


def _sklearn_kernel_factory(name: str, params: dict):
    from sklearn.gaussian_process.kernels import Matern, PairwiseKernel

    if name == "matern":
        return Matern(**params)
    return PairwiseKernel(metric=name, **params)


class KRRApprox:
    """Approximates an encoder (x_hat -> z_hat) by kernel ridge regression."""

    sklearn_kernel = {"rbf": "wrapper", "gaussian": "wrapper", "laplacian": "wrapper", "matern": "Matern"}
    falkon_kernel = {"rbf": GaussianKernel, "gaussian": GaussianKernel, "laplacian": LaplacianKernel,
                     "matern": MaternKernel}
    default_kernel_params = {
        "falkon": {"rbf": {"sigma": 1}, "gaussian": {"sigma": 1}, "laplacian": {"sigma": 1},
                   "matern": {"sigma": 1, "nu": 0.5}},
        "sklearn": {"rbf": {}, "gaussian": {}, "laplacian": {}, "matern": {}},
    }

    def __init__(self, method: str = "sklearn", kernel: str = "rbf", M: int = 100, kernel_params: dict = None,
                 penalization: float = 1e-6, maxiter: int = 20, falkon_options: dict = None,
                 mean_center: bool = False, unit_std: bool = False):
        self.method = method
        self.kernel = kernel
        if self.method.lower() not in self.default_kernel_params:
            raise NotImplementedError("%s not implemented. Choices: sklearn and falkon" % (self.method))
        self.kernel_params = kernel_params if kernel_params else self.default_kernel_params[self.method][self.kernel]
        self._make_kernel()
        self.penalization = penalization
        self.M = M
        self.maxiter = maxiter
        self.falkon_options = falkon_options if falkon_options else {}
        self.mean_center = mean_center
        self.unit_std = unit_std
        self.pre_process_ = None
        if mean_center or unit_std:
            self.pre_process_ = self._new_scaler()

    def _new_scaler(self):
        from sklearn.preprocessing import StandardScaler

        return StandardScaler(with_mean=self.mean_center, with_std=self.unit_std, copy=True)

    # ------------------------------------------------------------------ set-up
    def _make_kernel(self):
        m, k = self.method.lower(), self.kernel.lower()
        if m == "sklearn":
            self.kernel_ = _sklearn_kernel_factory(k, self.kernel_params) if k in self.sklearn_kernel else None
            if self.kernel_ is None:
                raise KeyError(k)
        elif m == "falkon":
            self.kernel_ = self.falkon_kernel[k](**self.kernel_params)
        else:
            raise NotImplementedError("%s not implemented. Choices: sklearn and falkon" % (self.method))
        return True

    def _setup_clf(self):
        m = self.method.lower()
        if m == "sklearn":
            return self._setup_sklearn_clf()
        if m == "falkon":
            return self._setup_falkon_clf()
        raise NotImplementedError("%s not implemented. Choices: sklearn and falkon" % (self.method))

    def _setup_sklearn_clf(self):
        from sklearn.kernel_ridge import KernelRidge

        self.ridge_clf_ = KernelRidge(kernel="precomputed", alpha=self.penalization)
        return True

    def _setup_falkon_clf(self):
        self.ridge_clf_ = Falkon(kernel=self.kernel_, penalty=self.penalization, M=self.M, maxiter=self.maxiter,
                                 options=FalkonOptions(**self.falkon_options))
        return True

    # ------------------------------------------------------------------ fit / transform
    def fit(self, X: torch.Tensor, y: torch.Tensor):
        """Regress every column of y on X.  X, y: samples in rows; never modified."""
        self._setup_clf()
        if self.pre_process_ is not None:
            self.pre_process_.fit(X)
            data = torch.Tensor(self.pre_process_.transform(torch.Tensor(X)))
        else:
            data = X
        self.training_data_ = data
        self.training_label_ = y
        if self.method == "sklearn":
            self.ridge_clf_.fit(self.kernel_(self.training_data_), y)
            self.sample_weights_ = torch.Tensor(self.ridge_clf_.dual_coef_)
            self.ridge_samples_idx_ = np.arange(self.training_data_.shape[0])
        elif self.method == "falkon":
            self.ridge_clf_.fit(self.training_data_, y)
            self.sample_weights_ = self.ridge_clf_.alpha_
            self.training_data_ = self.ridge_clf_.ny_points_
        gc.collect()
        return self

    def anchors(self):
        """Points the kernel machine is expanded on (Nystrom centres for falkon, all samples for sklearn)."""
        return self.training_data_

    def transform(self, X: torch.Tensor):
        """Out-of-sample extension: predicted embedding of X (samples in rows)."""
        if self.pre_process_ is not None:
            X = torch.Tensor(self.pre_process_.transform(X))
        if self.method == "sklearn":
            return self.ridge_clf_.predict(self.kernel_(X, self.training_data_))
        if self.method == "falkon":
            return self.ridge_clf_.predict(X)
        raise NotImplementedError("%s not implemented. Choices: sklearn and falkon" % (self.method))

    # ------------------------------------------------------------------ persistence
    def save(self, folder: str = "."):
        """params.pkl + sample_{anchors,weights}.{pt,csv}, the reference's file set (krr_approx.py:320-367)."""
        os.makedirs(folder, exist_ok=True)
        params = {"method": self.method, "kernel": self.kernel_, "M": self.M, "penalization": self.penalization,
                  "mean_center": self.mean_center, "unit_std": self.unit_std}
        params.update(self.kernel_params)
        with open(os.path.join(folder, "params.pkl"), "wb") as f:
            pickle.dump(params, f)
        anchors = torch.Tensor(self.anchors())
        weights = torch.Tensor(self.sample_weights_)
        with open(os.path.join(folder, "sample_anchors.pt"), "wb") as f:
            torch.save(anchors, f)
        with open(os.path.join(folder, "sample_weights.pt"), "wb") as f:
            torch.save(weights, f)
        np.savetxt(os.path.join(folder, "sample_weights.csv"), weights.detach().cpu().numpy())
        np.savetxt(os.path.join(folder, "sample_anchors.csv"), anchors.detach().cpu().numpy())
        return True

    def load(folder: str = "."):
        """Rebuild an instance from `save` output; usable for `transform` without refitting."""
        with open(os.path.join(folder, "params.pkl"), "rb") as f:
            params = pickle.load(f)
        clf = KRRApprox(**{k: v for k, v in params.items() if k in _PERSISTED_PARAMS})
        clf.kernel_ = params["kernel"]
        with open(os.path.join(folder, "sample_weights.pt"), "rb") as f:
            clf.sample_weights_ = torch.load(f)
        with open(os.path.join(folder, "sample_anchors.pt"), "rb") as f:
            clf.training_data_ = torch.load(f)
        clf._setup_clf()
        if clf.method.lower() == "falkon":
            clf.ridge_clf_.kernel = clf.kernel_
        clf.ridge_clf_.ny_points_ = clf.training_data_
        clf.ridge_clf_.alpha_ = clf.sample_weights_
        return clf
