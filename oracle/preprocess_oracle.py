"""
TEST INFRASTRUCTURE -- float64 numpy restatement of the reference's input preprocessing chain
(`SobolevAlignment._memmap_log_processing` / `_frobenius_normalisation`,
/root/reference/sobolev_alignment/sobolev_alignment.py:830-902).  Only tests/ may import it.

Parity pin: the StandardScaler step is checked against sklearn's own `StandardScaler` (the class the
reference instantiates at :856) in tests/test_oracle.py; the rest is three numpy lines of the reference.
"""
import numpy as np


class PreprocessOracle:
    def __init__(self, log_input=False, mean_center=False, unit_std=False, frob_norm_source=False):
        self.log_input, self.mean_center, self.unit_std = log_input, mean_center, unit_std
        self.frob_norm_source = frob_norm_source
        self.frob_norm_param = None

    def _frobenius(self, data_source, X):
        # sobolev_alignment.py:885-902
        if not self.frob_norm_source:
            return X
        if data_source == "source":
            self.frob_norm_param = float(np.mean(np.linalg.norm(X, axis=1)))
            return X
        return X * self.frob_norm_param / float(np.mean(np.linalg.norm(X, axis=1)))

    def fit_transform(self, data_source, X):
        X = np.asarray(X, dtype=np.float64)
        if self.log_input:                                   # :851-862
            X = np.log10(X + 1)
            Xraw = X
            # StandardScaler(with_mean, with_std).fit_transform: population variance
            if self.mean_center:
                X = X - X.mean(0, keepdims=True)
            if self.unit_std:
                # sklearn.preprocessing._data: scale_ = sqrt(var_), with features that are constant up to
                # round-off (_is_constant_feature) or have a scale below 10 eps (_handle_zeros_in_scale) left at 1
                n = X.shape[0]
                eps = np.finfo(np.float64).eps
                var = Xraw.var(0)
                mean = Xraw.mean(0)
                constant = var <= n * eps * var + (n * mean * eps) ** 2
                sd = np.sqrt(var)
                sd = np.where(constant | (sd < 10 * eps), 1.0, sd)
                X = X / sd
        return self._frobenius(data_source, X)
