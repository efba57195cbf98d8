"""
`MultiKRRApprox`: averages several fitted `KRRApprox` (the prototype at
/root/reference/sobolev_alignment/multi_krr_approx.py:14-54).  A pure consumer of
`transform` / `anchors` / `sample_weights_` / `kernel_`.
"""

from __future__ import annotations

import numpy as np
import torch


class MultiKRRApprox:
    def __init__(self):
        self.krr_regressors = []

    def add_clf(self, clf):
        self.krr_regressors.append(clf)

    def predict(self, X: torch.Tensor):
        """Mean of the member predictions."""
        preds = np.stack([clf.transform(torch.Tensor(X)).detach().cpu().numpy() for clf in self.krr_regressors])
        return torch.mean(torch.Tensor(preds), axis=0)

    def transform(self, X: torch.Tensor):
        return self.predict(X)

    def anchors(self):
        return self.anchors_

    def process_clfs(self):
        """Stack anchors and (1/k-scaled) weights so the ensemble is one kernel machine."""
        k = len(self.krr_regressors)
        self.anchors_ = torch.cat([clf.anchors() for clf in self.krr_regressors])
        self.sample_weights_ = torch.cat([clf.sample_weights_ for clf in self.krr_regressors]) / k
        self.kernel_ = self.krr_regressors[0].kernel_
