"""
Input preprocessing of the artificial samples on the device -- the step immediately before `KRRApprox.fit`
in the reference (`SobolevAlignment._memmap_log_processing` and `_frobenius_normalisation`,
/root/reference/sobolev_alignment/sobolev_alignment.py:830-902).

    log_input:         x <- log10(x + 1)
                       x <- StandardScaler(with_mean=mean_center, with_std=unit_std).fit_transform(x)
    frob_norm_source:  source: remember the mean row norm; target: rescale to the source's mean row norm

The reference does this with numpy on the host (three passes over a 12 GB array, then re-saves it as a
memmap).  Here the counts are streamed to the device once (pinned, chunked -- a read-only memmap is fine, it
is never written) and the chain is two streaming kernels (`sa_colstats`, `sa_transform_rows`) plus one
scalar multiply; the result stays on the device as the CUDA tensor `KRRApprox(method="falkon").fit` takes.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _native as nat


class InputPreprocessor:
    def __init__(self, log_input: bool = False, mean_center: bool = False, unit_std: bool = False,
                 frob_norm_source: bool = False, device=None, chunk_rows: int = 65536):
        if not torch.cuda.is_available():
            raise nat.NativeLibraryError("InputPreprocessor needs a CUDA device (there is no CPU path)")
        nat.load()
        self.log_input, self.mean_center, self.unit_std = bool(log_input), bool(mean_center), bool(unit_std)
        self.frob_norm_source = bool(frob_norm_source)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.chunk_rows = int(chunk_rows)
        self.frob_norm_param_ = None      # SobolevAlignment._frob_norm_param
        self.mean_, self.scale_ = {}, {}  # per data source, like the per-call StandardScaler of the reference

    # ------------------------------------------------------------------ staging
    def _to_device(self, X) -> torch.Tensor:
        if torch.is_tensor(X) and X.is_cuda:
            return X.to(self.device, torch.float32).contiguous().clone()
        if torch.is_tensor(X):
            Xt = X
        else:
            import warnings

            with warnings.catch_warnings():   # a read-only memmap is fine: it is only ever read
                warnings.simplefilter("ignore", UserWarning)
                Xt = torch.as_tensor(np.asarray(X))
        n, p = Xt.shape
        out = torch.empty(n, p, dtype=torch.float32, device=self.device)
        rows = max(1, min(self.chunk_rows, n))
        pinned = [torch.empty(rows, p, dtype=torch.float32).pin_memory() for _ in range(2)]
        events = [None, None]
        stream = torch.cuda.current_stream(self.device)
        for i, s in enumerate(range(0, n, rows)):
            e, b = min(n, s + rows), i & 1
            if events[b] is not None:
                events[b].synchronize()
            pinned[b][: e - s].copy_(Xt[s:e])          # never writes into the caller's (maybe memmap) data
            out[s:e].copy_(pinned[b][: e - s], non_blocking=True)
            events[b] = torch.cuda.Event()
            events[b].record(stream)
        for ev in events:
            if ev is not None:
                ev.synchronize()
        return out

    # ------------------------------------------------------------------ the chain
    def fit_transform(self, data_source: str, X) -> torch.Tensor:
        """X: counts [n x p] (numpy / memmap / CPU or CUDA tensor).  Returns a float32 CUDA tensor [n x p]."""
        if data_source not in ("source", "target"):
            raise ValueError("data_source must be 'source' or 'target'")
        Xd = self._to_device(X)
        n, p = Xd.shape
        shift = scale = None
        if self.log_input and (self.mean_center or self.unit_std) and n > 0:
            mean, var = nat.colstats(Xd, log10p1=True)
            if self.mean_center:
                shift = mean.to(torch.float32).contiguous()
            if self.unit_std:
                # sklearn's StandardScaler: features constant up to round-off, or with a scale below 10 eps, keep scale 1
                eps = torch.finfo(torch.float64).eps
                sd = var.sqrt()
                constant = var <= n * eps * var + (n * mean * eps) ** 2
                sd = torch.where(constant | (sd < 10 * eps), torch.ones_like(sd), sd)
                scale = (1.0 / sd).to(torch.float32).contiguous()
            self.mean_[data_source] = mean if self.mean_center else None
            self.scale_[data_source] = sd if self.unit_std else None
        need_norm = self.frob_norm_source
        if self.log_input or need_norm:
            norm_sum = nat.transform_rows(Xd, log10p1=self.log_input, shift=shift, scale=scale, gain=1.0,
                                          want_row_norms=need_norm)
        if need_norm and n > 0:
            mean_norm = float(norm_sum) / n
            if data_source == "source":
                self.frob_norm_param_ = mean_norm
            else:
                if self.frob_norm_param_ is None:
                    raise RuntimeError("frob_norm_source: the source must be processed before the target")
                nat.axpby(self.frob_norm_param_ / mean_norm, Xd.view(-1), 0.0, Xd.view(-1))
        return Xd
