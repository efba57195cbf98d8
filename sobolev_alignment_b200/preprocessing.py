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


class DeferredRows:
    """
    Rows whose preprocessing has been DECIDED but not applied: raw counts resident on the device plus the
    affine map that follows the log,  y = (log10(x + 1) - shift) * scale * gain.  `KRRApprox(method="falkon").fit`
    / `Falkon.fit` take it in place of the fp32 tensor: the map is folded into the operand preparation
    (`sa_prep_log`: counts -> fp16 rows + norms in ONE pass), so the preprocessed fp32 matrix (12 GB at
    BASELINE config 3) is never written.  Only the M centre rows are materialised (`rows`), because
    `anchors()` has to hand them back.
    """

    is_cuda = True
    dtype = torch.float32

    def __init__(self, counts: torch.Tensor, log10p1: bool, shift, scale, gain: float):
        self.counts, self.log10p1 = counts, bool(log10p1)
        self.shift, self.scale, self.gain = shift, scale, float(gain)

    @property
    def shape(self):
        return self.counts.shape

    @property
    def device(self):
        return self.counts.device

    def dim(self):
        return 2

    def rows(self, idx) -> torch.Tensor:
        """Materialised preprocessed rows idx (float32 CUDA tensor)."""
        idx = torch.as_tensor(idx, device=self.counts.device)
        out = self.counts[idx].contiguous()
        nat.transform_rows(out, self.log10p1, self.shift, self.scale, self.gain)
        return out

    def materialize(self) -> torch.Tensor:
        out = self.counts.clone()
        nat.transform_rows(out, self.log10p1, self.shift, self.scale, self.gain)
        return out

    def folded(self, frame_shift, frame_scale):
        """
        (shift', scale') with  ((f(x) - shift) scale gain - frame_shift) frame_scale == (f(x) - shift') scale'
        per column: what `sa_prep_log` needs next to the frame's power-of-two gain.
        """
        p = self.counts.shape[1]
        dev = self.counts.device
        a = torch.full((p,), self.gain, dtype=torch.float64, device=dev)          # y = (f - shift) a
        if self.scale is not None:
            a = a * self.scale.double()
        sh = torch.zeros(p, dtype=torch.float64, device=dev) if self.shift is None else self.shift.double()
        fs = torch.zeros(p, dtype=torch.float64, device=dev) if frame_shift is None else frame_shift.double()
        sc = a if frame_scale is None else a * frame_scale.double()
        shift = sh + fs / a
        return shift.float().contiguous(), sc.float().contiguous()


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
        return self._run(data_source, X, deferred=False)

    def fit_prepare(self, data_source: str, X) -> DeferredRows:
        """
        Same statistics as `fit_transform` (column mean / variance, mean row norm), but the transform itself is
        deferred: the result feeds `KRRApprox.fit` directly and the chain runs inside the operand preparation.
        Passes over the resident counts: column statistics, row norms (Frobenius only), and later the single
        counts -> fp16 pass; nothing is written back.
        """
        return self._run(data_source, X, deferred=True)

    def _run(self, data_source: str, X, deferred: bool):
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
        gain = 1.0
        if self.log_input or need_norm:
            norm_sum = nat.transform_rows(Xd, log10p1=self.log_input, shift=shift, scale=scale, gain=1.0,
                                          want_row_norms=need_norm, in_place=not deferred)
        if need_norm and n > 0:
            mean_norm = float(norm_sum) / n
            if data_source == "source":
                self.frob_norm_param_ = mean_norm
            else:
                if self.frob_norm_param_ is None:
                    raise RuntimeError("frob_norm_source: the source must be processed before the target")
                gain = self.frob_norm_param_ / mean_norm
                if not deferred:
                    nat.axpby(gain, Xd.view(-1), 0.0, Xd.view(-1))
        if deferred:
            return DeferredRows(Xd, self.log_input, shift, scale, gain)
        return Xd
