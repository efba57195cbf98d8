"""
`falkon.Falkon` look-alike: the estimator /root/reference/sobolev_alignment/krr_approx.py:262-268
constructs, fits (:227), queries (`alpha_`, `ny_points_`, :232,284), predicts with (:314) and
re-hydrates by ASSIGNING `ny_points_` / `alpha_` after `load` (:402-404).
"""

from __future__ import annotations

import time
import warnings

import numpy as np
import torch

from ..engine import CudaOps
from ..solver import Comm, FalkonSolver, Predictor
from .options import FalkonOptions


_pinned_fetch = {}      # (device index) -> pinned staging buffer of the centre read-back, reused across fits


class _HostFetch:
    """
    Device -> host copy of a tensor on its own stream and thread, overlapped with whatever the caller enqueues.
    The D2H itself goes through a cached PINNED buffer (a pageable cudaMemcpy is staged by the driver in blocking
    chunks and, from a second thread, stalls the main thread's kernel launches: 0.3 s per fit at 8 GPUs); the
    pinned -> pageable copy that follows is plain host work.
    """

    def __init__(self, t: torch.Tensor):
        import threading

        self._out, self._err = None, None
        main = torch.cuda.current_stream(t.device)
        side = torch.cuda.Stream(t.device)
        side.wait_stream(main)          # the tensor is complete on the caller's stream
        t.record_stream(side)

        def run():
            try:
                torch.cuda.set_device(t.device)
                key = t.device.index
                buf = _pinned_fetch.get(key)
                if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
                    buf = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
                    _pinned_fetch[key] = buf
                view = buf[: t.numel()].view(t.shape)
                with torch.cuda.stream(side):
                    view.copy_(t, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(side)
                done.synchronize()
                self._out = view.clone()          # pageable result the caller owns; the pinned buffer is reused
            except BaseException as exc:  # noqa: BLE001 - re-raised by result()
                self._err = exc

        self._thread = threading.Thread(target=run, name="sobolev-b200-fetch")
        self._thread.start()

    def result(self) -> torch.Tensor:
        self._thread.join()
        if self._err is not None:
            raise self._err
        return self._out


def _gather_rows(X: torch.Tensor, idx: np.ndarray) -> torch.Tensor:
    """X[idx] -- for host tensors split over a few threads (torchrun pins OMP to one thread; index_select drops the GIL)."""
    index = torch.from_numpy(np.ascontiguousarray(idx)).to(X.device)
    if X.is_cuda or len(idx) * X.shape[1] < (1 << 22):
        return X[index]
    from concurrent.futures import ThreadPoolExecutor

    T = 8
    out = torch.empty(len(idx), X.shape[1], dtype=X.dtype)
    step = (len(idx) + T - 1) // T
    with ThreadPoolExecutor(max_workers=T) as pool:
        futs = [pool.submit(torch.index_select, X, 0, index[a: a + step], out=out[a: a + step])
                for a in range(0, len(idx), step)]
        for f in futs:
            f.result()
    return out


class Falkon:
    def __init__(self, kernel, penalty: float, M: int, center_selection="uniform", maxiter: int = 20,
                 seed=None, error_fn=None, error_every=1, weight_fn=None, options: FalkonOptions | None = None):
        self.kernel = kernel
        self.penalty = penalty
        self.M = M
        self.maxiter = maxiter
        self.seed = seed
        self.center_selection = center_selection
        self.error_fn = error_fn
        self.error_every = error_every
        if weight_fn is not None:
            raise NotImplementedError("weight_fn is not supported (the reference never passes it)")
        self.weight_fn = None
        self.options = options if options is not None else FalkonOptions()
        self.alpha_ = None
        self.ny_points_ = None
        self.fit_times_ = None
        self._predictor = None
        self._predictor_key = None

    # ------------------------------------------------------------------ centre selection
    def _select_centres(self, n: int) -> np.ndarray | None:
        cs = self.center_selection
        if not isinstance(cs, str):  # explicit row indices (parity tests feed oracle and GPU the same rows)
            return np.asarray(torch.as_tensor(cs).cpu(), dtype=np.int64)
        if cs != "uniform":
            raise ValueError(f"center_selection {cs!r} not supported")
        if self.M >= n:
            if self.M > n:
                warnings.warn(f"Number of centers M={self.M} exceeds the number of samples n={n}: using all samples")
            return np.arange(n, dtype=np.int64)
        seed = self.seed if self.seed is not None else self.options.center_seed
        return np.sort(np.random.default_rng(seed).choice(n, size=self.M, replace=False))

    def _ops(self) -> CudaOps:
        return CudaOps(device=self.options.device, stage_rows=self.options.stage_rows)

    # ------------------------------------------------------------------ fit / predict
    def fit(self, X: torch.Tensor, Y: torch.Tensor, Xts=None, Yts=None, warm_start=None):
        if warm_start is not None:
            raise NotImplementedError("warm_start is not supported")
        deferred = hasattr(X, "folded")      # preprocessing.DeferredRows: counts + the affine map still to apply
        X, Y = (X if deferred else torch.as_tensor(X)), torch.as_tensor(Y)
        if X.dim() != 2:
            raise ValueError("X must be 2-dimensional")
        if Y.dim() == 1:
            Y = Y.reshape(-1, 1)
        if X.shape[0] != Y.shape[0]:
            raise ValueError(f"X and Y have different sample counts ({X.shape[0]} vs {Y.shape[0]})")
        if X.dtype != Y.dtype:
            raise TypeError(f"X and Y must have the same dtype ({X.dtype} vs {Y.dtype})")
        t0 = time.perf_counter()
        ops = self._ops()
        opt = self.options
        comm = Comm(enabled=bool(opt.row_sharded))
        n_local = X.shape[0]
        if comm.enabled and comm.world > 1:
            centres, n_total = self._gather_centres(X, comm, ops)
            t_gather = time.perf_counter() - t0
        else:
            t_gather = 0.0
            idx = self._select_centres(n_local)
            centres = X.rows(idx) if deferred else _gather_rows(X, idx)
            n_total = n_local
            t_gather = time.perf_counter() - t0
        # Row-sharded fit from host inputs: `ny_points_` has to be a host tensor again (anchors() contract).
        # Its 600 MB read-back (C3) runs on a helper thread / side stream while the solver works.
        # results live where the inputs live; deferred rows always sit on the device, so Y decides for them
        on_device = Y.is_cuda if deferred else X.is_cuda
        fetch = None
        if centres.is_cuda and not on_device:
            fetch = _HostFetch(centres)
        solver = FalkonSolver(ops, self.kernel.spec(), self.penalty, self.maxiter, pc_eps=opt.pc_eps(),
                              cg_eps=float(opt.cg_epsilon_32), cg_tol=float(opt.cg_tolerance),
                              full_grad_every=int(opt.cg_full_gradient_every), comm=comm,
                              kernel_blocks="recompute" if opt.never_store_kernel else str(opt.kernel_blocks),
                              kstore_budget=int(opt.kstore_budget))
        alpha = solver.fit(X, Y, centres, n_total=n_total)
        self.solver_stats_ = {"times": dict(solver.times), "residuals": list(solver.residuals),
                              "n_iter": solver.n_iter_, "n_kmv": solver.n_kmv_, "n_total": n_total}
        solver.release()
        t1 = time.perf_counter()
        self.alpha_ = alpha.to(Y.dtype) if on_device else alpha.to(Y.dtype).cpu()
        if fetch is not None:
            self.ny_points_ = fetch.result()
        else:
            self.ny_points_ = centres if centres.is_cuda == on_device else (centres.cuda() if on_device else centres.cpu())
        self.fit_times_ = [time.perf_counter() - t0]
        # host-side wall clock outside the solver: centre selection / exchange, and the read-back of the results
        self.solver_stats_["times"].update({"select_centres": t_gather, "read_back": time.perf_counter() - t1,
                                            "estimator_wall": self.fit_times_[0]})
        self._predictor = None
        if self.error_fn is not None and Xts is not None and Yts is not None:
            err = self.error_fn(Yts, self.predict(Xts))
            print(f"Iteration {solver.n_iter_:3d} - Elapsed {self.fit_times_[-1]:.2f}s - validation error: {err}",
                  flush=True)
        return self

    def _gather_centres(self, X, comm, ops):
        """Row-sharded fit: draw global row indices identically on every rank, exchange the rows."""
        dist = comm.dist
        n_local = X.shape[0]
        counts = torch.zeros(comm.world, dtype=torch.int64, device=ops.device)
        counts[comm.rank] = n_local
        dist.all_reduce(counts, group=comm.group)
        counts = counts.cpu().numpy()
        offsets = np.concatenate([[0], np.cumsum(counts)])
        n_total = int(offsets[-1])
        if self.seed is None and self.options.center_seed is None and isinstance(self.center_selection, str):
            raise ValueError("row_sharded fits need Falkon(seed=...) or FalkonOptions(center_seed=...) so that "
                             "all ranks draw the same centres")
        idx = self._select_centres(n_total)
        mine = idx[(idx >= offsets[comm.rank]) & (idx < offsets[comm.rank + 1])] - offsets[comm.rank]
        rows = X.rows(mine) if hasattr(X, "folded") else ops.to_device(_gather_rows(X, mine))
        p = X.shape[1]
        full = ops.zeros(len(idx), p)
        pos = np.nonzero((idx >= offsets[comm.rank]) & (idx < offsets[comm.rank + 1]))[0]
        if len(pos):
            full[torch.from_numpy(pos).to(ops.device)] = rows
        dist.all_reduce(full, group=comm.group)  # disjoint supports: the sum is the concatenation
        return full, n_total

    def _get_predictor(self):
        if self.alpha_ is None or self.ny_points_ is None:
            raise RuntimeError("Falkon has not been trained. `predict` must be called after `fit`.")
        key = (id(self.alpha_), id(self.ny_points_), self.alpha_.data_ptr(), self.ny_points_.data_ptr())
        if self._predictor is None or self._predictor_key != key:
            alpha = torch.as_tensor(self.alpha_)
            if alpha.dim() == 1:
                alpha = alpha.reshape(-1, 1)
            self._predictor = Predictor(self._ops(), self.kernel.spec(), torch.as_tensor(self.ny_points_), alpha)
            self._predictor_key = key
        return self._predictor

    def predict(self, X: torch.Tensor) -> torch.Tensor:
        X = torch.as_tensor(X)
        out = self._get_predictor().predict(X)
        out = out.to(X.dtype) if X.dtype in (torch.float32, torch.float64) else out
        return out if X.is_cuda else out.cpu()

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_predictor"] = None
        state["_predictor_key"] = None
        return state

    def __repr__(self):
        return (f"Falkon(kernel={self.kernel}, penalty={self.penalty}, M={self.M}, maxiter={self.maxiter}, "
                f"options={self.options})")
