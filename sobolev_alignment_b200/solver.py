"""
FALKON host algorithm (Rudi, Carratino, Rosasco 2017, Alg. 1; Meanti et al. 2020) over the device
engine.  This is the part of the third-party `falkon` package the reference reaches through
`Falkon.fit` / `Falkon.predict` (/root/reference/sobolev_alignment/krr_approx.py:227,314); SURVEY
section 3.1 lists the steps.

    T = chol(K_MM + eps M I)  (upper)    stored as the lower factor L  = T^T
    A = chol(T T^T / M + lambda I)       stored as the lower factor LA = A^T
    rhs  = A^-T T^-T K_nM^T (Y / n)
    BHB u = A^-T ( T^-T K_nM^T (K_nM T^-1 A^-1 u) / n + lambda A^-1 u )
    beta = CG(BHB, rhs) ;  alpha = T^-1 A^-1 beta

The two products with K_nM are two launches of the fused `kmv` kernel (K_nM is recomputed, never
stored).  Rows of X may be sharded over ranks: each rank computes K_g^T (K_g z) for its rows and the
M x d partial sums are all-reduced once per product pair; everything else is replicated and
deterministic, so ranks stay bit-identical without a broadcast.
"""

from __future__ import annotations

import threading
import time

import torch

from . import distributed
from .engine import KernelSpec, pad_cols

TILE = 128
MAX_COLS = 32   # right-hand-side columns per CG run / fused-kernel launch


class Comm:
    """Sum all-reduce over the ranks that hold row shards (identity for a single process)."""

    def __init__(self, group=None, enabled: bool = False):
        self.enabled = enabled
        self.group = group
        if enabled:
            import torch.distributed as dist

            self.dist = dist
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.enabled and self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def broadcast_async(self, t: torch.Tensor, src: int):
        """Start a broadcast of `t` from group rank `src`; returns a handle whose wait() orders the current stream."""
        if not (self.enabled and self.world > 1):
            return _Done()
        gsrc = self.dist.get_global_rank(self.group, src) if self.group is not None else src
        return self.dist.broadcast(t, src=gsrc, group=self.group, async_op=True)

    def sum_int(self, v: int, device) -> int:
        if not (self.enabled and self.world > 1):
            return v
        t = torch.tensor([v], dtype=torch.int64, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return int(t.item())


class _Done:
    def wait(self):
        return True


class Preconditioner:
    """
    Lower Cholesky factors L (= T^T) and LA (= A^T) of the Nystrom preconditioner, kept in the packed
    fp16 hi/lo tile form the triangular solves stream (`ops.trsm_pack`); the fp32 squares are released
    as soon as they are packed (the packed pair is half their size).
    """

    def __init__(self, ops, spec: KernelSpec, frame, Cprep, penalty: float, pc_eps: float, comm=None,
                 K: torch.Tensor | None = None):
        M = Cprep.n
        Mp = (M + TILE - 1) // TILE * TILE
        self.ops, self.M, self.Mp = ops, M, Mp
        nblk = Mp // TILE
        distributed = comm is not None and comm.enabled and comm.world > 1
        if K is None:
            # K_MM, padded with an identity block (leaves factors and solves of the leading block unchanged)
            K = ops.zeros(Mp, Mp)
            ops.kdense(spec, frame, Cprep, Cprep, K, symmetric=True)
            if Mp > M:
                K.diagonal()[M:] = 1.0
        ops.add_diag(K, M, pc_eps * M)
        L_inv = ops.empty(nblk, TILE, TILE)
        self.info = ops.zeros(2, dtype=torch.int32)
        self._potrf(K, L_inv, self.info[0:1], comm if distributed else None)   # in place: K now holds L
        held = [K, L_inv]
        del K, L_inv        # `held` is the only reference: _finish releases the fp32 factor once it is packed
        self._finish(held, penalty, comm if distributed else None)

    def _potrf(self, A, inv, info, comm):
        if comm is not None and hasattr(self.ops, "potrf_panel"):
            distributed.potrf(self.ops, A, inv, info, comm)     # outer panels dealt out over the ranks
        else:
            self.ops.potrf(A, inv, info)

    def _finish(self, held, penalty, comm):
        """From the factor L of K_MM + eps M I: G = L^T L / M + lambda I, its factor, and the packed forms."""
        ops, M, Mp = self.ops, self.M, self.Mp
        L, L_inv = held
        nblk = Mp // TILE
        # A A^T ... : G = L^T L / M + lambda I  (= T T^T / M + lambda I), then its Cholesky factor
        if comm is not None and getattr(ops, "ltl_sharded", False):
            # L^T L has no sequential chain: every rank forms its share of the 512-row slabs and the
            # parts are summed over NVLink (the all-reduce result is identical on all ranks)
            G = ops.zeros(Mp, Mp)
            ops.ltl(L, G, 1.0 / M, 0.0, comm.rank, comm.world)
            comm.all_reduce(G)
            ops.add_diag(G, Mp, penalty)
        else:
            G = ops.empty(Mp, Mp)
            ops.ltl(L, G, 1.0 / M, penalty)
        self.Lp = ops.trsm_pack(L, L_inv)
        held.clear()
        del L, L_inv
        LA_inv = ops.empty(nblk, TILE, TILE)
        self._potrf(G, LA_inv, self.info[1:2], comm)
        self.LAp = ops.trsm_pack(G, LA_inv)

    def check(self):
        info = self.ops.to_host(self.info).tolist()
        if info[0] or info[1]:
            which = "K_MM + eps*M*I" if info[0] else "T T^T / M + lambda*I"
            raise RuntimeError(
                f"Cholesky of {which} failed at pivot {max(info) - 1}: matrix not positive definite "
                "(increase penalization / check for duplicated centres)")

    # in-place solves on [Mp x dp] buffers
    def invT(self, v):
        return self.ops.trsm(self.Lp, v, True)

    def invTt(self, v):
        return self.ops.trsm(self.Lp, v, False)

    def invA(self, v):
        return self.ops.trsm(self.LAp, v, True)

    def invAt(self, v):
        return self.ops.trsm(self.LAp, v, False)

    # the operator always applies the factors in pairs: one launch with two interleaved chains
    def invA_invT(self, v, z):
        """v <- A^-1 v (in place);  z <- T^-1 v."""
        return self.ops.trsm_pair(self.LAp, self.Lp, v, z, None, 0.0, True)

    def invTt_invAt(self, w, out, v=None, lam=0.0):
        """w <- T^-T w (in place);  out <- A^-T (w + lam v)."""
        return self.ops.trsm_pair(self.Lp, self.LAp, w, out, v, lam, False)


class FalkonSolver:
    def __init__(self, ops, spec: KernelSpec, penalty: float, maxiter: int = 20, pc_eps: float = 1e-5,
                 cg_eps: float = 1e-7, cg_tol: float = 1e-7, full_grad_every: int = 10, comm: Comm | None = None):
        self.ops, self.spec = ops, spec
        self.penalty, self.maxiter = float(penalty), int(maxiter)
        self.pc_eps, self.cg_eps, self.cg_tol = float(pc_eps), float(cg_eps), float(cg_tol)
        self.full_grad_every = int(full_grad_every)
        self.comm = comm or Comm()
        self.times = {}
        self.residuals = []
        self.n_iter_ = 0
        self.n_kmv_ = 0

    # ------------------------------------------------------------------ products with K_nM
    def _knm_t_knm(self, z, out):
        """out[:M] = sum over local rows of K_g^T (K_g z[:M]), all-reduced over ranks."""
        ops, M = self.ops, self.M
        self.t_buf.zero_()
        ops.kmv(self.spec, self.frame, self.Xprep, self.Cprep, z[:M], self.t_buf)
        out.zero_()
        ops.kmv(self.spec, self.frame, self.Cprep, self.Xprep, self.t_buf, out[:M])
        self.n_kmv_ += 2
        self.comm.all_reduce(out)
        return out

    def _bhb(self, u, out):
        """out = BHB u  (u is left untouched; w1/w2 are scratch)."""
        ops, pre = self.ops, self.pre
        v, z, w = self.w1, self.w2, self.w3
        v.copy_(u)
        pre.invA_invT(v, z)               # v = A^-1 u ; z = T^-1 v
        self._knm_t_knm(z, w)             # w = K_nM^T K_nM z
        ops.axpby(1.0 / self.n_total, w, 0.0, w)   # w /= n
        pre.invTt_invAt(w, out, v, self.penalty)    # out = A^-T (T^-T w + lambda v)
        return out

    # ------------------------------------------------------------------ fit
    def fit(self, X, Y, centres, n_total: int | None = None):
        """
        X [n x p], Y [n x d]: this rank's rows (CPU or CUDA).  centres [M x p]: identical on all ranks.
        Returns alpha [M x d] (float32, on the engine's device).
        """
        ops = self.ops
        t0 = time.perf_counter()
        n, d = Y.shape
        self.n_total = int(n_total) if n_total is not None else self.comm.sum_int(n, ops.device)
        C = ops.to_device(centres)
        M = self.M = C.shape[0]
        self.frame = ops.make_frame(C, self.spec)
        self.Cprep = ops.prepare(C, self.frame)
        del C
        # Host inputs are staged (pinned copy, H2D, fp16 preparation) on a side stream WHILE the main
        # stream factorises the preconditioner, which only needs the centres.
        overlap = hasattr(ops, "staging_stream") and not (torch.is_tensor(X) and X.is_cuda)
        if overlap:
            main = torch.cuda.current_stream(ops.device)
            side = ops.staging_stream()
            side.wait_stream(main)                      # frame / prepared centres are ready
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record(main)
            # The preconditioner is ~2000 kernel launches: the enqueueing thread blocks on the launch
            # queue for most of its GPU time, so the staging loop runs on a helper thread (CUDA calls
            # release the GIL) instead of behind it.
            staged = {}

            def _stage():
                try:
                    torch.cuda.set_device(ops.device)
                    with torch.cuda.stream(side):
                        ev[2].record(side)
                        staged["X"] = ops.prepare(X, self.frame)
                        staged["Y"] = ops.to_device(Y)
                        ev[3].record(side)
                except BaseException as exc:  # noqa: BLE001 - re-raised on the calling thread
                    staged["error"] = exc

            helper = threading.Thread(target=_stage, name="sobolev-b200-staging")
            helper.start()
        self.pre = pre = Preconditioner(ops, self.spec, self.frame, self.Cprep, self.penalty, self.pc_eps, self.comm)
        Mp = pre.Mp
        if overlap:
            ev[1].record(main)
            helper.join()
            if "error" in staged:
                raise staged["error"]
            self.Xprep, Yfull = staged["X"], staged["Y"]
            for t in (self.Xprep.data, self.Xprep.norms, Yfull):
                t.record_stream(main)
            main.wait_stream(side)
            self._phase_events = ev
        else:
            ops.synchronize()
            self.times["preconditioner"] = time.perf_counter() - t0
            t1 = time.perf_counter()
            self.Xprep = ops.prepare(X, self.frame)
            Yfull = ops.to_device(Y)
            ops.synchronize()
            self.times["stage_inputs"] = time.perf_counter() - t1

        t2 = time.perf_counter()
        if overlap:
            ev[4].record(main)
        # the right-hand sides are solved simultaneously, at most MAX_COLS columns per CG run (the fused
        # kernel carries up to 32 columns of V); preconditioner and prepared rows are shared by the runs
        pre.check()
        out = ops.zeros(M, d)
        all_res, n_iter = [], 0
        for c0 in range(0, d, MAX_COLS):
            c1 = min(d, c0 + MAX_COLS)
            dp = pad_cols(c1 - c0)
            Yd = ops.zeros(n, dp)
            Yd[:, : c1 - c0] = Yfull[:, c0:c1]
            alpha = self._solve_columns(Yd, Mp, dp)
            out[:, c0:c1] = alpha[:M, : c1 - c0]
            all_res.append(list(self.residuals))
            n_iter = max(n_iter, self.n_iter_)
        del Yfull
        self.n_iter_ = n_iter
        self.residuals = [max(r[i] for r in all_res if i < len(r)) for i in range(max(len(r) for r in all_res))]
        if overlap:
            self._phase_events[5].record(main)
        ops.synchronize()
        if overlap:
            ev = self._phase_events
            self.times["preconditioner"] = ev[0].elapsed_time(ev[1]) * 1e-3
            self.times["stage_inputs"] = ev[2].elapsed_time(ev[3]) * 1e-3   # concurrent with the preconditioner
            self.times["staging_overlapped"] = True
            self.times["cg"] = ev[4].elapsed_time(ev[5]) * 1e-3
            del self._phase_events
        else:
            self.times["cg"] = time.perf_counter() - t2
        self.times["fit"] = time.perf_counter() - t0
        return out

    def _solve_columns(self, Yd, Mp, dp):
        """FALKON solve for the (padded) columns Yd [n x dp]; returns alpha [Mp x dp] (a scratch buffer)."""
        ops, pre, M = self.ops, self.pre, self.M
        n = Yd.shape[0]
        self.t_buf = ops.empty(n, dp)
        self.w1, self.w2, self.w3 = ops.zeros(Mp, dp), ops.zeros(Mp, dp), ops.zeros(Mp, dp)
        B = ops.zeros(Mp, dp)
        Bw = self.w3
        # rhs = A^-T T^-T K_nM^T (Y / n)
        ops.axpby(1.0 / self.n_total, Yd, 0.0, Yd)     # Y / n
        Bw.zero_()
        ops.kmv(self.spec, self.frame, self.Cprep, self.Xprep, Yd, Bw[:M])
        self.n_kmv_ += 1
        self.comm.all_reduce(Bw)
        pre.invTt_invAt(Bw, B)            # B = A^-T T^-T K_nM^T (Y / n)
        beta = self._conjugate_gradient(B, Mp, dp)
        alpha = self.w2
        pre.invA_invT(beta, alpha)        # alpha = T^-1 A^-1 beta
        return alpha

    def _conjugate_gradient(self, B, Mp, dp):
        """
        Batched-column CG (falkon `ConjugateGradient.solve`).  Convergence is decided by the kernels (a
        device flag freezes X once max_c sqrt(r_c^T r_c) < cg_tol); the host reads that flag LAG
        iterations late, so no iteration ends in a device synchronisation, and -- the flag being a
        function of bit-identical replicated state -- every rank leaves the loop at the same iteration.
        """
        ops = self.ops
        LAG = 2
        X = ops.zeros(Mp, dp)
        R = B.clone()
        P = B.clone()
        AP = ops.zeros(Mp, dp)
        rs_old, rs_new, pAp = ops.empty(dp), ops.empty(dp), ops.empty(dp)
        cgs = ops.cg_scratch(self.maxiter, dp)
        ops.coldot(R, R, rs_old, cgs)
        done = 0
        for it in range(self.maxiter):
            if it >= LAG and cgs.stopped_at(it - LAG):
                break
            self._bhb(P, AP)
            ops.coldot(P, AP, pAp, cgs)
            # Falkon recomputes the residual from scratch every `full_grad_every` iterations.  On the very
            # last iteration that product pair would only feed the reported residual (X is final), so
            # the recurrence is used there instead: same alpha, two fused-kernel launches fewer.
            if (it + 1) % self.full_grad_every == 0 and it + 1 < self.maxiter:
                ops.cg_update(P, None, X, None, rs_old, pAp, self.cg_eps, None, cgs)   # X += alpha P
                self._bhb(X, R)                                                       # R = B - BHB X
                ops.axpby(1.0, B, -1.0, R)
                ops.coldot(R, R, rs_new, cgs, hist_row=it, tol=self.cg_tol)
            else:
                ops.cg_update(P, AP, X, R, rs_old, pAp, self.cg_eps, rs_new, cgs, hist_row=it, tol=self.cg_tol)
            cgs.poll(it)
            done = it + 1
            ops.cg_direction(R, P, rs_new, rs_old, self.cg_eps, cgs)
            rs_old, rs_new = rs_new, rs_old
        # one read of the residual history; iterations after the first converged one were no-ops
        hist = cgs.history()[:done].double().clamp_min(0).max(dim=1).values.sqrt().tolist()
        self.residuals = []
        for r in hist:
            self.residuals.append(r)
            if r < self.cg_tol:
                break
        self.n_iter_ = len(self.residuals)
        return X

    # ------------------------------------------------------------------ release
    def release(self):
        for name in ("pre", "Xprep", "t_buf", "w1", "w2", "w3"):
            if hasattr(self, name):
                delattr(self, name)


class Predictor:
    """K(X, C) alpha with the centres and coefficients resident on the device."""

    def __init__(self, ops, spec: KernelSpec, centres: torch.Tensor, alpha: torch.Tensor):
        self.ops, self.spec = ops, spec
        C = ops.to_device(centres)
        self.frame = ops.make_frame(C, spec)
        self.Cprep = ops.prepare(C, self.frame)
        self.d = alpha.shape[1]
        alpha_d = ops.to_device(alpha)
        self.alpha_chunks = []      # (first column, padded coefficient block) per <= MAX_COLS columns
        for c0 in range(0, self.d, MAX_COLS):
            c1 = min(self.d, c0 + MAX_COLS)
            blk = ops.zeros(C.shape[0], pad_cols(c1 - c0))
            blk[:, : c1 - c0] = alpha_d[:, c0:c1]
            self.alpha_chunks.append((c0, c1, blk))

    def predict(self, X: torch.Tensor, row_chunk: int = 131072) -> torch.Tensor:
        """K(X, C) alpha; host rows are staged one chunk ahead of the fused kernel (engine.stream_rows)."""
        ops = self.ops
        out = ops.zeros(X.shape[0], self.d)

        def one_chunk(Xp, s):
            for c0, c1, blk in self.alpha_chunks:
                part = ops.zeros(Xp.n, blk.shape[1])
                ops.kmv(self.spec, self.frame, Xp, self.Cprep, blk, part)
                out[s: s + Xp.n, c0:c1] = part[:, : c1 - c0]

        stream_rows(ops, X, self.frame, one_chunk, row_chunk)
        return out


def stream_rows(ops, X, frame, fn, row_chunk: int = 131072):
    """fn(prepared_chunk, first_row) over row chunks of X, pipelined by the engine when it can."""
    impl = getattr(ops, "stream_rows", None)
    if impl is not None:
        return impl(X, frame, fn, row_chunk)
    for s in range(0, X.shape[0], row_chunk):
        fn(ops.prepare(X[s: s + row_chunk], frame), s)
