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

    With several penalties (the model-selection grid, krr_model_selection.py:68-274) T -- which does not
    depend on lambda -- is factored and packed once and G0 = T T^T / M is kept: per lambda only
    A = chol(G0 + lambda I) is new (`LAp[k]`).
    """

    def __init__(self, ops, spec: KernelSpec, frame, Cprep, penalties, pc_eps: float, comm=None):
        M = Cprep.n
        Mp = (M + TILE - 1) // TILE * TILE
        self.ops, self.M, self.Mp = ops, M, Mp
        self.penalties = [float(x) for x in (penalties if isinstance(penalties, (list, tuple)) else [penalties])]
        nblk = Mp // TILE
        distributed = comm is not None and comm.enabled and comm.world > 1
        # K_MM, padded with an identity block (leaves factors and solves of the leading block unchanged)
        K = ops.zeros(Mp, Mp)
        ops.kdense(spec, frame, Cprep, Cprep, K, symmetric=True)
        if Mp > M:
            K.diagonal()[M:] = 1.0
        ops.add_diag(K, M, pc_eps * M)
        L_inv = ops.empty(nblk, TILE, TILE)
        self.info = ops.zeros(1 + len(self.penalties), dtype=torch.int32)
        self._potrf(K, L_inv, self.info[0:1], comm if distributed else None)   # in place: K now holds L
        held = [K, L_inv]
        del K, L_inv        # `held` is the only reference: _finish releases the fp32 factor once it is packed
        self._finish(held, comm if distributed else None)

    def _potrf(self, A, inv, info, comm):
        if comm is not None and hasattr(self.ops, "potrf_panel"):
            distributed.potrf(self.ops, A, inv, info, comm)     # outer panels dealt out over the ranks
        else:
            self.ops.potrf(A, inv, info)

    def _finish(self, held, comm):
        """From the factor L of K_MM + eps M I: G = L^T L / M + lambda I, its factor(s), and the packed forms."""
        ops, M, Mp = self.ops, self.M, self.Mp
        L, L_inv = held
        nblk = Mp // TILE
        lams = self.penalties
        first = lams[0] if len(lams) == 1 else 0.0     # several penalties: keep G0 = L^T L / M without the ridge
        # A A^T ... : G = L^T L / M + lambda I  (= T T^T / M + lambda I), then its Cholesky factor
        if comm is not None and getattr(ops, "ltl_sharded", False):
            # L^T L has no sequential chain: every rank forms its share of the 512-row slabs and the
            # parts are summed over NVLink (the all-reduce result is identical on all ranks)
            G = ops.zeros(Mp, Mp)
            ops.ltl(L, G, 1.0 / M, 0.0, comm.rank, comm.world)
            comm.all_reduce(G)
            if first != 0.0:
                ops.add_diag(G, Mp, first)
        else:
            G = ops.empty(Mp, Mp)
            ops.ltl(L, G, 1.0 / M, first)
        self.Lp = ops.trsm_pack(L, L_inv)
        held.clear()
        del L, L_inv
        self.LAp = []
        for k, lam in enumerate(lams):
            last = k == len(lams) - 1
            Gk = G if last else G.clone()               # the last factorisation may destroy G0
            if len(lams) > 1:
                ops.add_diag(Gk, Mp, lam)
            LA_inv = ops.empty(nblk, TILE, TILE)
            self._potrf(Gk, LA_inv, self.info[1 + k: 2 + k], comm)
            self.LAp.append(ops.trsm_pack(Gk, LA_inv))
            del Gk, LA_inv

    def check(self):
        info = self.ops.to_host(self.info).tolist()
        if any(info):
            bad = next(i for i, v in enumerate(info) if v)
            which = "K_MM + eps*M*I" if bad == 0 else f"T T^T / M + lambda*I (lambda = {self.penalties[bad - 1]:g})"
            raise RuntimeError(
                f"Cholesky of {which} failed at pivot {info[bad] - 1}: matrix not positive definite "
                "(increase penalization / check for duplicated centres)")

    # in-place solves on [Mp x dp] buffers
    def invT(self, v):
        return self.ops.trsm(self.Lp, v, True)

    def invTt(self, v):
        return self.ops.trsm(self.Lp, v, False)

    def invA(self, v, k=0):
        return self.ops.trsm(self.LAp[k], v, True)

    def invAt(self, v, k=0):
        return self.ops.trsm(self.LAp[k], v, False)

    # the operator always applies the factors in pairs: one launch with two interleaved chains
    def invA_invT(self, v, z, k=0):
        """v <- A^-1 v (in place);  z <- T^-1 v."""
        return self.ops.trsm_pair(self.LAp[k], self.Lp, v, z, None, 0.0, True)

    def invTt_invAt(self, w, out, v=None, lam=0.0, k=0):
        """w <- T^-T w (in place);  out <- A^-T (w + lam v)."""
        return self.ops.trsm_pair(self.Lp, self.LAp[k], w, out, v, lam, False)


class _CgColumns:
    """CG state of one penalty: [Mp x dp] vectors, per-column scalars, device-side convergence scratch."""

    def __init__(self, ops, B, maxiter):
        Mp, dp = B.shape
        self.B = B
        self.X = ops.zeros(Mp, dp)
        self.R = B.clone()
        self.P = B.clone()
        self.AP = ops.zeros(Mp, dp)
        self.rs_old, self.rs_new, self.pAp = ops.empty(dp), ops.empty(dp), ops.empty(dp)
        self.cgs = ops.cg_scratch(maxiter, dp)
        self.v, self.z, self.w = ops.zeros(Mp, dp), ops.zeros(Mp, dp), ops.zeros(Mp, dp)
        ops.coldot(self.R, self.R, self.rs_old, self.cgs)


class FalkonSolver:
    def __init__(self, ops, spec: KernelSpec, penalty, maxiter: int = 20, pc_eps: float = 1e-5,
                 cg_eps: float = 1e-7, cg_tol: float = 1e-7, full_grad_every: int = 10, comm: Comm | None = None,
                 kernel_blocks: str = "recompute", kstore_budget: int = 24 << 30):
        """
        kernel_blocks: how K_nM^T (K_nM z) is formed.  "recompute" (default, the north-star design): the fused
        kernel runs twice, K_nM never exists in HBM.  "store": K is formed ONCE per product pair -- row block by
        row block the forward product also leaves the block behind as packed fp16 hi/lo tiles in a transient
        scratch of at most `kstore_budget` bytes, and K^T t streams it back (HBM-bound) -- SURVEY section 7,
        design (B); falkon's `never_store_kernel=False`.
        """
        if kernel_blocks not in ("recompute", "store"):
            raise ValueError("kernel_blocks must be 'recompute' or 'store'")
        self.kernel_blocks, self.kstore_budget = kernel_blocks, int(kstore_budget)
        self.ops, self.spec = ops, spec
        self.penalties = [float(x) for x in (penalty if isinstance(penalty, (list, tuple)) else [penalty])]
        self.penalty, self.maxiter = self.penalties[0], int(maxiter)
        self.pc_eps, self.cg_eps, self.cg_tol = float(pc_eps), float(cg_eps), float(cg_tol)
        self.full_grad_every = int(full_grad_every)
        self.comm = comm or Comm()
        self.times = {}
        self.residuals = []
        self.n_iter_ = 0
        self.n_kmv_ = 0

    # ------------------------------------------------------------------ products with K_nM
    def _knm_t_knm(self, Z, out):
        """out[:M] = sum over local rows of K_g^T (K_g Z[:M]), all-reduced over ranks (any <= 32 columns)."""
        ops, M = self.ops, self.M
        self.t_buf.zero_()
        out.zero_()
        if self.kernel_blocks == "store" and hasattr(ops, "ktv"):
            for r0, r1, Xb in self._row_blocks():
                tb = self.t_buf[r0:r1]
                ops.kmv(self.spec, self.frame, Xb, self.Cprep, Z[:M], tb, kstore=self.kstore)
                ops.ktv(self.kstore, r1 - r0, M, tb, out[:M])
            self.n_kmv_ += 1
        else:
            ops.kmv(self.spec, self.frame, self.Xprep, self.Cprep, Z[:M], self.t_buf)
            ops.kmv(self.spec, self.frame, self.Cprep, self.Xprep, self.t_buf, out[:M])
            self.n_kmv_ += 2
        self.comm.all_reduce(out)
        return out

    def _row_blocks(self):
        """Row blocks of the prepared rows whose kernel block fits the scratch: whole waves of CTA pairs (148 x 128 rows)."""
        if getattr(self, "_blocks", None) is None:
            ops, n, M = self.ops, self.Xprep.n, self.M
            wave = getattr(ops, "wave_rows", 148 * 128)
            per_wave = ops.kstore_bytes(wave, M)
            budget = min(self.kstore_budget, int(0.5 * ops.free_bytes()))
            rows = max(1, budget // per_wave) * wave
            rows = min(rows, (n + 127) // 128 * 128)
            self.kstore = ops.zeros(ops.kstore_bytes(min(rows, n), M), dtype=torch.uint8)
            self._blocks = []
            for r0 in range(0, n, rows):
                r1 = min(n, r0 + rows)
                self._blocks.append((r0, r1, self.Xprep.rows(r0, r1)))
        return self._blocks

    def _bhb(self, states, src, dst):
        """
        dst_k = BHB_k src_k for every penalty k of the group: one K_nM^T K_nM product for all of them --
        the columns of the group ride side by side through the two fused-kernel launches.
        """
        ops, pre, M = self.ops, self.pre, self.M
        dc = self._group_cols                                 # real (unpadded) columns per penalty
        wide = len(states) > 1
        for j, (k, st) in enumerate(states):
            st.v.copy_(getattr(st, src))
            pre.invA_invT(st.v, st.z, k)                      # v = A_k^-1 u ; z = T^-1 v
            if wide:
                self.Zw[:, j * dc: (j + 1) * dc] = st.z[:, :dc]
        if wide:
            self._knm_t_knm(self.Zw, self.Ww)                 # W = K_nM^T K_nM [z_1 | z_2 | ...]
        else:
            self._knm_t_knm(states[0][1].z, states[0][1].w)
        for j, (k, st) in enumerate(states):
            if wide:
                st.w[:, :dc] = self.Ww[:, j * dc: (j + 1) * dc]
            ops.axpby(1.0 / self.n_total, st.w, 0.0, st.w)    # w /= n
            pre.invTt_invAt(st.w, getattr(st, dst), st.v, self.penalties[k], k)   # A_k^-T (T^-T w + lambda v)

    # ------------------------------------------------------------------ fit
    def fit(self, X, Y, centres, n_total: int | None = None):
        """
        X [n x p], Y [n x d]: this rank's rows (CPU or CUDA).  centres [M x p]: identical on all ranks.
        Returns alpha [M x d] (float32, on the engine's device) -- a list of them, one per penalty, when the
        solver was given a list of penalties.
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
        overlap = hasattr(ops, "staging_stream") and not getattr(X, "is_cuda", False)
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
            enqueue_only = getattr(ops, "stages_without_host_work", lambda _x: False)(X)
            pre_alloc = {}
            if enqueue_only:
                # memory of the staged operands comes from the MAIN stream's pool (the caching allocator keeps one pool
                # per stream: side-stream allocations could not reuse what the previous fit freed and went to cudaMalloc
                # -- 0.3 s of GB-sized allocations in front of the preconditioner)
                pre_alloc["out"] = ops.alloc_prepared(X.shape[0], X.shape[1])
                pre_alloc["dev_bufs"] = ops.stage_buffers(X.shape[0], X.shape[1])

            def _stage():
                try:
                    torch.cuda.set_device(ops.device)
                    with torch.cuda.stream(side):
                        ev[2].record(side)
                        staged["X"] = ops.prepare(X, self.frame, **pre_alloc)
                        staged["Y"] = ops.to_device(Y)
                        ev[3].record(side)
                    for b in pre_alloc.get("dev_bufs", ()):
                        b.record_stream(side)
                except BaseException as exc:  # noqa: BLE001 - re-raised on the calling thread
                    staged["error"] = exc

            helper = None
            if enqueue_only:
                _stage()        # pinned rows: enqueue-only, done before the preconditioner's launches pile up
                pre_alloc.clear()
            else:
                helper = threading.Thread(target=_stage, name="sobolev-b200-staging")
                helper.start()
        self.pre = pre = Preconditioner(ops, self.spec, self.frame, self.Cprep, self.penalties, self.pc_eps, self.comm)
        if overlap:
            ev[1].record(main)
            if helper is not None:
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
        pre.check()
        outs = self.solve_prepared(Yfull)
        del Yfull
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
        return outs if len(self.penalties) > 1 else outs[0]

    def solve_prepared(self, Yfull):
        """
        FALKON solves for every penalty with self.pre / self.Xprep / self.Cprep / self.frame in place.  The
        right-hand sides are solved simultaneously, at most MAX_COLS columns per fused-kernel launch: d > 32
        runs in column chunks, and as many penalties as fit share a launch (32 // dp of them).
        """
        ops, pre, M = self.ops, self.pre, self.M
        n, d = Yfull.shape
        Mp = pre.Mp
        nl = len(self.penalties)
        outs = [ops.zeros(M, d) for _ in range(nl)]
        traces = [[] for _ in range(nl)]
        for c0 in range(0, d, MAX_COLS):
            c1 = min(d, c0 + MAX_COLS)
            dp = pad_cols(c1 - c0)
            Yd = ops.zeros(n, dp)
            Yd[:, : c1 - c0] = Yfull[:, c0:c1]
            # T^-T K_nM^T (Y / n) does not depend on the penalty
            ops.axpby(1.0 / self.n_total, Yd, 0.0, Yd)
            tz = ops.zeros(Mp, dp)
            ops.kmv(self.spec, self.frame, self.Cprep, self.Xprep, Yd, tz[:M])
            self.n_kmv_ += 1
            self.comm.all_reduce(tz)
            del Yd
            group = max(1, MAX_COLS // (c1 - c0))     # penalties per fused launch: their REAL columns side by side
            self._group_cols = c1 - c0
            tzt = None
            if nl > 1:                       # T^-T of it is shared by all penalties too
                tzt = tz
                pre.invTt(tzt)
            for g0 in range(0, nl, group):
                ks = list(range(g0, min(nl, g0 + group)))
                self.t_buf = ops.empty(n, dp if len(ks) == 1 else pad_cols(len(ks) * (c1 - c0)))
                alphas = self._solve_group(tz, tzt, ks, Mp, dp)
                for k, alpha in zip(ks, alphas):
                    outs[k][:, c0:c1] = alpha[:M, : c1 - c0]
                    traces[k].append(list(self._group_residuals[k]))
        self.residuals_per_penalty = [
            [max(r[i] for r in tr if i < len(r)) for i in range(max(len(r) for r in tr))] for tr in traces]
        self.residuals = self.residuals_per_penalty[0]
        self.n_iter_ = max(len(r) for r in self.residuals_per_penalty)
        return outs

    def _solve_group(self, tz, tzt, ks, Mp, dp):
        """FALKON solve for the penalties `ks` (sharing the fused-kernel launches); returns [alpha_k] scratch buffers."""
        ops, pre = self.ops, self.pre
        if len(ks) > 1:
            wide = pad_cols(len(ks) * self._group_cols)
            self.Zw, self.Ww = ops.zeros(Mp, wide), ops.zeros(Mp, wide)
        states = []
        for k in ks:
            # B_k = A_k^-T T^-T K_nM^T (Y / n)
            B = ops.zeros(Mp, dp)
            if tzt is not None:
                B.copy_(tzt)
                pre.invAt(B, k)
            else:
                pre.invTt_invAt(tz, B, None, 0.0, k)
            states.append((k, _CgColumns(ops, B, self.maxiter)))
        self._conjugate_gradient(states)
        alphas = []
        self._group_residuals = {}
        for k, st in states:
            pre.invA_invT(st.X, st.z, k)          # alpha = T^-1 A_k^-1 beta
            alphas.append(st.z)
            hist = st.cgs.history()[: self._cg_done].double().clamp_min(0).max(dim=1).values.sqrt().tolist()
            res = []
            for r in hist:
                res.append(r)
                if r < self.cg_tol:
                    break
            self._group_residuals[k] = res
        return alphas

    def _conjugate_gradient(self, states):
        """
        Batched-column CG (falkon `ConjugateGradient.solve`) for a group of penalties in lock-step.
        Convergence is decided by the kernels (a device flag per penalty freezes its X once
        max_c sqrt(r_c^T r_c) < cg_tol); the host reads the flags LAG iterations late, so no iteration ends
        in a device synchronisation, and -- the flags being functions of bit-identical replicated state --
        every rank leaves the loop at the same iteration.
        """
        ops = self.ops
        LAG = 2
        done = 0
        for it in range(self.maxiter):
            if it >= LAG and all(st.cgs.stopped_at(it - LAG) for _, st in states):
                break
            self._bhb(states, "P", "AP")
            # Falkon recomputes the residual from scratch every `full_grad_every` iterations.  On the very
            # last iteration that product pair would only feed the reported residual (X is final), so
            # the recurrence is used there instead: same alpha, two fused-kernel launches fewer.
            full = (it + 1) % self.full_grad_every == 0 and it + 1 < self.maxiter
            for _, st in states:
                ops.coldot(st.P, st.AP, st.pAp, st.cgs)
                if full:
                    ops.cg_update(st.P, None, st.X, None, st.rs_old, st.pAp, self.cg_eps, None, st.cgs)   # X += alpha P
                else:
                    ops.cg_update(st.P, st.AP, st.X, st.R, st.rs_old, st.pAp, self.cg_eps, st.rs_new, st.cgs,
                                  hist_row=it, tol=self.cg_tol)
            if full:
                self._bhb(states, "X", "R")                                    # R = B - BHB X
                for _, st in states:
                    ops.axpby(1.0, st.B, -1.0, st.R)
                    ops.coldot(st.R, st.R, st.rs_new, st.cgs, hist_row=it, tol=self.cg_tol)
            done = it + 1
            for _, st in states:
                st.cgs.poll(it)
                ops.cg_direction(st.R, st.P, st.rs_new, st.rs_old, self.cg_eps, st.cgs)
                st.rs_old, st.rs_new = st.rs_new, st.rs_old
        self._cg_done = done

    # ------------------------------------------------------------------ release
    def release(self):
        for name in ("pre", "Xprep", "t_buf", "Zw", "Ww", "kstore", "_blocks"):
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
