"""
Device engine: the thin layer between the host algorithms (solver.py, falkon/) and the C ABI.

`CudaOps` is the ONLY compute backend the package ships.  It owns host<->device staging (pinned,
chunked, never writing into the caller's tensors -- the reference hands over read-only memmap-backed
tensors, /root/reference/sobolev_alignment/sobolev_alignment.py:613-616,922-925) and forwards every
numerical operation to libsobolev_b200.so.  The host algorithms are written against this small
interface so that the multi-rank host logic can be exercised on CPU by the test-suite with a
stand-in defined under tests/ (never importable from here).
"""

from __future__ import annotations

import math
import os

import torch

from . import _native as nat


class KernelSpec:
    """Kernel family + length-scale, resolved to what the CUDA epilogue needs."""

    __slots__ = ("family", "kid", "sigma", "sigma_vec")

    def __init__(self, family: str, sigma):
        if family not in nat.KERNEL_IDS:
            raise KeyError(family)
        self.family = family
        self.kid = nat.KERNEL_IDS[family]
        if torch.is_tensor(sigma) and sigma.numel() > 1:
            self.sigma = None
            self.sigma_vec = sigma.detach().to(torch.float32).flatten().cpu()
        else:
            self.sigma = float(sigma.item() if torch.is_tensor(sigma) else sigma)
            self.sigma_vec = None
            if not self.sigma > 0:
                raise ValueError("sigma must be positive")

    def kscale(self, gscale: float) -> float:
        """Constant applied to the distance between PREPARED rows (which carry the factor gscale)."""
        s = 1.0 if self.sigma is None else self.sigma
        if self.family == "gaussian":
            return 1.0 / (2.0 * s * s * gscale * gscale)
        return 1.0 / (s * gscale)


class Frame:
    """Affine map applied to rows before rounding to fp16: (x - shift) * scale * gscale."""

    __slots__ = ("shift", "scale", "gscale")

    def __init__(self, shift, scale, gscale):
        self.shift, self.scale, self.gscale = shift, scale, gscale


class CudaOps:
    name = "cuda"
    ltl_sharded = True     # L^T L can be formed in parts (sa_ltl_part) and summed over the ranks

    def __init__(self, device=None, stage_rows: int = 65536, copy_threads: int | None = None):
        if not torch.cuda.is_available():
            raise nat.NativeLibraryError(
                "sobolev_alignment_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        nat.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.stage_rows = int(stage_rows)
        self._pinned = None
        self._pinned_events = [None, None]   # last device-side use of each pinned staging buffer
        self._side = None
        self._pool = None
        self.copy_threads = max(1, min(8, (os.cpu_count() or 2) // 2)) if copy_threads is None else int(copy_threads)

    def staging_stream(self):
        """Side stream on which host inputs are copied / prepared while the main stream builds the preconditioner."""
        if self._side is None:
            # high priority: its short copy / prep kernels go first whenever a preconditioner launch ends
            self._side = torch.cuda.Stream(self.device, priority=-1)
        return self._side

    # ------------------------------------------------------------------ memory plumbing
    def zeros(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def to_device(self, t: torch.Tensor) -> torch.Tensor:
        """fp32 contiguous copy on the device (no-op when already there)."""
        if t.is_cuda:
            return t.to(device=self.device, dtype=torch.float32).contiguous()
        t = t.to(torch.float32).contiguous()
        return t.to(self.device, non_blocking=t.is_pinned())     # pinned source: stream-ordered, no host wait

    def to_host(self, t: torch.Tensor) -> torch.Tensor:
        return t.detach().to("cpu")

    # ------------------------------------------------------------------ frames
    def make_frame(self, ref_rows: torch.Tensor, spec: KernelSpec) -> Frame:
        """
        Centre on the column mean of `ref_rows` (distances are translation invariant; centring keeps
        |x|^2 small so |a|^2+|b|^2-2ab cancels less) and pick a power-of-two gain that puts the
        centred rows at unit rms per feature, well inside fp16's normal range.
        """
        ref = ref_rows.to(device=self.device, dtype=torch.float32)
        shift = ref.mean(0).contiguous()
        scale = None
        centred = ref - shift
        if spec.sigma_vec is not None:
            scale = (1.0 / spec.sigma_vec).to(self.device).contiguous()
            centred = centred * scale
        rms = float(centred.pow(2).mean().sqrt())
        gscale = 1.0 if not (rms > 0 and math.isfinite(rms)) else 2.0 ** round(-math.log2(rms))
        return Frame(shift, scale, gscale)

    # ------------------------------------------------------------------ operand preparation
    def _host_copy(self, dst: torch.Tensor, src: torch.Tensor):
        """
        dst (pinned) <- src (pageable, possibly memmap-backed), split over a few threads: one thread moves
        ~1-5 GB/s (page faults on a memmap), which would leave the H2D link idle (torch's copy releases the GIL).
        """
        n = src.shape[0]
        T = self.copy_threads
        if T <= 1 or n * src.shape[1] < (1 << 22):
            dst.copy_(src)
            return
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor

            self._pool = ThreadPoolExecutor(max_workers=T, thread_name_prefix="sobolev-b200-copy")
        step = (n + T - 1) // T
        futs = [self._pool.submit(dst[a: a + step].copy_, src[a: a + step]) for a in range(0, n, step)]
        for f in futs:
            f.result()

    @staticmethod
    def stages_without_host_work(X) -> bool:
        """True when staging X is enqueue-only (pinned, float32, contiguous host rows): no copy through our pinned buffers."""
        return (torch.is_tensor(X) and not X.is_cuda and X.dtype == torch.float32 and X.dim() == 2 and X.is_contiguous()
                and X.is_pinned())

    def alloc_prepared(self, n: int, p: int) -> nat.Prepared:
        return nat.alloc_prepared(n, p, self.device)

    def stage_buffers(self, n: int, p: int):
        """Device fp32 chunk buffers of the pinned-input staging path (allocate them on the stream that will free them)."""
        rows = max(1, min(self.stage_rows, n))
        return [torch.empty(rows, p, dtype=torch.float32, device=self.device) for _ in range(min(2, -(-n // rows)))]

    def prepare(self, X: torch.Tensor, frame: Frame, out: nat.Prepared | None = None, dev_bufs=None) -> nat.Prepared:
        """
        Rows of X (CPU or CUDA, any float dtype) -> prepared fp16 operand resident on the device (`out`, when
        given, must hold exactly X.shape[0] rows: the streaming callers reuse two buffers).
        """
        n, p = X.shape
        if out is not None and (out.n != n or out.p != p):
            raise ValueError("prepared buffer does not match the rows to prepare")
        if hasattr(X, "folded"):      # preprocessing.DeferredRows: counts -> fp16 rows + norms in one pass
            shift, scale = X.folded(frame.shift, frame.scale)
            return nat.prep(X.counts, shift, scale, frame.gscale, out=out, log10p1=X.log10p1)
        if X.is_cuda:
            Xd = X if X.dtype == torch.float32 and X.stride(1) == 1 else X.float().contiguous()
            return nat.prep(Xd, frame.shift, frame.scale, frame.gscale, out=out)
        if out is None:
            out = nat.alloc_prepared(n, p, self.device)
        if n == 0:
            return out
        rows = max(1, min(self.stage_rows, n))
        if self.stages_without_host_work(X):
            # pinned fp32 rows: H2D straight from the caller's buffer; everything is ordered by the stream (two
            # device chunk buffers reused in stream order), the host only enqueues
            if dev_bufs is None:
                dev_bufs = self.stage_buffers(n, p)
            for i, s in enumerate(range(0, n, rows)):
                e = min(n, s + rows)
                buf = dev_bufs[i & 1][: e - s]
                buf.copy_(X[s:e], non_blocking=True)
                nat.prep(buf, frame.shift, frame.scale, frame.gscale, out=out, row_offset=s)
            return out
        if self._pinned is None or self._pinned[0].shape[1] != p or self._pinned[0].shape[0] < rows:
            for ev in self._pinned_events:
                if ev is not None:
                    ev.synchronize()     # copies out of the old buffers are done before they are dropped
            self._pinned = [torch.empty(rows, p, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._pinned_events = [None, None]
        dev_bufs = [torch.empty(rows, p, dtype=torch.float32, device=self.device) for _ in range(2)]
        # The events outlive this call: the NEXT prepare() must not overwrite a pinned buffer whose
        # asynchronous H2D copy (enqueued by this call) is still in flight.
        events = self._pinned_events
        stream = torch.cuda.current_stream(self.device)
        for i, s in enumerate(range(0, n, rows)):
            e = min(n, s + rows)
            b = i & 1
            if events[b] is not None:
                events[b].synchronize()  # the pinned buffer (and device buffer b) is free again
            src = X[s:e]
            if src.is_pinned() and src.dtype == torch.float32 and src.is_contiguous():
                dev_bufs[b][: e - s].copy_(src, non_blocking=True)
            else:
                self._host_copy(self._pinned[b][: e - s], src)  # never writes into the caller's (maybe memmap) data
                dev_bufs[b][: e - s].copy_(self._pinned[b][: e - s], non_blocking=True)
            nat.prep(dev_bufs[b][: e - s], frame.shift, frame.scale, frame.gscale, out=out, row_offset=s)
            ev = torch.cuda.Event()
            ev.record(stream)
            events[b] = ev
        return out

    def stream_rows(self, X: torch.Tensor, frame: Frame, fn, row_chunk: int = 131072):
        """
        fn(prepared_chunk, first_row) for consecutive row chunks of X.  Host rows are staged (pinned copy,
        H2D, fp16 preparation) on the side stream one chunk AHEAD of the main stream's work on the previous
        chunk, in two alternating buffers -- the caller's kernels and the copies overlap.
        """
        n, p = X.shape
        if n == 0:
            return
        if X.is_cuda or n <= row_chunk:
            for s in range(0, n, row_chunk):
                fn(self.prepare(X[s: s + row_chunk], frame), s)
            return
        main = torch.cuda.current_stream(self.device)
        side = self.staging_stream()
        bufs = [nat.alloc_prepared(row_chunk, p, self.device) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        free = [None, None]
        side.wait_stream(main)              # frame tensors were produced on the main stream

        def stage(k):
            b, s = k & 1, k * row_chunk
            e = min(n, s + row_chunk)
            view = bufs[b] if e - s == row_chunk else bufs[b].head(e - s)
            with torch.cuda.stream(side):
                if free[b] is not None:
                    side.wait_event(free[b])
                self.prepare(X[s:e], frame, out=view)
                ready[b].record(side)
            return view

        nchunks = (n + row_chunk - 1) // row_chunk
        nxt = stage(0)
        for k in range(nchunks):
            cur, b = nxt, k & 1
            main.wait_event(ready[b])
            fn(cur, k * row_chunk)
            free[b] = torch.cuda.Event()
            free[b].record(main)
            if k + 1 < nchunks:
                nxt = stage(k + 1)
        main.wait_stream(side)

    # ------------------------------------------------------------------ numerical operations
    def kmv(self, spec: KernelSpec, frame: Frame, A, B, V, out, symmetric=False, kstore=None):
        return nat.kmv(spec.kid, spec.kscale(frame.gscale), A, B, V, out, symmetric, kstore)

    ktv = staticmethod(nat.ktv)
    kstore_bytes = staticmethod(nat.kstore_bytes)

    def free_bytes(self) -> int:
        return int(torch.cuda.mem_get_info(self.device)[0])

    def kdense(self, spec: KernelSpec, frame: Frame, A, B, out, symmetric=False):
        return nat.kdense(spec.kid, spec.kscale(frame.gscale), A, B, out, symmetric)

    potrf = staticmethod(nat.potrf)
    potrf_outer_blocks = staticmethod(nat.potrf_outer_blocks)
    potrf_begin = staticmethod(nat.potrf_begin)
    potrf_panel = staticmethod(nat.potrf_panel)
    potrf_split = staticmethod(nat.potrf_split)
    potrf_update = staticmethod(nat.potrf_update)
    potrf_update_cyclic = staticmethod(nat.potrf_update_cyclic)
    ltl = staticmethod(nat.ltl)
    add_diag = staticmethod(nat.add_diag)
    trsm_pack = staticmethod(nat.trsm_pack)
    trsm = staticmethod(nat.trsm)
    trsm_pair = staticmethod(nat.trsm_pair)
    coldot = staticmethod(nat.coldot)
    cg_update = staticmethod(nat.cg_update)
    cg_direction = staticmethod(nat.cg_direction)
    axpby = staticmethod(nat.axpby)

    def cg_scratch(self, maxiter: int, dp: int):
        return nat.CgScratch(self.device, maxiter, dp)

    def synchronize(self):
        torch.cuda.synchronize(self.device)


def pad_cols(d: int) -> int:
    """Right-hand sides are padded to a multiple of 4 columns (128-bit shared-memory reads)."""
    return max(4, nat.round_up(d, 4))
