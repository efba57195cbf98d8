"""
ctypes binding of libsobolev_b200.so (the C ABI in include/sobolev_b200.h).

This is the only compute backend of the package: if the shared library is missing, or a call is
made without a CUDA device, it raises -- there is no CPU or eager-PyTorch fallback.  PyTorch is
used for device memory and streams only; every pointer handed to the library is `tensor.data_ptr()`
and every launch goes to `torch.cuda.current_stream()`.
"""

from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

import torch

_LIB_NAME = "libsobolev_b200.so"
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", _LIB_NAME)

KERNEL_IDS = {"gaussian": 0, "laplacian": 1, "matern15": 2, "matern25": 3}
TILE = 128  # order granularity of the dense fp32 routines
NORM_PAD = 256  # norm arrays are padded to a multiple of this
FEAT_PAD = 64  # prepared rows are padded to a multiple of this


class NativeLibraryError(RuntimeError):
    pass


class LaunchStats:
    """
    Launch accounting for bench.py: `launches` counts the CUDA kernels this library enqueued; when
    `time_kmv` is set every fused-kernel launch is bracketed by CUDA events on the launching stream.
    """

    def __init__(self):
        self.reset()
        self.time_kmv = False

    def reset(self):
        self.launches = 0
        self.kmv_events = []   # (flops, start_event, end_event)
        self.trsm_events = []  # (bytes, start_event, end_event)
        self.ktv_events = []   # (bytes, start_event, end_event)

    def kmv_summary(self):
        """(total algorithmic flops, total seconds, launches) over the recorded fused-kernel launches."""
        fl = sum(f for f, _, _ in self.kmv_events)
        ms = sum(a.elapsed_time(b) for _, a, b in self.kmv_events)
        return fl, ms * 1e-3, len(self.kmv_events)

    def ktv_summary(self):
        by = sum(f for f, _, _ in self.ktv_events)
        ms = sum(a.elapsed_time(b) for _, a, b in self.ktv_events)
        return by, ms * 1e-3, len(self.ktv_events)

    def trsm_summary(self):
        by = sum(f for f, _, _ in self.trsm_events)
        ms = sum(a.elapsed_time(b) for _, a, b in self.trsm_events)
        return by, ms * 1e-3, len(self.trsm_events)


STATS = LaunchStats()


_lib = None

_SIGNATURES = {
    "sa_version": (c_int, []),
    "sa_last_error": (c_char_p, []),
    "sa_set_option": (c_int, [c_char_p, c_char_p]),
    "sa_device_info": (c_int, [c_void_p, c_void_p]),
    "sa_prep": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_float,
                        c_void_p, c_int64, c_void_p, c_int64]),
    "sa_prep_log": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_float,
                            c_void_p, c_int64, c_void_p, c_int64]),
    "sa_colstats": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "sa_transform_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_float,
                                  c_void_p, c_int64, c_void_p]),
    "sa_kmv_workspace_bytes": (c_int64, [c_int64, c_int]),
    "sa_kmv": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                       c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64]),
    "sa_kstore_bytes": (c_int64, [c_int64, c_int64]),
    "sa_kmv_store": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                             c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                             c_int64]),
    "sa_ktv_workspace_bytes": (c_int64, [c_int64]),
    "sa_ktv": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_int64]),
    "sa_kdense": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                          c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int]),
    "sa_linalg_workspace_bytes": (c_int64, [c_int64]),
    "sa_potrf": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64]),
    "sa_potrf_outer_blocks": (c_int, []),
    "sa_potrf_begin": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64]),
    "sa_potrf_panel": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int]),
    "sa_potrf_split": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64]),
    "sa_potrf_update": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_int64, c_int,
                                c_void_p, c_int64]),
    "sa_potrf_update_cyclic": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                       c_int64, c_int, c_void_p, c_int64]),
    "sa_ltl": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_float, c_float, c_void_p,
                       c_int64]),
    "sa_ltl_part": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_float, c_float, c_void_p,
                            c_int64, c_int, c_int]),
    "sa_add_diag": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float]),
    "sa_trsm_packed_bytes": (c_int64, [c_int64]),
    "sa_trsm_pack_workspace_bytes": (c_int64, [c_int64]),
    "sa_trsm_pack": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64]),
    "sa_trsm_workspace_bytes": (c_int64, [c_int64]),
    "sa_trsm_solve": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_int64]),
    "sa_trsm_solve_pair": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_float,
                                   c_int, c_int, c_void_p, c_int64]),
    "sa_cg_workspace_bytes": (c_int64, []),
    "sa_coldot": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_float]),
    "sa_cg_update": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_float]),
    "sa_cg_direction": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                c_void_p]),
    "sa_axpby": (c_int, [c_void_p, c_int64, c_float, c_void_p, c_float, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def library_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NativeLibraryError(
                f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `sobolev_alignment_b200/csrc/build.sh` (nvcc, sm_100a). There is no fallback path.")
        lib = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def set_option(name: str, value=None):
    """Tuning / diagnostic option of the library (SAB_*); None restores the environment's value."""
    _check(load().sa_set_option(name.encode(), None if value is None else str(value).encode()))


def _check(rc: int):
    if rc != 0:
        msg = load().sa_last_error()
        raise NativeLibraryError(f"libsobolev_b200 call failed ({rc}): {msg.decode() if msg else '?'}")


def _call(fn, device, *args):
    """
    Enqueue a library call on the CURRENT stream of `device` with that device current (the library
    launches on cudaGetDevice(); the tensors may live on a GPU that is not torch's current one).
    """
    if device.index is None or device.index == torch.cuda.current_device():
        _check(fn(torch.cuda.current_stream(device).cuda_stream, *args))
        return
    with torch.cuda.device(device):
        _check(fn(torch.cuda.current_stream(device).cuda_stream, *args))


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise NativeLibraryError(f"{name} must be a CUDA tensor (the native path has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------
class Prepared:
    """fp16 rows (centred, scaled, zero padded to a multiple of 64 features) + fp32 squared norms."""

    __slots__ = ("data", "norms", "n", "p", "p_pad")

    def __init__(self, data, norms, n, p):
        self.data, self.norms, self.n, self.p, self.p_pad = data, norms, n, p, data.shape[1]

    def rows(self, r0: int, r1: int) -> "Prepared":
        """View of rows [r0, r1) as the A operand of a fused product (r0 a multiple of 128)."""
        return Prepared(self.data[r0:r1], self.norms[r0:], r1 - r0, self.p)

    def head(self, m: int) -> "Prepared":
        """View of the first m rows (the norm padding of the view is re-zeroed by the prep that fills it)."""
        return Prepared(self.data[:m], self.norms[: round_up(max(m, 1), NORM_PAD)], m, self.p)


def prep(X: torch.Tensor, shift: torch.Tensor | None, scale: torch.Tensor | None, gscale: float,
         out: Prepared | None = None, row_offset: int = 0, log10p1: bool = False) -> Prepared:
    """sa_prep (sa_prep_log when `log10p1`: raw counts in) on a [n x p] fp32 CUDA tensor (row stride may exceed p)."""
    lib = load()
    if not X.is_cuda or X.dtype != torch.float32 or X.dim() != 2 or X.stride(1) != 1:
        raise NativeLibraryError("prep expects a 2-D float32 CUDA tensor with unit column stride")
    n, p = X.shape
    if out is None:
        p_pad = round_up(p, FEAT_PAD)
        data = torch.empty(n, p_pad, dtype=torch.float16, device=X.device)
        norms = torch.empty(round_up(max(n, 1), NORM_PAD), dtype=torch.float32, device=X.device)
        out = Prepared(data, norms, n, p)
        norms_len = norms.numel()
        dptr, nptr = data.data_ptr(), norms.data_ptr()
    else:
        # fill rows [row_offset, row_offset + n) of an existing buffer (chunked staging)
        dptr = out.data.data_ptr() + row_offset * out.p_pad * 2
        nptr = out.norms.data_ptr() + row_offset * 4
        last = row_offset + n == out.n
        norms_len = (out.norms.numel() - row_offset) if last else n
    if n == 0:
        return out
    _call(lib.sa_prep_log if log10p1 else lib.sa_prep, X.device, X.data_ptr(), n, p, X.stride(0), _ptr(shift), _ptr(scale),
                       float(gscale), dptr, out.p_pad, nptr, norms_len)
    STATS.launches += 1
    return out


def colstats(X: torch.Tensor, log10p1: bool):
    """Per-column (mean, population variance) of f(X), f = log10(x + 1) or identity; float64 CUDA tensors [p]."""
    lib = load()
    if not X.is_cuda or X.dtype != torch.float32 or X.dim() != 2 or X.stride(1) != 1:
        raise NativeLibraryError("colstats expects a 2-D float32 CUDA tensor with unit column stride")
    n, p = X.shape
    s1 = torch.zeros(p, dtype=torch.float64, device=X.device)
    s2 = torch.zeros(p, dtype=torch.float64, device=X.device)
    # pivot = f(first row): sums of (y - pivot) keep the variance free of cancellation
    pivot = torch.empty(1, p, dtype=torch.float32, device=X.device)
    _call(lib.sa_transform_rows, X.device, X.data_ptr(), 1, p, X.stride(0), int(log10p1), None, None, 1.0,
                                 pivot.data_ptr(), p, None)
    step = 65535 * 4096
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        _call(lib.sa_colstats, X.device, X[r0:r1].data_ptr(), r1 - r0, p, X.stride(0), int(log10p1),
                               pivot.data_ptr(), s1.data_ptr(), s2.data_ptr())
        STATS.launches += 2
    m = s1 / n
    var = (s2 / n - m * m).clamp_min_(0.0)
    return pivot.view(-1).double() + m, var


def transform_rows(X: torch.Tensor, log10p1: bool, shift: torch.Tensor | None, scale: torch.Tensor | None,
                   gain: float = 1.0, want_row_norms: bool = False, in_place: bool = True):
    """
    X <- (f(X) - shift) * scale * gain in place; returns sum_i |row_i|_2 of the result (float64 scalar tensor)
    if asked.  in_place=False: X is left untouched and only the row-norm sum is computed.
    """
    lib = load()
    if not X.is_cuda or X.dtype != torch.float32 or X.dim() != 2 or X.stride(1) != 1:
        raise NativeLibraryError("transform_rows expects a 2-D float32 CUDA tensor with unit column stride")
    n, p = X.shape
    acc = torch.zeros(1, dtype=torch.float64, device=X.device) if want_row_norms else None
    if n:
        _call(lib.sa_transform_rows, X.device, X.data_ptr(), n, p, X.stride(0), int(log10p1), _ptr(shift), _ptr(scale),
              float(gain), X.data_ptr() if in_place else None, X.stride(0), _ptr(acc))
        STATS.launches += 1
    return acc


def alloc_prepared(n: int, p: int, device) -> Prepared:
    p_pad = round_up(p, FEAT_PAD)
    data = torch.empty(n, p_pad, dtype=torch.float16, device=device)
    norms = torch.zeros(round_up(max(n, 1), NORM_PAD), dtype=torch.float32, device=device)
    return Prepared(data, norms, n, p)


_workspaces = {}
_workspaces_lock = threading.Lock()


def _workspace(device, nbytes: int) -> torch.Tensor:
    """
    Scratch of sa_kmv (V tiles), the Cholesky operand panels and the triangular-solve flags: one growing
    buffer per (device, stream) -- reuse is ordered by that stream, and work issued from another stream
    or thread gets a buffer of its own.
    """
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    with _workspaces_lock:
        ws = _workspaces.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            _workspaces[key] = ws
        return ws


def kmv(kernel_id: int, kscale: float, A: Prepared, B: Prepared, V: torch.Tensor, out: torch.Tensor,
        symmetric: bool = False, kstore: torch.Tensor | None = None):
    """
    out[a x dp] += k(A, B) @ V[b x dp]   (out must be zeroed by the caller for a plain product).
    `kstore` (uint8, >= kstore_bytes(a, b)): the kernel block is also left behind as packed fp16 hi/lo tiles for
    `ktv` (sa_kmv_store).
    """
    lib = load()
    _req(V, torch.float32, "V")
    _req(out, torch.float32, "out")
    dp = V.shape[1]
    if V.shape[0] != B.n or out.shape != (A.n, dp):
        raise ValueError(f"shape mismatch: V {tuple(V.shape)}, out {tuple(out.shape)}, a={A.n}, b={B.n}")
    if A.p_pad != B.p_pad:
        raise ValueError("A and B were prepared with different feature counts")
    ev = None
    if STATS.time_kmv:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(torch.cuda.current_stream(V.device))
    nbytes = int(lib.sa_kmv_workspace_bytes(B.n, dp))
    if nbytes < 0:
        raise NativeLibraryError(f"dp must be a multiple of 4 in [4, 32], got {dp}")
    ws = _workspace(V.device, nbytes)
    if kstore is None:
        _call(lib.sa_kmv, V.device, kernel_id, float(kscale), A.data.data_ptr(), A.norms.data_ptr(), A.n,
              A.p_pad, B.data.data_ptr(), B.norms.data_ptr(), B.n, B.p_pad, A.p_pad,
              V.data_ptr(), dp, out.data_ptr(), int(symmetric), ws.data_ptr(), ws.numel())
    else:
        _call(lib.sa_kmv_store, V.device, kernel_id, float(kscale), A.data.data_ptr(), A.norms.data_ptr(), A.n,
              A.p_pad, B.data.data_ptr(), B.norms.data_ptr(), B.n, B.p_pad, A.p_pad,
              V.data_ptr(), dp, out.data_ptr(), ws.data_ptr(), ws.numel(), kstore.data_ptr(), kstore.numel())
    if ev is not None:
        ev[1].record(torch.cuda.current_stream(V.device))
        # algorithmic work: 2 a b p + 2 a b d, with the caller's p and padded d
        STATS.kmv_events.append((2.0 * A.n * B.n * (A.p + dp), ev[0], ev[1]))
    STATS.launches += 3   # column maxima, fp16 split of V, fused kernel
    return out


def kstore_bytes(a: int, b: int) -> int:
    return int(load().sa_kstore_bytes(a, b))


def ktv(kstore: torch.Tensor, a: int, b: int, t: torch.Tensor, out: torch.Tensor):
    """out[b x dp] += k(A, B)^T t[a x dp] from the block a previous kmv(..., kstore=) left behind."""
    _req(t, torch.float32, "t")
    _req(out, torch.float32, "out")
    dp = t.shape[1]
    if t.shape[0] != a or out.shape != (b, dp):
        raise ValueError(f"shape mismatch: t {tuple(t.shape)}, out {tuple(out.shape)}, a={a}, b={b}")
    ev = None
    if STATS.time_kmv:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(torch.cuda.current_stream(t.device))
    ws = _workspace(t.device, int(load().sa_ktv_workspace_bytes(a)))
    _call(load().sa_ktv, t.device, kstore.data_ptr(), a, b, t.data_ptr(), dp, out.data_ptr(), ws.data_ptr(), ws.numel())
    if ev is not None:
        ev[1].record(torch.cuda.current_stream(t.device))
        STATS.ktv_events.append((4.0 * a * b, ev[0], ev[1]))     # the stored block is read once: 4 bytes per entry
    STATS.launches += 2
    return out


def kdense(kernel_id: int, kscale: float, A: Prepared, B: Prepared, out: torch.Tensor,
           symmetric: bool = False):
    """out[:a, :b] = k(A, B); `out` is a 2-D fp32 CUDA tensor whose row stride may exceed b."""
    lib = load()
    if not out.is_cuda or out.dtype != torch.float32 or out.stride(1) != 1:
        raise NativeLibraryError("kdense output must be a float32 CUDA tensor with unit column stride")
    if out.shape[0] < A.n or out.shape[1] < B.n:
        raise ValueError("output too small")
    _call(lib.sa_kdense, out.device, kernel_id, float(kscale), A.data.data_ptr(), A.norms.data_ptr(), A.n,
                         A.p_pad, B.data.data_ptr(), B.norms.data_ptr(), B.n, B.p_pad, A.p_pad,
                         out.data_ptr(), out.stride(0), int(symmetric))
    STATS.launches += 1
    return out


def _linalg_workspace(device, m: int) -> torch.Tensor:
    """Scratch for the fp16 hi/lo operand panels of the tensor-core Cholesky updates (shared with kmv's)."""
    nbytes = int(load().sa_linalg_workspace_bytes(m))
    if nbytes < 0:
        raise NativeLibraryError("bad matrix order")
    return _workspace(device, nbytes)


def potrf(A: torch.Tensor, invdiag: torch.Tensor, info: torch.Tensor):
    _req(A, torch.float32, "A")
    _req(invdiag, torch.float32, "invdiag")
    ws = _linalg_workspace(A.device, A.shape[0])
    _call(load().sa_potrf, A.device, A.data_ptr(), A.shape[0], A.stride(0), invdiag.data_ptr(),
                           info.data_ptr(), ws.data_ptr(), ws.numel())
    nblk = A.shape[0] // TILE
    STATS.launches += 2 + nblk + 3 * (nblk - 1)   # scale, diagonal blocks, (panel, split, update) per step


# ---- the factorisation in steps (row-sharded solver: outer panels dealt out over the ranks)
def potrf_outer_blocks() -> int:
    """Width of an outer panel in 128-blocks (0: the fp32 SIMT cross-check path is selected)."""
    return int(load().sa_potrf_outer_blocks())


def potrf_begin(A: torch.Tensor, info: torch.Tensor):
    _req(A, torch.float32, "A")
    ws = _linalg_workspace(A.device, A.shape[0])
    _call(load().sa_potrf_begin, A.device, A.data_ptr(), A.shape[0], A.stride(0), info.data_ptr(), ws.data_ptr(),
          ws.numel())
    STATS.launches += 2


def potrf_panel(A: torch.Tensor, invdiag: torch.Tensor, info: torch.Tensor, k0: int, kend: int):
    _call(load().sa_potrf_panel, A.device, A.data_ptr(), A.shape[0], A.stride(0), invdiag.data_ptr(), info.data_ptr(),
          int(k0), int(kend))
    STATS.launches += 3 * (kend - k0)


def potrf_split(m: int, panel: torch.Tensor, W: int):
    """`panel`: [rows x W] fp32 view (unit column stride) of the finished panel below its diagonal blocks."""
    if panel.dtype != torch.float32 or panel.stride(1) != 1 or panel.shape[1] != W:
        raise ValueError("panel must be a float32 [rows x W] view with unit column stride")
    ws = _linalg_workspace(panel.device, m)
    _call(load().sa_potrf_split, panel.device, m, panel.data_ptr(), panel.stride(0), panel.shape[0], int(W),
          ws.data_ptr(), ws.numel())
    STATS.launches += 1


def potrf_update(A: torch.Tensor, kend: int, c0: int, c1: int, panel: torch.Tensor, W: int):
    ws = _linalg_workspace(A.device, A.shape[0])
    _call(load().sa_potrf_update, A.device, A.data_ptr(), A.shape[0], A.stride(0), int(kend), int(c0), int(c1),
          panel.data_ptr(), panel.stride(0), int(W), ws.data_ptr(), ws.numel())
    STATS.launches += 3


def potrf_update_cyclic(A: torch.Tensor, kend: int, nb: int, first: int, stride: int, skip: int, panel: torch.Tensor,
                        W: int):
    """potrf_update for every outer panel first, first + stride, ... (except `skip`) in one library call."""
    ws = _linalg_workspace(A.device, A.shape[0])
    _call(load().sa_potrf_update_cyclic, A.device, A.data_ptr(), A.shape[0], A.stride(0), int(kend), int(nb), int(first),
          int(stride), int(skip), panel.data_ptr(), panel.stride(0), int(W), ws.data_ptr(), ws.numel())
    npan = (A.shape[0] // TILE + nb - 1) // nb
    STATS.launches += 3 * max(0, (npan - first + stride - 1) // stride)


def ltl(L: torch.Tensor, G: torch.Tensor, alpha: float, beta: float, part: int = 0, nparts: int = 1):
    """G = alpha L^T L + beta I; nparts > 1: this rank's share of the slabs accumulated into a zeroed G (no beta)."""
    _req(L, torch.float32, "L")
    _req(G, torch.float32, "G")
    ws = _linalg_workspace(L.device, L.shape[0])
    _call(load().sa_ltl_part, L.device, L.data_ptr(), L.shape[0], L.stride(0), G.data_ptr(), G.stride(0),
                              float(alpha), float(beta), ws.data_ptr(), ws.numel(), int(part), int(nparts))
    STATS.launches += 3 + 2 * (L.shape[0] // TILE)   # scale, (split, update) per block row, diagonal


def add_diag(A: torch.Tensor, m: int, value: float):
    _req(A, torch.float32, "A")
    _call(load().sa_add_diag, A.device, A.data_ptr(), m, A.stride(0), float(value))
    STATS.launches += 1


class PackedFactor:
    """Solve-ready form of a Cholesky factor (sa_trsm_pack): fp16 hi/lo tiles + inverse / pre-multiplied blocks."""

    __slots__ = ("data", "m")

    def __init__(self, data: torch.Tensor, m: int):
        self.data, self.m = data, m


def trsm_pack(L: torch.Tensor, invdiag: torch.Tensor) -> PackedFactor:
    """Pack the lower factor L (fp32 [m x m], m % 128 == 0) for the triangular solves; L may be freed afterwards."""
    _req(L, torch.float32, "L")
    _req(invdiag, torch.float32, "invdiag")
    m = L.shape[0]
    nbytes = int(load().sa_trsm_packed_bytes(m))
    if nbytes < 0:
        raise NativeLibraryError("matrix order must be a positive multiple of 128")
    data = torch.empty(nbytes, dtype=torch.uint8, device=L.device)
    ws = _workspace(L.device, int(load().sa_trsm_pack_workspace_bytes(m)))
    _call(load().sa_trsm_pack, L.device, L.data_ptr(), m, L.stride(0), invdiag.data_ptr(), data.data_ptr(),
          ws.data_ptr(), ws.numel())
    STATS.launches += 7
    return PackedFactor(data, m)


def _timed(device):
    if not STATS.time_kmv:
        return None
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev[0].record(torch.cuda.current_stream(device))
    return ev


def trsm(F: PackedFactor, B: torch.Tensor, transpose: bool):
    """B <- L^-1 B (transpose False) or L^-T B (True), in place, with a packed factor."""
    _req(B, torch.float32, "B")
    if B.shape[0] != F.m:
        raise ValueError("rhs rows must equal the (padded) matrix order")
    ev = _timed(B.device)
    ws = _workspace(B.device, int(load().sa_trsm_workspace_bytes(F.m)))
    _call(load().sa_trsm_solve, B.device, F.data.data_ptr(), F.m, B.data_ptr(), B.shape[1], int(transpose),
          ws.data_ptr(), ws.numel())
    if ev is not None:
        ev[1].record(torch.cuda.current_stream(B.device))
        STATS.trsm_events.append((F.m * (F.m + TILE) / 2 * 4.0, ev[0], ev[1]))   # the triangle is read once
    STATS.launches += 1
    return B


def trsm_pair(Fa: PackedFactor, Fb: PackedFactor, Xa: torch.Tensor, Xb: torch.Tensor, V: torch.Tensor | None,
              lam: float, transpose: bool):
    """Xa <- op(La)^-1 Xa ;  Xb <- op(Lb)^-1 (Xa + lam V)   -- two interleaved chained solves in one launch."""
    _req(Xa, torch.float32, "Xa")
    _req(Xb, torch.float32, "Xb")
    if V is not None:
        _req(V, torch.float32, "V")
    m = Fa.m
    if Fb.m != m or Xa.shape != Xb.shape or Xa.shape[0] != m or (V is not None and V.shape != Xa.shape):
        raise ValueError("shape mismatch")
    if Xa.data_ptr() == Xb.data_ptr():
        raise ValueError("Xa and Xb must be different buffers")
    ev = _timed(Xa.device)
    ws = _workspace(Xa.device, int(load().sa_trsm_workspace_bytes(m)))
    _call(load().sa_trsm_solve_pair, Xa.device, Fa.data.data_ptr(), Fb.data.data_ptr(), m, Xa.data_ptr(),
          Xb.data_ptr(), _ptr(V), float(lam), Xa.shape[1], int(transpose), ws.data_ptr(), ws.numel())
    if ev is not None:
        ev[1].record(torch.cuda.current_stream(Xa.device))
        STATS.trsm_events.append((2 * m * (m + TILE) / 2 * 4.0, ev[0], ev[1]))   # both triangles are read once
    STATS.launches += 1
    return Xb


class CgScratch:
    """
    Device scratch of one CG run: reduction slots of the deterministic column sums, the stop flag the
    kernels raise on convergence, the residual history [maxiter x dp], and pinned slots through which
    the host reads the flag a few iterations late (no per-iteration synchronisation).
    """

    def __init__(self, device, maxiter: int, dp: int):
        self.ws = torch.zeros(int(load().sa_cg_workspace_bytes()) // 4, dtype=torch.float32, device=device)
        self.stop = torch.zeros(1, dtype=torch.int32, device=device)
        self.hist = torch.zeros(max(1, maxiter), dp, dtype=torch.float32, device=device)
        self._pinned = torch.zeros(max(1, maxiter), dtype=torch.int32).pin_memory()
        self._events = {}

    def poll(self, slot: int):
        """Enqueue an asynchronous copy of the stop flag as it stands after the work enqueued so far."""
        self._pinned[slot : slot + 1].copy_(self.stop, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.stop.device))
        self._events[slot] = ev

    def stopped_at(self, slot: int) -> bool:
        """Value of the flag at poll `slot` (waits for that copy only; the same on every rank)."""
        ev = self._events.get(slot)
        if ev is None:
            return False
        ev.synchronize()
        return bool(self._pinned[slot].item())

    def history(self) -> torch.Tensor:
        return self.hist.detach().cpu()


def _hist_row(cgs: CgScratch, row):
    return None if row is None else cgs.hist.data_ptr() + row * cgs.hist.stride(0) * 4


def coldot(P: torch.Tensor, Q: torch.Tensor, out: torch.Tensor, cgs: CgScratch, hist_row=None, tol: float = 0.0):
    """out[c] = sum_i P[i,c] Q[i,c]; with `hist_row` the sums are also the residual of that iteration."""
    _req(P, torch.float32, "P")
    _req(Q, torch.float32, "Q")
    track = hist_row is not None
    _call(load().sa_coldot, P.device, P.data_ptr(), Q.data_ptr(), P.shape[0], P.shape[1], out.data_ptr(),
          cgs.ws.data_ptr(), _hist_row(cgs, hist_row), cgs.stop.data_ptr() if track else None, float(tol))
    STATS.launches += 1
    return out


def cg_update(P, AP, X, R, rs_old, pAp, eps: float, rs_new, cgs: CgScratch, hist_row=None, tol: float = 0.0):
    _req(P, torch.float32, "P")
    _call(load().sa_cg_update, P.device, P.shape[0], P.shape[1], P.data_ptr(), _ptr(AP), X.data_ptr(),
          _ptr(R), rs_old.data_ptr(), pAp.data_ptr(), float(eps), _ptr(rs_new), cgs.ws.data_ptr(),
          _hist_row(cgs, hist_row), cgs.stop.data_ptr(), float(tol))
    STATS.launches += 1


def cg_direction(R, P, rs_new, rs_old, eps: float, cgs: CgScratch):
    _call(load().sa_cg_direction, P.device, P.shape[0], P.shape[1], R.data_ptr(), P.data_ptr(),
          rs_new.data_ptr(), rs_old.data_ptr(), float(eps), cgs.stop.data_ptr())
    STATS.launches += 1


def axpby(a: float, X: torch.Tensor, b: float, Y: torch.Tensor):
    _req(X, torch.float32, "X")
    _req(Y, torch.float32, "Y")
    if X.numel() != Y.numel():
        raise ValueError("size mismatch")
    _call(load().sa_axpby, X.device, X.numel(), float(a), X.data_ptr(), float(b), Y.data_ptr())
    STATS.launches += 1
    return Y


def device_info():
    sms, smem = c_int(0), c_int(0)
    _check(load().sa_device_info(ctypes.byref(sms), ctypes.byref(smem)))
    return sms.value, smem.value
