"""
Multi-GPU pieces of the Nystrom preconditioner ("not sharded ... the Amdahl term").

The fused `kmv` products divide by the number of ranks; what is left replicated -- the two Cholesky
factorisations of the M x M matrices -- then bounds the speed-up (0.50 s of a 2.16 s fit on 8 GPUs in
round 1).  `potrf` below deals the 512-column outer panels of the two-level factorisation out over the
ranks (block-cyclic): the owner factors a panel, the panel travels by one NCCL broadcast over NVLink, and
every rank applies it -- on the tensor cores -- only to the outer panels it owns.  Every rank ends with
the complete factor (it has received every panel), bit-identical on all ranks, so the triangular solves
stay replicated without a further exchange.

One step of look-ahead hides the sequential part: the owner of panel j + 1 brings that panel up to date
first, factors it and sends it off BEFORE applying panel j to the rest of its share, so the broadcast and
the next in-panel factorisation overlap everybody's trailing updates.

Reference interface: falkon `FalkonPreconditioner.init` (potrf) [external dependency], reached from
/root/reference/sobolev_alignment/krr_approx.py:227.
"""

from __future__ import annotations

import torch

TILE = 128


def potrf(ops, A: torch.Tensor, invdiag: torch.Tensor, info: torch.Tensor, comm) -> None:
    """In-place lower Cholesky of A (replicated, identical on all ranks) with the work shared by the ranks of `comm`."""
    nb = ops.potrf_outer_blocks() if hasattr(ops, "potrf_outer_blocks") else 0
    Mp = A.shape[0]
    nblk = Mp // TILE
    npan = 0 if nb == 0 else (nblk + nb - 1) // nb
    if comm is None or not comm.enabled or comm.world == 1 or npan < 2:
        ops.potrf(A, invdiag, info)
        return
    R, me = comm.world, comm.rank
    Wmax = nb * TILE
    tail = nb * TILE * TILE + 4                      # inverse diagonal blocks + info word
    bufs = [ops.empty(Mp * Wmax + tail) for _ in range(2)]

    def span(j):
        k0 = j * nb
        return k0, min(nblk, k0 + nb)

    def payload(j):
        """(flat message, [rows x W] panel view, inverse-block view, info view) of panel j inside its buffer."""
        k0, kend = span(j)
        rows, W = Mp - k0 * TILE, (kend - k0) * TILE
        buf = bufs[j % 2]
        n_inv = (kend - k0) * TILE * TILE
        msg = buf[: rows * W + n_inv + 4]
        return (msg, msg[: rows * W].view(rows, W), msg[rows * W: rows * W + n_inv].view(kend - k0, TILE, TILE),
                msg[rows * W + n_inv:].view(torch.int32) if msg.dtype == torch.float32 else msg[rows * W + n_inv:])

    def factor_and_pack(j):
        k0, kend = span(j)
        ops.potrf_panel(A, invdiag, info, k0, kend)
        msg, pv, iv, nfo = payload(j)
        pv.copy_(A[k0 * TILE:, k0 * TILE: kend * TILE])
        iv.copy_(invdiag[k0:kend])
        nfo[:1].copy_(info.to(nfo.dtype))

    ops.potrf_begin(A, info)
    if me == 0 % R:
        factor_and_pack(0)
    handle = comm.broadcast_async(payload(0)[0], 0 % R)
    for j in range(npan):
        k0, kend = span(j)
        W = (kend - k0) * TILE
        handle.wait()
        msg, pv, iv, nfo = payload(j)
        if j % R != me:
            # the finished panel is part of the factor every rank keeps
            A[k0 * TILE:, k0 * TILE: kend * TILE].copy_(pv)
            invdiag[k0:kend].copy_(iv)
            got = nfo[:1].to(info.dtype)
            info.copy_(torch.where(info != 0, info, got))
        if kend == nblk:
            break
        below = pv[W:]                                # rows under the panel's own diagonal blocks
        ops.potrf_split(Mp, below, W)
        ahead = (j + 1) % R == me
        if ahead:
            c0, c1 = span(j + 1)
            ops.potrf_update(A, kend, c0, c1, below, W)
            factor_and_pack(j + 1)
        handle = comm.broadcast_async(payload(j + 1)[0], (j + 1) % R)
        # the rest of this rank's share: outer panels first, first + R, ... beyond j + 1 (one library call)
        first = j + 1 + ((me - (j + 1)) % R)
        if first < npan:
            ops.potrf_update_cyclic(A, kend, nb, first, R, j + 1 if ahead else -1, below, W)
