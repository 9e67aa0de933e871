"""Member-sharded ensembles over the GPUs of one node (one process per GPU).

Forward runs need no communication: each rank simulates its block of members.
The analysis has ONE exchange step:

* ``all_gather`` of the small predicted-data block ``Eo`` ``(N_local,p) -> (N,p)``;
* ``all_to_all`` re-sharding the parameter ensemble from member rows
  ``(N_local, M)`` to parameter columns ``(N, M_local)``, so that every rank can
  run the update (global ES or per-parameter local analysis) on its columns
  with all members present, and a second ``all_to_all`` back.

The collectives are ``torch.distributed`` (NCCL over NVLink on the GPU box;
gloo on CPU in the tests).
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _all_to_all(recv, send):
    """List all_to_all; gloo has no alltoall, so fall back to paired send/recv there."""
    if dist.get_backend() == "nccl":
        dist.all_to_all(recv, send)
        return
    rank, size = world()
    recv[rank].copy_(send[rank])
    ops = []
    for r in range(size):
        if r != rank:
            ops.append(dist.P2POp(dist.isend, send[r], r))
            ops.append(dist.P2POp(dist.irecv, recv[r], r))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def member_slice(N, rank=None, size=None):
    """Block partition of N members: rank r owns [lo, hi)."""
    r, s = world()
    rank = r if rank is None else rank
    size = s if size is None else size
    base, rem = divmod(N, size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def column_slice(M, rank=None, size=None):
    return member_slice(M, rank, size)


def gather_members(x_local, N):
    """(N_local, p) -> (N, p), members in rank order (uneven blocks allowed)."""
    rank, size = world()
    if size == 1:
        return x_local
    counts = [hi - lo for lo, hi in (member_slice(N, r, size) for r in range(size))]
    nmax = max(counts)  # equal-size blocks for the collective; uneven tails are padded
    pad = torch.zeros((nmax, *x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    parts = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(parts, pad)
    return torch.cat([q[:n] for q, n in zip(parts, counts)], 0)


def members_to_columns(E_local, N):
    """(N_local, M) member rows -> (N, M_local) parameter columns (one all_to_all)."""
    rank, size = world()
    if size == 1:
        return E_local
    M = E_local.shape[1]
    send = [E_local[:, slice(*column_slice(M, r, size))].contiguous() for r in range(size)]
    lo, hi = column_slice(M, rank, size)
    recv = [torch.empty((mhi - mlo, hi - lo), dtype=E_local.dtype, device=E_local.device)
            for mlo, mhi in (member_slice(N, r, size) for r in range(size))]
    _all_to_all(recv, send)
    return torch.cat(recv, 0)


def columns_to_members(E_cols, N, M):
    """(N, M_local) parameter columns -> (N_local, M) member rows (the inverse all_to_all)."""
    rank, size = world()
    if size == 1:
        return E_cols
    send = [E_cols[slice(*member_slice(N, r, size))].contiguous() for r in range(size)]
    lo, hi = member_slice(N, rank, size)
    recv = [torch.empty((hi - lo, chi - clo), dtype=E_cols.dtype, device=E_cols.device)
            for clo, chi in (column_slice(M, r, size) for r in range(size))]
    _all_to_all(recv, send)
    return torch.cat(recv, 1)


def sharded_update(update_fn, E_local, Eo_local, N, **kw):
    """Run ``update_fn(E_cols, Eo_full, **kw)`` on this rank's parameter columns.

    ``update_fn`` is e.g. ``analysis.ens_update0`` or ``ens_update0_loc`` (whose
    ``taper`` keyword must then already be restricted to this rank's columns);
    ``perturbs`` in ``kw`` is the FULL ``(N,p)`` block, identical on every rank.
    Returns the updated ``(N_local, M)`` block.
    """
    M = E_local.shape[1]
    Eo = gather_members(Eo_local, N)
    E_cols = members_to_columns(E_local, N)
    E_cols = update_fn(E_cols, Eo, **kw)
    return columns_to_members(E_cols, N, M)
