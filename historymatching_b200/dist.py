"""Member-sharded ensembles over the GPUs of one node (one process per GPU).

Forward runs need no communication: each rank simulates its block of members.
The analysis has ONE exchange step:

* ``all_gather`` of the small predicted-data block ``Eo`` ``(N_local,p) -> (N,p)``;
* an all-to-all re-sharding the parameter ensemble from member rows
  ``(N_local, M)`` to parameter columns ``(N, M_local)``, so that every rank can
  run the update (global ES, per-parameter local analysis, or the recomposition
  ``x0 + W X0`` of the iterative smoother) on its columns with all members
  present, and a second all-to-all back.

The collectives are ``torch.distributed`` (NCCL over NVLink on the GPU box; gloo
on CPU in the tests).  The all-to-all is ONE ``all_to_all_single`` on
pre-allocated buffers (``Resharder``): member rows -> columns needs one strided
pack (``hm_copy2d``) and lands directly in the ``(N, M_local)`` layout; columns ->
member rows sends contiguous row blocks and needs one strided unpack.

Product API of the sharded cycle (the loop the reference writes in its cells,
``HistoryMatch.py:652, 906-944, 961``):

* ``es_update_sharded`` / ``les_update_sharded``  - one analysis step;
* ``es_mda_sharded``   - ``Na`` x (forward run of the local members, sharded ES update);
* ``ies_sharded``      - iterative smoother: the ``N x N`` weight matrix ``W`` is
  replicated (every rank takes the same Gauss-Newton step from the gathered
  predicted data), the recomposition ``E = x0 + W X0`` runs on this rank's
  parameter columns, the all-to-all returns member rows for the next forward run;
* ``iles_sharded``     - localised iterative smoother: the per-parameter weight matrices live
  on this rank's parameter columns, step and recomposition are local
  (``HistoryMatch.py:1007-1064``).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


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


def _all_to_all_single(out, inp, out_splits, in_splits):
    """``dist.all_to_all_single`` on flat buffers; gloo has no all-to-all: paired send / recv there."""
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(out, inp, out_splits, in_splits)
        return
    rank, size = world()
    oo = np.concatenate([[0], np.cumsum(out_splits)])
    io = np.concatenate([[0], np.cumsum(in_splits)])
    out[oo[rank]:oo[rank + 1]].copy_(inp[io[rank]:io[rank + 1]])
    ops = []
    for r in range(size):
        if r != rank:
            if in_splits[r]:
                ops.append(dist.P2POp(dist.isend, inp[io[r]:io[r + 1]], r))
            if out_splits[r]:
                ops.append(dist.P2POp(dist.irecv, out[oo[r]:oo[r + 1]], r))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()


def _copy2d(dst, src):
    """``dst[:, :] = src[:, :]`` for 2-D float64 views with unit column stride: ``hm_copy2d`` on the GPU."""
    if dst.numel() == 0:
        return
    if dst.is_cuda:
        from . import _lib

        assert dst.dtype == torch.float64 and src.dtype == torch.float64 and dst.stride(1) == 1 and src.stride(1) == 1
        ctx = _lib.Context.get(dst.device.index if dst.device.index is not None else torch.cuda.current_device())
        ctx.use_torch_stream()
        _lib.check(ctx.lib.hm_copy2d(ctx.handle, dst.shape[0], dst.shape[1], C.c_void_p(src.data_ptr()), src.stride(0),
                                     C.c_void_p(dst.data_ptr()), dst.stride(0)))
    else:
        dst.copy_(src)


class Resharder:
    """Member rows ``(N_local, M)`` <-> parameter columns ``(N, M_local)`` with pre-allocated exchange buffers."""

    _cache: dict = {}

    def __init__(self, N, M, device, dtype=torch.float64):
        self.N, self.M = N, M
        self.rank, self.size = world()
        self.members = [member_slice(N, r, self.size) for r in range(self.size)]
        self.columns = [column_slice(M, r, self.size) for r in range(self.size)]
        self.n_loc = self.members[self.rank][1] - self.members[self.rank][0]
        self.m_loc = self.columns[self.rank][1] - self.columns[self.rank][0]
        # flat exchange buffers: [chunk for rank 0 | chunk for rank 1 | ...]
        self.rows_buf = torch.empty(self.n_loc * M, dtype=dtype, device=device)   # this rank's members, all columns
        self.cols_buf = torch.empty(N * self.m_loc, dtype=dtype, device=device)   # all members, this rank's columns
        self.rows_splits = [self.n_loc * (hi - lo) for lo, hi in self.columns]     # per column owner
        self.cols_splits = [(hi - lo) * self.m_loc for lo, hi in self.members]     # per member owner

    @classmethod
    def get(cls, N, M, device, dtype=torch.float64):
        key = (N, M, str(device), dtype, world())
        if key not in cls._cache:
            cls._cache[key] = Resharder(N, M, device, dtype)
        return cls._cache[key]

    def to_columns(self, E_local):
        """``(N_local, M)`` -> ``(N, M_local)``; the result is a view of the exchange buffer (valid until the next call)."""
        if self.size == 1:
            return E_local
        assert E_local.shape == (self.n_loc, self.M)
        off = 0
        for (lo, hi), n in zip(self.columns, self.rows_splits):   # pack: column block of every destination, contiguous
            _copy2d(self.rows_buf[off:off + n].view(self.n_loc, hi - lo), E_local[:, lo:hi])
            off += n
        _all_to_all_single(self.cols_buf, self.rows_buf, self.cols_splits, self.rows_splits)
        return self.cols_buf.view(self.N, self.m_loc)   # source ranks in member order: already (N, M_local) row-major

    def to_members(self, E_cols, out=None):
        """``(N, M_local)`` -> ``(N_local, M)`` (into ``out`` if given)."""
        if self.size == 1:
            return E_cols
        assert E_cols.shape == (self.N, self.m_loc) and E_cols.is_contiguous()
        # row blocks of every destination are contiguous in E_cols: send in place
        _all_to_all_single(self.rows_buf, E_cols.reshape(-1), self.rows_splits, self.cols_splits)
        if out is None:
            out = torch.empty((self.n_loc, self.M), dtype=E_cols.dtype, device=E_cols.device)
        off = 0
        for (lo, hi), n in zip(self.columns, self.rows_splits):   # unpack: column block received from every owner
            _copy2d(out[:, lo:hi], self.rows_buf[off:off + n].view(self.n_loc, hi - lo))
            off += n
        return out


def gather_members(x_local, N):
    """(N_local, p) -> (N, p), members in rank order (uneven blocks allowed)."""
    rank, size = world()
    if size == 1:
        return x_local
    counts = [hi - lo for lo, hi in (member_slice(N, r, size) for r in range(size))]
    tail = x_local.shape[1:]
    if len(set(counts)) == 1:
        out = torch.empty((N, *tail), dtype=x_local.dtype, device=x_local.device)
        dist.all_gather_into_tensor(out, x_local.contiguous())
        return out
    nmax = max(counts)  # equal-size blocks for the collective; uneven tails are padded
    pad = torch.zeros((nmax, *tail), dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    parts = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(parts, pad)
    return torch.cat([q[:n] for q, n in zip(parts, counts)], 0)


def members_to_columns(E_local, N):
    """(N_local, M) member rows -> (N, M_local) parameter columns (one all-to-all)."""
    if world()[1] == 1:
        return E_local
    return Resharder.get(N, E_local.shape[1], E_local.device, E_local.dtype).to_columns(E_local.contiguous())


def columns_to_members(E_cols, N, M):
    """(N, M_local) parameter columns -> (N_local, M) member rows (the inverse all-to-all)."""
    if world()[1] == 1:
        return E_cols
    return Resharder.get(N, M, E_cols.device, E_cols.dtype).to_members(E_cols.contiguous())


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


# ---- the sharded cycle -----------------------------------------------------------------------------------------
def es_update_sharded(E_local, Eo_local, N, obs, perturbs, decorr):
    """``ens_update0`` (``HistoryMatch.py:578-586``) of a member-sharded ensemble; ``perturbs`` is the full ``(N,p)`` block."""
    from . import analysis as ha

    return sharded_update(ha.ens_update0, E_local, Eo_local, N, obs=obs, perturbs=perturbs, decorr=decorr)


def les_update_sharded(E_local, Eo_local, N, obs, perturbs, decorr, taper_cols):
    """``ens_update0_loc`` (``HistoryMatch.py:774-797``); ``taper_cols`` = the taper rows of this rank's parameter columns."""
    from . import analysis as ha

    return sharded_update(ha.ens_update0_loc, E_local, Eo_local, N, obs=obs, perturbs=perturbs, decorr=decorr,
                          taper=taper_cols)


def _f64_on(x, device):
    """float64 tensor of ``x`` (numpy array or tensor on any device) on ``device``."""
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.float64)
    return torch.as_tensor(np.asarray(x, float), device=device)


def _normal_block(N, p, seed, device):
    """The same standard-normal ``(N,p)`` block on every rank (a seeded generator, not the rank's global stream)."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    return torch.randn(N, p, dtype=torch.float64, generator=g).to(device)


def es_mda_sharded(forward, E_local, N, obs, R12, alphas, decorr=None, seed=0, perturbs=None):
    """ES-MDA of a member-sharded ensemble: ``len(alphas)`` x (forward run of this rank's members,
    sharded ES update with ``R -> alpha_i R``).

    ``forward``   callable ``(N_local, M) -> (N_local, p)`` (e.g. ``lambda E: case.forward(E)[0]``);
    ``perturbs``  optional list of standard-normal ``(N,p)`` blocks, identical on every rank; default: drawn from
                  ``seed + i`` on every rank.
    Returns ``(E_local, stats)`` with ``stats["Eo"]`` the local predicted data of every pass.
    """
    alphas = np.asarray(alphas, float)
    if not np.isclose(np.sum(1 / alphas), 1.0):
        raise ValueError("ES-MDA inflation factors must satisfy sum(1/alpha) == 1")
    dev = E_local.device
    R12T = torch.as_tensor(np.ascontiguousarray(np.asarray(R12, float).T), device=dev)
    if decorr is None:
        decorr = np.linalg.inv(np.asarray(R12, float).T)
    dec = torch.as_tensor(np.asarray(decorr, float), device=dev)
    y = _f64_on(obs, dev)
    p = y.shape[0]
    stats = dict(Eo=[])
    E = E_local
    for i, a in enumerate(alphas):
        Eo = forward(E)
        stats["Eo"].append(Eo)
        Z = _normal_block(N, p, seed + i, dev) if perturbs is None else torch.as_tensor(perturbs[i], device=dev)
        E = es_update_sharded(E, Eo, N, y, np.sqrt(a) * (Z @ R12T), dec / np.sqrt(a))
    return E, stats


def ies_sharded(forward, E_local, N, obs, perturbs, decorr, xStep=1.0, iMax=4):
    """Iterative ensemble smoother (``HistoryMatch.py:906-944``) of a member-sharded ensemble.

    ``W (N,N)`` is replicated: every rank gathers the predicted data and takes the same Gauss-Newton step
    (``hm_ies_step``); anomalies ``X0`` and mean ``x0`` live on this rank's parameter columns, where the
    recomposition ``E = x0 + W X0`` is one FP64 tensor-core GEMM; an all-to-all returns member rows for the next
    forward run.  ``perturbs`` is the full ``(N,p)`` block, identical on every rank.
    Returns ``(E_local, stats)``; ``stats["Eo"]`` holds the gathered predicted data ``(N,p)`` of every iteration.
    """
    from . import _lib
    from . import analysis as ha

    dev = E_local.device
    M = E_local.shape[1]
    if dev.type != "cuda":
        raise _lib.HmError("ies_sharded runs on CUDA tensors (no CPU fallback)")
    y, pert, dec = (_f64_on(x, dev) for x in (obs, perturbs, decorr))
    p = y.shape[0]
    ctx = _lib.Context.get(dev.index if dev.index is not None else torch.cuda.current_device())
    ctx.use_torch_stream()
    rs = Resharder.get(N, M, dev)
    E_cols = members_to_columns(E_local.contiguous(), N)
    m_loc = E_cols.shape[1]
    X0 = torch.empty((N, m_loc), dtype=torch.float64, device=dev)
    x0 = torch.empty(m_loc, dtype=torch.float64, device=dev)
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(ctx.lib.hm_center(ctx.handle, N, m_loc, ptr(E_cols), m_loc, ptr(X0), m_loc, ptr(x0), 0))
    W = torch.eye(N, dtype=torch.float64, device=dev)
    stats = dict(Eo=[])
    E = E_local
    for it in range(iMax + 1):
        Ec = ha._recompose(ctx, x0, W, X0)
        E = rs.to_members(Ec) if rs.size > 1 else Ec
        if it == iMax:
            break
        Eo = gather_members(forward(E), N)
        stats["Eo"].append(Eo)
        ctx.use_torch_stream()
        _lib.check(ctx.lib.hm_ies_step(ctx.handle, N, p, ptr(W), ptr(Eo.contiguous()), ptr(y), ptr(pert), ptr(dec),
                                       float(xStep)))
    return E, stats


def iles_sharded(forward, E_local, N, obs, perturbs, decorr, taper_cols, xStep=1.0, iMax=4):
    """Localised iterative ensemble smoother (``HistoryMatch.py:1007-1064``) of a member-sharded ensemble.

    The per-parameter weight matrices ``Ws (m_loc, N, N)``, the anomalies and the mean live on this rank's parameter
    columns (``taper_cols`` = the ``(m_loc, p)`` taper rows of those columns), where the batched Gauss-Newton step
    (``hm_iles_step``) and the recomposition (``hm_iles_recompose``) are fully local; per iteration the predicted data
    are all-gathered and one all-to-all returns member rows for the forward run.
    Returns ``(E_local, stats)``; ``stats["Eo"]`` holds the gathered predicted data ``(N,p)`` of every iteration.
    """
    from . import _lib

    dev = E_local.device
    M = E_local.shape[1]
    if dev.type != "cuda":
        raise _lib.HmError("iles_sharded runs on CUDA tensors (no CPU fallback)")
    y, pert, dec, tap = (_f64_on(x, dev).contiguous() for x in (obs, perturbs, decorr, taper_cols))
    p = y.shape[0]
    ctx = _lib.Context.get(dev.index if dev.index is not None else torch.cuda.current_device())
    ctx.use_torch_stream()
    rs = Resharder.get(N, M, dev)
    E_cols = members_to_columns(E_local.contiguous(), N).contiguous()
    m_loc = E_cols.shape[1]
    if tuple(tap.shape) != (m_loc, p):
        raise ValueError(f"taper_cols must be ({m_loc}, {p}): the taper rows of this rank's parameter columns")
    X0 = torch.empty((N, m_loc), dtype=torch.float64, device=dev)
    x0 = torch.empty(m_loc, dtype=torch.float64, device=dev)
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(ctx.lib.hm_center(ctx.handle, N, m_loc, ptr(E_cols), m_loc, ptr(X0), m_loc, ptr(x0), 0))
    Ws = torch.eye(N, dtype=torch.float64, device=dev).repeat(m_loc, 1, 1).contiguous()
    stats = dict(Eo=[])
    E = E_local
    for it in range(iMax + 1):
        Ec = torch.empty((N, m_loc), dtype=torch.float64, device=dev)
        ctx.use_torch_stream()
        _lib.check(ctx.lib.hm_iles_recompose(ctx.handle, N, m_loc, ptr(Ws), ptr(X0), ptr(x0), ptr(Ec)))
        E = rs.to_members(Ec) if rs.size > 1 else Ec
        if it == iMax:
            break
        Eo = gather_members(forward(E), N).contiguous()
        stats["Eo"].append(Eo)
        ctx.use_torch_stream()
        _lib.check(ctx.lib.hm_iles_step(ctx.handle, N, m_loc, p, ptr(Ws), ptr(Eo), ptr(y), ptr(pert), ptr(dec), ptr(tap),
                                        float(xStep)))
    return E, stats
