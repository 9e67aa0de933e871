"""Ensemble forward run: host side of ``hm_sim_batch`` (include/hm_b200.h).

Replaces the per-member multiprocessing map ``utils.apply(comp1, ...)``
(reference ``tools/utils.py:155-242``, ``HistoryMatch.py:358-387``) by one
batched launch sequence on the GPU.  Inputs may be

* torch CUDA float64 tensors -> device path (``hm_sim_batch``), results stay on
  the device;
* numpy arrays -> host path (``hm_sim_batch_host``), the library stages the
  data through its own device workspace, results are numpy arrays.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib


@dataclass
class GridSpec:
    """Grid and fluid of ``ResSim(Nx=, Ny=, Lx=, Ly=)`` (``HistoryMatch.py:97``)."""

    Nx: int
    Ny: int
    Lx: float = 1.0
    Ly: float = 1.0
    vw: float = 1.0
    vo: float = 1.0
    swc: float = 0.0
    sor: float = 0.0

    @property
    def M(self):
        return self.Nx * self.Ny


@dataclass
class SimResult:
    S_last: object
    obs: object = None
    S_hist: object = None
    P_last: object = None
    status: object = None
    substeps: object = None
    cg_iters: object = None
    stats: dict = field(default_factory=dict)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


_LANES: dict = {}  # (device, lane) -> (Context, torch stream): the extra lanes' library contexts and streams


def uses_cluster_transport(grid: GridSpec, sat_block=0) -> bool:
    """Whether hm_sim_batch takes the streamed path with an on-chip transport kernel for this grid: the temporally blocked
    kernel (csrc/hm_transport.cu: row length a multiple of 64) or the cluster kernel (the tiling rule of csrc/hm_sim.cu:
    row tiles of <= 2048 cells, even height, at most 16 tiles per member)."""
    if sat_block in (1, 5, 6) or (grid.M <= 2048 and sat_block == 0):
        return False
    if sat_block in (0, 7) and grid.Ny % 64 == 0 and grid.Ny <= 2048:
        return True
    R = max(1, min(grid.Nx, 2048 // grid.Ny))
    if 1 < R < grid.Nx:
        R &= ~1
    return -(-grid.Nx // R) <= 16


def run_ensemble(grid: GridSpec, K, well_cell, well_rate, S0, dt, n_steps, *, n_members=None, lanes=1, **kw) -> SimResult:
    """Run ``n_steps`` of the simulator for every ensemble member (see ``_run_ensemble_one`` for the arguments).

    ``lanes = L > 1`` (device path only): the members are split into L contiguous shares that run concurrently from L
    host threads, each with its own library context (workspace) and CUDA stream.  The on-chip transport kernels are
    bound by the FP64 pipe and leave HBM idle, the pressure solve is HBM / latency bound and leaves the FP64 pipe idle:
    with two lanes one share's pressure solve overlaps with the other's transport (measured at 128^2 x 1024 members
    with k_sat_tb: forward run 1105 -> 1019 ms, bench value +8 .. +15 %; results bit-identical,
    tests/test_sim_gpu.py::test_concurrent_lanes_are_bit_identical).  ``lanes = 0``: 2 where an on-chip transport kernel
    runs and the ensemble has at least 256 members, else 1.  ``stats`` are summed over the lanes (phase times therefore
    add up to more than the wall time).
    """
    use_torch = _is_torch(K)
    counts = [x.shape[0] for x, nd in ((K, 2), (S0, 2), (well_cell, 2), (well_rate, 3))
              if hasattr(x, "ndim") and x.ndim >= nd and x.shape[0] > 1]
    N = n_members or (counts[0] if counts else 1)
    if lanes == 0:
        lanes = 2 if (use_torch and N >= 256 and uses_cluster_transport(grid, kw.get("sat_block", 0))) else 1
    if lanes <= 1 or not use_torch or N < 2 * lanes or kw.get("ctx") is not None:
        return _run_ensemble_one(grid, K, well_cell, well_rate, S0, dt, n_steps, n_members=n_members, **kw)

    import threading

    import torch

    dev = K.device
    di = dev.index if dev.index is not None else torch.cuda.current_device()
    main = torch.cuda.current_stream(di)

    def share(x, nd, lo, hi):  # per-member arrays are sliced, shared ones passed through
        if hasattr(x, "ndim") and x.ndim >= nd and x.shape[0] == N and N > 1:
            return x[lo:hi]
        return x

    bounds = [(N * i // lanes, N * (i + 1) // lanes) for i in range(lanes)]
    # outputs of the whole ensemble, allocated once; every lane writes its own member range
    full = _run_ensemble_one(grid, K, well_cell, well_rate, S0, dt, n_steps, n_members=N, _alloc_only=True, **kw)
    results, errors = [None] * lanes, [None] * lanes

    def work(i):
        lo, hi = bounds[i]
        try:
            torch.cuda.set_device(di)  # a new thread starts on device 0
            if (di, i) not in _LANES:
                _LANES[(di, i)] = (_lib.Context(di), torch.cuda.Stream(device=di))
            ctx, stream = _LANES[(di, i)]
            stream.wait_stream(main)
            with torch.cuda.stream(stream):
                out = SimResult(**{f: (getattr(full, f)[lo:hi] if getattr(full, f) is not None and f != "stats" else None)
                                   for f in ("S_last", "obs", "S_hist", "P_last", "status", "substeps", "cg_iters")})
                results[i] = _run_ensemble_one(grid, share(K, 2, lo, hi), share(well_cell, 2, lo, hi),
                                               share(well_rate, 3, lo, hi), share(S0, 2, lo, hi), dt, n_steps,
                                               n_members=hi - lo, _out=out, **{**kw, "ctx": ctx})
        except BaseException as e:  # re-raised in the caller's thread
            errors[i] = e

    threads = [threading.Thread(target=work, args=(i,)) for i in range(lanes)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i in range(lanes):
        main.wait_stream(_LANES[(di, i)][1]) if (di, i) in _LANES else None
    for e in errors:
        if e is not None:
            raise e
    stats = {}
    for r in results:
        for k, v in r.stats.items():
            if k == "phase_ms":
                stats.setdefault(k, {})
                for ph, ms in v.items():
                    stats[k][ph] = stats[k].get(ph, 0.0) + ms
            elif k in ("sat_resident_ctas", "sat_tb_cluster", "sat_tb_strips", "sat_tb_halo"):
                stats[k] = max(stats.get(k, 0), v)
            else:
                stats[k] = stats.get(k, 0) + v
    stats["lanes"] = lanes
    full.stats = stats
    return full


def _run_ensemble_one(grid: GridSpec, K, well_cell, well_rate, S0, dt, n_steps, *, n_members=None, _out=None,
                      _alloc_only=False,
                 obs_cell=None, por=None, history=False, pressure=False, want_substeps=False,
                 cg_rtol=0.0, cg_max_iter=0, chunk_members=0, precond=0, mg_switch_iters=0, sat_block=0, warm_start=0,
                 tb_cluster_rows=0, tb_halo=0, k_transform=None, ctx=None) -> SimResult:
    """Run ``n_steps`` of the simulator for every ensemble member.

    K          (M,) shared isotropic; (N,M) isotropic; (N,2,M) anisotropic (Kx, Ky);
               a leading dimension of 1 means "shared by all members"
    well_cell  (nW,) shared or (N,nW) int32 flat cell indices
    well_rate  signed rates (+inj, -prd): (nW,) constant & shared; (nT,nW) shared
               schedule; (N,1,nW) / (N,nT,nW) per member
    S0         (M,) shared or (N,M)
    k_transform  ``(a, b)``: ``K`` holds the notebooks' log-permeability parameter ``x`` and the permeability is
               ``a + exp(b x)`` (``perm_transf``, ``HistoryMatch.py:137-138``: ``(0.1, 5)``), evaluated inside the kernels
    history    True: ``S_hist (N, n_steps+1, M)`` (row 0 = S0, the reference's ``ResSim.sim`` output);
               an int k > 1: every k-th step and the last one, ``(N, 1 + ceil(n_steps/k), M)``
    """
    M = grid.M
    use_torch = _is_torch(K)
    if use_torch:
        import torch

        dev = K.device
        if dev.type != "cuda":
            raise _lib.HmError("torch inputs must live on a CUDA device (no CPU fallback)")

        def as_f64(x):
            return torch.as_tensor(x, dtype=torch.float64, device=dev).contiguous()

        def as_i32(x):
            return torch.as_tensor(x, dtype=torch.int32, device=dev).contiguous()

        def empty(shape, dtype=torch.float64):
            return torch.empty(shape, dtype=dtype, device=dev)

        def ptr(x):
            return C.c_void_p(x.data_ptr()) if x is not None else None

        i32 = torch.int32
    else:
        def as_f64(x):
            return np.ascontiguousarray(x, dtype=np.float64)

        def as_i32(x):
            return np.ascontiguousarray(x, dtype=np.int32)

        def empty(shape, dtype=np.float64):
            return np.empty(shape, dtype=dtype)

        def ptr(x):
            return C.c_void_p(x.ctypes.data) if x is not None else None

        i32 = np.int32

    K = as_f64(K)
    S0 = as_f64(S0)
    well_cell = as_i32(well_cell)
    well_rate = as_f64(well_rate)

    # ---- shapes -> strides ---------------------------------------------------------
    if K.ndim == 1:          # (M,)
        assert K.shape == (M,)
        K_ms, K_cs, nK = 0, 0, None
    elif K.ndim == 2:        # (N,M)
        assert K.shape[1] == M
        K_ms, K_cs, nK = M, 0, K.shape[0]
    else:                    # (N,2,M)
        assert K.shape[1:] == (2, M)
        K_ms, K_cs, nK = 2 * M, M, K.shape[0]
    if nK == 1 and n_members not in (None, 1):
        K_ms, nK = 0, None
    S0_ms, nS = (0, None) if S0.ndim == 1 else (M, S0.shape[0])
    wc_ms, nWc = (0, None) if well_cell.ndim == 1 else (well_cell.shape[1], well_cell.shape[0])
    nW = well_cell.shape[-1]
    if well_rate.ndim == 1:
        wr_ms, wr_ss, nWr = 0, 0, None
    elif well_rate.ndim == 2:
        assert well_rate.shape[0] in (1, n_steps)
        wr_ms, wr_ss, nWr = 0, (nW if well_rate.shape[0] > 1 else 0), None
    else:
        assert well_rate.shape[1] in (1, n_steps)
        wr_ss = nW if well_rate.shape[1] > 1 else 0
        wr_ms, nWr = well_rate.shape[1] * nW, well_rate.shape[0]
    assert well_rate.shape[-1] == nW
    counts = {c for c in (nK, nS, nWc, nWr, n_members) if c is not None}
    if len(counts) > 1:
        raise ValueError(f"inconsistent ensemble sizes {counts}")
    N = counts.pop() if counts else 1

    n_obs = 0
    if obs_cell is not None:
        obs_cell = as_i32(obs_cell)
        n_obs = obs_cell.shape[0]
    if por is not None:
        por = as_f64(por).reshape(-1)

    hist_stride = int(history) if (history is not True and history) else (1 if history else 0)
    n_hist = n_steps + 1 if hist_stride <= 1 else 1 + -(-n_steps // hist_stride)
    if _out is not None:  # a lane of run_ensemble: views into the outputs of the whole ensemble
        res = _out
    else:
        res = SimResult(S_last=empty((N, M)))
        res.obs = empty((N, n_steps, n_obs)) if n_obs else None
        res.S_hist = empty((N, n_hist, M)) if history else None
        res.P_last = empty((N, M)) if pressure else None
        res.status = empty((N,), i32)
        res.substeps = empty((N, n_steps), i32) if want_substeps else None
        res.cg_iters = empty((N, n_steps), i32) if want_substeps else None
    if _alloc_only:
        return res

    d = _lib.SimDesc()
    d.n_members, d.Nx, d.Ny, d.Lx, d.Ly = N, grid.Nx, grid.Ny, grid.Lx, grid.Ly
    d.vw, d.vo, d.swc, d.sor = grid.vw, grid.vo, grid.swc, grid.sor
    d.K, d.K_member_stride, d.K_comp_stride = ptr(K), K_ms, K_cs
    d.por = ptr(por)
    d.n_wells = nW
    d.well_cell, d.well_cell_member_stride = ptr(well_cell), wc_ms
    d.well_rate, d.well_rate_member_stride, d.well_rate_step_stride = ptr(well_rate), wr_ms, wr_ss
    d.S0, d.S0_member_stride = ptr(S0), S0_ms
    d.dt, d.n_steps = float(dt), int(n_steps)
    d.n_obs, d.obs_cell = n_obs, ptr(obs_cell)
    d.S_last, d.S_hist, d.obs, d.P_last = ptr(res.S_last), ptr(res.S_hist), ptr(res.obs), ptr(res.P_last)
    d.status, d.substeps, d.cg_iters = ptr(res.status), ptr(res.substeps), ptr(res.cg_iters)
    d.cg_rtol, d.cg_max_iter, d.chunk_members = float(cg_rtol), int(cg_max_iter), int(chunk_members)
    d.precond = int(precond)
    d.mg_switch_iters = int(mg_switch_iters)
    d.sat_block = int(sat_block)
    d.hist_stride = hist_stride
    d.warm_start = int(warm_start)
    d.tb_cluster_rows, d.tb_halo = int(tb_cluster_rows), int(tb_halo)
    if k_transform is not None:
        d.K_transform, d.K_a, d.K_b = 1, float(k_transform[0]), float(k_transform[1])

    if use_torch:
        ctx = ctx or _lib.Context.get(dev.index if dev.index is not None else 0)
        ctx.use_torch_stream()
        _lib.check(ctx.lib.hm_sim_batch(ctx.handle, C.byref(d)))
    else:
        ctx = ctx or _lib.Context.get()
        _lib.check(ctx.lib.hm_sim_batch_host(ctx.handle, C.byref(d)))

    st = _lib.SimStats()
    _lib.check(ctx.lib.hm_sim_get_stats(ctx.handle, C.byref(st)))
    ph = (C.c_double * 5)()
    _lib.check(ctx.lib.hm_sim_get_phase_ms(ctx.handle, ph))
    res.stats = dict(
        cg_iterations=st.cg_iterations, sat_substeps=st.sat_substeps,
        kernel_launches=st.kernel_launches, cg_kernel_launches=st.cg_kernel_launches,
        sat_kernel_launches=st.sat_kernel_launches, mg_fp64_fallbacks=st.mg_fp64_fallbacks, cg_restarts=st.cg_restarts, sat_resident_ctas=st.sat_resident_ctas,
        sat_tb_cluster=st.sat_tb_cluster, sat_tb_strips=st.sat_tb_strips, sat_tb_halo=st.sat_tb_halo,
        sat_cell_updates=st.sat_cell_updates,
        phase_ms=dict(zip(("setup", "cg", "flux", "saturation", "obs"), list(ph))),
    )
    return res
