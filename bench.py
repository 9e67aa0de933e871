#!/usr/bin/env python
"""Benchmark of the history-matching hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one ES-MDA assimilation pass over the synthetic ensemble: the
forward run of every member for ``nTime`` = 40 simulator steps (pressure solve +
CFL sub-stepped transport + observation gather) followed by the ES update
(``alpha = 4``, one of the ``Na = 4`` passes) of the whole parameter ensemble.
``value`` = member*steps per second = members * nTime * K / time.

Workload (``config.workload``): BASELINE config "ES-MDA on a 128x128 synthetic
permeability grid, 1024 members, 1 B200"; with ``--gpus N`` every GPU holds 1024
members (weak scaling, members sharded, NCCL all-gather of the predicted data and
all-to-all re-shard of the parameter matrix inside the update).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "A": dict(Nx=20, Ny=20, members=40, nTime=40, name="HistoryMatch.ipynb default: 20x20 grid, 40 members"),
    "A200": dict(Nx=20, Ny=20, members=200, nTime=40, name="HistoryMatch.ipynb grid with a 200-member ensemble (BASELINE config 1)"),
    "C": dict(Nx=128, Ny=128, members=1024, nTime=40,
              name="ES-MDA pass, 128x128 synthetic permeability grid, 1024 members per GPU"),
    "D": dict(Nx=512, Ny=512, members=512, nTime=40,
              name="ES-MDA pass, 512x512 grid, 512 members per GPU (4096 on 8 GPUs)"),
}
ALPHA = 4.0
PEAKS_FALLBACK = dict(hbm_gbs=6650.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C", choices=list(WORKLOADS))
    ap.add_argument("--members", type=int, default=0, help="override members per GPU (development)")
    ap.add_argument("--ntime", type=int, default=0, help="override simulator steps per pass (development)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-update-bench", action="store_true")
    ap.add_argument("--no-config-d", action="store_true", help="skip the short BASELINE-config-4 (512x512) run")
    ap.add_argument("--config-d-ntime", type=int, default=2, help="simulator steps of the short config-D run")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU time of the baseline sample")
    ap.add_argument("--sat-block", type=int, default=0, help="transport kernel variant (hm_sim_desc.sat_block)")
    ap.add_argument("--precond", type=int, default=0, help="pressure preconditioner (hm_sim_desc.precond)")
    ap.add_argument("--lanes", type=int, default=0,
                    help="concurrent member shares (host threads / library contexts / streams) of a forward run.  0 (default) = "
                         "the product's automatic choice (run_ensemble(lanes=0)): 2 where an on-chip transport kernel runs and "
                         "the ensemble has >= 256 members - one share's HBM-bound pressure solve overlaps with the other's "
                         "FP64-bound transport (bit-identical results; measured at config C: value +8 .. +15 %).  With several "
                         "lanes the per-phase CUDA-event times include the other lane's kernels, so phases_ms_per_step and the "
                         "roofline are taken from ONE additional single-lane pass outside the timed region (`attribution`)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return PEAKS_FALLBACK, "fallback"


def workload_config(wl, world):
    """The `config` object of a bench line: identical for the GPU arm and the reference arm of the same workload."""
    M = wl["Nx"] * wl["Ny"]
    return dict(workload=wl["name"], grid=[wl["Nx"], wl["Ny"]], members_per_gpu=wl["members"],
                members=wl["members"] * max(1, world), nTime=wl["nTime"], p=4 * wl["nTime"],
                update="ES (one ES-MDA pass, alpha=4)", parallelism=f"members sharded x{max(1, world)}",
                l2=("inputs larger than L2 (working set %.1f GB per GPU)" % (13 * wl["members"] * M * 8 / 1e9)
                    if 13 * wl["members"] * M * 8 > 256e6 else
                    "working set %.1f MB, smaller than L2: the fused small-grid kernel keeps a member's whole state in shared "
                    "memory for the run and touches HBM / L2 once per pass (inputs in, results out), so the L2 state between "
                    "timed iterations is immaterial" % (13 * wl["members"] * M * 8 / 1e6)))


# ---- clocks ----------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region.  In-process NVML (nvidia_ml_py): spawning nvidia-smi
    every 0.2 s stalls kernel launches for milliseconds (its NVML start-up takes driver locks), which tripled the
    measured time of the 14 ms notebook-size workload.  nvidia-smi is the fallback when NVML cannot be loaded."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.times, self.stop_flag = index, [], [], threading.Event()
        self.period = float(os.environ.get("HM_BENCH_CLOCK_PERIOD", "0.5"))  # seconds between NVML samples
        self.nvml = self.handle = None
        try:
            import pynvml

            pynvml.nvmlInit()
            try:
                import torch

                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = int(get(h))
        masks = (0x8, 0x40, 0x20, 0x4)  # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        return [str(sm), str(mx)] + ["Active" if bits & m else "Not Active" for m in masks]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    smp = self._sample_nvml()
                    self.times.append(time.perf_counter())
                    self.samples.append(smp)
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                    self.times.append(time.perf_counter())
                    self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def window(self, t0, t1):
        """Keep the samples taken inside the timed region [t0, t1]; a region shorter than the sampling period keeps
        the sample nearest to it (the sampler runs from before the warm-up passes, so the GPU is already under load)."""
        n = min(len(self.times), len(self.samples))
        inside = [i for i in range(n) if t0 <= self.times[i] <= t1]
        if not inside and n:
            inside = [min(range(n), key=lambda i: min(abs(self.times[i] - t0), abs(self.times[i] - t1)))]
        self.samples = [self.samples[i] for i in inside]

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v == "Active"})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


# ---- CPU reference path (oracle port of the reference's scipy/numpy path) ---------------------
def _cpu_member(args):
    import threadpoolctl

    threadpoolctl.threadpool_limits(1)  # the reference pins BLAS to 1 thread per worker (tools/utils.py:207-209)
    from oracle import ressim as orr

    Nx, Ny, logk, dt, nT = args
    m = orr.notebook_model(Nx, Ny)
    prd = m.xy2ind(*m.prd_xy.T)
    return orr.forward_member(m, logk, dt, nT, np.zeros(m.Nxy), prd)[1]


def cpu_forward_sample(wl, seconds, seed=0):
    """Time the oracle forward run on the host cores on a bounded sample; member*steps/s."""
    import multiprocessing as mp

    from historymatching_b200.dropin.tools import geostat
    from historymatching_b200.sim import GridSpec

    cores = min(mp.cpu_count(), 64)
    grid = GridSpec(wl["Nx"], wl["Ny"], 2.0, 1.0)
    # cost model from SURVEY.md section 6: ~0.215 s per member-step per core at 128^2 (x (cells/16384)^1.7)
    per = 0.215 * (grid.M / 16384.0) ** 1.7 if grid.M > 1000 else 0.0016
    nT = int(max(1, min(wl["nTime"], seconds / per)))
    members = cores if per * nT * 1.0 < seconds else max(1, int(cores * seconds / (per * nT)))
    members = max(1, min(members, cores * max(1, int(seconds / (per * nT)))))
    logk = geostat.gaussian_fields_separable(grid, members, r=0.8, rng=np.random.RandomState(seed))
    # hard wall-clock guard: the sample is sized from a cost model, and a 512^2 member-step is about a minute per core
    limit = max(60.0, 12.0 * seconds)
    with mp.get_context("fork").Pool(cores) as pool:
        try:
            warm = (grid.Nx, grid.Ny, logk[0], 0.025, 1) if grid.M <= 20000 else (20, 20, np.zeros(400), 0.025, 1)
            pool.map_async(_cpu_member, [warm] * cores).get(timeout=limit)  # warm the workers (imports, first solve)
            t0 = time.perf_counter()
            pool.map_async(_cpu_member, [(grid.Nx, grid.Ny, lk, 0.025, nT) for lk in logk], chunksize=1).get(timeout=limit)
            dt = time.perf_counter() - t0
        except mp.TimeoutError:
            pool.terminate()
            return None, cores, f"CPU sample ({members} members x {nT} steps of the {grid.Nx}x{grid.Ny} forward run) did not finish within {limit:.0f} s"
    return members * nT / dt, cores, f"{members} members x {nT} steps of the {grid.Nx}x{grid.Ny} forward run, oracle (scipy spsolve + explicit upwind), one process per core, BLAS pinned to 1 thread"


def run_reference(args, wl, rank):
    if rank != 0:
        return
    vals = []
    sample = ""
    cores = 0
    per_step = max(5.0, min(args.cpu_seconds, 240.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_forward_sample(wl, per_step, seed=i)
        if i >= args.warmup and v is not None:
            vals.append(v)
    if not vals:
        print(json.dumps(dict(impl="reference", unavailable=sample)))
        return
    value = float(np.mean(vals))
    line = dict(metric="ensemble forward-sim member*steps/s", value=value, unit="member*steps/s", impl="reference",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * wl["members"] * wl["nTime"] / value, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=workload_config(wl, args.gpus),
                note="the forward run (the part that scales with the ensemble) timed on a bounded sample of members x steps "
                     "on the host cores; the ES update of the oracle is timed in the GPU arm's `update.es_cpu_ms`",
                cpu_baseline=dict(value=value, unit="member*steps/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="member*steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ---- the notebook-cell path: forward_model -> tools.utils.apply -> comp1 -> ResSim.sim ------------------------------
def dropin_forward_bench(wl, reps=5):
    """The call a notebook user makes (HistoryMatch.py:358-387, cells as in tests/test_notebook_flow_gpu.py): every member
    deep-copies the model, sets its permeability and calls ``ResSim.sim``; the drop-in ``apply`` runs the members on
    threads whose ``sim`` calls rendezvous into ONE ``hm_sim_batch_host`` (host buffers in, full saturation history
    out).  Returns member*steps/s and the wall time per ensemble run, next to the same run through the array API."""
    import copy

    import historymatching_b200 as hmb

    hmb.activate()
    import TPFA_ResSim as simulator
    from tools import geostat, utils
    from tools.utils import apply

    Nx, Ny, N, nTime, dt = wl["Nx"], wl["Ny"], wl["members"], wl["nTime"], 0.025
    model = simulator.ResSim(Nx=Nx, Ny=Ny, Lx=2, Ly=1)
    near01 = np.array([0.12, 0.87])
    model.prd_xy = [[x, y] for y in model.Ly * near01 for x in model.Lx * near01]
    model.inj_xy = [[model.Lx / 2, model.Ly / 2]]
    model.inj_rates = [[1]]
    model.prd_rates = np.ones((4, 1)) / 4
    prod_inds = model.xy2ind(*model.prd_xy.T)
    wsat0 = np.zeros(model.Nxy)
    np.random.seed(1)
    prior = np.clip(geostat.gaussian_fields(model.mesh, N, r=0.8), -2.2, 2.2)

    def set_perm(model, log_perm_array):
        p = (0.1 + np.exp(5 * log_perm_array)).reshape(model.shape)
        model.K = np.stack([p, p])

    def comp1(perm, wsat0=wsat0):
        new_model = copy.deepcopy(model)
        set_perm(new_model, perm)
        wsats = new_model.sim(dt, nTime, wsat0, pbar=False)
        prods = np.array([x[prod_inds] for x in wsats[1:]])
        return wsats, prods

    def forward_model(*args, **kwargs):
        output = apply(comp1, *args, pbar=False, **kwargs)
        return [np.asarray(y) for y in zip(*output)]

    utils.nCPU = "auto"
    forward_model(prior)  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        wsats, prods = forward_model(prior)
    cells_s = (time.perf_counter() - t0) / reps
    # the same ensemble through the array API with host buffers (no per-member Python objects)
    from historymatching_b200.sim import GridSpec, run_ensemble

    grid = GridSpec(Nx, Ny, 2.0, 1.0)
    cells = np.concatenate([model.xy2ind(*model.inj_xy.T), prod_inds]).astype(np.int32)
    rates = np.array([1.0, -0.25, -0.25, -0.25, -0.25])
    K = 0.1 + np.exp(5 * prior)
    run_ensemble(grid, K, cells, rates, wsat0, dt, nTime, obs_cell=prod_inds.astype(np.int32), history=True)
    t0 = time.perf_counter()
    for _ in range(reps):
        res = run_ensemble(grid, K, cells, rates, wsat0, dt, nTime, obs_cell=prod_inds.astype(np.int32), history=True)
    array_s = (time.perf_counter() - t0) / reps
    assert np.array_equal(res.S_hist, wsats)
    utils.nCPU = 1
    return dict(value=N * nTime / cells_s, unit="member*steps/s", ms_per_ensemble_run=1e3 * cells_s,
                array_api_host_buffers_ms=1e3 * array_s, members=N,
                path="forward_model -> tools.utils.apply(comp1) -> copy.deepcopy(model), set_perm, ResSim.sim -> collector -> "
                     "one hm_sim_batch_host; full (N, nTime+1, M) history returned to the cells",
                d2h_bytes_per_run=int(8 * N * (nTime + 1) * Nx * Ny))


# ---- update wall times + FP64 GEMM roofline (second half of the BASELINE metric) -----------------------
def update_benchmarks(case, E0, Eo, noisy, pert, dec, cpu=True, reps=5):
    """ES / LES / IES-iteration wall times at the bench ensemble size (CUDA events, inputs resident) and the
    FP64 tensor-core GEMM of the IES recomposition E = x0 + W X0 against a cuBLAS DGEMM timed in the same run."""
    import ctypes as C

    import torch

    from historymatching_b200 import _lib
    from historymatching_b200 import analysis as ha

    dev = E0.device
    N, M = E0.shape
    p = Eo.shape[1]
    ctx = _lib.Context.get(dev.index or 0)
    ctx.use_torch_stream()

    def timed(fn, n=reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = dict(N=N, M=M, p=p)
    out["es_ms"] = timed(lambda: ha.ens_update0(E0, Eo, noisy, pert, dec))
    # localised ES: bump taper of radius 1.2 between cell centres and the (well, time) observations (HM:863)
    g = case.grid
    ix, iy = np.divmod(np.arange(M), g.Ny)
    xy_prm = np.stack([(ix + 0.5) * g.Lx / g.Nx, (iy + 0.5) * g.Ly / g.Ny], 1)
    xy_obs = np.tile(xy_prm[case.obs_cell], (case.nTime, 1))
    taper = ha.bump_taper(torch.as_tensor(xy_prm, device=dev), torch.as_tensor(xy_obs, device=dev), 1.2)
    out["les_ms"] = timed(lambda: ha.ens_update0_loc(E0, Eo, noisy, pert, dec, taper), n=2)
    # one IES iteration of algebra: hm_ies_step (N x N solve) + recomposition GEMM (HM:920-944)
    W = torch.eye(N, dtype=torch.float64, device=dev)
    X0 = E0 - E0.mean(0, keepdim=True)
    x0 = E0.mean(0)

    def ies_iter():
        _lib.check(ctx.lib.hm_ies_step(ctx.handle, N, p, C.c_void_p(W.data_ptr()), C.c_void_p(Eo.data_ptr()),
                                       C.c_void_p(noisy.data_ptr()), C.c_void_p(pert.data_ptr()),
                                       C.c_void_p(dec.data_ptr()), 0.4))
        return ha._recompose(ctx, x0, W, X0)

    out["ies_iter_ms"] = timed(ies_iter)
    # FP64 GEMM roofline: the recomposition product W (N x N) @ X0 (N x M), 2 N^2 M flops
    Cm = torch.empty_like(X0)

    def gemm():
        _lib.check(ctx.lib.hm_dgemm(ctx.handle, 0, 0, N, M, N, 1.0, C.c_void_p(W.data_ptr()), N,
                                    C.c_void_p(X0.data_ptr()), M, 0.0, C.c_void_p(Cm.data_ptr()), M))

    ms = timed(gemm)
    A = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    ms_peak = timed(lambda: torch.matmul(A, A), n=3)
    ms_same = timed(lambda: torch.matmul(W, X0))
    peak = 2 * 4096**3 / ms_peak / 1e9
    ach = 2.0 * N * N * M / ms / 1e9
    out["dgemm"] = dict(kernel="k_dgemm (mma.sync m8n8k4 f64, DMMA) W@X0 %dx%dx%d" % (N, M, N), bound="tensor",
                        achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak,
                        peak_source="cuBLAS DGEMM 4096^3 timed in this run",
                        cublas_same_shape=2.0 * N * N * M / ms_same / 1e9)
    if cpu:  # the reference's numpy path (oracle, pinned to the reference's own functions by the golden vectors)
        from oracle import analysis as oa

        En, Eon, yn, pn, dn = (x.cpu().numpy() for x in (E0, Eo, noisy, pert, dec))
        t0 = time.perf_counter()
        oa.ens_update0(En, Eon, yn, pn, dn)
        out["es_cpu_ms"] = 1e3 * (time.perf_counter() - t0)
    return out


def cycle_benchmarks(case, E0, noisy, pert, dec, reps=3):
    """Wall times of the complete iterative cycles of the notebook (BASELINE configs 1 and 2) at the bench ensemble
    size, forward re-runs included: IES and ILES with ``xStep=0.4, iMax=10`` (HM:961, 1075-1077: 10 forward runs + 10
    Gauss-Newton steps each; ILES with the bump taper of radius 1.2) and ES-MDA with ``Na = 4`` (4 forward runs + 4
    updates).  Inputs resident on the device, CUDA events around the whole call."""
    import torch

    from historymatching_b200 import analysis as ha

    dev = E0.device
    N, M = E0.shape
    g = case.grid
    ix, iy = np.divmod(np.arange(M), g.Ny)
    xy_prm = np.stack([(ix + 0.5) * g.Lx / g.Nx, (iy + 0.5) * g.Ly / g.Ny], 1)
    xy_obs = np.tile(xy_prm[case.obs_cell], (case.nTime, 1))
    taper = ha.bump_taper(torch.as_tensor(xy_prm, device=dev), torch.as_tensor(xy_obs, device=dev), 1.2)
    fwd = lambda X: case.forward(X.contiguous())[0]  # noqa: E731
    rng = np.random.RandomState(11)
    Zs = [rng.randn(N, case.p) for _ in range(4)]

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = dict(N=N, M=M, p=case.p, nTime=case.nTime, forward_runs=dict(ies=10, iles=10, es_mda=4))
    out["ies_ms"] = timed(lambda: ha.IES(E0, fwd, noisy, pert, dec, xStep=0.4, iMax=10))
    out["iles_ms"] = timed(lambda: ha.ILES(E0, fwd, noisy, pert, dec, taper, xStep=0.4, iMax=10))
    out["es_mda_ms"] = timed(lambda: ha.es_mda(E0, fwd, noisy, case.R12, [4.0] * 4, perturbs=Zs))
    out["forward_ms"] = timed(lambda: fwd(E0))
    return out


# ---- GPU arm ------------------------------------------------------------------------------------
def main():
    args = parse()
    wl = dict(WORKLOADS[args.workload])
    if args.members:
        wl["members"] = args.members
    if args.ntime:
        wl["nTime"] = args.ntime
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL writes its version / debug lines to stdout by default; stdout carries the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    line = measure(args, wl, rank, world, local_rank, dev, args.steps, args.warmup, full=True)
    if args.workload == "C" and not args.no_config_d and not args.members and not args.ntime:
        # BASELINE config 4 (the north-star size): 512 x 512, 512 members per GPU (4096 on 8 GPUs), a SHORT run - nTime
        # simulator steps instead of 40, stated in the block - with the sharded ES update over NCCL included
        wd = dict(WORKLOADS["D"], nTime=args.config_d_ntime)
        d = measure(args, wd, rank, world, local_rank, dev, steps=1, warmup=1, full=False)
        if rank == 0:
            per_step_s = d["ms_per_step"] / 1e3 / wd["nTime"]
            d["projected"] = dict(
                seconds_per_simulator_step=per_step_s,
                seconds_per_40_step_pass=40 * (per_step_s - d["update_ms"] / 1e3 / wd["nTime"]) + d["update_ms"] / 1e3,
                seconds_per_es_mda_cycle_Na4=4 * (40 * (per_step_s - d["update_ms"] / 1e3 / wd["nTime"]) + d["update_ms"] / 1e3),
                note="forward time is linear in the number of simulator steps (the sub-step count per step is constant, "
                     "SURVEY.md A.5); one complete 40-step pass measured on 8 GPUs: profiles/bench_r2_configD_8gpu_full.json")
            line["config_d"] = d
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure(args, wl, rank, world, local_rank, dev, steps, warmup, full):
    """One workload: warm-up passes, `steps` timed passes, the roofline of its dominant kernel.  full: also the end-to-end
    number through host buffers, the update benchmarks and the CPU baseline (the headline workload)."""
    import torch
    import torch.distributed as dist

    args = argparse.Namespace(**{**vars(args), "steps": steps, "warmup": warmup})
    if not full:
        args.no_e2e = args.no_update_bench = args.no_cpu_baseline = True
    from historymatching_b200 import _lib
    from historymatching_b200 import analysis as ha
    from historymatching_b200 import dist as hd
    from historymatching_b200.dropin.tools import geostat
    from historymatching_b200.workflow import HistoryMatchCase

    case = HistoryMatchCase(wl["Nx"], wl["Ny"], 2.0, 1.0, 0.025, wl["nTime"])
    M, p = case.grid.M, case.p
    N_loc = wl["members"]
    N = N_loc * world
    lo, hi = hd.member_slice(N, rank, world)

    # synthetic inputs: separable Gaussian-variogram prior (r = 0.8), truth drawn first (seed 1)
    truth = geostat.gaussian_fields_separable(case.grid, 1, r=0.8, rng=np.random.RandomState(1), device=dev)
    E0 = geostat.gaussian_fields_separable(case.grid, N_loc, r=0.8, rng=np.random.RandomState(100 + rank), device=dev)
    obs_truth, _ = case.forward(truth)
    g = torch.Generator(device=dev).manual_seed(7)
    R12T = torch.as_tensor(case.R12.T.copy(), device=dev)
    noisy = (obs_truth[0] + torch.randn(p, dtype=torch.float64, device=dev, generator=g) @ R12T).clamp(0, 1)
    Z = torch.randn(N, p, dtype=torch.float64, device=dev, generator=g)  # same stream on every rank
    pert = np.sqrt(ALPHA) * (Z @ R12T)
    dec = torch.as_tensor(case.decorr / np.sqrt(ALPHA), device=dev)
    ctx = _lib.Context.get(local_rank)

    last = {}
    t_start = time.perf_counter()

    def log(msg):
        if os.environ.get("HM_BENCH_LOG") and rank == 0:
            print(f"[bench +{time.perf_counter() - t_start:7.1f}s] {msg}", file=sys.stderr, flush=True)

    def one_pass(E):
        f0 = torch.cuda.Event(enable_timing=True)
        f0.record()
        Eo, res = case.forward(E, want_substeps=True, sat_block=args.sat_block, precond=args.precond, lanes=args.lanes)
        last["res"] = res
        last["fwd0"] = f0
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()  # the ranks' forward runs end at slightly different times: keep that wait out of update_ms
        t0.record()
        h0 = time.perf_counter()
        post = hd.es_update_sharded(E, Eo, N, noisy, pert, dec)
        t1.record()
        if os.environ.get("HM_BENCH_LOG"):
            h1 = time.perf_counter()
            torch.cuda.synchronize()
            log("update: host call %.2f ms, with sync %.2f ms, events %.2f ms" % (1e3 * (h1 - h0), 1e3 * (time.perf_counter() - h0), t0.elapsed_time(t1)))
            ms_ = torch.cuda.memory_stats()
            log("allocator: cudaMalloc %d, cudaFree %d, retries %d, reserved %.0f MB" % (
                ms_.get("num_device_alloc", 0), ms_.get("num_device_free", 0), ms_.get("num_alloc_retries", 0),
                ms_.get("reserved_bytes.all.current", 0) / 1e6))
        last["upd"] = (t0, t1)
        return post, Eo

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log("inputs ready")
    # the clock sampler starts before the warm-up: its first NVML queries take driver locks for milliseconds, which
    # must not land inside a timed region that may itself be only milliseconds long (workload A)
    sampler = ClockSampler(local_rank)
    sampler.start()
    post = Eo = res = None
    for _ in range(args.warmup):
        # the same references as in the timed loop: the previous pass's posterior AND its SimResult (`res`) stay alive
        # while the next pass runs, so torch's caching allocator reaches its steady state here - otherwise the second
        # timed pass is the first one whose update cannot reuse the freed result block and pays two cudaMalloc
        # (0.7 - 75 ms inside update_ms; found with HM_BENCH_LOG=1)
        post, Eo = one_pass(E0)
        torch.cuda.synchronize()
        res = last["res"]
        # the bookkeeping reductions of the timed loop too: their first call loads a torch kernel module (tens of ms)
        int(res.cg_iters.sum()), int(res.substeps.sum())
        log("warm-up pass done")
    barrier()
    t_region0 = time.perf_counter()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    phase = dict(setup=0.0, cg=0.0, flux=0.0, saturation=0.0, obs=0.0)
    upd_ms, fwd_ms, cg_member_iters, sat_member_substeps = 0.0, 0.0, 0, 0
    upd_steps = []
    stats_acc = dict(cg_kernel_launches=0, sat_kernel_launches=0, mg_fp64_fallbacks=0, kernel_launches=0, sat_cell_updates=0)
    torch.cuda.nvtx.range_push("timed")  # lets ncu select the timed region (--nvtx --nvtx-include "timed/")
    for _ in range(args.steps):
        post, Eo = one_pass(E0)
        torch.cuda.synchronize()
        res = last["res"]
        for k in phase:
            phase[k] += res.stats["phase_ms"][k]
        upd_steps.append(last["upd"][0].elapsed_time(last["upd"][1]))
        upd_ms += upd_steps[-1]
        fwd_ms += last["fwd0"].elapsed_time(last["upd"][0])
        cg_member_iters += int(res.cg_iters.sum())
        sat_member_substeps += int(res.substeps.sum())
        for k in stats_acc:
            stats_acc[k] += res.stats[k]
    ev1.record()
    torch.cuda.nvtx.range_pop()
    log("timed passes done")
    barrier()
    t_region1 = time.perf_counter()
    sampler.stop_flag.set()
    sampler.join()
    sampler.window(t_region0, t_region1)
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    launches = ctx.launch_count() - l0
    if int(last["res"].stats.get("lanes", 1)) > 1:  # the lanes launch from their own library contexts
        launches += stats_acc["kernel_launches"]
    value = N * wl["nTime"] * args.steps / (ms / 1e3)
    bad = int((last["res"].status != 0).sum())
    lanes_used = int(last["res"].stats.get("lanes", 1))
    attr_steps = args.steps
    attribution = "timed passes"
    if lanes_used > 1:
        # phases / roofline of the kernels themselves: one single-lane pass, outside the timed region
        _, res = case.forward(E0, want_substeps=True, sat_block=args.sat_block, precond=args.precond, lanes=1)
        torch.cuda.synchronize()
        phase = dict(res.stats["phase_ms"])
        cg_member_iters, sat_member_substeps = int(res.cg_iters.sum()), int(res.substeps.sum())
        stats_acc = {k: res.stats[k] for k in stats_acc}
        last["res"] = res
        attr_steps = 1
        attribution = "one additional single-lane pass outside the timed region (the timed passes run %d lanes)" % lanes_used

    # ---- end to end through the host-buffer API -------------------------------------------
    e2e = None
    if not args.no_e2e:
        x_host = E0.cpu().pin_memory()
        pert_host, noisy_host = pert.cpu().pin_memory(), noisy.cpu().pin_memory()
        out_host = torch.empty_like(x_host).pin_memory()
        eo_host = torch.empty((N_loc, p), dtype=torch.float64).pin_memory()

        def e2e_pass():
            E = x_host.to(dev, non_blocking=True)
            pr = pert_host.to(dev, non_blocking=True)
            ob = noisy_host.to(dev, non_blocking=True)
            Eo, _ = case.forward(E, sat_block=args.sat_block, precond=args.precond, lanes=args.lanes)
            post = hd.es_update_sharded(E, Eo, N, ob, pr, dec)
            out_host.copy_(post, non_blocking=True)
            eo_host.copy_(Eo, non_blocking=True)

        e2e_pass()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            e2e_pass()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = dict(value=N * wl["nTime"] * n_e2e / float(dt), unit="member*steps/s",
                   h2d_bytes_per_step=int(8 * (x_host.numel() + pert_host.numel() + noisy_host.numel())),
                   d2h_bytes_per_step=int(8 * (out_host.numel() + eo_host.numel())))

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel ------------------------------------------------------
    # Algorithmic HBM bytes: a transport sub-step streams 32 B/cell
    # (S in, S out, two face fluxes); one multigrid-PCG iteration on level 0 = k_mg_down 50 +
    # k_mg_up 60 + k_cg_spmv 48 + k_cg_update 48 B/cell per ACTIVE member-iteration.
    pk, pk_kind = peaks()
    hbm = float(pk.get("hbm_gbs", PEAKS_FALLBACK["hbm_gbs"]))
    pcg_name = "MG-PCG iteration (k_mg_down+k_mg_onchip+k_mg_up+k_cg_spmv+k_cg_update)"
    st_last = last["res"].stats
    tb = int(st_last.get("sat_tb_cluster", 0)) > 0            # temporally blocked kernel k_sat_tb (hm_transport.cu)
    cluster = not tb and stats_acc["sat_kernel_launches"] <= attr_steps * wl["nTime"]  # k_sat_cluster: one launch per time step
    stream_name = ("k_sat_stream (one sub-step per launch, tile staged by bulk copies)"
                   if wl["Ny"] % 2 == 0 and args.sat_block != 5 else "k_sat_substep (one sub-step per launch, plain loads)")
    sat_name = "k_sat_cluster (all CFL sub-steps of a time step, register/DSMEM resident)" if cluster else stream_name
    if tb:
        sat_name = ("k_sat_tb (temporally blocked: %d-CTA clusters, 4096 cells per CTA in registers, %s)"
                    % (st_last["sat_tb_cluster"], "all CFL sub-steps of a time step in one launch" if st_last["sat_tb_strips"] == 1
                       else "%d overlapping row strips per member, %d sub-steps per HBM round trip"
                       % (st_last["sat_tb_strips"], st_last["sat_tb_halo"])))
    # FP64 cycle: k_mg_down 50 + k_mg_up 60 + k_cg_spmv 48 + k_cg_update 48 = 206 B; FP32 cycle (the default): the cycle's
    # operators and iterates are 4-byte, k_mg_down 25 (r 8, 1/diag 4, TX TY 8, x 4, coarse rhs 1) + k_mg_up 33 = 154 B
    pcg_bytes_per_cell = 154.0 if (args.precond in (0, 3) and stats_acc["mg_fp64_fallbacks"] == 0) else 206.0
    cg_bytes = pcg_bytes_per_cell * M * cg_member_iters
    sat_bytes = 32.0 * M * sat_member_substeps
    cands = {
        pcg_name: (cg_bytes, phase["cg"], stats_acc["cg_kernel_launches"] / 6),
        sat_name: (sat_bytes, phase["saturation"], stats_acc["sat_kernel_launches"]),
    }
    # The dominant KERNEL: the transport phase is one kernel, the pressure phase five per iteration (the
    # largest of them ~30 % of the phase, profiles/launches_*): transport dominates unless it is < 0.35 x CG.
    fused = stats_acc["sat_kernel_launches"] == 0  # small grids: the whole simulator is one kernel (hm_small.cu)
    dom = sat_name if (phase["saturation"] > 0.35 * phase["cg"] and not fused) else pcg_name
    if fused:
        dom = pcg_name = "k_sim_small (whole forward run of a member in one CTA, shared-memory resident; latency bound)"
        cands[dom] = (cg_bytes + sat_bytes, phase["cg"], attr_steps)
    b, t_ms, n_launch = cands[dom]
    achieved = b / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
    roofline = dict(bound="hbm", kernel=dom, achieved=achieved, peak=hbm, unit="GB/s", frac=achieved / hbm,
                    traffic=None, peak_source=pk_kind + (" burst" if pk_kind == "measured" else ""),
                    algorithmic_bytes_per_launch=b / max(1, n_launch), avg_launch_ms=t_ms / max(1, n_launch),
                    share_of_step=(t_ms / sum(phase.values())) if lanes_used > 1 else t_ms / ms)
    if dom == sat_name and (cluster or tb):
        # On-chip transport kernels: the streaming model above counts 32 B per cell and sub-step, the kernels touch HBM once
        # per time step (k_sat_cluster, k_sat_tb with one strip: 40 B/cell: S in, 3 flux reads incl. pads, S out) or once per
        # round of k sub-steps (k_sat_tb on row strips): frac > 1 is on-chip reuse.  What binds them is on-chip; peaks
        # MEASURED on this GPU (profiles/tools/probe_fp64.cu, profiles/probe_r1.txt): FP64 pipe 61.5 lanes/clk/SM (DFMA: 2.08
        # cycles per warp instruction and SM sub-partition), shared-memory crossbar 128 B/clk/SM.
        nts_mean = sat_member_substeps / max(1, N_loc * wl["nTime"] * attr_steps)
        sm_hz = (sampler.summary()["sm_mhz"] or 1965.0) * 1e6
        sm_n = 148
        resident = int(st_last.get("sat_resident_ctas", 0))
        if tb:
            # per cell and sub-step: 11 FP64 instructions (7 fw + 4 face FMAs), 20 B through shared memory (1 store + 1.5
            # loads of fw); redundant halo-row updates of overlapping strips are executed work, not useful work
            fp64_per_cell, smem_per_cell = 11.0, 20.0
            executed = stats_acc["sat_cell_updates"]
            useful = float(M) * sat_member_substeps
            strips, halo = int(st_last["sat_tb_strips"]), int(st_last["sat_tb_halo"])
            rounds = 1.0 if strips == 1 else nts_mean / halo
            rows_loaded = executed / max(useful, 1.0)                    # strip rows loaded per grid row
            hbm_actual = 8.0 * M * N_loc * rounds * (3.0 * rows_loaded + 1.0)  # S + 2 fluxes in (with overlap), S out
            roofline["redundancy"] = executed / max(useful, 1.0)
            roofline["cell_updates_per_s"] = dict(useful=useful / (t_ms * 1e-3), executed=executed / (t_ms * 1e-3))
        else:
            fp64_per_cell, smem_per_cell = 13.0, 32.0
            executed = useful = float(M) * sat_member_substeps
            hbm_actual = 40.0 * M * N_loc
        roofline["on_chip_reuse_factor"] = 32.0 * M * N_loc * nts_mean / hbm_actual
        roofline["hbm_bytes_per_time_step_actual"] = hbm_actual
        roofline["hbm_actual_GBps"] = hbm_actual * wl["nTime"] * attr_steps / (t_ms * 1e-3) / 1e9
        roofline["fp64_pipe_frac"] = fp64_per_cell * executed / (t_ms * 1e-3) / (sm_n * 61.5 * sm_hz)
        smem_peak = sm_n * 128.0 * sm_hz / 1e9
        smem_ach = smem_per_cell * executed / (t_ms * 1e-3) / 1e9
        roofline["smem_crossbar"] = dict(achieved=smem_ach, peak=smem_peak, unit="GB/s", frac=smem_ach / smem_peak)
        roofline["resident_ctas"] = resident
        roofline["sm_coverage"] = min(1.0, resident / sm_n) if resident else None
        roofline["note"] = ("streaming model of SURVEY 8(d) (32 B per cell and sub-step); the kernel keeps S and the face "
                            "coefficients in registers for many sub-steps, so frac > 1 is the on-chip reuse factor at work.  "
                            "What binds it is on-chip: the FP64 pipe (fp64_pipe_frac, of all 148 SMs), the shared-memory "
                            "crossbar (smem_crossbar.frac) and the SMs its clusters can occupy (sm_coverage)")
        prof = os.path.join(ROOT, "profiles", "ncu_k_sat_tb.json" if tb else "ncu_k_sat_cluster.json")
        if os.path.exists(prof) and (wl["Nx"], wl["Ny"], N_loc) == (128, 128, 1024):
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch at exactly this configuration
            # (ncu --set full capture, profiles/ncu_k_sat_*.txt)
            roofline["traffic"] = json.load(open(prof))[0]["traffic_MB"] * 1e6
    other = sat_name if dom != sat_name else pcg_name
    ob, ot, _ = cands[other]

    line = dict(
        metric="ensemble forward-sim member*steps/s", value=value, unit="member*steps/s", n_gpus=world,
        steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f64", data="synthetic",
        config=workload_config(wl, world), lanes_per_gpu=lanes_used,
        update_ms=upd_ms / args.steps, update_ms_steps=[round(u, 3) for u in upd_steps], forward_ms=fwd_ms / args.steps,
        phases_ms_per_step={k: v / attr_steps for k, v in phase.items()}, attribution=attribution,
        secondary_kernel=dict(kernel=other, achieved=(ob / (ot * 1e-3) / 1e9 if ot > 0 else 0.0), unit="GB/s"),
        members_failed=bad, gpu_launches=int(launches), clocks=sampler.summary(), roofline=roofline,
        pressure=dict(precond=args.precond, iterations_per_solve=cg_member_iters / max(1, N_loc * wl["nTime"] * attr_steps),
                      mg_fp64_fallbacks=stats_acc["mg_fp64_fallbacks"]),
    )
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_update_bench:
        line["update"] = update_benchmarks(case, E0, Eo, noisy, pert, dec, cpu=not args.no_cpu_baseline)
    if world == 1 and full and wl["Nx"] * wl["Ny"] <= 4096 and not args.no_update_bench:
        # the iterative cycles of the notebook incl. their forward re-runs (BASELINE configs 1, 2), undamped perturbations
        line["cycles"] = cycle_benchmarks(case, E0, noisy, pert / np.sqrt(ALPHA), dec * np.sqrt(ALPHA))
    if world == 1 and full and wl["Nx"] * wl["Ny"] <= 4096 and not args.no_e2e:
        line["e2e_dropin"] = dropin_forward_bench(wl)  # the notebook-cell path (BASELINE configs 1, 2, 5)
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample = cpu_forward_sample(wl, args.cpu_seconds)
        line["cpu_baseline"] = dict(value=v, unit="member*steps/s", cores=cores, kind="port", sample=sample)
    return line


if __name__ == "__main__":
    main()
