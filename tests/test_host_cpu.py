"""CPU tests (-m "not gpu"): host-side logic of the drop-ins, the C-ABI surface,
and the golden vectors through the drop-in numpy helpers.  No GPU compute."""

import copy
import ctypes
import os
import re
import sys

import numpy as np
import pytest

import historymatching_b200 as hmb

hmb.activate()
import tools.localization as loc  # noqa: E402
import TPFA_ResSim as simulator  # noqa: E402
from tools import geostat, utils  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from historymatching_b200 import _lib

    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "hm_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(hm_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.hm_version() >= 100
    # struct mirrors match the header's field order
    fields = re.search(r"typedef struct hm_sim_desc \{(.*?)\} hm_sim_desc;", header, flags=re.S).group(1)
    fields = re.sub(r"/\*.*?\*/", "", fields, flags=re.S)
    names = []
    for decl in fields.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names += [n.strip().lstrip("*") for n in decl.split()[-1:]] if "," not in decl else [
            n.strip().lstrip("*") for n in re.sub(r"^\s*(const\s+)?\w+\s*\**", "", decl).split(",")]
    assert names == [f[0] for f in _lib.SimDesc._fields_]
    stats = re.search(r"typedef struct hm_sim_stats \{(.*?)\} hm_sim_stats;", header, flags=re.S).group(1)
    stats = re.sub(r"/\*.*?\*/", "", stats, flags=re.S)
    assert re.findall(r"int64_t\s+(\w+)\s*;", stats) == [f[0] for f in _lib.SimStats._fields_]


def test_history_rows_of_strided_history():
    """Row bookkeeping of hm_sim_desc.hist_stride (mirrors hist_rows / hist_row of csrc/hm_sim_common.cuh)."""
    for n_steps in (1, 4, 6, 7, 40):
        for k in (2, 3, 5, 50):
            rows = sorted(set(range(0, n_steps + 1, k)) | {n_steps})
            assert len(rows) == 1 + -(-n_steps // k)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from historymatching_b200 import _lib
    from historymatching_b200 import analysis as ha

    with pytest.raises(_lib.HmError):
        _lib.Context(0)
    with pytest.raises(_lib.HmError):
        ha.ens_update0(np.zeros((3, 4)), np.zeros((3, 2)), np.zeros(2), np.zeros((3, 2)), np.eye(2))
    m = simulator.ResSim(Nx=4, Ny=4)
    m.inj_xy, m.prd_xy, m.inj_rates, m.prd_rates = [[0.1, 0.1]], [[0.9, 0.9]], [[1]], [[1]]
    with pytest.raises(_lib.HmError):
        m.sim(0.1, 1, np.zeros(16))


def test_utils_primitives_golden(golden):
    g = golden("primitives.npz")
    X, x = utils.center(g["E"])
    np.testing.assert_array_equal(X, g["center_X"])
    np.testing.assert_array_equal(x, g["center_x"])
    np.testing.assert_array_equal(utils.center(g["E"], rescale=True)[0], g["center_X_rescaled"])
    np.testing.assert_array_equal(utils.cov(g["E"], g["b"]), g["cov"])
    np.testing.assert_array_equal(utils.corr(g["E"], g["b"][:, 0]), g["corr"])
    np.testing.assert_allclose(utils.rinv(g["A"], 0.1), g["rinv_tikh"], rtol=1e-12)
    np.testing.assert_allclose(utils.rinv(g["A"], 0.1, tikh=False), g["rinv_trunc"], rtol=1e-12)
    np.testing.assert_array_equal(loc.pairwise_distances(g["pts_a"], g["pts_b"]), g["pd_ab"])
    np.testing.assert_array_equal(loc.pairwise_distances(g["pts_a"], g["pts_b"], domain=(2, 1)), g["pd_periodic"])
    np.testing.assert_array_equal(loc.bump(g["dist"]), g["bump1"])
    np.testing.assert_array_equal(loc.bump(g["dist"], 10), g["bump_sharp"])
    np.testing.assert_array_equal(geostat.variogram_gauss(np.array([0.0, 1.0, 2.0]), 1, n=0.1, a=1), g["variogram"])


def test_localization_doctests():
    A = np.arange(4)[:, None]
    np.testing.assert_array_equal(loc.pairwise_distances(A, [[2]]).T, [[2.0, 1.0, 0.0, 1.0]])
    np.testing.assert_array_equal(loc.pairwise_distances(A, domain=(4,))[1], [1.0, 0.0, 1.0, 2.0])
    np.testing.assert_array_equal(loc.pairwise_distances(np.arange(4)), [[0.0]])
    batches = loc.rectangular_partitioning([4, 13], [2, 4])
    assert sorted(np.concatenate(batches)) == list(range(52))


def test_prior_bit_exact_and_stream_order(golden):
    """seed(1) -> truth -> prior, HistoryMatch.py:78,167,290: identical draws, no extra numbers."""
    g = golden("prior_20x20_seed1.npz")
    model = simulator.ResSim(Nx=20, Ny=20, Lx=2, Ly=1)
    np.random.seed(1)
    truth = geostat.gaussian_fields(model.mesh, 1, r=0.8)
    prior = geostat.gaussian_fields(model.mesh, 40, r=0.8)
    nxt = np.random.randn()
    np.testing.assert_array_equal(truth, g["truth"])
    np.testing.assert_array_equal(prior, g["prior"])
    np.random.seed(1)
    np.random.randn(41, 400)
    assert nxt == np.random.randn()


def test_separable_prior_statistics():
    model = simulator.ResSim(Nx=24, Ny=16, Lx=2, Ly=1)
    F = geostat.gaussian_fields_separable(model, N=3000, r=0.8, rng=np.random.RandomState(0))
    assert F.shape == (3000, 24 * 16)
    C = np.cov(F.T)
    X = geostat.vectorize(*model.mesh)
    want = 1 - geostat.variogram_gauss(geostat.dist_euclid(X), 0.8)
    assert np.abs(C - want).max() < 0.12
    assert abs(F.var() - 1) < 0.05


def test_ressim_surface_and_validation():
    model = simulator.ResSim(Nx=20, Ny=20, Lx=2, Ly=1, name="Base")
    assert model.shape == (20, 20) and model.Nxy == 400 and model.domain[1] == (2, 1)
    assert model.mesh[0].shape == (20, 20) and model.name == "Base"
    # K setter forms used by the notebooks: (2,Nx,Ny), (1,Nxy), (Nxy,)
    p = np.random.RandomState(0).rand(20, 20) + 0.5
    for val in (np.stack([p, p]), p.reshape(1, -1), p.ravel()):
        model.K = val
        assert model.K.shape == (2, 20, 20)
        np.testing.assert_array_equal(model.K[0], p)
    with pytest.raises(ValueError):
        model.K = -p
    # wells: list of [x, y], flat arrays; collocated with cell centres
    near01 = np.array([0.12, 0.87])
    model.prd_xy = [[x, y] for y in model.Ly * near01 for x in model.Lx * near01]
    model.inj_xy = np.array([1.0, 0.5])
    assert model.prd_xy.shape == (4, 2) and model.inj_xy.shape == (1, 2) and model.nPrd == 4 and model.nInj == 1
    np.testing.assert_allclose(model.inj_xy, [[1.05, 0.525]])
    x, y = model.prd_xy.T
    np.testing.assert_array_equal(model.xy2ind(x, y), model.xy2ind(*model.ind2xy(model.xy2ind(x, y))))
    assert model.ind2xy(np.arange(400)).shape == (2, 400) and model.ind2xy(7).shape == (2,)
    assert model.sub2ind(3, 4) == 64 and tuple(model.sub2xy(0, 0)) == (0.05, 0.025)
    with pytest.raises(ValueError):
        model.inj_xy = [[2.5, 0.5]]
    with pytest.raises(ValueError):
        model.inj_xy = [[np.nan, 0.5]]
    model.inj_rates = [[1]]
    model.prd_rates = np.ones((4, 1)) / 4
    rates, cells = model._schedule(5)
    assert rates.shape == (5, 5) and cells.dtype == np.int32 and np.allclose(rates.sum(1), 0)
    assert model.actual_rates["inj"].shape == (1, 5) and model.actual_rates["prd"].shape == (4, 5)
    model.prd_rates = np.ones(4)  # 1-D form (Optimise.py:643-645)
    assert model.prd_rates.shape == (4, 1)
    with pytest.raises(ValueError):  # unbalanced: raises at run (HistoryMatch.py:182-184)
        model._schedule(5)
    m2 = copy.deepcopy(model)
    m2.K = 2 * model.K
    assert not np.array_equal(m2.K, model.K)
    import dill

    assert dill.loads(dill.dumps(model)).Nxy == 400


def test_apply_contract_serial_and_threaded():
    calls = []

    def f(a, b, c=0):
        calls.append(a)
        return a + b.sum() + c

    A, B, Cc = np.arange(5), np.ones((5, 3)), np.arange(5) * 10
    for n in (1, False, "auto", 4, True, None):
        utils.nCPU = n
        out = utils.apply(f, A, B, c=Cc, pbar=False)
        assert out == [a + 3 + 10 * a for a in range(5)]
    with pytest.raises(ValueError):
        utils.apply(f, A, B[:4], pbar=False)
    utils.nCPU = "auto"

    def boom(a):
        if a == 2:
            raise KeyError("member 2")
        return a

    with pytest.raises(KeyError):
        utils.apply(boom, A, pbar=False)
    # pbar forms: str, dict, existing tqdm (re-used, not closed)
    bar = utils.progbar(total=3, disable=True)
    for pb in ("desc", dict(desc="x", leave=False, disable=True), bar):
        assert utils.apply(lambda a: a, A, pbar=pb) == list(A)
    utils.nCPU = 1


def test_collector_batches_sim_calls(monkeypatch):
    """apply() gathers the members' ResSim.sim calls into ONE batched request list."""
    batches = []

    def fake_run(reqs):
        batches.append(len(reqs))
        for r in reqs:
            if r.S0[0] < 0:
                r.error = RuntimeError("bad member")
            else:
                r.result = np.tile(r.S0, (r.nSteps + 1, 1)) + r.K[0].mean()
            r.done = True

    monkeypatch.setattr(simulator, "run_requests", fake_run)
    model = simulator.ResSim(Nx=4, Ny=4, Lx=1, Ly=1)
    model.inj_xy, model.prd_xy, model.inj_rates, model.prd_rates = [[0.1, 0.1]], [[0.9, 0.9]], [[1]], [[1]]

    def comp(k, s0):
        m = copy.deepcopy(model)
        m.K = np.full(16, k)
        try:
            w = m.sim(0.1, 2, s0, pbar=False)
        except RuntimeError:
            return -1.0
        return w[-1, 0]

    ks = np.arange(1.0, 9.0)
    s0 = np.zeros((8, 16))
    s0[3] = -1
    utils.nCPU = "auto"
    out = utils.apply(comp, ks, s0, pbar=False)
    assert batches == [8]
    assert out == [1.0, 2.0, 3.0, -1.0, 5.0, 6.0, 7.0, 8.0]
    # members that never reach sim() (invalid parameters raise first, Optimise.py:548-555)
    def comp2(k):
        m = copy.deepcopy(model)
        try:
            if k > 6:
                m.inj_xy = [[5.0, 5.0]]
            return m.sim(0.1, 1, np.zeros(16), pbar=False)[-1, 0] + k
        except ValueError:
            return 0.0

    batches.clear()
    out = utils.apply(comp2, ks, pbar=False)
    assert batches == [6] and out[-2:] == [0.0, 0.0] and out[0] == 2.0
    utils.nCPU = 1
    batches.clear()
    utils.apply(comp2, ks[:3], pbar=False)
    assert batches == [1, 1, 1]


def test_enopt_on_quadratic():
    from tools import enopt

    utils.nCPU = 1
    np.random.seed(3)
    obj = lambda u: -np.sum((u - np.array([1.0, -2.0])) ** 2)  # noqa: E731
    path, objs, info = enopt.GD(obj, np.zeros(2), enopt.nabla_ens(0.1, nEns=12), quiet=True)
    assert np.linalg.norm(path[-1] - [1, -2]) < 0.05 and objs[-1] > objs[0]


def test_enopt_matches_reference_golden(golden):
    """The drop-in tools/enopt.py (nabla_ens, backtracker, GD) and utils.gaussian_noise against the trajectory the
    reference's own tools/enopt.py produced on the same seed (tests/golden/enopt.npz, make_golden.py)."""
    from tools import enopt

    g = golden("enopt.npz")
    utils.nCPU = 1
    target = np.array([1.0, -2.0])
    obj = lambda u: -np.sum((u - target) ** 2)  # noqa: E731
    for tag, precond in (("lls", False), ("precond", True)):
        np.random.seed(3)
        path, objs, info = enopt.GD(obj, np.zeros(2), enopt.nabla_ens(0.1, nEns=12, precond=precond), nIter=6, quiet=True)
        np.testing.assert_array_equal(np.asarray(path, float), g[f"{tag}_path"])
        np.testing.assert_array_equal(np.asarray(objs, float), g[f"{tag}_objs"])
        np.testing.assert_array_equal(np.array([i["grad"] for i in info if "grad" in i]), g[f"{tag}_grads"])
    np.random.seed(5)
    np.testing.assert_array_equal(utils.gaussian_noise(4, 3, 0.5), g["noise_scalar"])
    L = np.linalg.cholesky(np.array([[2.0, 0.3, 0], [0.3, 1, 0.1], [0, 0.1, 0.5]]))
    np.testing.assert_array_equal(utils.gaussian_noise(4, 3, L), g["noise_chol"])


def test_workflow_case_matches_notebook_setup(golden):
    """HistoryMatchCase (the product's packaged notebook set-up): observation-error model equal to the notebook's cell
    statements (tests/golden/obs_error.npz), wells collocated like ResSim.xy2ind, obs index = well + nPrd * t."""
    from historymatching_b200.workflow import HistoryMatchCase

    g = golden("obs_error.npz")
    case = HistoryMatchCase(20, 20, 2.0, 1.0, 0.025, int(g["nTime"]))
    np.testing.assert_array_equal(case.R, g["R"])
    np.testing.assert_array_equal(case.R12, g["R12"])
    np.testing.assert_allclose(case.decorr @ case.R12.T, np.eye(case.p), atol=1e-12)
    model = simulator.ResSim(Nx=20, Ny=20, Lx=2, Ly=1)
    near01 = np.array([0.12, 0.87])
    prd_xy = [[x, y] for y in 1.0 * near01 for x in 2.0 * near01]            # HistoryMatch.py:179-181
    np.testing.assert_array_equal(case.obs_cell, model.xy2ind(*np.array(prd_xy).T))
    np.testing.assert_array_equal(case.well_cell[:1], model.xy2ind(1.0, 0.5))
    assert case.p == 4 * int(g["nTime"]) and case.well_rate.sum() == 0


def test_header_is_plain_c_and_matches_the_ctypes_mirrors(tmp_path):
    """include/hm_b200.h compiles as strict C99, a C program links against libhm_b200.so, and the struct sizes the C
    compiler sees equal the ctypes mirrors (the boundary a non-Python host would bind)."""
    import shutil
    import subprocess

    from historymatching_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    _lib.load()
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "hm_b200.h"\n'
        "int main(void) {\n"
        "    hm_sim_desc d; hm_sim_stats st; hm_ctx* ctx = NULL;\n"
        "    memset(&d, 0, sizeof d); memset(&st, 0, sizeof st);\n"
        "    int rc = hm_ctx_create(0, &ctx);\n"
        '    printf("%d %d %zu %zu\\n", hm_version(), rc, sizeof d, sizeof st);\n'
        "    if (rc == 0) hm_ctx_destroy(ctx);\n"
        "    return 0;\n}\n")
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-lhm_b200", f"-Wl,-rpath,{libdir}"], check=True)
    version, rc, size_desc, size_stats = map(int, subprocess.run([str(exe)], capture_output=True, text=True,
                                                                  check=True).stdout.split())
    assert version >= 100
    assert rc in (0, -4)  # HM_OK on a GPU box, HM_ERR_NO_DEVICE here: never a silent CPU path
    assert size_desc == ctypes.sizeof(_lib.SimDesc) and size_stats == ctypes.sizeof(_lib.SimStats)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints ONE JSON line with the contract's
    keys, the GPU arm's metric / unit / workload, and loads no CUDA code; the GPU arm's helpers agree on the workload text."""
    import json
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "A", "--steps", "1",
                          "--warmup", "0", "--cpu-seconds", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "member*steps/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    sys.path.insert(0, root)
    import bench

    wl = dict(bench.WORKLOADS["A"])
    assert bench.workload_config(wl, 1)["workload"] == d["config"]["workload"]
    assert set(bench.WORKLOADS) >= {"A", "A200", "C", "D"}
    assert bench.WORKLOADS["C"]["members"] == 1024 and bench.WORKLOADS["D"]["Nx"] == 512


def test_collector_scheduling_edge_cases(monkeypatch):
    """The baton scheduler behind apply(): members that call sim() several times (one batch per round), chunks of
    `max_batch` members, a nested apply() inside a member (runs in place, no second rendezvous), exceptions, and the
    order of the members' side effects (one member runs at a time, in input order, up to its first sim())."""
    batches, trace = [], []

    def fake_run(reqs):
        batches.append(len(reqs))
        for r in reqs:
            r.result = np.tile(r.S0, (r.nSteps + 1, 1)) + r.K[0].mean()
            r.done = True

    monkeypatch.setattr(simulator, "run_requests", fake_run)
    model = simulator.ResSim(Nx=4, Ny=4, Lx=1, Ly=1)
    model.inj_xy, model.prd_xy, model.inj_rates, model.prd_rates = [[0.1, 0.1]], [[0.9, 0.9]], [[1]], [[1]]

    def two_runs(k):
        trace.append(("start", k))
        m = copy.deepcopy(model)
        m.K = np.full(16, k)
        a = m.sim(0.1, 1, np.zeros(16), pbar=False)[-1, 0]
        trace.append(("mid", k))
        if k % 2:  # odd members run a second simulation from the first one's result
            a = m.sim(0.1, 1, np.full(16, a), pbar=False)[-1, 0]
        return a

    utils.nCPU = "auto"
    ks = np.arange(1.0, 7.0)
    out = utils.apply(two_runs, ks, pbar=False)
    assert out == [2.0, 2.0, 6.0, 4.0, 10.0, 6.0]
    assert batches == [6, 3]                                     # all members, then the odd ones
    assert trace[:6] == [("start", k) for k in ks]               # in input order, each up to its first sim()
    assert [t for t in trace[6:]] == [("mid", k) for k in ks]    # and resumed in the same order

    monkeypatch.setattr(utils, "max_batch", 4)                   # chunks of 4 members
    batches.clear()
    out = utils.apply(two_runs, np.arange(2.0, 22.0, 2.0), pbar=False)
    assert batches == [4, 4, 2] and out == list(np.arange(2.0, 22.0, 2.0))
    monkeypatch.setattr(utils, "max_batch", 512)

    def nested(k):  # a member that maps over sub-cases itself: the inner apply runs in place on the member's thread
        inner = utils.apply(lambda j: two_runs(2.0 * j), np.arange(1.0, 4.0), pbar=False)
        return k + sum(inner)

    batches.clear()
    out = utils.apply(nested, np.arange(3.0), pbar=False)
    assert out == [12.0, 13.0, 14.0]
    assert batches == [3, 3, 3]                                  # the outer members' inner calls still rendezvous

    def failing(k):
        if k == 2:
            raise KeyError("member 2")
        return two_runs(k)

    with pytest.raises(KeyError):
        utils.apply(failing, np.arange(1.0, 5.0), pbar=False)
    batches.clear()
    assert utils.apply(two_runs, np.array([4.0, 8.0]), pbar=False) == [4.0, 8.0]  # the workers are reusable after an error
    utils.nCPU = 1


def test_well_schedule_cache_sees_in_place_changes():
    """ResSim caches the (validated) well schedule per instance and shares it with deep copies; assignments AND in-place
    edits of the well arrays invalidate it."""
    model = simulator.ResSim(Nx=4, Ny=4, Lx=1, Ly=1)
    model.inj_xy, model.prd_xy = [[0.1, 0.1]], [[0.9, 0.9], [0.1, 0.9]]
    model.inj_rates, model.prd_rates = [[1.0]], [[0.5], [0.5]]
    q0, c0 = model._schedule(3)
    assert q0.shape == (3, 3) and np.allclose(q0[0], [1, -0.5, -0.5]) and list(c0) == list(model.xy2ind(*np.array([[0.1, 0.1], [0.9, 0.9], [0.1, 0.9]]).T))
    twin = copy.deepcopy(model)
    q1, _ = twin._schedule(3)
    assert q1 is q0 and not q0.flags.writeable                      # shared, read-only
    assert twin._schedule(2)[0].shape == (2, 3)                     # another step count: rebuilt
    twin.inj_rates[0] = 2.0                                         # in place: unbalanced now
    with pytest.raises(ValueError):
        twin._schedule(3)
    twin.prd_rates[:] = 1.0
    assert np.allclose(twin._schedule(3)[0][0], [2, -1, -1]) and np.allclose(twin.actual_rates["inj"], 2.0)
    assert np.allclose(model._schedule(3)[0][0], [1, -0.5, -0.5])   # the original is untouched
    twin.prd_xy = [[0.9, 0.1], [0.1, 0.9]]                          # assignment: new cells
    assert list(twin._schedule(3)[1]) != list(c0)
    model.actual_rates["inj"][:] = 7.0                              # the user's copy is theirs to edit
    assert np.allclose(copy.deepcopy(model)._schedule(3)[0][0], [1, -0.5, -0.5])


def test_robust_increments_host_logic():
    """enopt_fast.robust_increments against the formulas of ens_eval_duplex (Optimise.py:833-853) with a stand-in
    objective: what is paired with what, ONE batch per call (2 nEns members for StoSAG), the mean model."""
    from historymatching_b200.enopt_fast import robust_increments

    rng = np.random.RandomState(2)
    n, du, dx = 5, 2, 7
    U, X, u = rng.randn(n, du), rng.randn(n, dx), rng.randn(du)
    calls = []

    def obj_ux(uu, xx):  # a scalar objective of (control, uncertain parameter)
        return float(np.sin(uu).sum() * np.cos(xx).sum() + uu[0] * xx[1])

    def npv_batch(params):
        calls.append(len(params))
        assert all(set(q) == {"inj_xy", "K"} for q in params)
        return np.array([obj_ux(q["inj_xy"], q["K"]) for q in params])

    paired = robust_increments(npv_batch, "Paired", u, U, X)
    np.testing.assert_allclose(paired, [obj_ux(U[i], X[i]) for i in range(n)])
    stosag = robust_increments(npv_batch, "StoSAG", u, U, X)
    np.testing.assert_allclose(stosag, [obj_ux(U[i], X[i]) - obj_ux(u, X[i]) for i in range(n)])
    x1 = X.mean(0)
    for kind in ("Mean-model", "Fragile"):
        np.testing.assert_allclose(robust_increments(npv_batch, kind, u, U, X), [obj_ux(U[i], x1) for i in range(n)])
    assert calls == [n, 2 * n, n, n]
    with pytest.raises(ValueError):
        robust_increments(npv_batch, "Regular", u, U, X)
    other = robust_increments(lambda ps: np.array([q["rates"].sum() + q["perm"].sum() for q in ps]), "Paired", u, U, X,
                              param_u="rates", param_x="perm")
    np.testing.assert_allclose(other, U.sum(1) + X.sum(1))
