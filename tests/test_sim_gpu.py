"""GPU parity: hm_sim_batch (CUDA, through the C ABI) vs the CPU oracle.

Tolerances (north_star: "within a stated relative tolerance, e.g. 1e-8 FP64"):
pressures 1e-8 of the pressure range, saturations / production curves atol 1e-8
(saturations are O(1)); CFL sub-step counts must agree exactly.

Double-precision floor: the face flux is (P_a - P_b) * T with P = O(1), so it
carries an absolute error ~ eps * T_max in BOTH implementations, and the oracle's
direct solve has a forward error ~ cond(A) * eps.  For permeabilities up to ~1e5
(|log-perm| <= 2.3 under K = 0.1 + exp(5x)) this stays below 1e-9 and the 1e-8
tolerance holds; for the rare members of the notebook prior with K up to 3e8
(x = 3.9) neither implementation is accurate to 1e-8 and the tolerance is
stated as 4 * eps * K_max * nSteps (test_extreme_contrast_floor).
"""

import numpy as np
import pytest

from oracle import ressim as orr

pytestmark = pytest.mark.gpu

SAT_TOL = 1e-8


def _setup(Nx, Ny, N, seed, rough=1.0, clip=2.2):
    from historymatching_b200.sim import GridSpec

    m = orr.notebook_model(Nx, Ny)
    rng = np.random.RandomState(seed)
    # smooth-ish random log-permeability with the notebook's contrast (K = 0.1 + exp(5x))
    x = rng.randn(N, Nx // 4 + 2, Ny // 4 + 2)
    x = np.kron(x, np.ones((4, 4)))[:, :Nx, :Ny]
    for _ in range(3):
        x = 0.25 * (np.roll(x, 1, 1) + np.roll(x, -1, 1) + np.roll(x, 1, 2) + np.roll(x, -1, 2))
    x = np.clip(rough * x / x.std(), -clip, clip)
    logk = x.reshape(N, -1)
    grid = GridSpec(Nx=Nx, Ny=Ny, Lx=m.Lx, Ly=m.Ly)
    inj = m.xy2ind(*m.inj_xy.T)
    prd = m.xy2ind(*m.prd_xy.T)
    cells = np.concatenate([inj, prd]).astype(np.int32)
    rates = np.concatenate([m.inj_rates[:, 0], -m.prd_rates[:, 0]])
    return m, grid, logk, cells, rates, prd.astype(np.int32)


def _oracle(m, logk, dt, nT, S0, prd):
    outs = [orr.forward_member(m, lk, dt, nT, S0 if S0.ndim == 1 else S0[i], prd) for i, lk in enumerate(logk)]
    return np.array([o[0] for o in outs]), np.array([o[1] for o in outs])


# sat_block 0 = automatic: grids of <= 2048 cells run in the fused one-CTA-per-member kernel (hm_small.cu),
# larger ones on the streamed path; 2 = streamed path with the cluster transport kernel; 1 = streamed path
# with the streaming transport kernel (bulk-copy staged when the row length is even); 5 = streamed path with the
# plain-load streaming transport kernel.
@pytest.mark.parametrize("Nx,Ny,N,nT,sat_block", [
    (20, 20, 6, 40, 0), (20, 20, 6, 40, 2), (33, 17, 3, 5, 0), (33, 17, 3, 5, 2), (33, 17, 3, 5, 1),
    (45, 45, 2, 3, 0), (26, 30, 2, 4, 0), (64, 64, 2, 3, 0),
    # streaming transport: 1 = bulk-copy staged kernel (even row length), 5 = plain-load kernel
    (26, 30, 2, 4, 1), (64, 64, 2, 3, 1), (26, 30, 2, 4, 5), (70, 36, 2, 2, 1),
    # cluster transport kernel: general (predicated) path, compile-time row length 64, run-time row length 256
    (100, 100, 2, 2, 0), (96, 64, 2, 2, 0), (40, 256, 2, 2, 0),
    # streaming kernels on several ragged tiles; vectorised multigrid kernels (row length 128) with a ragged last tile
    (150, 60, 1, 2, 1), (150, 61, 1, 2, 1), (72, 128, 2, 2, 0), (150, 60, 1, 2, 6), (64, 64, 2, 3, 6)])
def test_forward_ensemble_matches_oracle(Nx, Ny, N, nT, sat_block):
    from historymatching_b200.sim import run_ensemble

    m, grid, logk, cells, rates, prd = _setup(Nx, Ny, N, seed=Nx + Ny)
    dt = 0.025
    S0 = np.zeros(grid.M)
    res = run_ensemble(grid, orr.perm_transf(logk), cells, rates, S0, dt, nT, obs_cell=prd,
                       history=True, pressure=True, want_substeps=True, sat_block=sat_block)
    assert not res.status.any()
    assert (res.stats["kernel_launches"] == 1) == (sat_block == 0 and Nx * Ny <= 2048)  # fused kernel: one launch
    wsats, prods = _oracle(m, logk, dt, nT, S0, prd)
    np.testing.assert_array_equal(res.S_hist[:, 0], 0.0)
    np.testing.assert_allclose(res.S_hist, wsats, rtol=0, atol=SAT_TOL)
    np.testing.assert_allclose(res.obs, prods, rtol=0, atol=SAT_TOL)
    np.testing.assert_allclose(res.S_last, wsats[:, -1], rtol=0, atol=SAT_TOL)
    # sub-step counts are integers of the flow field: must match exactly
    mm = orr.notebook_model(Nx, Ny)
    p = orr.perm_transf(logk[0]).reshape(mm.shape)
    mm.K = np.stack([p, p])
    _, aux = mm.sim(dt, nT, S0, return_aux=True)
    np.testing.assert_array_equal(res.substeps[0], aux["Nts"])
    P = aux["P"][-1]
    np.testing.assert_allclose(res.P_last[0], P, rtol=0, atol=1e-8 * np.abs(P).max())


def test_device_path_equals_host_path():
    import torch

    from historymatching_b200.sim import run_ensemble

    m, grid, logk, cells, rates, prd = _setup(20, 20, 4, seed=3)
    K = orr.perm_transf(logk)
    host = run_ensemble(grid, K, cells, rates, np.zeros(grid.M), 0.025, 4, obs_cell=prd)
    dev = run_ensemble(grid, torch.as_tensor(K, device="cuda"), cells, rates, np.zeros(grid.M), 0.025, 4,
                       obs_cell=prd)
    np.testing.assert_array_equal(host.S_last, dev.S_last.cpu().numpy())
    np.testing.assert_array_equal(host.obs, dev.obs.cpu().numpy())
    assert dev.stats["kernel_launches"] > 0


def test_restart_and_per_member_state():
    """forward_model(perm, wsat.curnt) (HistoryMatch.py:1227): per-member initial saturation."""
    from historymatching_b200.sim import run_ensemble

    m, grid, logk, cells, rates, prd = _setup(20, 20, 3, seed=11)
    K = orr.perm_transf(logk)
    full = run_ensemble(grid, K, cells, rates, np.zeros(grid.M), 0.025, 8, obs_cell=prd)
    first = run_ensemble(grid, K, cells, rates, np.zeros(grid.M), 0.025, 4, obs_cell=prd)
    second = run_ensemble(grid, K, cells, rates, first.S_last, 0.025, 4, obs_cell=prd)
    np.testing.assert_allclose(second.S_last, full.S_last, rtol=0, atol=1e-9)  # cold vs warm CG start
    np.testing.assert_allclose(second.obs, full.obs[:, 4:], rtol=0, atol=1e-9)


def test_per_member_wells_and_schedules():
    """EnOpt-style batches (Optimise.py:112-125): wells / rates differ per member."""
    from historymatching_b200.sim import GridSpec, run_ensemble

    Nx = Ny = 16
    nT, dt = 6, 0.025
    grid = GridSpec(Nx=Nx, Ny=Ny, Lx=2.0, Ly=1.0)
    rng = np.random.RandomState(5)
    K = np.exp(rng.randn(Nx * Ny) * 0.5)
    N = 3
    wc = np.zeros((N, 3), np.int32)
    wr = np.zeros((N, nT, 3))
    S_ref = []
    for i in range(N):
        m = orr.OracleResSim(Nx, Ny, 2.0, 1.0)
        m.K = np.stack([K.reshape(Nx, Ny)] * 2)
        m.inj_xy = m.ind2xy(np.array([rng.randint(Nx * Ny)])).T
        m.prd_xy = m.ind2xy(rng.choice(Nx * Ny, 2, replace=False)).T
        inj = 0.5 + rng.rand(1, nT)
        split = rng.rand(1, nT)
        m.inj_rates = inj
        m.prd_rates = np.concatenate([inj * split, inj * (1 - split)])
        wc[i] = np.concatenate([m.xy2ind(*m.inj_xy.T), m.xy2ind(*m.prd_xy.T)])
        wr[i] = np.concatenate([m.inj_rates, -m.prd_rates]).T
        S_ref.append(m.sim(dt, nT, np.zeros(Nx * Ny)))
    res = run_ensemble(grid, K, wc, wr, np.zeros(Nx * Ny), dt, nT, history=True, n_members=N)
    assert not res.status.any()
    np.testing.assert_allclose(res.S_hist, np.array(S_ref), rtol=0, atol=SAT_TOL)


def test_full_size_properties():
    """BASELINE config sizes (128^2): size-independent invariants instead of the oracle."""
    import torch

    from historymatching_b200.sim import run_ensemble

    m, grid, logk, cells, rates, prd = _setup(128, 128, 8, seed=1)
    K = torch.as_tensor(orr.perm_transf(logk), device="cuda")
    nT, dt = 3, 0.025
    res = run_ensemble(grid, K, cells, rates, torch.zeros(grid.M, dtype=torch.float64, device="cuda"), dt, nT,
                       obs_cell=prd, history=True, want_substeps=True)
    S = res.S_hist.cpu().numpy()
    assert not res.status.cpu().numpy().any()
    assert (res.substeps.cpu().numpy() == 615).all()
    assert S.min() >= 0 and S.max() < 1
    vol = S.sum(-1) * (grid.Lx / grid.Nx) * (grid.Ly / grid.Ny)
    np.testing.assert_allclose(vol, np.broadcast_to(dt * np.arange(nT + 1), vol.shape), rtol=1e-10, atol=1e-13)
    assert np.abs(res.obs.cpu().numpy()).max() < 1e-12  # no breakthrough yet


def test_config_d_size_properties():
    """BASELINE config D grid (512^2, temporally blocked transport on row strips, 3 streamed multigrid levels): invariants
    (the comparison with the oracle at this size: tests/test_transport_tb_gpu.py::test_config_d_grid_matches_oracle)."""
    import torch

    from historymatching_b200.sim import run_ensemble

    m, grid, logk, cells, rates, prd = _setup(512, 512, 2, seed=2)
    K = torch.as_tensor(orr.perm_transf(logk), device="cuda")
    dt = 0.025
    res = run_ensemble(grid, K, cells, rates, torch.zeros(grid.M, dtype=torch.float64, device="cuda"), dt, 1,
                       obs_cell=prd, want_substeps=True)
    S = res.S_last.cpu().numpy()
    assert not res.status.cpu().numpy().any()
    assert (res.substeps.cpu().numpy() == 9831).all()
    assert res.stats["sat_tb_strips"] > 1                 # a member does not fit one cluster: overlapping row strips,
    assert res.stats["sat_kernel_launches"] == -(-9831 // res.stats["sat_tb_halo"])  # one launch per round of sub-steps
    assert S.min() >= 0 and S.max() < 1
    vol = S.sum(-1) * (grid.Lx / grid.Nx) * (grid.Ly / grid.Ny)
    np.testing.assert_allclose(vol, dt, rtol=1e-9)     # water balance before breakthrough
    assert np.abs(res.obs.cpu().numpy()).max() < 1e-12


def test_extreme_contrast_floor(golden):
    """The notebook's own seed-1 prior (HistoryMatch.py:78,290) contains K up to 2.8e8."""
    from historymatching_b200.sim import run_ensemble

    m, grid, _, cells, rates, prd = _setup(20, 20, 1, seed=0)
    logk = golden("prior_20x20_seed1.npz")["prior"][[2, 6, 36, 0, 1]]
    nT, dt = 40, 0.025
    S0 = np.zeros(grid.M)
    res = run_ensemble(grid, orr.perm_transf(logk), cells, rates, S0, dt, nT, obs_cell=prd, history=True)
    assert not res.status.any()
    eps = np.finfo(float).eps
    for i, lk in enumerate(logk):
        for refine in (0, 3):
            mm = orr.notebook_model(20, 20)
            p = orr.perm_transf(lk).reshape(mm.shape)
            mm.K = np.stack([p, p])
            mm.refine = refine
            ref = mm.sim(dt, nT, S0)
            tol = max(SAT_TOL, 4 * eps * p.max() * nT)
            err = np.abs(res.S_hist[i] - ref).max()
            assert err <= tol, (i, refine, err, tol)


@pytest.mark.parametrize("sat_block", [0, 1, 2, 5])
def test_non_default_fluid_and_porosity(sat_block):
    """Viscosity ratio, irreducible saturations and a porosity field (both transport kernels)."""
    from historymatching_b200.sim import GridSpec, run_ensemble

    m, grid, logk, cells, rates, prd = _setup(24, 16, 3, seed=9)
    grid = GridSpec(24, 16, m.Lx, m.Ly, vw=0.8, vo=1.3, swc=0.1, sor=0.15)
    rng = np.random.RandomState(0)
    por = 0.6 + 0.4 * rng.rand(24 * 16)
    S0 = np.full(grid.M, 0.1)
    nT, dt = 6, 0.02
    res = run_ensemble(grid, orr.perm_transf(logk), cells, rates, S0, dt, nT, obs_cell=prd, history=True, por=por,
                       want_substeps=True, sat_block=sat_block)
    assert not res.status.any()
    for i in range(3):
        om = orr.OracleResSim(24, 16, m.Lx, m.Ly, vw=0.8, vo=1.3, swc=0.1, sor=0.15)
        p = orr.perm_transf(logk[i]).reshape(om.shape)
        om.K = np.stack([p, p])
        om.por = por.reshape(om.shape)
        om.inj_xy, om.prd_xy, om.inj_rates, om.prd_rates = m.inj_xy, m.prd_xy, m.inj_rates, m.prd_rates
        ref, aux = om.sim(dt, nT, S0, return_aux=True)
        np.testing.assert_array_equal(res.substeps[i], aux["Nts"])
        np.testing.assert_allclose(res.S_hist[i], ref, rtol=0, atol=SAT_TOL)


@pytest.mark.parametrize("precond", [0, 1, 2, 3, 4])
def test_pressure_preconditioners_agree_with_oracle(precond):
    """Every preconditioner (FP32 / FP64 V-cycle, W-cycle, Jacobi) solves to the same FP64 tolerance."""
    from historymatching_b200.sim import run_ensemble

    m, grid, logk, cells, rates, prd = _setup(64, 64, 2, seed=11)
    dt, nT = 0.025, 3
    res = run_ensemble(grid, orr.perm_transf(logk), cells, rates, np.zeros(grid.M), dt, nT, obs_cell=prd,
                       history=True, pressure=True, want_substeps=True, precond=precond)
    assert not res.status.any()
    # precond 0 may switch to the FP64 cycle on this blocky field (it needs > 40 iterations); the others never do
    assert res.stats["mg_fp64_fallbacks"] <= (nT if precond == 0 else 0)
    wsats, _ = _oracle(m, logk, dt, nT, np.zeros(grid.M), prd)
    np.testing.assert_allclose(res.S_hist, wsats, rtol=0, atol=SAT_TOL)


def test_fp32_cycle_falls_back_to_fp64():
    """precond 0: a solve that is still running after mg_switch_iters iterations restarts CG with the FP64
    cycle (here forced after 2 iterations); the result is unchanged and the rest of the run stays FP64."""
    from historymatching_b200.sim import run_ensemble

    m, grid, logk, cells, rates, prd = _setup(64, 48, 3, seed=5)
    dt, nT = 0.025, 3
    res = run_ensemble(grid, orr.perm_transf(logk), cells, rates, np.zeros(grid.M), dt, nT, obs_cell=prd,
                       history=True, want_substeps=True, mg_switch_iters=2)
    assert not res.status.any()
    assert res.stats["mg_fp64_fallbacks"] == 1  # sticky: only the first solve switches
    wsats, _ = _oracle(m, logk, dt, nT, np.zeros(grid.M), prd)
    np.testing.assert_allclose(res.S_hist, wsats, rtol=0, atol=SAT_TOL)


@pytest.mark.parametrize("Nx,Ny", [(64, 64), (128, 128), (48, 64)])
def test_half_tile_cluster_kernel_matches_default(Nx, Ny):
    """sat_block 4 (1024-cell tiles, two CTAs per SM) against the default cluster kernel: same sub-step counts,
    saturations equal to rounding (the two kernels evaluate the same expressions)."""
    from historymatching_b200.sim import run_ensemble

    m, grid, logk, cells, rates, prd = _setup(Nx, Ny, 3, seed=Nx)
    dt, nT = 0.025, 2
    kw = dict(obs_cell=prd, want_substeps=True)
    a = run_ensemble(grid, orr.perm_transf(logk), cells, rates, np.zeros(grid.M), dt, nT, sat_block=2, **kw)
    b = run_ensemble(grid, orr.perm_transf(logk), cells, rates, np.zeros(grid.M), dt, nT, sat_block=4, **kw)
    assert not a.status.any() and not b.status.any()
    np.testing.assert_array_equal(a.substeps, b.substeps)
    np.testing.assert_allclose(b.S_last, a.S_last, rtol=0, atol=1e-12)


@pytest.mark.parametrize("Nx,Ny,nT,stride", [(20, 20, 7, 3), (20, 20, 6, 3), (64, 64, 5, 2), (64, 64, 4, 5)])
def test_strided_history(Nx, Ny, nT, stride):
    """history=k keeps rows 0, k, 2k, ... and the last step of the full history (fused and streamed path)."""
    from historymatching_b200.sim import run_ensemble

    m, grid, logk, cells, rates, prd = _setup(Nx, Ny, 3, seed=2)
    args = (grid, orr.perm_transf(logk), cells, rates, np.zeros(grid.M), 0.025, nT)
    full = run_ensemble(*args, history=True)
    part = run_ensemble(*args, history=stride)
    rows = sorted(set(range(0, nT + 1, stride)) | {nT})
    assert part.S_hist.shape == (3, len(rows), grid.M)
    np.testing.assert_array_equal(part.S_hist, full.S_hist[:, rows])
    np.testing.assert_array_equal(part.S_last, full.S_last)


def test_concurrent_lanes_are_bit_identical():
    """lanes=2 / 3: member shares run from concurrent host threads, contexts and streams; same bits, merged stats."""
    import torch

    from historymatching_b200.sim import run_ensemble

    # a mild field: every solve converges within the FP32 cycle and below the verification threshold, so each member's
    # path is independent of which members share its batch (the FP64 fallback and the true-residual check are taken
    # per batch; with them the lanes agree to the solver tolerance instead of bit for bit)
    m, grid, logk, cells, rates, prd = _setup(64, 64, 7, seed=4, rough=0.3)
    K = torch.as_tensor(orr.perm_transf(logk), device="cuda")
    S0 = torch.zeros(grid.M, dtype=torch.float64, device="cuda")
    kw = dict(obs_cell=prd, history=2, pressure=True, want_substeps=True)
    one = run_ensemble(grid, K, cells, rates, S0, 0.025, 3, **kw)
    for lanes in (2, 3):
        many = run_ensemble(grid, K, cells, rates, S0, 0.025, 3, lanes=lanes, **kw)
        assert many.stats["lanes"] == lanes
        for f in ("S_last", "obs", "S_hist", "P_last", "status", "substeps", "cg_iters"):
            assert torch.equal(getattr(many, f), getattr(one, f)), f
        assert many.stats["sat_kernel_launches"] == lanes * one.stats["sat_kernel_launches"]
        assert many.stats["mg_fp64_fallbacks"] == 0 and many.stats["cg_restarts"] == 0
    # per-member wells and initial state are split with the members
    wc = torch.as_tensor(np.tile(cells, (7, 1)), device="cuda")
    wr = torch.as_tensor(np.tile(rates, (7, 1, 1)), device="cuda")
    S0m = torch.zeros(7, grid.M, dtype=torch.float64, device="cuda")
    per = run_ensemble(grid, K, wc, wr, S0m, 0.025, 3, lanes=2, **kw)
    assert torch.equal(per.S_last, one.S_last)


@pytest.mark.parametrize("sat_block", [0, 2])
def test_buckley_leverett_known_answer(sat_block):
    """The CUDA path against an answer that does not come from the oracle: the Buckley-Leverett profile of a 1-D
    displacement (tests/test_oracle_golden.py::test_sim_converges_to_buckley_leverett), on the fused and the streamed path."""
    from scipy.optimize import brentq

    from historymatching_b200.sim import GridSpec, run_ensemble

    Nx, T, nT = 200, 0.5, 10
    hx = 1.0 / Nx
    grid = GridSpec(Nx=Nx, Ny=2, Lx=1.0, Ly=1.0)
    cells = np.array([0, 1, 2 * (Nx - 1), 2 * (Nx - 1) + 1], np.int32)     # injectors in row 0, producers in the last row
    rates = np.array([0.5, 0.5, -0.5, -0.5])
    res = run_ensemble(grid, np.ones(grid.M), cells, rates, np.zeros(grid.M), T / nT, nT, sat_block=sat_block)
    assert not res.status.any()
    S = res.S_last[0].reshape(Nx, 2)
    assert np.abs(S[:, 0] - S[:, 1]).max() < 1e-8
    x = (np.arange(Nx) + 0.5) * hx
    df = lambda s: 2 * s * (1 - s) / (s * s + (1 - s) ** 2) ** 2              # noqa: E731
    s_shock = 1 / np.sqrt(2)
    v_shock = (s_shock**2 / (s_shock**2 + (1 - s_shock) ** 2)) / s_shock
    exact = np.array([brentq(lambda s: df(s) - xi, s_shock, 1.0) if xi < v_shock else 0.0 for xi in x / T])
    assert np.abs(S[:, 0] - exact).mean() < 0.009                           # first-order upwind at h = 1/200
    assert abs(x[np.argmax(S[:, 0] < 0.35)] - v_shock * T) < 0.012
    assert abs(S[:, 0].sum() * hx - T) < 1e-10                              # injected volume


@pytest.mark.parametrize("Nx,Ny", [(20, 20), (64, 64)])
def test_perm_transf_inside_the_kernels(Nx, Ny):
    """hm_sim_desc.K_transform: the forward run takes the log-permeability parameter x and evaluates the notebook's
    perm_transf 0.1 + exp(5 x) (HistoryMatch.py:137-138) where the transmissibilities are built (fused and streamed
    path); same result as handing it the transformed field (exp differs from numpy's in the last bit at most)."""
    from historymatching_b200.sim import run_ensemble

    m, grid, logk, cells, rates, prd = _setup(Nx, Ny, 3, seed=8)
    args = (cells, rates, np.zeros(grid.M), 0.025, 3)
    a = run_ensemble(grid, orr.perm_transf(logk), *args, obs_cell=prd, pressure=True)
    b = run_ensemble(grid, logk, *args, obs_cell=prd, pressure=True, k_transform=(0.1, 5.0))
    assert not b.status.any()
    # K differs in the last bit, each pressure solve is accurate to the CG tolerance: the stated parity tolerance applies
    np.testing.assert_allclose(b.S_last, a.S_last, rtol=0, atol=SAT_TOL)
    np.testing.assert_allclose(b.obs, a.obs, rtol=0, atol=SAT_TOL)
    np.testing.assert_allclose(b.P_last, a.P_last, rtol=0, atol=1e-8 * np.abs(a.P_last).max())
    # and against the oracle (which transforms on the host)
    wsats, _ = _oracle(m, logk, 0.025, 3, np.zeros(grid.M), prd)
    np.testing.assert_allclose(b.S_last, wsats[:, -1], rtol=0, atol=SAT_TOL)
