"""Pin the oracle: golden vectors from the reference's own functions
(tests/golden/make_golden.py), the reference doctest values and the
notebook's in-cell self-checks."""

import numpy as np
import pytest

from oracle import analysis as oa
from oracle import ressim as orr

TOL = dict(rtol=1e-11, atol=1e-12)


def test_primitives_match_reference(golden):
    g = golden("primitives.npz")
    X, x = oa.center(g["E"])
    np.testing.assert_array_equal(X, g["center_X"])
    np.testing.assert_array_equal(x, g["center_x"])
    np.testing.assert_array_equal(oa.center(g["E"], rescale=True)[0], g["center_X_rescaled"])
    np.testing.assert_allclose(oa.cov(g["E"], g["b"]), g["cov"], **TOL)
    np.testing.assert_allclose(oa.corr(g["E"], g["b"][:, 0]), g["corr"], **TOL)
    np.testing.assert_allclose(oa.rinv(g["A"], 0.1, tikh=True), g["rinv_tikh"], **TOL)
    np.testing.assert_allclose(oa.rinv(g["A"], 0.1, tikh=False), g["rinv_trunc"], **TOL)
    np.testing.assert_array_equal(oa.pairwise_distances(g["pts_a"], g["pts_b"]), g["pd_ab"])
    np.testing.assert_array_equal(oa.pairwise_distances(g["pts_a"]), g["pd_aa"])
    np.testing.assert_array_equal(oa.pairwise_distances(g["pts_a"], g["pts_b"], domain=(2, 1)), g["pd_periodic"])
    np.testing.assert_array_equal(oa.bump(g["dist"]), g["bump1"])
    np.testing.assert_array_equal(oa.bump(g["dist"], 10), g["bump_sharp"])
    np.testing.assert_array_equal(oa.bump(g["dist"], 0.1), g["bump_soft"])


def test_reference_doctest_values():
    # tools/geostat.py:19-22
    np.testing.assert_allclose(oa.variogram_gauss(np.array([0, 1, 2]), 1, n=0.1, a=1),
                               [0.0, 0.6689085, 0.98351593], atol=5e-9)
    # tools/localization.py:31-60
    A = [[0, 0], [0, 1], [1, 0], [1, 1]]
    D = oa.pairwise_distances(A)
    assert np.allclose(D[0], [0, 1, 1, 2**0.5])
    A1 = np.arange(4)[:, None]
    np.testing.assert_array_equal(oa.pairwise_distances(A1, [[2]]).T, [[2.0, 1.0, 0.0, 1.0]])
    np.testing.assert_array_equal(oa.pairwise_distances(A1, domain=(4,))[0], [0.0, 1.0, 2.0, 1.0])
    np.testing.assert_array_equal(oa.pairwise_distances(np.arange(4)), [[0.0]])


def test_prior_bit_exact_same_stream(golden):
    """HistoryMatch.py:78,167,290: seed(1) -> truth -> 40-member prior."""
    g = golden("prior_20x20_seed1.npz")
    m = orr.OracleResSim(Nx=20, Ny=20, Lx=2, Ly=1)
    np.random.seed(1)
    truth = oa.gaussian_fields(m.mesh, 1, r=0.8)
    prior = oa.gaussian_fields(m.mesh, 40, r=0.8)
    np.testing.assert_array_equal(truth, g["truth"])
    np.testing.assert_array_equal(prior, g["prior"])


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_updates_match_reference(golden, tag):
    g = golden("updates.npz")
    H = g[f"{tag}_H"]

    def fwd(X):
        return np.tanh(X @ H) + 0.1 * (X @ H)

    kw = {k: g[f"{tag}_{k}"] for k in ("prior_ens", "obs", "perturbs", "decorr")}
    Eo, taper = g[f"{tag}_obs_ens"], g[f"{tag}_taper"]
    np.testing.assert_allclose(oa.ens_update0(obs_ens=Eo, **kw), g[f"{tag}_ES"], **TOL)
    np.testing.assert_allclose(oa.ens_update0_loc(obs_ens=Eo, taper=taper, **kw), g[f"{tag}_LES"], **TOL)
    E, st = oa.IES(obs_ens=fwd, xStep=0.6, iMax=3, **kw)
    np.testing.assert_allclose(E, g[f"{tag}_IES"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(np.array(st.E), g[f"{tag}_IES_E"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(np.array(st.Eo), g[f"{tag}_IES_Eo"], rtol=1e-9, atol=1e-10)
    E, st = oa.ILES(obs_ens=fwd, taper=taper, xStep=0.6, iMax=3, **kw)
    np.testing.assert_allclose(E, g[f"{tag}_ILES"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(np.array(st.E), g[f"{tag}_ILES_E"], rtol=1e-9, atol=1e-10)


def test_notebook_self_checks(golden):
    """HistoryMatch.py:598-612, 811-822, 949-951, 1069-1071."""
    g = golden("gauss_gauss.npz")
    kw = {k: g[k] for k in ("prior_ens", "obs", "perturbs", "decorr")}
    E = kw["prior_ens"]
    post = oa.ens_update0(obs_ens=E, **kw)
    np.testing.assert_allclose(post, g["post"], **TOL)
    post_loc = oa.ens_update0_loc(obs_ens=E, taper=np.eye(3), **kw)
    np.testing.assert_allclose(post_loc, g["post_loc"], **TOL)
    assert np.allclose(oa.ens_update0_loc(obs_ens=E, taper=np.ones((3, 3)), **kw), post)
    assert np.allclose(oa.IES(obs_ens=lambda x: x, **kw)[0], post)
    assert np.allclose(oa.ILES(obs_ens=lambda x: x, taper=np.eye(3), **kw)[0], post_loc)
    # posterior moments ~ N(1, I) up to sampling error
    assert np.allclose(post.mean(0), 1.0, atol=0.2)
    assert np.allclose(np.cov(post.T), np.eye(3), atol=0.25)


def test_es_mda_single_pass_is_es(golden):
    g = golden("updates.npz")
    H = g["small_H"]
    kw = {k: g[f"small_{k}"] for k in ("prior_ens", "obs", "decorr")}
    p = len(kw["obs"])
    R12 = np.linalg.inv(kw["decorr"]).T
    Z = np.linalg.solve(R12, g["small_perturbs"].T).T
    E, _ = oa.es_mda(kw["prior_ens"], lambda X: np.tanh(X @ H) + 0.1 * (X @ H), kw["obs"], R12, [1.0],
                     decorr=kw["decorr"], perturbs=[Z])
    np.testing.assert_allclose(E, g["small_ES"], rtol=1e-9, atol=1e-10)
    assert p == Z.shape[1]


# ---- simulator oracle: self-consistency invariants (parity unpinned) -------------------------
@pytest.fixture(scope="module")
def sim_run(golden):
    m = orr.notebook_model(20, 20)
    x = golden("prior_20x20_seed1.npz")["truth"][0]
    p = orr.perm_transf(x).reshape(m.shape)
    m.K = np.stack([p, p])
    S, aux = m.sim(0.025, 40, np.zeros(m.Nxy), return_aux=True)
    return m, S, aux


def test_sim_row0_bounds_and_substeps(sim_run):
    m, S, aux = sim_run
    assert S.shape == (41, 400) and not S[0].any()
    assert S.min() >= 0 and S.max() < 1
    assert set(aux["Nts"]) == {15}


def test_sim_water_balance(sim_run):
    m, S, _ = sim_run
    prod = m.xy2ind(*m.prd_xy.T)
    vol = S.sum(1) * m.h2
    t = 0.025 * np.arange(41)
    # before breakthrough every injected volume stays in the reservoir
    early = S[:, prod].max(1) < 1e-12
    assert early[:10].all()
    np.testing.assert_allclose(vol[early], t[early] * 1.0, rtol=1e-12, atol=1e-14)
    assert np.all(np.diff(vol) > 0) and vol[-1] <= t[-1] + 1e-12


def test_sim_symmetry_homogeneous():
    m = orr.notebook_model(21, 21)  # odd grid: the injector sits in the centre cell
    S = m.sim(0.025, 6, np.zeros(m.Nxy))[-1].reshape(m.shape)
    assert S.max() > 0.5
    np.testing.assert_allclose(S, S[::-1, :], atol=1e-12)
    np.testing.assert_allclose(S, S[:, ::-1], atol=1e-12)


def test_sim_pressure_small_dense():
    m = orr.OracleResSim(4, 4, 1.0, 1.0)
    m.K = np.exp(np.random.RandomState(0).randn(2, 4, 4))
    q = np.zeros(16)
    q[5], q[10] = 1.0, -1.0
    S = np.full(16, 0.3)
    P, Vx, Vy = m.pressure_step(S, q)
    lw, lo = m.mobilities(S)
    TX, TY = m.transmissibilities((lw + lo).reshape(4, 4) * m.K)
    A = m.pressure_matrix(TX, TY).toarray()
    np.testing.assert_allclose(A, A.T, atol=1e-14)
    np.testing.assert_allclose(np.linalg.solve(A, q), P.ravel(), rtol=1e-10, atol=1e-13)
    div = Vx[1:] - Vx[:-1] + Vy[:, 1:] - Vy[:, :-1]
    np.testing.assert_allclose(div.ravel(), q, atol=1e-10)


def test_sim_rejects_unbalanced_and_outside():
    m = orr.notebook_model(8, 8)
    m.inj_rates = np.array([[2.0]])
    with pytest.raises(ValueError):
        m.sim(0.025, 1, np.zeros(64))
    with pytest.raises(ValueError):
        m.xy2ind(np.array([2.5]), np.array([0.5]))


def test_obs_error_model_matches_notebook_cell(golden):
    """oracle.obs_error_model against the notebook's own statements (HistoryMatch.py:243-247, 259), executed by
    tests/golden/make_golden.py: the temporally correlated observation-error covariance and its Cholesky factor."""
    g = golden("obs_error.npz")
    R, R12 = oa.obs_error_model(int(g["nTime"]), int(g["nPrd"]))
    np.testing.assert_array_equal(R, g["R"])
    np.testing.assert_array_equal(R12, g["R12"])


def test_sim_converges_to_buckley_leverett():
    """Known answer independent of the (absent) simulator package: 1-D displacement with quadratic relative
    permeabilities and unit viscosity ratio has the Buckley-Leverett solution - a rarefaction behind a shock of
    height 1/sqrt(2) moving at f(s*)/s* = 1.2071 pore volumes per unit time.  The restated scheme (first-order
    upwind, CFL sub-stepped, sequential splitting) must converge to it: front position and L1 error shrink with h,
    mass is conserved exactly."""
    from scipy.optimize import brentq

    def run(Nx, T=0.5, nT=10):
        m = orr.OracleResSim(Nx=Nx, Ny=2, Lx=1.0, Ly=1.0)   # two identical rows: a 1-D flow
        m.K = np.ones((2, Nx, 2))
        hx = 1.0 / Nx
        m.inj_xy = np.array([[hx / 2, 0.25], [hx / 2, 0.75]])
        m.prd_xy = np.array([[1 - hx / 2, 0.25], [1 - hx / 2, 0.75]])
        m.inj_rates = m.prd_rates = np.array([[0.5], [0.5]])
        S = m.sim(T / nT, nT, np.zeros(2 * Nx))[-1].reshape(Nx, 2)
        assert np.abs(S[:, 0] - S[:, 1]).max() < 1e-8
        x = (np.arange(Nx) + 0.5) * hx
        df = lambda s: 2 * s * (1 - s) / (s * s + (1 - s) ** 2) ** 2          # noqa: E731
        s_shock = 1 / np.sqrt(2)
        v_shock = (s_shock**2 / (s_shock**2 + (1 - s_shock) ** 2)) / s_shock
        exact = np.array([brentq(lambda s: df(s) - xi, s_shock, 1.0) if xi < v_shock else 0.0 for xi in x / T])
        front = x[np.argmax(S[:, 0] < 0.35)]
        return abs(front - v_shock * T), np.abs(S[:, 0] - exact).mean(), S[:, 0].sum() * hx

    # Darcy's law for the initial state (S = 0: total mobility 1): a unit Darcy velocity needs dP/dx = -1 exactly
    m = orr.OracleResSim(Nx=50, Ny=2, Lx=1.0, Ly=1.0)
    m.K = np.ones((2, 50, 2))
    m.inj_xy, m.prd_xy = np.array([[0.01, 0.25], [0.01, 0.75]]), np.array([[0.99, 0.25], [0.99, 0.75]])
    m.inj_rates = m.prd_rates = np.array([[0.5], [0.5]])
    P, Vx, Vy = m.pressure_step(np.zeros(100), m.source_field(0))
    np.testing.assert_allclose(np.diff(P, axis=0), -1.0 / 50, rtol=1e-10)
    np.testing.assert_allclose(Vx[1:-1], 0.5, rtol=1e-10)
    assert np.abs(Vy).max() < 1e-12

    e100, e200 = run(100), run(200)
    assert e100[0] < 0.02 and e200[0] < 0.012 and e200[0] < e100[0]      # front position converges
    assert e100[1] < 0.02 and e200[1] < 0.009 and e200[1] < 0.6 * e100[1]   # L1 error converges
    assert abs(e100[2] - 0.5) < 1e-10 and abs(e200[2] - 0.5) < 1e-10     # injected volume = t * rate
