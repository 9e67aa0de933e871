"""GPU parity of the analysis path (through the C ABI) vs golden vectors from the
reference's own functions and vs the oracle.  Tolerance: rtol 1e-8 as stated by
north_star for posterior ensembles (observed ~1e-13)."""

import numpy as np
import pytest

from oracle import analysis as oa

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-8, atol=1e-10)


def _case(g, tag):
    return {k: g[f"{tag}_{k}"] for k in ("prior_ens", "obs", "perturbs", "decorr")}


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_es_les_against_reference_golden(golden, tag):
    from historymatching_b200 import analysis as ha

    g = golden("updates.npz")
    kw = _case(g, tag)
    Eo, taper = g[f"{tag}_obs_ens"], g[f"{tag}_taper"]
    np.testing.assert_allclose(ha.ens_update0(obs_ens=Eo, **kw), g[f"{tag}_ES"], **TOL)
    np.testing.assert_allclose(ha.ens_update0_loc(obs_ens=Eo, taper=taper, **kw), g[f"{tag}_LES"], **TOL)


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_ies_against_reference_golden(golden, tag):
    from historymatching_b200 import analysis as ha

    g = golden("updates.npz")
    H = g[f"{tag}_H"]

    def fwd(X):
        return np.tanh(X @ H) + 0.1 * (X @ H)

    E, st = ha.IES(obs_ens=fwd, xStep=0.6, iMax=3, **_case(g, tag))
    np.testing.assert_allclose(E, g[f"{tag}_IES"], **TOL)
    np.testing.assert_allclose(np.array(st.E), g[f"{tag}_IES_E"], **TOL)
    np.testing.assert_allclose(np.array(st.Eo), g[f"{tag}_IES_Eo"], **TOL)


def test_notebook_self_checks_on_gpu(golden):
    """HistoryMatch.py:598-612, 811-822, 949-951."""
    from historymatching_b200 import analysis as ha

    g = golden("gauss_gauss.npz")
    kw = {k: g[k] for k in ("prior_ens", "obs", "perturbs", "decorr")}
    E = kw["prior_ens"]
    post = ha.ens_update0(obs_ens=E, **kw)
    np.testing.assert_allclose(post, g["post"], **TOL)
    np.testing.assert_allclose(ha.ens_update0_loc(obs_ens=E, taper=np.eye(3), **kw), g["post_loc"], **TOL)
    np.testing.assert_allclose(ha.ens_update0_loc(obs_ens=E, taper=np.ones((3, 3)), **kw), post, **TOL)
    np.testing.assert_allclose(ha.IES(obs_ens=lambda x: x, **kw)[0], post, rtol=1e-7, atol=1e-9)


def _hm_case(N, M, p, seed):
    rng = np.random.RandomState(seed)
    E = rng.randn(N, M)
    H = rng.randn(M, p) / np.sqrt(M)
    nT = p // 4
    R, R12 = oa.obs_error_model(nT, 4)
    Eo = np.tanh(E @ H)
    return dict(prior_ens=E, obs_ens=Eo, obs=np.tanh(rng.randn(M) @ H), perturbs=rng.randn(N, p) @ R12.T,
                decorr=np.linalg.inv(R12.T)), R12, H


@pytest.mark.parametrize("N,M,p", [(40, 400, 160), (200, 400, 160), (64, 1000, 48), (33, 257, 20)])
def test_es_les_notebook_sizes_vs_oracle(N, M, p):
    from historymatching_b200 import analysis as ha

    kw, _, _ = _hm_case(N, M, p, seed=N)
    np.testing.assert_allclose(ha.ens_update0(**kw), oa.ens_update0(**kw), **TOL)
    rng = np.random.RandomState(1)
    xy_prm = rng.rand(M, 2) * [2, 1]
    xy_obs = np.tile(rng.rand(4, 2) * [2, 1], (p // 4, 1))
    taper = oa.bump(oa.pairwise_distances(xy_prm, xy_obs) / 1.2)
    taper_d = ha.bump_taper(xy_prm, xy_obs, 1.2)
    np.testing.assert_allclose(taper_d, taper, rtol=1e-10, atol=1e-14)  # 1/(1-x^2) is ill-conditioned near the edge
    np.testing.assert_allclose(ha.ens_update0_loc(taper=taper, **kw), oa.ens_update0_loc(taper=taper, **kw), **TOL)


def test_les_edge_cases():
    from historymatching_b200 import analysis as ha

    kw, _, _ = _hm_case(20, 50, 16, seed=4)
    # no active observation anywhere: posterior == prior, bit for bit
    np.testing.assert_array_equal(ha.ens_update0_loc(taper=np.zeros((50, 16)), **kw), kw["prior_ens"])
    # taper just below / above the 1e-2 cut on sqrt(taper) (HistoryMatch.py:786)
    t = np.zeros((50, 16))
    t[:, 3] = (1e-2) ** 2 * 0.99
    t[:, 5] = (1e-2) ** 2 * 1.01
    np.testing.assert_allclose(ha.ens_update0_loc(taper=t, **kw), oa.ens_update0_loc(taper=t, **kw), **TOL)


def test_es_mda_vs_oracle_and_single_pass_identity():
    from historymatching_b200 import analysis as ha

    kw, R12, H = _hm_case(30, 120, 24, seed=9)

    def fwd(X):
        return np.tanh(X @ H)

    rng = np.random.RandomState(3)
    Z = [rng.randn(30, 24) for _ in range(4)]
    E_gpu, st = ha.es_mda(kw["prior_ens"], fwd, kw["obs"], R12, [4, 4, 4, 4], perturbs=Z)
    E_ref, _ = oa.es_mda(kw["prior_ens"], fwd, kw["obs"], R12, [4, 4, 4, 4], perturbs=Z)
    np.testing.assert_allclose(E_gpu, E_ref, **TOL)
    assert len(st.E) == 4 and len(st.Eo) == 4
    one, _ = ha.es_mda(kw["prior_ens"], fwd, kw["obs"], R12, [1.0], perturbs=Z[:1])
    es = ha.ens_update0(kw["prior_ens"], fwd(kw["prior_ens"]), kw["obs"], Z[0] @ R12.T, np.linalg.inv(R12.T))
    np.testing.assert_allclose(one, es, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("m,n,k,tA,tB", [(160, 300, 40, True, False), (40, 400, 160, False, False),
                                         (129, 130, 17, False, True), (1, 160, 160, False, False),
                                         (257, 65, 1000, True, True), (64, 160, 33, False, False)])
def test_dgemm_dmma(m, n, k, tA, tB):
    import ctypes as C

    import torch

    from historymatching_b200 import _lib

    rng = np.random.RandomState(m + n + k)
    A = rng.randn(*((k, m) if tA else (m, k)))
    B = rng.randn(*((n, k) if tB else (k, n)))
    Cm = rng.randn(m, n)
    ref = 0.7 * (A.T if tA else A) @ (B.T if tB else B) - 0.3 * Cm
    dA, dB, dC = (torch.as_tensor(x, device="cuda") for x in (A, B, Cm))
    ctx = _lib.Context.get(0)
    ctx.use_torch_stream()
    _lib.check(ctx.lib.hm_dgemm(ctx.handle, int(tA), int(tB), m, n, k, 0.7, C.c_void_p(dA.data_ptr()), A.shape[1],
                                C.c_void_p(dB.data_ptr()), B.shape[1], -0.3, C.c_void_p(dC.data_ptr()), n))
    np.testing.assert_allclose(dC.cpu().numpy(), ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("m,n,k,tA,tB,off", [(131, 77, 45, False, False, 0), (131, 77, 45, True, True, 0),
                                             (33, 129, 19, True, False, 0), (70, 35, 35, False, True, 0),
                                             (64, 64, 32, False, False, 1), (1, 3, 1, True, True, 0),
                                             (40, 50, 0, False, False, 0)])
def test_dgemm_submatrix_views(m, n, k, tA, tB, off):
    """Operands that are windows of larger row-major arrays: even leading dimensions with odd extents (the last
    16-byte copy of a row / column is half outside the matrix and must be zero-filled), a base that is only 8-byte
    aligned (`off` = 1: the 8-byte copy path), and k = 0 (C = beta C)."""
    import ctypes as C

    import torch

    from historymatching_b200 import _lib

    rng = np.random.RandomState(7 * m + n + k)
    ra, ca = (k, m) if tA else (m, k)
    rb, cb = (n, k) if tB else (k, n)
    bigA, bigB = rng.randn(ra + 3, ca + 6 + ca % 2), rng.randn(rb + 2, cb + 4 + cb % 2)  # even leading dimensions
    assert bigA.shape[1] % 2 == 0 and bigB.shape[1] % 2 == 0
    A, B = bigA[1:1 + ra, off:off + ca], bigB[2:2 + rb, off:off + cb]
    Cm = rng.randn(m, n)
    ref = 0.7 * (A.T if tA else A) @ (B.T if tB else B) - 0.3 * Cm
    dA, dB, dC = (torch.as_tensor(x, device="cuda") for x in (bigA, bigB, Cm))
    ctx = _lib.Context.get(0)
    ctx.use_torch_stream()
    pa = dA.data_ptr() + 8 * (bigA.shape[1] + off)
    pb = dB.data_ptr() + 8 * (2 * bigB.shape[1] + off)
    _lib.check(ctx.lib.hm_dgemm(ctx.handle, int(tA), int(tB), m, n, k, 0.7, C.c_void_p(pa), bigA.shape[1],
                                C.c_void_p(pb), bigB.shape[1], -0.3, C.c_void_p(dC.data_ptr()), n))
    np.testing.assert_allclose(dC.cpu().numpy(), ref, rtol=1e-12, atol=1e-12)


def test_es_update_large_linear_gaussian_property():
    """BASELINE config-C shape (N=1024, M=16384, p=160): linearity in the innovations and
    agreement with the reference-order formula on a column subset."""
    import torch

    from historymatching_b200 import analysis as ha

    N, M, p = 1024, 16384, 160
    gen = torch.Generator(device="cuda").manual_seed(0)
    E = torch.randn(N, M, dtype=torch.float64, device="cuda", generator=gen)
    Eo = torch.randn(N, p, dtype=torch.float64, device="cuda", generator=gen)
    pert = 0.1 * torch.randn(N, p, dtype=torch.float64, device="cuda", generator=gen)
    obs = torch.randn(p, dtype=torch.float64, device="cuda", generator=gen)
    R, R12 = oa.obs_error_model(40, 4)
    dec = np.linalg.inv(R12.T)
    post = ha.ens_update0(E, Eo, obs, pert, dec)
    cols = slice(100, 164)
    ref = oa.ens_update0(E[:, cols].cpu().numpy(), Eo.cpu().numpy(), obs.cpu().numpy(), pert.cpu().numpy(), dec)
    np.testing.assert_allclose(post[:, cols].cpu().numpy(), ref, **TOL)
    # the update is affine in obs: post(obs1) - post(obs0) is the same for every perturbation set
    d1 = ha.ens_update0(E, Eo, obs + 1.0, pert, dec) - post
    d2 = ha.ens_update0(E, Eo, obs + 1.0, 0 * pert, dec) - ha.ens_update0(E, Eo, obs, 0 * pert, dec)
    assert float((d1 - d2).abs().max()) < 1e-9


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_iles_against_reference_golden(golden, tag):
    """ILES (HistoryMatch.py:1007-1064): batched per-parameter Gauss-Newton steps."""
    from historymatching_b200 import analysis as ha

    g = golden("updates.npz")
    H = g[f"{tag}_H"]

    def fwd(X):
        return np.tanh(X @ H) + 0.1 * (X @ H)

    E, st = ha.ILES(obs_ens=fwd, taper=g[f"{tag}_taper"], xStep=0.6, iMax=3, **_case(g, tag))
    np.testing.assert_allclose(E, g[f"{tag}_ILES"], **TOL)
    np.testing.assert_allclose(np.array(st.E), g[f"{tag}_ILES_E"], **TOL)


def test_iles_notebook_self_check_and_size(golden):
    """HistoryMatch.py:1069-1071: ILES(taper=eye) reproduces the local ES; notebook size N=40, M=400, p=160."""
    from historymatching_b200 import analysis as ha

    g = golden("gauss_gauss.npz")
    kw = {k: g[k][:48] if k in ("prior_ens", "perturbs") else g[k] for k in ("prior_ens", "obs", "perturbs", "decorr")}
    ref = oa.ens_update0_loc(obs_ens=kw["prior_ens"], taper=np.eye(3), **kw)
    out, _ = ha.ILES(obs_ens=lambda x: x, taper=np.eye(3), **kw)
    np.testing.assert_allclose(out, ref, rtol=1e-7, atol=1e-9)

    kw, _, H = _hm_case(40, 400, 160, seed=2)
    kw.pop("obs_ens")
    rng = np.random.RandomState(1)
    xy_prm = rng.rand(400, 2) * [2, 1]
    xy_obs = np.tile(rng.rand(4, 2) * [2, 1], (40, 1))
    taper = oa.bump(oa.pairwise_distances(xy_prm, xy_obs) / 1.2)

    def fwd(X):
        return np.tanh(X @ H)

    E_gpu, _ = ha.ILES(obs_ens=fwd, taper=taper, xStep=0.4, iMax=2, **kw)
    E_ref, _ = oa.ILES(obs_ens=fwd, taper=taper, xStep=0.4, iMax=2, **kw)
    np.testing.assert_allclose(E_gpu, E_ref, **TOL)


@pytest.mark.parametrize("N,M,q", [(40, 400, 1), (200, 400, 1), (64, 1000, 5), (33, 257, 19), (2, 7, 1)])
def test_cov_corr_fields_vs_oracle(N, M, q):
    """hm_corr against the oracle's utils.cov / utils.corr (tools/utils.py:31-55); b 1-D as in the notebook
    (HistoryMatch.py:741, 831) and 2-D (cov for every q; corr for q > 1 with the (M,q) broadcasting)."""
    import torch

    from historymatching_b200 import analysis as ha
    from historymatching_b200.dropin.tools import utils as dutils

    rng = np.random.RandomState(N + M + q)
    a = rng.randn(N, M) * (1 + 10 * rng.rand(M)) + 5 * rng.randn(M)
    b = a[:, : max(q, 1)] @ rng.randn(max(q, 1), q) + rng.randn(N, q)
    if q == 1:
        b1 = b[:, 0]
        np.testing.assert_allclose(ha.cov(a, b1), oa.cov(a, b1), rtol=1e-11, atol=1e-12)
        got = ha.corr(a, b1)
        assert got.shape == (M,)
        np.testing.assert_allclose(got, oa.corr(a, b1), rtol=1e-11, atol=1e-13)
        # device-resident inputs through the drop-in tools.utils
        got_t = dutils.corr(torch.as_tensor(a, device="cuda"), torch.as_tensor(b1, device="cuda"))
        assert got_t.is_cuda
        np.testing.assert_array_equal(got_t.cpu().numpy(), got)
        assert int(got_t.argmax()) == int(np.argmax(oa.corr(a, b1)))  # xy_max_corr, HistoryMatch.py:832
    else:
        np.testing.assert_allclose(ha.cov(a, b), oa.cov(a, b), rtol=1e-11, atol=1e-12)
        ref = (oa.cov(a, b) / np.std(a, axis=0, ddof=1)[:, None] / np.std(b, axis=0, ddof=1)[None, :]).clip(-999, 999)
        np.testing.assert_allclose(ha.corr(a, b), ref, rtol=1e-11, atol=1e-13)


def test_corr_ill_defined_column():
    """A constant column of a has zero spread: 0/0 -> NaN as in numpy (np.clip keeps NaN)."""
    from historymatching_b200 import analysis as ha

    rng = np.random.RandomState(0)
    a = rng.randn(10, 6)
    a[:, 2] = 1.5
    b = rng.randn(10)
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = oa.corr(a, b)
    got = ha.corr(a, b)
    assert np.isnan(got[2]) and np.isnan(ref[2])
    np.testing.assert_allclose(np.delete(got, 2), np.delete(ref, 2), rtol=1e-11)


def test_cov_corr_against_reference_golden(golden):
    """hm_corr against the outputs of the reference's own utils.cov / utils.corr (tests/golden/primitives.npz)."""
    from historymatching_b200 import analysis as ha

    g = golden("primitives.npz")
    np.testing.assert_allclose(ha.cov(g["E"], g["b"]), g["cov"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(ha.corr(g["E"], g["b"][:, 0]), g["corr"], rtol=1e-11, atol=1e-13)


# ---- device entry points that round 1 exercised only indirectly ------------------------------------------------
@pytest.mark.parametrize("rescale", [False, True])
def test_center_device_vs_reference(golden, rescale):
    """analysis.center on the device (hm_center), with and without the sqrt(N/(N-1)) rescaling (tools/utils.py:10-28)."""
    import torch

    from historymatching_b200 import analysis as ha

    rng = np.random.RandomState(3)
    E = rng.randn(37, 3, 50) * 3 + rng.randn(3, 50)
    X, x = ha.center(E, rescale=rescale)
    Xr, xr = oa.center(E, rescale=rescale)
    np.testing.assert_allclose(X, Xr, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(x, xr, rtol=1e-13, atol=1e-13)
    Xd, xd = ha.center(torch.as_tensor(E[:, 0], device="cuda"), rescale=rescale)   # CUDA tensors in -> CUDA tensors out
    assert Xd.is_cuda and xd.is_cuda
    Xr, xr = oa.center(E[:, 0], rescale=rescale)
    np.testing.assert_allclose(Xd.cpu().numpy(), Xr, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(xd.cpu().numpy(), xr, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("sharpness", [1, 10, 0.5])
def test_bump_taper_sharpness(sharpness):
    """hm_taper_bump with sharpness != 1 (loc.bump(distances, sharpness), tools/localization.py:86-92)."""
    from historymatching_b200 import analysis as ha

    rng = np.random.RandomState(5)
    xy_prm = rng.rand(300, 2) * [2, 1]
    xy_obs = rng.rand(12, 2) * [2, 1]
    ref = oa.bump(oa.pairwise_distances(xy_prm, xy_obs) / 0.7, sharpness)
    np.testing.assert_allclose(ha.bump_taper(xy_prm, xy_obs, 0.7, sharpness), ref, rtol=1e-10, atol=1e-14)


def test_es_update_host_entry_point():
    """hm_es_update_host: the call a non-CUDA host makes (host pointers in, posterior written in place)."""
    import ctypes as C

    from historymatching_b200 import _lib

    kw, _, _ = _hm_case(24, 130, 16, seed=2)
    E = np.ascontiguousarray(kw["prior_ens"]).copy()
    ctx = _lib.Context.get()
    ptr = lambda a: C.c_void_p(np.ascontiguousarray(a).ctypes.data)  # noqa: E731
    Eo, obs, pert, dec = (np.ascontiguousarray(kw[k]) for k in ("obs_ens", "obs", "perturbs", "decorr"))
    _lib.check(ctx.lib.hm_es_update_host(ctx.handle, 24, 130, 16, C.c_void_p(E.ctypes.data), ptr(Eo), ptr(obs), ptr(pert),
                                         ptr(dec)))
    np.testing.assert_allclose(E, oa.ens_update0(**kw), **TOL)


def test_separable_prior_on_device():
    """gaussian_fields_separable(device=...): the two products Fx Z Fy^T through hm_dgemm equal the numpy einsum on the
    same draw; the fields have the prescribed covariance structure (unit variance, correlation exp(-3 d^2 / r^2))."""
    import torch

    from historymatching_b200.dropin.tools import geostat
    from historymatching_b200.sim import GridSpec

    grid = GridSpec(48, 40, 2.0, 1.0)
    r = 0.8
    hx, hy = grid.Lx / grid.Nx, grid.Ly / grid.Ny
    Fx = geostat._factor_1d((np.arange(grid.Nx) + 0.5) * hx, r)
    Fy = geostat._factor_1d((np.arange(grid.Ny) + 0.5) * hy, r)
    Z = torch.randn((7, grid.Nx, grid.Ny), dtype=torch.float64, device="cuda")
    out = geostat.separable_apply(Fx, Z, Fy).cpu().numpy()
    np.testing.assert_allclose(out, np.einsum("ia,nab,jb->nij", Fx, Z.cpu().numpy(), Fy), rtol=1e-12, atol=1e-12)
    F = geostat.gaussian_fields_separable(grid, 4000, r=r, rng=np.random.RandomState(3), device="cuda")
    assert F.is_cuda and F.shape == (4000, grid.M)
    F = F.cpu().numpy().reshape(4000, grid.Nx, grid.Ny)
    assert abs(F.var(0).mean() - 1) < 0.05
    c = np.mean(F[:, 10, 10] * F[:, 16, 10])   # 6 cells apart in x
    assert abs(c - np.exp(-3 * (6 * hx) ** 2 / r**2)) < 0.06


# ---- sizes beyond the shared-memory workspaces (global-memory workspace path of the same kernels) ----------------------
def test_les_more_observations_than_shared_memory_holds():
    """ens_update0_loc with p = 240 > 165 active observations per parameter (a p x p tapered system does not fit 227 KB):
    the local-analysis kernel runs from its global-memory workspace; same result as the reference algorithm."""
    from historymatching_b200 import analysis as ha

    kw, _, _ = _hm_case(40, 150, 240, seed=21)
    rng = np.random.RandomState(2)
    xy_prm = rng.rand(150, 2) * [2, 1]
    xy_obs = np.tile(rng.rand(4, 2) * [2, 1], (60, 1))
    taper = oa.bump(oa.pairwise_distances(xy_prm, xy_obs) / 2.5)      # wide taper: (almost) every observation active
    assert (np.sqrt(taper) > 1e-2).sum(1).max() > 200
    np.testing.assert_allclose(ha.ens_update0_loc(taper=taper, **kw), oa.ens_update0_loc(taper=taper, **kw), **TOL)


def test_iles_notebook_ensemble_of_200():
    """ILES at N = 200 members, p = 160 (BASELINE config 1 speaks of a ~200-member ensemble; SURVEY 8 size A "also N=200"):
    3 N^2 + N p doubles = 1.2 MB per parameter, far beyond shared memory - global-memory workspace path."""
    from historymatching_b200 import analysis as ha

    N, M, p = 200, 60, 160
    kw, _, H = _hm_case(N, M, p, seed=5)
    kw.pop("obs_ens")
    rng = np.random.RandomState(4)
    xy_prm = rng.rand(M, 2) * [2, 1]
    xy_obs = np.tile(rng.rand(4, 2) * [2, 1], (p // 4, 1))
    taper = oa.bump(oa.pairwise_distances(xy_prm, xy_obs) / 1.2)

    def fwd(X):
        return np.tanh(X @ H)

    E, st = ha.ILES(obs_ens=fwd, taper=taper, xStep=0.4, iMax=2, **kw)
    E_ref, st_ref = oa.ILES(obs_ens=fwd, taper=taper, xStep=0.4, iMax=2, **kw)
    np.testing.assert_allclose(E, E_ref, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(np.array(st.Eo), np.array(st_ref.Eo), rtol=1e-7, atol=1e-9)
