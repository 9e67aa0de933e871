"""The notebook's forward-model cells, written the way a user writes them
(HistoryMatch.py:97-225, 358-387, 635-652, 958-961), run unchanged on top of the
drop-in ``TPFA_ResSim`` / ``tools`` packages and compared with the oracle."""

import copy

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import analysis as oa
from oracle import ressim as orr

pytestmark = pytest.mark.gpu


def test_history_match_cells_run_unchanged():
    import historymatching_b200 as hmb

    hmb.activate()
    import TPFA_ResSim as simulator
    from tools import geostat, utils
    from tools.utils import apply

    from historymatching_b200 import analysis as ha

    np.random.seed(1)
    model = simulator.ResSim(Nx=20, Ny=20, Lx=2, Ly=1)

    def perm_transf(x):
        return 0.1 + np.exp(5 * x)

    def set_perm(model, log_perm_array):
        p = perm_transf(log_perm_array).reshape(model.shape)
        model.K = np.stack([p, p])

    truth = geostat.gaussian_fields(model.mesh, 1, r=0.8)
    set_perm(model, truth)
    near01 = np.array([0.12, 0.87])
    xy_4corners = [[x, y] for y in model.Ly * near01 for x in model.Lx * near01]
    nPrd = len(xy_4corners)
    model.prd_xy = xy_4corners
    model.inj_xy = [[model.Lx / 2, model.Ly / 2]]
    model.inj_rates = [[1]]
    model.prd_rates = np.ones((nPrd, 1)) / nPrd
    prod_inds = model.xy2ind(*model.prd_xy.T)

    def obs_model(water_sat):
        return water_sat[prod_inds]

    T, dt = 0.25, 0.025
    nTime = round(T / dt)
    wsat0 = np.zeros(model.Nxy)
    wsat_truth = model.sim(dt, nTime, wsat0)
    prod_truth = np.array([obs_model(x) for x in wsat_truth[1:]])
    assert wsat_truth.shape == (nTime + 1, 400)

    N = 12
    prior = np.clip(geostat.gaussian_fields(model.mesh, N, r=0.8), -2.2, 2.2)

    def comp1(perm, wsat0=wsat0):
        new_model = copy.deepcopy(model)
        set_perm(new_model, perm)
        wsats = new_model.sim(dt, nTime, wsat0, pbar=False)
        prods = np.array([obs_model(x) for x in wsats[1:]])
        return wsats, prods

    def forward_model(*args, leave=True, desc="Ens-run", **kwargs):
        output = apply(comp1, *args, pbar=dict(leave=leave, desc=desc, disable=True), **kwargs)
        return [np.asarray(y) for y in zip(*output)]

    utils.nCPU = "auto"
    wsat_prior, prod_prior = forward_model(prior)
    assert wsat_prior.shape == (N, nTime + 1, 400) and prod_prior.shape == (N, nTime, nPrd)

    # oracle
    om = orr.notebook_model(20, 20)
    ref = [orr.forward_member(om, x, dt, nTime, wsat0, prod_inds) for x in prior]
    np.testing.assert_allclose(wsat_prior, np.array([r[0] for r in ref]), rtol=0, atol=1e-8)
    np.testing.assert_allclose(prod_prior, np.array([r[1] for r in ref]), rtol=0, atol=1e-8)
    tr = orr.forward_member(om, np.clip(truth[0], -9, 9), dt, nTime, wsat0, prod_inds)
    np.testing.assert_allclose(wsat_truth, tr[0], rtol=0, atol=1e-7)

    # restart with per-member state (HistoryMatch.py:1227) == serial loop (nCPU=1)
    futr_batched = forward_model(prior, wsat_prior[:, -1])[1]
    utils.nCPU = 1
    futr_serial = forward_model(prior[:3], wsat_prior[:3, -1])[1]
    np.testing.assert_allclose(futr_batched[:3], futr_serial, rtol=0, atol=1e-12)

    # ES and IES on top (HistoryMatch.py:635-652, 958-961)
    def vect(x):
        *n, a, b = x.shape
        return x.reshape(n + [a * b])

    R, R12 = oa.obs_error_model(nTime, nPrd)
    noisy = (prod_truth + (R12 @ np.random.randn(nTime * nPrd)).reshape(nTime, nPrd)).clip(0, 1)
    setup0 = dict(obs_ens=vect(prod_prior), obs=vect(noisy), perturbs=np.random.randn(N, nPrd * nTime) @ R12.T,
                  decorr=sla.inv(R12.T))
    np.testing.assert_allclose(ha.ens_update0(prior, **setup0), oa.ens_update0(prior, **setup0), rtol=1e-8, atol=1e-10)
    utils.nCPU = "auto"
    setupI = dict(setup0, obs_ens=lambda x: vect(forward_model(x, leave=False)[1]))
    E_gpu, st = ha.IES(prior, **setupI, xStep=0.4, iMax=2)
    E_ref, st_ref = oa.IES(prior, **dict(setup0, obs_ens=lambda X: np.array(
        [orr.forward_member(om, x, dt, nTime, wsat0, prod_inds)[1].ravel() for x in X])), xStep=0.4, iMax=2)
    np.testing.assert_allclose(E_gpu, E_ref, rtol=1e-6, atol=1e-7)
    assert len(st.E) == 2 and st.Eo[0].shape == (N, nPrd * nTime)
    utils.nCPU = 1
