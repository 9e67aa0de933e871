"""EnOpt-style batches (Optimise.py:112-210, 428-466): every member re-configures the wells / rates
of a deep copy of the model and runs it; the batch lands on the GPU as ONE ensemble launch with
per-member well lists, invalid members are penalised without disturbing the rest."""

import copy

import numpy as np
import pytest

from oracle import ressim as orr

pytestmark = pytest.mark.gpu


def test_npv_batches_through_collector():
    import historymatching_b200 as hmb

    hmb.activate()
    import TPFA_ResSim as simulator
    from tools import utils
    from tools.utils import apply

    rng = np.random.RandomState(23)
    model = simulator.ResSim(Nx=16, Ny=16, Lx=2, Ly=1, name="Base model")
    K = 0.1 + np.exp(1.5 * rng.randn(1, 256))
    model.K = K
    near01 = np.array([0.12, 0.87])
    rate0 = 1.5
    model.inj_xy = [[model.Lx / 2, model.Ly / 2]]
    model.prd_xy = [[x, y] for y in model.Ly * near01 for x in model.Lx * near01]
    model.inj_rates = rate0 * np.ones((1, 1))
    model.prd_rates = rate0 * np.ones((4, 1)) / 4
    wsat0 = np.zeros(model.Nxy)
    dt, nTime = 0.025, 6

    def remake(model, **params):
        model = copy.deepcopy(model)
        for k, v in params.items():
            setattr(model, k, v)
        return model

    def npv(model, **params):
        try:
            model = remake(model, **params)
            wsats = model.sim(dt, nTime, wsat0, pbar=False)
            s = wsats[:, model.xy2ind(*model.prd_xy.T)]
            prd_sat = ((s[:-1] + s[1:]) / 2).T
            oil = dt * model.actual_rates["prd"] * (1 - prd_sat)
            value = 100 * oil.sum() - 20 * dt * model.actual_rates["inj"].sum()
        except Exception:
            value = 0
        return value

    def obj(xys):
        return npv(model, inj_xy=xys)

    U = np.array([[1.0, 0.5], [0.3, 0.2], [1.7, 0.9], [2.6, 0.5], [0.9, 0.1], [-0.1, 0.3]])
    utils.nCPU = "auto"
    batched = apply(obj, U, pbar=False)
    utils.nCPU = 1
    serial = apply(obj, U, pbar=False)
    assert batched[3] == 0 and batched[5] == 0  # outside the domain: penalised, not fatal (Optimise.py:548-555)
    np.testing.assert_allclose(batched, serial, rtol=1e-10)

    # oracle for one valid member, incl. time-dependent rates (Optimise.py:745-767)
    sched = rate0 * (0.5 + rng.rand(1, nTime))
    m2 = remake(model, inj_xy=[0.3, 0.2], inj_rates=sched, prd_rates=np.tile(sched / 4, (4, 1)))
    got = m2.sim(dt, nTime, wsat0, pbar=False)
    om = orr.OracleResSim(16, 16, 2.0, 1.0)
    om.K = model.K
    om.inj_xy, om.prd_xy = m2.inj_xy, m2.prd_xy
    om.inj_rates, om.prd_rates = m2.inj_rates, m2.prd_rates
    np.testing.assert_allclose(got, om.sim(dt, nTime, wsat0), rtol=0, atol=1e-8)
    assert m2.actual_rates["inj"].shape == (1, nTime) and m2.actual_rates["prd"].shape == (4, nTime)


def test_enopt_gd_on_the_gpu_path():
    """BASELINE config 5 in miniature: tools.enopt.GD maximises an NPV-like objective over the injector position;
    the control ensembles of the gradient estimate and the trial steps of the line search run as batched forward
    runs (utils.nCPU = "auto").  Batched and serial (one member per launch) runs follow the same trajectory."""
    import historymatching_b200 as hmb

    hmb.activate()
    import TPFA_ResSim as simulator
    from tools import enopt, utils

    model = simulator.ResSim(Nx=16, Ny=16, Lx=2, Ly=1)
    model.K = 0.1 + np.exp(1.2 * np.random.RandomState(5).randn(1, 256))
    near01 = np.array([0.12, 0.87])
    model.inj_xy = [[0.5, 0.3]]
    model.prd_xy = [[x, y] for y in model.Ly * near01 for x in model.Lx * near01]
    model.inj_rates = np.ones((1, 1))
    model.prd_rates = np.ones((4, 1)) / 4
    dt, nTime = 0.025, 5

    def obj(xy):
        try:
            m = copy.deepcopy(model)
            m.inj_xy = xy
            wsats = m.sim(dt, nTime, np.zeros(m.Nxy), pbar=False)
            s = wsats[:, m.xy2ind(*m.prd_xy.T)]
            return float(100 * (dt * m.actual_rates["prd"] * (1 - ((s[:-1] + s[1:]) / 2).T)).sum()
                         + 30 * wsats[-1].mean())   # produced oil + sweep
        except Exception:
            return 0.0

    runs = {}
    for mode in ("auto", 1):
        utils.nCPU = mode
        np.random.seed(11)
        path, objs, info = enopt.GD(obj, np.array([0.5, 0.3]), enopt.nabla_ens(0.05, nEns=8),
                                    enopt.backtracker(xSteps=(0.2, 0.1, 0.05), nCPU=3), nIter=3, quiet=True)
        runs[mode] = (np.asarray(path, float), np.asarray(objs, float))
    utils.nCPU = 1
    path, objs = runs["auto"]
    assert len(objs) >= 2 and np.all(np.diff(objs) > 0)          # every accepted step improves the objective
    np.testing.assert_allclose(runs[1][1], objs, rtol=1e-9)
    np.testing.assert_allclose(runs[1][0], path, rtol=1e-9, atol=1e-12)
