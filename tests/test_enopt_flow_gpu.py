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


def _optimise_cells(simulator, Nx=16, dt=0.025, nTime=6):
    """The notebook's cells (Optimise.py:64-210), verbatim apart from the grid size."""
    rng = np.random.RandomState(3)
    model = simulator.ResSim(Nx=Nx, Ny=Nx, Lx=2, Ly=1, name="Base model")
    model.K = 0.1 + np.exp(1.2 * rng.randn(1, model.Nxy))
    near01 = np.array([0.12, 0.87])
    rate0 = 1.5
    model.inj_xy = [[model.Lx / 2, model.Ly / 2]]
    model.prd_xy = [[x, y] for y in model.Ly * near01 for x in model.Lx * near01]
    model.inj_rates = rate0 * np.ones((1, 1))
    model.prd_rates = rate0 * np.ones((4, 1)) / 4
    wsat0 = np.zeros(model.Nxy)
    OneYear = 0.1
    price = {"inj": 20, "oil": 100, "turbo": 1, "wat": 6, "diffs": 1, "fixed": 0.8 * dt / OneYear, "/well": 0.3 * dt / OneYear}
    discounts = 0.96 ** (dt / OneYear * np.arange(nTime))

    def remake(model, **params):
        model = copy.deepcopy(model)
        for k, v in params.items():
            setattr(model, k, v)
        return model

    def prd_sats(model, wsats):
        s = wsats[:, model.xy2ind(*model.prd_xy.T)]
        return (s[:-1] + s[+1:]) / 2

    def accounting(model, wsats):
        prd_wsats = prd_sats(model, wsats).T
        inj_rates = model.actual_rates["inj"]
        prd_rates = model.actual_rates["prd"]
        inj_volumes = dt * inj_rates * 1
        oil_volumes = dt * prd_rates * (1 - prd_wsats)
        wat_volumes = dt * prd_rates * prd_wsats
        inj_total = inj_volumes.sum(0) @ discounts
        oil_total = oil_volumes.sum(0) @ discounts
        wat_total = wat_volumes.sum(0) @ discounts
        values = {}
        values["oil"] = +price["oil"] * oil_total
        values["inj"] = -price["inj"] * inj_total
        values["wat"] = -price["wat"] * wat_total
        excess = (prd_rates.sum(0) - rate0).clip(0)
        diffs = np.diff(inj_rates, 1)
        values["pwell"] = -price["/well"] * np.sum(prd_rates != 0)
        values["iwell"] = -price["/well"] * np.sum(inj_rates != 0)
        values["turbo"] = -price["turbo"] * excess.sum() ** 2 * dt
        values["diffs"] = -price["diffs"] * (np.abs(diffs) ** 0.1).sum()
        return values

    def npv(model, **params):
        try:
            model = remake(model, **params)
            wsats = model.sim(dt, nTime, wsat0, pbar=False)
            ledgr = accounting(model, wsats)
            value = sum(ledgr.values())
            other = dict(model=model, wsats=wsats, ledgr=ledgr)
        except Exception:
            value = 0
            other = None
        return value, other

    return dict(model=model, npv=npv, wsat0=wsat0, price=price, discounts=discounts, rate0=rate0, dt=dt, nTime=nTime, rng=rng)


def test_robust_objective_duplex_batches_and_device_accounting():
    """Optimise.py:833-853 (ens_eval_duplex: StoSAG / Paired / Mean-model): members differ in the injector position AND in
    the permeability field.  (1) the notebook's own cells through the drop-in apply: the batched run equals the serial
    one; (2) the explicit fast path (EnsembleNPV: one hm_sim_batch, accounting on the device) reproduces the cells' npv,
    its ledger, the penalty for invalid members and the duplex increments."""
    import historymatching_b200 as hmb

    hmb.activate()
    import TPFA_ResSim as simulator
    from tools import utils
    from tools.enopt import nabla_ens
    from tools.utils import apply

    from historymatching_b200.enopt_fast import LEDGER_KEYS, EnsembleNPV, robust_increments

    c = _optimise_cells(simulator)
    model, npv, rng = c["model"], c["npv"], c["rng"]
    nEns = 7
    uq_ens = 0.1 + np.exp(1.2 * rng.randn(nEns, model.Nxy))          # uncertain permeability (Optimise.py:887-888)

    def obj1(u, x):
        return npv(model, inj_xy=u, K=x)[0]

    def ens_eval_duplex(self, obj, u, U, pbar):                      # the notebook's cell, verbatim
        if self.robustly == "Paired":
            dJ = apply(self.obj_ux, U, x=self.X, pbar=pbar)
        elif self.robustly == "StoSAG":
            uu = np.tile(u, (self.nEns, 1))
            JU = apply(self.obj_ux, U, x=self.X, pbar=pbar)
            Ju = apply(self.obj_ux, uu, x=self.X, pbar=pbar)
            dJ = np.asarray(JU) - Ju
        elif self.robustly in ["Mean-model", "Fragile"]:
            x1 = np.tile(self.X.mean(0), (self.nEns, 1))
            dJ = apply(self.obj_ux, U, x=x1, pbar=pbar)
        else:
            dJ = apply(obj, U, pbar=pbar)
        return dJ

    u = np.array([1.0, 0.5])
    U = u + 0.2 * rng.randn(nEns, 2)
    U[2] = [2.7, 0.4]                                               # outside the domain: value 0, not fatal
    fast = EnsembleNPV(model, c["dt"], c["nTime"], c["wsat0"], c["price"], c["discounts"], c["rate0"])
    for robustly in ("Paired", "StoSAG", "Mean-model"):
        nab = nabla_ens(0.1, nEns=nEns, robustly=robustly, obj_ux=obj1, X=uq_ens)
        utils.nCPU = "auto"
        batched = np.asarray(ens_eval_duplex(nab, None, u, U, False), float)
        utils.nCPU = 1
        serial = np.asarray(ens_eval_duplex(nab, None, u, U, False), float)
        np.testing.assert_allclose(batched, serial, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(robust_increments(fast, robustly, u, U, uq_ens), serial, rtol=1e-9, atol=1e-9)
        if robustly == "Paired":
            assert batched[2] == 0 and np.all(batched[[0, 1, 3]] != 0)
    # the ledger of one member, entry by entry, incl. time-dependent rates (Optimise.py:745-767)
    sched = c["rate0"] * (0.5 + rng.rand(1, c["nTime"]))
    params = dict(inj_xy=[0.3, 0.2], inj_rates=sched, prd_rates=np.tile(sched / 4, (4, 1)), K=uq_ens[1])
    value, other = npv(model, **params)
    vals, led = fast([params, dict(inj_xy=[-1.0, 0.2])], ledgers=True)
    assert vals[1] == 0 and not led[1].any()
    np.testing.assert_allclose(vals[0], value, rtol=1e-10)
    np.testing.assert_allclose(led[0], [other["ledgr"][k] for k in LEDGER_KEYS], rtol=1e-9, atol=1e-12)
