"""EnOpt objective batches without the per-member host round trip.

The notebook's ``npv`` (``Optimise.py:112-125``) re-configures a deep copy of the model per control vector, runs it
and turns the saturations at the producers plus ``model.actual_rates`` into discounted cash flow (``accounting``,
``prd_sats``: ``Optimise.py:170-210``).  Through the drop-in ``tools.utils.apply`` those calls already land on the GPU
as one batch, but every member still pulls its full ``(nTime+1, Nxy)`` saturation history to the host, of which the
accounting reads ``nPrd`` cells per step.

``EnsembleNPV`` is the explicit fast path: ONE ``hm_sim_batch`` for the whole control / uncertainty batch with
per-member wells, rates and permeability, observation gather on the device at the union of the members' producer
cells, and the ledger of ``accounting`` evaluated for all members at once on the device; only the ``(N,)`` values
(and, if asked, the ledgers) come back.  Members with invalid parameters (well outside the domain, unbalanced
rates, non-positive permeability) get value 0 like the notebook's ``try / except`` (``Optimise.py:119-124,
548-555``) without disturbing the rest of the batch.  ``robust_increments`` is the batched form of the notebook's
``ens_eval_duplex`` (StoSAG / Paired / Mean-model, ``Optimise.py:833-853``).
"""

from __future__ import annotations

import copy

import numpy as np

from .sim import GridSpec, run_ensemble

LEDGER_KEYS = ("oil", "inj", "wat", "pwell", "iwell", "turbo", "diffs")


class EnsembleNPV:
    """Batched ``npv(model, **params)[0]`` for the drop-in ``ResSim`` ``model``.

    ``price``      the notebook's price table (``Optimise.py:150-158``): keys ``inj, oil, wat, turbo, diffs, /well``;
    ``discounts``  ``(nTime,)`` discount factors (``Optimise.py:159``);
    ``rate0``      nominal total rate of the "turbo" penalty (``Optimise.py:193``).
    """

    def __init__(self, model, dt, nTime, wsat0, price, discounts, rate0, device="cuda"):
        self.model, self.dt, self.nTime = model, float(dt), int(nTime)
        self.wsat0 = np.asarray(wsat0, float).reshape(-1)
        self.price, self.discounts, self.rate0 = dict(price), np.asarray(discounts, float), float(rate0)
        self.device = device
        self.grid = GridSpec(model.Nx, model.Ny, float(model.Lx), float(model.Ly), model.vw, model.vo, model.swc, model.sor)

    # ---- host side: the members' configurations (cheap; the notebook's remake + the drop-in's validation) ------------
    def _configure(self, params_list):
        cells, rates, Ks, valid = [], [], [], []
        nW = None
        for params in params_list:
            try:
                m = copy.copy(self.model)              # shallow: setattr replaces whole arrays, as remake() does on a deepcopy
                for k, v in params.items():
                    setattr(m, k, v)                   # validates and collocates the wells; raises outside the domain
                q, c = m._schedule(self.nTime)         # (nTime, nW) signed rates; raises when unbalanced
                if nW is not None and len(c) != nW:
                    raise ValueError("members of one batch must have the same number of wells")
                nW = len(c)
                cells.append(c), rates.append(q), Ks.append(m.K.reshape(2, -1)), valid.append((m.nInj, True))
            except Exception:  # noqa: BLE001 - invalid parameters are penalised, not fatal (Optimise.py:119-124)
                cells.append(None), rates.append(None), Ks.append(None), valid.append((0, False))
        return cells, rates, Ks, valid

    def __call__(self, params_list, ledgers=False):
        """Values ``(N,)`` of the batch (numpy); with ``ledgers=True`` also the ``(N, 7)`` ledger in ``LEDGER_KEYS`` order."""
        import torch

        N = len(params_list)
        cells, rates, Ks, valid = self._configure(params_list)
        ok = [i for i in range(N) if valid[i][1]]
        values = np.zeros(N)
        ledger = np.zeros((N, len(LEDGER_KEYS)))
        if not ok:
            return (values, ledger) if ledgers else values
        nInj = valid[ok[0]][0]
        if any(valid[i][0] != nInj for i in ok):
            raise ValueError("members of one batch must have the same number of injectors")
        dev = self.device
        wc = torch.as_tensor(np.stack([cells[i] for i in ok]), dtype=torch.int32, device=dev)          # (n, nW)
        wr = torch.as_tensor(np.stack([rates[i] for i in ok]), dtype=torch.float64, device=dev)        # (n, nT, nW)
        K = np.stack([Ks[i] for i in ok])
        K = torch.as_tensor(K[:, 0] if all(np.array_equal(k[0], k[1]) for k in K) else K, dtype=torch.float64, device=dev)
        if K.shape[0] > 1 and bool((K == K[:1]).all()):
            K = K[:1]
        S0 = torch.as_tensor(self.wsat0, device=dev)
        # observations: the union of the members' producer cells, gathered on the device every step
        prd_cells = wc[:, nInj:]
        uniq, inv = torch.unique(prd_cells.reshape(-1), return_inverse=True)
        res = run_ensemble(self.grid, K, wc, wr, S0, self.dt, self.nTime, obs_cell=uniq.to(torch.int32), n_members=len(ok))
        inv = inv.reshape(prd_cells.shape)                                                              # (n, nPrd)
        s_end = torch.gather(res.obs, 2, inv[:, None, :].expand(-1, self.nTime, -1))                   # (n, nT, nPrd)
        s_start = torch.cat([S0[prd_cells.long()][:, None, :], s_end[:, :-1]], 1)
        prd_wsat = 0.5 * (s_start + s_end)                                                              # prd_sats: trapezoid
        # accounting (Optimise.py:170-200); rates as the simulator used them
        inj_r, prd_r = wr[:, :, :nInj], -wr[:, :, nInj:]                                                # (n, nT, n*)
        disc = torch.as_tensor(self.discounts, device=dev)
        dt, price = self.dt, self.price
        inj_total = (dt * inj_r).sum(2) @ disc
        oil_total = (dt * prd_r * (1 - prd_wsat)).sum(2) @ disc
        wat_total = (dt * prd_r * prd_wsat).sum(2) @ disc
        excess = (prd_r.sum(2) - self.rate0).clamp(min=0)
        diffs = inj_r[:, 1:] - inj_r[:, :-1]
        led = torch.stack([
            price["oil"] * oil_total, -price["inj"] * inj_total, -price["wat"] * wat_total,
            -price["/well"] * (prd_r != 0).sum((1, 2)).double(), -price["/well"] * (inj_r != 0).sum((1, 2)).double(),
            -price["turbo"] * excess.sum(1) ** 2 * dt, -price["diffs"] * (diffs.abs() ** 0.1).sum((1, 2))], 1)
        status = res.status.cpu().numpy()
        led = led.cpu().numpy()
        for j, i in enumerate(ok):
            if status[j] == 0:                          # a failed member is penalised like an exception in npv()
                ledger[i] = led[j]
                values[i] = led[j].sum()
        return (values, ledger) if ledgers else values


def robust_increments(npv_batch, robustly, u, U, X, param_u="inj_xy", param_x="K"):
    """Batched ``ens_eval_duplex`` (``Optimise.py:833-853``): objective increments of the control ensemble ``U`` under the
    uncertainty ensemble ``X`` with ONE forward batch per ``apply`` of the notebook.

    ``Paired``: member ``i`` pairs control ``U[i]`` with uncertainty ``X[i]``; ``StoSAG``: the same minus the value of
    the unperturbed control ``u`` under ``X[i]`` (both halves in one batch of ``2 nEns`` members); ``Mean-model`` /
    ``Fragile``: every control under the mean of ``X``.
    """
    U, X = np.asarray(U, float), np.asarray(X, float)
    n = len(U)
    if robustly == "Paired":
        return npv_batch([{param_u: U[i], param_x: X[i]} for i in range(n)])
    if robustly == "StoSAG":
        both = npv_batch([{param_u: U[i], param_x: X[i]} for i in range(n)] +
                         [{param_u: np.asarray(u, float), param_x: X[i]} for i in range(n)])
        return both[:n] - both[n:]
    if robustly in ("Mean-model", "Fragile"):
        x1 = X.mean(0)
        return npv_batch([{param_u: U[i], param_x: x1} for i in range(n)])
    raise ValueError(f"unknown robust treatment {robustly!r}")
