"""History-matching case: the notebook's problem setup as one object whose
``forward`` and ``es_update`` run on the GPU (single device or member-sharded
over the ranks of a ``torch.distributed`` job).

Mirrors the notebook cells: model / wells (``HistoryMatch.py:97,177-190``),
``perm_transf`` (``:137-138``), ``forward_model`` (``:383-387``), the obs-error
model ``R``, ``R12`` (``:243-259``) and ``hm_setup0`` (``:635-640``).
"""

from __future__ import annotations

import numpy as np

from . import analysis as ha
from .sim import GridSpec, run_ensemble


def notebook_wells(grid: GridSpec):
    """Injector at the centre, 4 producers near the corners, collocated with cells."""
    near01 = np.array([0.12, 0.87])
    xy = [[grid.Lx / 2, grid.Ly / 2]] + [[x, y] for y in grid.Ly * near01 for x in grid.Lx * near01]
    xy = np.array(xy)
    ix = np.minimum((xy[:, 0] / grid.Lx * grid.Nx).astype(int), grid.Nx - 1)
    iy = np.minimum((xy[:, 1] / grid.Ly * grid.Ny).astype(int), grid.Ny - 1)
    cells = (ix * grid.Ny + iy).astype(np.int32)
    rates = np.array([1.0, -0.25, -0.25, -0.25, -0.25])
    return cells, rates


def check_status(status):
    """Warn about members whose forward run did not complete cleanly (``hm_sim_desc.status`` bits): a pressure solve that
    hit the iteration limit, or a non-finite saturation.  The run is not aborted - an EnOpt batch may contain members
    outside the admissible domain (``Optimise.py:548-555``) - but it must not pass silently."""
    import warnings

    bad = status.nonzero()
    bad = bad[0] if isinstance(bad, tuple) else bad.reshape(-1)
    if len(bad):
        first = [int(i) for i in bad[:8]]
        warnings.warn(f"forward run: {len(bad)} member(s) with a non-zero status (first: {first}); bit 1 = pressure solve "
                      f"not converged, bit 2 = non-finite saturation", RuntimeWarning, stacklevel=3)


class HistoryMatchCase:
    def __init__(self, Nx=20, Ny=20, Lx=2.0, Ly=1.0, dt=0.025, nTime=40, device=None):
        import scipy.linalg as sla

        self.grid = GridSpec(Nx, Ny, Lx, Ly)
        self.dt, self.nTime = dt, nTime
        self.well_cell, self.well_rate = notebook_wells(self.grid)
        self.obs_cell = self.well_cell[1:].copy()
        self.nPrd = len(self.obs_cell)
        self.p = self.nPrd * nTime
        c = np.exp(-np.arange(nTime) / 2)
        c[c < 1e-2] = 0
        self.R = np.kron(1e-2 * sla.toeplitz(c), np.eye(self.nPrd))
        self.R12 = sla.cholesky(self.R, lower=True)
        self.decorr = sla.inv(self.R12.T)
        self.device = device
        self._dev_cache = {}

    @staticmethod
    def perm_transf(x):
        """``0.1 + exp(5 x)`` (``HistoryMatch.py:137-138``) for numpy arrays or torch tensors."""
        if type(x).__module__.startswith("torch"):
            return 0.1 + (5 * x).exp()
        return 0.1 + np.exp(5 * x)

    def _const(self, name, arr, like):
        """Small constant arrays, uploaded once per device."""
        import torch

        key = (name, like.device)
        if key not in self._dev_cache:
            self._dev_cache[key] = torch.as_tensor(arr, device=like.device)
        return self._dev_cache[key]

    def forward(self, logperm, S0=None, history=False, **kw):
        """``forward_model`` for the whole ensemble: ``(N,M)`` log-perm -> predicted data ``(N,p)``.

        Returns ``(obs (N, nTime*nPrd), SimResult)``; tensors stay on the device for CUDA input.
        """
        is_t = type(logperm).__module__.startswith("torch")
        K = logperm  # perm_transf (0.1 + exp(5 x)) is evaluated where the transmissibilities are built
        kw.setdefault("k_transform", (0.1, 5.0))
        if S0 is None:
            S0 = np.zeros(self.grid.M)
            if is_t:
                S0 = self._const("S0", S0, logperm)
        if is_t:
            wc = self._const("wc", self.well_cell, logperm)
            wr = self._const("wr", self.well_rate, logperm)
            oc = self._const("oc", self.obs_cell, logperm)
        else:
            wc, wr, oc = self.well_cell, self.well_rate, self.obs_cell
        res = run_ensemble(self.grid, K, wc, wr, S0, self.dt, self.nTime, obs_cell=oc, history=history,
                           n_members=K.shape[0], **kw)
        check_status(res.status)
        return res.obs.reshape(K.shape[0], self.p), res

    def es_update(self, E, Eo, obs, Z, alpha=1.0):
        """ES / one ES-MDA pass: perturbs = sqrt(alpha) Z R12^T, decorr / sqrt(alpha)."""
        is_t = type(E).__module__.startswith("torch")
        if is_t:
            R12T = self._const("R12T", self.R12.T.copy(), E)
            dec = self._const("decorr", self.decorr, E) / np.sqrt(alpha)
            pert = np.sqrt(alpha) * (Z @ R12T)
        else:
            dec = self.decorr / np.sqrt(alpha)
            pert = np.sqrt(alpha) * (Z @ self.R12.T)
        return ha.ens_update0(E, Eo, obs, pert, dec)
