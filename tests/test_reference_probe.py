"""Probe for the real simulator.

The reference's simulator is the third-party package ``TPFA_ResSim`` pinned at
``git+https://github.com/patnr/TPFA-ResSim.git@adc89536`` (``/root/reference/requirements.txt:1``); it is neither
vendored in the reference tree nor installed in the build container, so ``oracle/ressim.py`` restates the published
scheme and the simulator half of the parity claim is UNPINNED.  This test closes the gap the
moment the package is available: if a ``TPFA_ResSim`` that is not this repo's drop-in can be imported - from
site-packages or from ``baseline/_ref/`` (the path the repo reserves for a driver-provided install) - the oracle is
diffed against it on the notebook's own case (``HistoryMatch.py:97,177-190,219-225``) and must agree to 1e-10.
Otherwise the test is skipped and says why.
"""

import importlib
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_real_package():
    """A ``TPFA_ResSim`` outside this repo's drop-in directory, or ``(None, reason)``."""
    here = os.path.join(ROOT, "historymatching_b200", "dropin")
    candidates = [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(here)]
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref):
        candidates.insert(0, ref)
    for base in candidates:
        init = os.path.join(base or ".", "TPFA_ResSim", "__init__.py")
        if os.path.exists(init) and not os.path.abspath(init).startswith(os.path.abspath(here)):
            return base, None
    return None, ("TPFA_ResSim@adc89536 (requirements.txt:1 of the reference) is not installed: not in site-packages, "
                  "not under baseline/_ref/; there is no network to fetch it.  The simulator oracle stays pinned by "
                  "invariants and the Buckley-Leverett known answer only (parity unpinned).")


def test_oracle_against_real_tpfa_ressim_if_present():
    base, reason = _find_real_package()
    if base is None:
        pytest.skip(reason)
    saved = {k: v for k, v in sys.modules.items() if k == "TPFA_ResSim" or k.startswith("TPFA_ResSim.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, base)
    try:
        real = importlib.import_module("TPFA_ResSim")
        from oracle import ressim as orr

        for Nx, Ny, nT in ((20, 20, 40), (32, 24, 10)):
            om = orr.notebook_model(Nx, Ny)
            rm = real.ResSim(Nx=Nx, Ny=Ny, Lx=om.Lx, Ly=om.Ly)
            rng = np.random.RandomState(Nx)
            K = 0.1 + np.exp(2.0 * rng.randn(Nx, Ny))
            for mdl in (om, rm):
                mdl.K = np.stack([K, K])
                mdl.inj_xy, mdl.prd_xy = om.inj_xy.tolist(), om.prd_xy.tolist()
                mdl.inj_rates, mdl.prd_rates = om.inj_rates.copy(), om.prd_rates.copy()
            S0 = np.zeros(Nx * Ny)
            ours = om.sim(0.025, nT, S0)
            theirs = rm.sim(0.025, nT, S0, pbar=False)
            assert theirs.shape == ours.shape
            err = np.abs(np.asarray(theirs) - ours).max()
            assert err < 1e-10, f"oracle differs from TPFA_ResSim at {Nx}x{Ny}: max|dS| = {err}"
            np.testing.assert_array_equal(om.xy2ind(*om.prd_xy.T), rm.xy2ind(*np.asarray(rm.prd_xy).T))
    finally:
        sys.path.remove(base)
        for k in [k for k in sys.modules if k == "TPFA_ResSim" or k.startswith("TPFA_ResSim.")]:
            del sys.modules[k]
        sys.modules.update(saved)
