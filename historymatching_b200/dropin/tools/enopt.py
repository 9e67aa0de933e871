"""Ensemble-based optimisation (EnOpt) host logic: drop-in for the reference's ``notebooks/tools/enopt.py``.

Only the surface the notebooks use is kept - ``nabla_ens``, ``backtracker``, ``GD``, ``split`` with the same
constructor fields, call signatures and return values - and the arithmetic is arranged so that a run on the same
seed reproduces the reference's trajectory bit for bit (``tests/golden/enopt.npz``).  What matters for the hot
path: every batch of objective evaluations goes through ``tools.utils.apply``,
whose collector turns the members' ``ResSim.sim`` calls into ONE ``hm_sim_batch``; the regression and the step
control below are a few flops and stay numpy.
"""

from __future__ import annotations

import contextlib
import multiprocessing
from dataclasses import dataclass

import numpy as np

from tools import utils

_HALVINGS = tuple(0.5 ** (k + 1) for k in range(8))


@dataclass
class nabla_ens:
    """Gradient of ``obj`` at ``u`` estimated by linear regression on an ensemble of perturbed controls.

    ``chol``: Cholesky factor of the perturbation covariance (or a scalar standard deviation); ``nEns``: ensemble
    size; ``precond``: return the covariance-preconditioned gradient ``C_u g`` (cross-covariance of controls and
    objective) instead of the regularised least-squares solution.  ``robustly``, ``obj_ux``, ``X`` are carried for
    the notebooks' robust-objective variants, which override ``ens_eval``.
    """

    # dataclass fields with class-level defaults, as in the reference (``nabla_ens.nEns``, ``dataclasses.replace`` work)
    chol: float = 1.0
    nEns: int = 10
    precond: bool = False
    robustly: None = None
    obj_ux: None = None
    X: None = None

    def ens_eval(self, obj, u, U, pbar):
        """Objective of every perturbed control: one batched forward run behind ``utils.apply``."""
        return utils.apply(obj, U, pbar=pbar)

    def __call__(self, obj, u, pbar=None):
        anomalies = utils.center(utils.gaussian_noise(self.nEns, len(u), self.chol))[0]
        responses = self.ens_eval(obj, u, u + anomalies, pbar)
        return _regress(anomalies, responses, self.nEns, self.precond)


def _regress(dU, dJ, n_ens, precond):
    """Slope of ``dJ`` against the control anomalies ``dU``: cross-covariance, or Tikhonov pseudo-inverse (10 %)."""
    if precond:
        return dU.T @ dJ / (n_ens - 1)
    return utils.rinv(dU, reg=0.1, tikh=True) @ dJ


def split(arr, step):
    """``arr`` cut into consecutive pieces of ``step`` items; no ``step``: one piece per available worker."""
    size = step or max(1, multiprocessing.cpu_count() - 1)
    return [arr[start:start + size] for start in range(0, len(arr), size)]


@dataclass
class backtracker:
    """Line search: try the step lengths ``xSteps`` in order, accept the first admissible improvement.

    ``sign`` = +1 maximises, -1 minimises; an improvement is admissible if it exceeds ``rtol * max(1e-8, |J0|)``.
    The trials of one piece (``split(xSteps, nCPU)``) are evaluated together - as one batch on the GPU.
    Returns ``(u1, J1, info)`` or ``None`` when every trial is declined.
    """

    sign: int = +1
    xSteps: tuple = _HALVINGS
    rtol: float = 1e-8
    nCPU: int = None

    def __call__(self, obj, u0, J0, search_direction, pbar):
        threshold = max(1e-8, abs(J0)) * self.rtol
        pbar.reset(len(self.xSteps))

        def probe(length):
            candidate = u0 + self.sign * length * search_direction
            value = obj(candidate)
            return candidate, value, value - J0

        for piece in split(self.xSteps, self.nCPU):
            for candidate, value, gain in utils.apply(probe, piece, pbar=False):
                pbar.update()
                if self.sign * gain > threshold:
                    return candidate, value, dict(nDeclined=pbar.n)
        return None


@contextlib.contextmanager
def _bars(n_iter, quiet):
    """The three progress bars of a run (outer loop, gradient ensemble, line search) and terse array printing."""
    with contextlib.ExitStack() as stack:
        outer = stack.enter_context(utils.progbar(total=n_iter, desc="⏳ GD running", leave=True, disable=quiet))
        grad = stack.enter_context(utils.progbar(total=10000, desc="→ grad. comp.", leave=False, disable=quiet))
        search = stack.enter_context(utils.progbar(total=10000, desc="→ line_search", leave=False, disable=quiet))
        stack.enter_context(np.printoptions(precision=2, threshold=2, edgeitems=1))
        yield outer, grad, search


def GD(objective, u, nabla=nabla_ens(), line_search=backtracker(), nrmlz=True, nIter=100, quiet=False):
    """Steepest ascent / descent driven by ``nabla`` (gradient estimate) and ``line_search`` (step control).

    Returns three arrays (a generator of them, as the reference does): the accepted controls, their objective values
    and one info dict per accepted state (``grad`` - normalised in place when ``nrmlz`` -, ``nDeclined``, and on the
    last one ``cause`` / ``nIter``).
    """
    history = [[u, objective(u), {}]]
    with _bars(nIter, quiet) as (bar_outer, bar_grad, bar_search):
        done = 0
        verdict = "❌ GD ran out of iters"
        while done < nIter:
            here, value, notes = history[-1]
            bar_outer.set_postfix(u=f"{here}", obj=f"{value:.3g}📈")
            direction = nabla(objective, here, bar_grad)
            notes["grad"] = direction
            if nrmlz:
                direction /= np.sqrt(np.mean(direction**2))
            accepted = line_search(objective, here, value, direction, bar_search)
            bar_outer.update()
            if not accepted:
                verdict = "✅ GD converged"
                break
            history.append(accepted)
            done += 1
        else:
            done = nIter - 1 if nIter else 0
        notes = history[-1][2] if nIter == 0 else notes
        notes["cause"] = verdict
        notes["nIter"] = done
        bar_outer.set_description(verdict)
    return (np.asarray(column) for column in zip(*history))
