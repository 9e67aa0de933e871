"""EnOpt driver - drop-in for reference ``notebooks/tools/enopt.py``.

Host logic only: every ``apply(obj, U)`` batch of
objective evaluations lands on the GPU through the collector in
``tools.utils.apply``; the gradient / line-search arithmetic is tiny and stays numpy.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from tools import utils
from tools.utils import apply, center, progbar


@dataclass
class nabla_ens:
    """Ensemble (LLS-regression) gradient estimate (``tools/enopt.py:11-35``)."""

    chol: float = 1.0
    nEns: int = 10
    precond: bool = False
    robustly: None = None
    obj_ux: None = None
    X: None = None

    def __call__(self, obj, u, pbar=None):
        U = utils.gaussian_noise(self.nEns, len(u), self.chol)
        dU = center(U)[0]
        dJ = self.ens_eval(obj, u, u + dU, pbar)
        if self.precond:
            return dU.T @ dJ / (self.nEns - 1)
        return utils.rinv(dU, reg=0.1, tikh=True) @ dJ

    def ens_eval(self, obj, u, U, pbar):
        return apply(obj, U, pbar=pbar)


def split(arr, step):
    """Consecutive segments of length ``step`` (default: cpu_count()-1) (``tools/enopt.py:64-72``)."""
    if not step:
        import multiprocessing

        step = max(1, multiprocessing.cpu_count() - 1)
    return [arr[i:i + step] for i in range(0, len(arr), step)]


@dataclass
class backtracker:
    """Shrink the step until the objective improves admissibly (``tools/enopt.py:38-61``)."""

    sign: int = +1
    xSteps: tuple = tuple(0.5 ** (i + 1) for i in range(8))
    rtol: float = 1e-8
    nCPU: int = None

    def __call__(self, obj, u0, J0, search_direction, pbar):
        atol = max(1e-8, abs(J0)) * self.rtol
        pbar.reset(len(self.xSteps))

        def trial(xStep):
            u1 = u0 + self.sign * xStep * search_direction
            J1 = obj(u1)
            return u1, J1, J1 - J0

        for steps in split(self.xSteps, self.nCPU):
            for u1, J1, dJ in apply(trial, steps, pbar=False):
                pbar.update()
                if self.sign * dJ > atol:
                    return u1, J1, dict(nDeclined=pbar.n)


def GD(objective, u, nabla=nabla_ens(), line_search=backtracker(), nrmlz=True, nIter=100, quiet=False):
    """Steepest ascent/descent with ensemble gradients (``tools/enopt.py:75-107``)."""
    with (progbar(total=nIter, desc="⏳ GD running", leave=True, disable=quiet) as pbar_gd,
          progbar(total=10000, desc="→ grad. comp.", leave=False, disable=quiet) as pbar_en,
          progbar(total=10000, desc="→ line_search", leave=False, disable=quiet) as pbar_ls,
          np.printoptions(precision=2, threshold=2, edgeitems=1)):
        states = [[u, objective(u), {}]]
        itr = 0
        for itr in range(nIter):
            u, J, info = states[-1]
            pbar_gd.set_postfix(u=f"{u}", obj=f"{J:.3g}📈")
            grad = nabla(objective, u, pbar_en)
            info["grad"] = grad
            if nrmlz:
                grad /= np.sqrt(np.mean(grad**2))
            updated = line_search(objective, u, J, grad, pbar_ls)
            pbar_gd.update()
            if updated:
                states.append(updated)
            else:
                info["cause"] = "✅ GD converged"
                break
        else:
            info["cause"] = "❌ GD ran out of iters"
        info["nIter"] = itr
        pbar_gd.set_description(info["cause"])
    return (np.asarray(arr) for arr in zip(*states))
