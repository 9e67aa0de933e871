"""Common tools - drop-in for reference ``notebooks/tools/utils.py``.

Same names, signatures and return conventions.  The one piece that changes
underneath is ``apply``: where the reference fans the members out over a
``pathos`` process pool (``tools/utils.py:201-224``), this version runs them on
light host threads whose ``ResSim.sim`` calls are collected into ONE batched GPU
ensemble run (see ``TPFA_ResSim.Collector``).
"""

from __future__ import annotations

import threading

import numpy as np
import numpy.random as rnd
import scipy.linalg as sla

nCPU = 1
"Parallelisation switch, as in the reference: an int > 1, True, None or 'auto' enable the batched GPU map; 1 / False give a plain for-loop (each sim a batch of one)."

max_batch = 512
"Additional knob (never required): largest number of members gathered into one batched launch."


def center(E, axis=0, rescale=False):
    """Anomalies of ``E`` along ``axis`` and the mean (``tools/utils.py:10-28``)."""
    mu = np.mean(E, axis=axis, keepdims=True)
    X = E - mu
    if rescale:
        n = E.shape[axis]
        X *= np.sqrt(n / (n - 1))
    return X, mu.squeeze()


def _on_device(x):
    return type(x).__module__.startswith("torch")


def cov(a, b):
    """Cross-covariance of two samples with equal ensemble size (``tools/utils.py:31-39``).
    Torch (CUDA) inputs are reduced on the device (``hm_corr``), numpy inputs on the host as in the reference."""
    if _on_device(a):
        from historymatching_b200 import analysis

        return analysis.cov(a, b)
    return center(a)[0].T @ center(b)[0] / (len(b) - 1)


def corr(a, b):
    """Cross-correlation built on ``cov``, clipped to +-999 (``tools/utils.py:42-55``)."""
    if _on_device(a):
        from historymatching_b200 import analysis

        return analysis.corr(a, b)
    C = cov(a, b)
    sa = np.std(a.T, axis=-1, ddof=1)
    sb = np.std(b, axis=0, ddof=1, keepdims=True)
    return (C / sa / sb).clip(-999, 999)


def gaussian_noise(N, M, L):
    """Zero-mean Gaussian ensemble; ``L`` is a Cholesky factor or a scalar std (``tools/utils.py:58-67``)."""
    try:
        return rnd.randn(N, len(L)) @ L.T
    except TypeError:
        return rnd.randn(N, M) * L


def rinv(A, reg, tikh=True, nMax=None):
    """Tikhonov-regularised or truncated pseudo-inverse (``tools/utils.py:70-90``)."""
    U, s, VT = sla.svd(A, full_matrices=False)
    thr = reg * s[0]
    if tikh:
        s1 = s / (s**2 + thr**2)
    else:
        keep = s >= thr
        s1 = np.zeros_like(s)
        s1[keep] = 1 / s[keep]
    if nMax:
        s1[nMax:] = 0
    return (VT.T * s1) @ U.T


def emph(text):
    return f"\033[1m{text}\033[0m"


def _mnorm(x, axis=0):
    """RMS: L2 norm with a mean instead of a sum (``tools/utils.py:124-127``)."""
    return np.sqrt(np.mean(x * x, axis))


def print_RMSMs(series, ref):
    """RMS error and deviation of each ensemble mean wrt ``series[ref]`` (``tools/utils.py:97-121``)."""
    x = series[ref]
    if x.shape[0] != 1:
        x = x[None, :]
    header = "Series    rms err  rms dev"
    print(header, "-" * len(header), sep="\n")
    for k, y in series.items():
        if y.ndim < x.ndim:
            y = y[None, :]
            assert y.shape == x.shape
        err = x - y.mean(0)
        dev = y - y.mean(0)
        print(f"{k:8}: {_mnorm(err, None):6.4f}   {_mnorm(dev, None):6.4f}")


def pCircle(degree, Lx, Ly, p=4, norm_val=0.87):
    """Point at angle ``degree`` on the p-norm circle, scaled to the domain (``tools/utils.py:130-143``)."""
    t = 2 * np.pi * degree / 360
    c, s = np.cos(t), np.sin(t)
    scale = norm_val / (np.abs(c) ** p + np.abs(s) ** p) ** (1 / p)
    x = np.round(Lx / 2 * (1 + scale * c), 2)
    y = np.round(Ly / 2 * (1 + scale * s), 2)
    return x, y


def mesh2list(*arrs):
    """``np.meshgrid`` output -> (nPoints, nDim) (``tools/utils.py:146-148``)."""
    return np.stack(arrs, -1).reshape(-1, len(arrs))


def progbar(*args, **kwargs):
    """``tqdm`` with the reference's bar format (``tools/utils.py:245-269``)."""
    kwargs.setdefault("bar_format", "{l_bar}|{bar}| {n_fmt}/{total_fmt}, ⏱️ {elapsed}, {rate_fmt}{postfix}")
    from tqdm.auto import tqdm

    return tqdm(*args, **kwargs)


def _parallel_enabled():
    is_int = type(nCPU) == int  # noqa: E721  (bool is not an int here, as in the reference)
    if not is_int and nCPU in [True, None, "auto"]:
        return True
    return bool(nCPU) and nCPU > 1


def _inside_member():
    from TPFA_ResSim import inside_member

    return inside_member()


def apply(fun, *args, pbar=True, **kwargs):
    """Apply ``fun`` along axis 0 of every positional and keyword argument.

    Contract of reference ``tools/utils.py:155-242``: arguments are zipped with
    ``strict=True``, the output is an ordered list, ``pbar`` may be a bool, a
    description string, a dict of tqdm options or an existing tqdm bar, and
    ``fun.nCalls`` is bumped when present.  Exceptions raised by ``fun`` propagate.
    """
    nPositional = len(args)
    args = list(args) + list(kwargs.values())
    inputs = list(zip(*args, strict=True))

    def _fun(x):
        positional, named = x[:nPositional], x[nPositional:]
        return fun(*positional, **dict(zip(kwargs, named)))

    if "tqdm" in str(type(pbar)).lower():
        pbar.do_close = False
        pbar.reset(total=len(args[0]))
    elif pbar:
        kws = dict(total=len(args[0]), desc=f"map({getattr(fun, '__name__', 'fun')}, ...)", leave=True)
        if isinstance(pbar, str):
            kws["desc"] = pbar
        elif isinstance(pbar, dict):
            kws.update(pbar)
        pbar = progbar(**kws)
    else:
        pbar = progbar(disable=True)

    if _parallel_enabled() and len(inputs) > 1 and not _inside_member():
        # NB: threads share memory, so a `fun.nCalls` counter is bumped by the calls
        # themselves (the reference adds len(inputs) only because child processes cannot).
        output = _batched_map(_fun, inputs, pbar)
    else:
        output = []
        for x in inputs:
            output.append(_fun(x))
            pbar.update()

    pbar.refresh()
    if getattr(pbar, "do_close", True):
        pbar.close()
    return output


class _Member(threading.Thread):
    """Persistent worker: one member of a chunk runs on it as a coroutine of ``TPFA_ResSim.Collector`` (creating 40
    threads per ensemble run costs more than the 20 x 20 forward run itself).  The thread blocks on ``lock`` whenever it
    does not hold the collector's baton; ``task`` is set by the dispatching thread before the baton first arrives."""

    def __init__(self):
        super().__init__(daemon=True, name="hm-member")
        self.lock = threading.Lock()
        self.lock.acquire()
        self.task = None
        self.start()

    def run(self):
        while True:
            self.lock.acquire()  # the baton arrives: run this chunk's member
            collector, me, work = self.task
            self.task = None
            collector.attach(me)
            try:
                work(me)
            finally:
                collector.detach()
                collector.finish()


_members = []
_members_lock = threading.Lock()  # one chunk at a time uses the workers


def _batched_map(_fun, inputs, pbar):
    """Run the members as coroutines on worker threads; their ``ResSim.sim`` calls rendezvous into batched GPU runs."""
    from TPFA_ResSim import Collector

    output = [None] * len(inputs)
    errors = [None] * len(inputs)
    with _members_lock:
        for lo in range(0, len(inputs), max_batch):
            n = min(max_batch, len(inputs) - lo)
            while len(_members) < n:
                _members.append(_Member())

            def work(me, lo=lo):
                try:
                    output[lo + me] = _fun(inputs[lo + me])
                except BaseException as e:  # noqa: BLE001 - re-raised in the caller below
                    errors[lo + me] = e
                pbar.update()

            collector = Collector([m.lock for m in _members[:n]])
            for me in range(n):
                _members[me].task = (collector, me, work)
            collector.run()
            for e in errors:
                if e is not None:
                    raise e
    return output
