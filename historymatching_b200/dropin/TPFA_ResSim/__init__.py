"""B200-native drop-in for the ``TPFA_ResSim`` simulator package.

The reference imports ``TPFA_ResSim`` (``HistoryMatch.py:88``, ``Optimise.py:46``;
pinned at ``requirements.txt:1``) and uses its ``ResSim`` object as listed in
SURVEY.md section 8(b) Seam 1.  This class keeps that surface - constructor keywords,
the ``K`` / ``inj_xy`` / ``prd_xy`` / ``inj_rates`` / ``prd_rates`` setters, grid
helpers and ``sim(dt, nSteps, S0, pbar)`` - but ``sim`` runs on the GPU through
``hm_sim_batch`` (include/hm_b200.h).  When called inside
``tools.utils.apply`` (which runs the members' callables on light threads under a
*collector*), the calls of all members are gathered and executed as ONE batched
ensemble launch; outside of it a call is a batch of one.
"""

from __future__ import annotations

import threading

import numpy as np

from .grid import Grid2D
from .plotting import Plot2D

__all__ = ["ResSim", "Grid2D"]

_tls = threading.local()


class ResSim(Grid2D, Plot2D):
    """2-D two-phase incompressible TPFA reservoir simulator (``HistoryMatch.py:93-97``)."""

    def __init__(self, Lx=1.0, Ly=1.0, Nx=32, Ny=32, name="", vw=1.0, vo=1.0, swc=0.0, sor=0.0):
        Grid2D.__init__(self, Lx=Lx, Ly=Ly, Nx=Nx, Ny=Ny)
        object.__setattr__(self, "name", name)
        self.vw, self.vo, self.swc, self.sor = vw, vo, swc, sor
        self.K = np.ones((2, self.Nx, self.Ny))
        self.por = np.ones(self.shape)
        self.inj_xy = np.zeros((0, 2))
        self.prd_xy = np.zeros((0, 2))
        self.inj_rates = np.zeros((0, 1))
        self.prd_rates = np.zeros((0, 1))
        self.actual_rates = {}

    # Validation / normalisation of the settable attributes lives here because the
    # notebooks configure models with plain ``setattr`` (``Optimise.py:133-136``).
    def __setattr__(self, key, val):
        if key == "K":
            val = np.asarray(val, float)
            if val.size == self.Nxy:  # (Nxy,), (1,Nxy), (Nx,Ny): isotropic
                val = np.broadcast_to(val.reshape(self.shape), (2, *self.shape))
            val = np.array(val.reshape((2, *self.shape)))
            if not np.all(val > 0):
                raise ValueError("permeability must be positive")
        elif key in ("inj_xy", "prd_xy"):
            val = np.array(val, float).reshape((-1, 2))
            cells = np.zeros(0, np.int32)
            if len(val):  # collocate with cell centres; raises outside the domain
                cells = self.xy2ind(*val.T)
                val = self.ind2xy(cells).T
            object.__setattr__(self, "_" + key[:3] + "_cells", np.asarray(cells, np.int32))
        elif key in ("inj_rates", "prd_rates"):
            val = np.array(val, float)
            if val.ndim == 1:
                val = val[:, None]
            if val.ndim != 2 or not np.all(np.isfinite(val)):
                raise ValueError(f"{key} must be a finite (nWell, nTime|1) array")
        object.__setattr__(self, key, val)

    def __deepcopy__(self, memo):
        """``copy.deepcopy(model)`` is what every member's forward run starts with (``HistoryMatch.py:360``,
        ``Optimise.py:133``): copy the arrays directly instead of walking the object graph (~10x faster)."""
        new = object.__new__(type(self))
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_sched":  # the cached well schedule: read-only arrays, shared
                pass
            elif isinstance(v, np.ndarray):
                v = v.copy()
            elif isinstance(v, dict) and all(isinstance(x, np.ndarray) for x in v.values()):
                v = {kk: x.copy() for kk, x in v.items()}
            elif not isinstance(v, (int, float, str, bool, type(None))):
                import copy

                v = copy.deepcopy(v, memo)
            object.__setattr__(new, k, v)
        return new

    @property
    def nInj(self):
        return len(self.inj_xy)

    @property
    def nPrd(self):
        return len(self.prd_xy)

    # ---- request assembly ------------------------------------------------------------
    def _schedule(self, nSteps):
        """Signed well rates (nSteps, nW) and cells (nW,); validates the balance.

        The members of an ensemble are deep copies of one model with the same wells, so the result is cached on the
        instance (and travels with ``copy.deepcopy``); the key holds the bytes of the four small well arrays, so a
        change made in place (``model.inj_rates[0] = 2``) is seen as well as an assignment."""
        key = (nSteps, self.inj_xy.tobytes(), self.prd_xy.tobytes(), self.inj_rates.tobytes(), self.prd_rates.tobytes(),
               self.inj_rates.shape, self.prd_rates.shape)
        cached = self.__dict__.get("_sched")
        if cached is not None and cached[0] == key:
            self.actual_rates = {k: v.copy() for k, v in cached[3].items()}
            return cached[1], cached[2]
        rates = []
        for kind, sign in (("inj", 1.0), ("prd", -1.0)):
            xy, r = getattr(self, f"{kind}_xy"), getattr(self, f"{kind}_rates")
            if len(xy) != len(r):
                raise ValueError(f"{kind}_xy and {kind}_rates disagree on the number of wells")
            if r.shape[1] not in (1, nSteps) and r.shape[1] < nSteps:
                raise ValueError(f"{kind}_rates has {r.shape[1]} time columns, need 1 or >= {nSteps}")
            r = np.broadcast_to(r, (len(r), nSteps)) if r.shape[1] == 1 else r[:, :nSteps]
            rates.append(sign * r)
        q = np.concatenate(rates).T  # (nSteps, nW)
        if np.abs(q.sum(1)).max(initial=0.0) > 1e-8:  # np.allclose(q.sum(1), 0) without its overhead
            raise ValueError("total injection must equal total production at every time step")
        cells = np.concatenate([self._inj_cells, self._prd_cells])  # collocated when the wells were set
        self.actual_rates = {"inj": np.array(rates[0]), "prd": -np.array(rates[1])}
        q = np.ascontiguousarray(q)
        q.setflags(write=False)
        cells.setflags(write=False)
        object.__setattr__(self, "_sched", (key, q, cells, {k: v.copy() for k, v in self.actual_rates.items()}))
        return q, cells

    def sim(self, dt, nSteps, S0, pbar=True, leave=False):
        """Run ``nSteps`` steps of length ``dt`` from saturation ``S0``.

        Returns ``(nSteps+1, Nxy)`` float64, row 0 = ``S0`` (``HistoryMatch.py:224-225``).
        """
        S0 = np.asarray(S0, float).reshape(-1)
        if S0.shape != (self.Nxy,):
            raise ValueError("S0 must have Nxy entries")
        rates, cells = self._schedule(int(nSteps))
        req = _Request(self, float(dt), int(nSteps), S0, rates, cells)
        collector = getattr(_tls, "collector", None)
        if collector is not None:
            return collector.submit(req)
        run_requests([req])
        return req.take()


class _Request:
    __slots__ = ("key", "K", "por", "S0", "rates", "cells", "dt", "nSteps", "result", "error", "done", "grid")

    def __init__(self, model, dt, nSteps, S0, rates, cells):
        from historymatching_b200.sim import GridSpec

        self.grid = GridSpec(model.Nx, model.Ny, float(model.Lx), float(model.Ly), model.vw, model.vo,
                             model.swc, model.sor)
        self.K = model.K.reshape(2, -1)
        por = np.asarray(model.por, float).reshape(-1)
        self.por = None if np.all(por == 1.0) else por
        self.S0, self.rates, self.cells, self.dt, self.nSteps = S0, rates, cells, dt, nSteps
        self.key = (model.Nx, model.Ny, float(model.Lx), float(model.Ly), model.vw, model.vo, model.swc,
                    model.sor, dt, nSteps, len(cells), None if self.por is None else self.por.tobytes())
        self.result = self.error = None
        self.done = False

    def take(self):
        if self.error is not None:
            raise self.error
        return self.result


def run_requests(reqs):
    """Execute simulation requests, one batched GPU call per group of compatible requests."""
    from historymatching_b200.sim import run_ensemble

    groups = {}
    for r in reqs:
        groups.setdefault(r.key, []).append(r)
    for grp in groups.values():
        g0 = grp[0]
        try:
            K = np.stack([r.K for r in grp])                       # (n,2,M)
            if all(np.array_equal(r.K[0], r.K[1]) for r in grp):
                K = np.ascontiguousarray(K[:, 0])                  # isotropic: (n,M)
            res = run_ensemble(
                g0.grid, K, np.stack([r.cells for r in grp]), np.stack([r.rates for r in grp]),
                np.stack([r.S0 for r in grp]), g0.dt, g0.nSteps, por=g0.por, history=True,
                n_members=len(grp),
            )
            for i, r in enumerate(grp):
                st = int(res.status[i])
                if st:
                    r.error = RuntimeError(f"simulation failed for this member (status bits {st})")
                else:
                    r.result = res.S_hist[i]
        except Exception as e:  # launch / allocation failure: every member of the group fails
            for r in grp:
                r.error = e
        for r in grp:
            r.done = True


class Collector:
    """Cooperative scheduler of the member threads of one ``apply`` chunk.

    The members' own cell code is plain Python under the GIL, so running it on many threads at once buys nothing and
    costs interpreter hand-offs, condition-variable wake-ups of the whole herd and cache misses (measured on an 8-core
    host, 200 members of the 20 x 20 notebook case, GPU call stubbed out: 250 ms per ensemble run with free-running
    threads against 88 ms for the members run one after the other).  The threads are therefore used as *coroutines*:
    exactly one of them holds the **baton** and runs; every thread owns a lock it blocks on while it does not.  A member
    that reaches ``ResSim.sim`` parks its request and passes the baton to the next member that can run; when none can
    and requests are parked, the holder runs them as ONE batched GPU call and the parked members resume in order; when
    nothing is left the dispatching thread is released.  The scheduler state is only ever touched by the baton holder,
    so it needs no mutex (a lock release / acquire pair orders the memory accesses of the two threads).
    """

    def __init__(self, locks):
        import collections

        self.locks = locks                        # one per member, held (= blocked) unless the member has the baton
        self.idle = threading.Lock()              # the dispatching thread blocks on this one
        self.idle.acquire()
        self.runq = collections.deque(range(len(locks)))
        self.parked = []                          # (member, request) in arrival order

    def run(self):
        """Called by the dispatching thread: start the first member, return when every member has finished."""
        self._pass(None)
        self.idle.acquire()

    def _pass(self, me):
        while True:
            if self.runq:
                nxt = self.runq.popleft()
                if nxt == me:
                    return                        # it is my turn again
                self.locks[nxt].release()
                break
            if self.parked:                       # every unfinished member is parked in sim(): one batched run
                batch, self.parked = self.parked, []
                run_requests([r for _, r in batch])
                self.runq.extend(i for i, _ in batch)
                continue
            self.idle.release()                   # nothing left to run
            break
        if me is not None:
            self.locks[me].acquire()              # until the baton comes back

    def submit(self, req):
        me = _tls.member
        self.parked.append((me, req))
        self._pass(me)
        return req.take()

    def finish(self):
        self._pass(None)

    def attach(self, me):
        _tls.collector, _tls.member = self, me

    @staticmethod
    def detach():
        _tls.collector = None


def inside_member():
    """True on a member thread of a running ``apply`` chunk (a nested ``apply`` must not wait for that pool)."""
    return getattr(_tls, "collector", None) is not None
