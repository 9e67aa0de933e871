"""CPU oracle: 2-D two-phase incompressible TPFA reservoir simulator.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: the
algorithm lives in the un-vendored dependency ``TPFA-ResSim@adc89536``
(``/root/reference/requirements.txt:1``); this is a numpy/scipy restatement of
the published scheme it implements (Aarnes, Gimse & Lie, "An introduction to
the numerics of flow in porous media using Matlab": TPFA.m / RelPerm.m /
Upstream.m / GenA.m), following SURVEY.md Appendix A.  It is anchored on the
reference's call sites: ``HistoryMatch.py:97`` (ctor), ``:164`` (K),
``:187-190`` (wells), ``:209`` (xy2ind), ``:224`` / ``:362`` (sim);
``Optimise.py:64-89,116,175-176``.

Index convention: fields are ``(Nx, Ny)`` arrays, flat cell index
``c = ix*Ny + iy`` (C-order ravel; ``HistoryMatch.py:163``).
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve


class OracleResSim:
    """Minimal simulator object with the attribute surface the notebooks use."""

    def __init__(self, Nx, Ny, Lx=1.0, Ly=1.0, vw=1.0, vo=1.0, swc=0.0, sor=0.0):
        self.Nx, self.Ny, self.Lx, self.Ly = int(Nx), int(Ny), float(Lx), float(Ly)
        self.shape = (self.Nx, self.Ny)
        self.Nxy = self.Nx * self.Ny
        self.hx, self.hy = self.Lx / self.Nx, self.Ly / self.Ny
        self.h2 = self.hx * self.hy
        self.vw, self.vo, self.swc, self.sor = vw, vo, swc, sor
        self.K = np.ones((2, Nx, Ny))
        self.por = np.ones(self.shape)
        self.inj_xy = np.zeros((0, 2))
        self.prd_xy = np.zeros((0, 2))
        self.inj_rates = np.zeros((0, 1))
        self.prd_rates = np.zeros((0, 1))
        # 0 = the reference's numerical path (one SuperLU solve).  k > 0 adds k rounds of
        # iterative refinement with an extended-precision residual: the direct solve alone
        # carries a forward error ~ cond(A)*eps, up to 1e-6 for K contrasts of 1e8+.
        self.refine = 0

    # ---- grid (Appendix A.1) -------------------------------------------------
    def xy2ind(self, x, y):
        """Cell index containing the point (x, y); raises outside the domain."""
        x, y = np.asarray(x, float), np.asarray(y, float)
        if np.any((x < 0) | (x > self.Lx) | (y < 0) | (y > self.Ly)):
            raise ValueError("point outside the domain")
        ix = np.minimum((x / self.Lx * self.Nx).astype(int), self.Nx - 1)
        iy = np.minimum((y / self.Ly * self.Ny).astype(int), self.Ny - 1)
        return ix * self.Ny + iy

    def ind2xy(self, ind):
        ind = np.asarray(ind)
        ix, iy = ind // self.Ny, ind % self.Ny
        return np.array([(ix + 0.5) * self.hx, (iy + 0.5) * self.hy])

    @property
    def mesh(self):
        xs = np.linspace(0, self.Lx, self.Nx, endpoint=False) + self.hx / 2
        ys = np.linspace(0, self.Ly, self.Ny, endpoint=False) + self.hy / 2
        return np.meshgrid(xs, ys, indexing="ij")

    # ---- wells (Appendix A.1) ------------------------------------------------
    def source_field(self, k):
        """q[c] for time index k: +rate at injectors, -rate at producers."""
        q = np.zeros(self.Nxy)
        for xy, rates, sign in ((self.inj_xy, self.inj_rates, +1.0),
                                (self.prd_xy, self.prd_rates, -1.0)):
            xy = np.asarray(xy, float).reshape(-1, 2)
            rates = np.asarray(rates, float)
            if rates.ndim == 1:
                rates = rates[:, None]
            r = rates[:, 0] if rates.shape[1] == 1 else rates[:, k]
            np.add.at(q, self.xy2ind(xy[:, 0], xy[:, 1]), sign * r)
        if not np.isclose(q.sum(), 0.0):
            raise ValueError("injection and production do not balance")
        return q

    # ---- fluid ----------------------------------------------------------------
    def mobilities(self, s):
        """RelPerm.m: quadratic relative permeabilities over viscosity."""
        se = (s - self.swc) / (1 - self.swc - self.sor)
        return se**2 / self.vw, (1 - se) ** 2 / self.vo

    # ---- pressure (Appendix A.2) ----------------------------------------------
    def transmissibilities(self, KM):
        """TPFA.m: harmonic-average face transmissibilities, zero on the boundary."""
        Nx, Ny = self.shape
        L = 1.0 / KM
        TX = np.zeros((Nx + 1, Ny))
        TY = np.zeros((Nx, Ny + 1))
        TX[1:-1, :] = 2 * self.hy / self.hx / (L[0, :-1, :] + L[0, 1:, :])
        TY[:, 1:-1] = 2 * self.hx / self.hy / (L[1, :, :-1] + L[1, :, 1:])
        return TX, TY

    def pressure_matrix(self, TX, TY):
        Nx, Ny = self.shape
        x1, x2 = TX[:-1, :].ravel(), TX[1:, :].ravel()
        y1, y2 = TY[:, :-1].ravel(), TY[:, 1:].ravel()
        diag = y1 + y2 + x1 + x2
        # Pin the otherwise singular Neumann problem (TPFA.m: A(1,1) += sum(K(:,1,1)))
        diag[0] += self.K[0, 0, 0] + self.K[1, 0, 0]
        A = sp.spdiags([-x2, -y2, diag, -y1, -x1], [-Ny, -1, 0, 1, Ny], self.Nxy, self.Nxy)
        return A.tocsc()

    def residual_ext(self, TX, TY, q, u):
        """q - A u evaluated in extended precision (difference form of the stencil)."""
        ld = np.longdouble
        P = np.asarray(u, ld).reshape(self.shape)
        TXl, TYl = TX.astype(ld), TY.astype(ld)
        Au = np.zeros(self.shape, ld)
        fx = (P[:-1, :] - P[1:, :]) * TXl[1:-1, :]   # flux from ix to ix+1
        fy = (P[:, :-1] - P[:, 1:]) * TYl[:, 1:-1]
        Au[:-1, :] += fx
        Au[1:, :] -= fx
        Au[:, :-1] += fy
        Au[:, 1:] -= fy
        Au[0, 0] += (ld(self.K[0, 0, 0]) + ld(self.K[1, 0, 0])) * P[0, 0]
        return q.astype(ld) - Au.ravel()

    def pressure_step(self, S, q):
        lw, lo = self.mobilities(S)
        KM = (lw + lo).reshape(self.shape) * self.K
        TX, TY = self.transmissibilities(KM)
        A = self.pressure_matrix(TX, TY)
        if self.refine:
            from scipy.sparse.linalg import splu

            lu = splu(A)
            u = lu.solve(q).astype(np.longdouble)
            for _ in range(self.refine):
                u = u + lu.solve(np.asarray(self.residual_ext(TX, TY, q, u), float))
            u = np.asarray(u, float)
        else:
            u = spsolve(A, q)
        P = u.reshape(self.shape)
        Vx = np.zeros_like(TX)
        Vy = np.zeros_like(TY)
        Vx[1:-1, :] = (P[:-1, :] - P[1:, :]) * TX[1:-1, :]
        Vy[:, 1:-1] = (P[:, :-1] - P[:, 1:]) * TY[:, 1:-1]
        return P, Vx, Vy

    # ---- saturation (Appendix A.3) ---------------------------------------------
    def cfl_substeps(self, q, Vx, Vy, T):
        pv = self.h2 * self.por.ravel()
        fi = np.maximum(q, 0)
        XP, XN = np.maximum(Vx, 0), np.minimum(Vx, 0)
        YP, YN = np.maximum(Vy, 0), np.minimum(Vy, 0)
        Vi = XP[:-1] + YP[:, :-1] - XN[1:] - YN[:, 1:]
        with np.errstate(divide="ignore"):
            pm = np.min(pv / (Vi.ravel() + fi))
        cfl = ((1 - (self.swc + self.sor)) / 3) * pm
        Nts = int(np.ceil(T / cfl))
        return Nts, (T / Nts) / pv

    def upwind_matrix(self, q, Vx, Vy):
        Ny = self.Ny
        fp = np.minimum(q, 0)
        x1 = np.minimum(Vx, 0)[:-1, :].ravel()
        x2 = np.maximum(Vx, 0)[1:, :].ravel()
        y1 = np.minimum(Vy, 0)[:, :-1].ravel()
        y2 = np.maximum(Vy, 0)[:, 1:].ravel()
        diag = fp + y1 - y2 + x1 - x2
        return sp.spdiags([x2, y2, diag, -y1, -x1], [-Ny, -1, 0, 1, Ny], self.Nxy, self.Nxy).tocsr()

    def saturation_step(self, S, q, Vx, Vy, T):
        Nts, dtx = self.cfl_substeps(q, Vx, Vy, T)
        fi = np.maximum(q, 0)
        B = sp.diags(dtx) @ self.upwind_matrix(q, Vx, Vy)
        src = fi * dtx
        for _ in range(Nts):
            lw, lo = self.mobilities(S)
            fw = lw / (lw + lo)
            S = S + (B @ fw + src)
        return S, Nts

    # ---- driver (Appendix A.4) ---------------------------------------------------
    def sim(self, dt, nSteps, S0, return_aux=False):
        out = np.zeros((nSteps + 1, self.Nxy))
        out[0] = S0
        aux = dict(P=[], Nts=[])
        for k in range(nSteps):
            q = self.source_field(k)
            P, Vx, Vy = self.pressure_step(out[k], q)
            out[k + 1], Nts = self.saturation_step(out[k], q, Vx, Vy, dt)
            aux["P"].append(P.ravel())
            aux["Nts"].append(Nts)
        return (out, aux) if return_aux else out


def notebook_model(Nx=20, Ny=20, Lx=2.0, Ly=1.0):
    """The well/grid setup of ``HistoryMatch.py:97,177-190``."""
    m = OracleResSim(Nx=Nx, Ny=Ny, Lx=Lx, Ly=Ly)
    near01 = np.array([0.12, 0.87])
    m.prd_xy = np.array([[x, y] for y in Ly * near01 for x in Lx * near01])
    m.inj_xy = np.array([[Lx / 2, Ly / 2]])
    m.inj_rates = np.array([[1.0]])
    m.prd_rates = np.ones((4, 1)) / 4
    # wells are collocated with cell centres (HistoryMatch.py:197)
    m.prd_xy = m.ind2xy(m.xy2ind(*m.prd_xy.T)).T
    m.inj_xy = m.ind2xy(m.xy2ind(*m.inj_xy.T)).T
    return m


def perm_transf(x):
    """``HistoryMatch.py:137-138``."""
    return 0.1 + np.exp(5 * x)


def forward_member(model, log_perm, dt, nSteps, S0, obs_cells):
    """``comp1`` (``HistoryMatch.py:358-364``) for one member on the oracle."""
    import copy

    m = copy.copy(model)
    p = perm_transf(np.asarray(log_perm)).reshape(model.shape)
    m.K = np.stack([p, p])
    wsats = m.sim(dt, nSteps, S0)
    return wsats, wsats[1:, obs_cells]
