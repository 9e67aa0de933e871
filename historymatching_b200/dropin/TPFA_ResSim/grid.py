"""2-D rectangular grid helpers of the simulator object.

Drop-in for ``TPFA_ResSim.grid.Grid2D`` (imported at reference
``tools/geostat.py:103``); member contracts inferred from the notebook call
sites (SURVEY.md section 8(b) Seam 1): ``shape``, ``Nxy``, ``mesh``, ``domain``,
``xy2ind``, ``ind2xy``, ``sub2ind``, ``sub2xy`` (``HistoryMatch.py:152,163,209,
479,700-701``; ``Optimise.py:451,465``).
"""

from __future__ import annotations

import numpy as np


class Grid2D:
    def __init__(self, Lx=1.0, Ly=1.0, Nx=32, Ny=32):
        self.Lx, self.Ly, self.Nx, self.Ny = Lx, Ly, int(Nx), int(Ny)

    # derived quantities are properties so that deepcopy / setattr stay trivial
    @property
    def shape(self):
        return (self.Nx, self.Ny)

    @property
    def Nxy(self):
        return self.Nx * self.Ny

    @property
    def hx(self):
        return self.Lx / self.Nx

    @property
    def hy(self):
        return self.Ly / self.Ny

    @property
    def h2(self):
        return self.hx * self.hy

    @property
    def domain(self):
        return ((0, 0), (self.Lx, self.Ly))

    @property
    def mesh(self):
        """Cell-centre coordinates, C-ravel order = flat cell index."""
        xx = np.linspace(0, self.Lx, self.Nx, endpoint=False) + self.hx / 2
        yy = np.linspace(0, self.Ly, self.Ny, endpoint=False) + self.hy / 2
        return np.meshgrid(xx, yy, indexing="ij")

    def sub2ind(self, ix, iy):
        return np.ravel_multi_index((np.asarray(ix), np.asarray(iy)), self.shape)

    def ind2sub(self, ind):
        return np.unravel_index(np.asarray(ind), self.shape)

    def xy2sub(self, x, y):
        x, y = np.asarray(x, float), np.asarray(y, float)
        if np.any((x < 0) | (x > self.Lx) | (y < 0) | (y > self.Ly)) or not (
            np.all(np.isfinite(x)) and np.all(np.isfinite(y))
        ):
            raise ValueError("coordinates outside the model domain")
        ix = np.minimum((x / self.Lx * self.Nx).astype(int), self.Nx - 1)
        iy = np.minimum((y / self.Ly * self.Ny).astype(int), self.Ny - 1)
        return ix, iy

    def sub2xy(self, ix, iy):
        return np.array([(np.asarray(ix) + 0.5) * self.hx, (np.asarray(iy) + 0.5) * self.hy])

    def xy2ind(self, x, y):
        return self.sub2ind(*self.xy2sub(x, y))

    def ind2xy(self, ind):
        return self.sub2xy(*self.ind2sub(ind))
