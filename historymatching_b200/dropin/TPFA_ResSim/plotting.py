"""Plotting hand-off of the simulator object (OUT OF SCOPE beyond a stub).

The reference's figures (``TPFA_ResSim.plotting``; used at ``HistoryMatch.py:201,
233,275,851`` and ``tools/plotting.py:18``) need matplotlib, which this image
does not have.  ``styles`` is provided because ``tools/plotting.py`` imports it.
"""

from __future__ import annotations

import numpy as np

styles = {
    "default": dict(title="", levels=10, cmap="jet"),
    "oil": dict(title="Oil saturation", levels=np.linspace(0 - 1e-7, 1 + 1e-7, 20), cmap="viridis"),
    "pperm": dict(title="Pre-perm", levels=np.linspace(-4, 4, 21), cmap="jet"),
    "perm": dict(title="Perm", levels=None, cmap="jet", locator="log"),
    "corr": dict(title="Correlations", levels=np.linspace(-1.00001, 1.00001, 20), cmap="bwr"),
    "NPV": dict(title="NPV", levels=12, cmap="inferno"),
}


class Plot2D:
    def _need_mpl(self):
        try:
            import matplotlib  # noqa: F401
        except ImportError as e:  # pragma: no cover - image has no matplotlib
            raise ImportError("plotting needs matplotlib, which is not installed") from e

    def plt_field(self, ax, Z, style="default", wells=True, argmax=False, colorbar=True, labels=True,
                  grid=False, finalize=True, **kwargs):
        self._need_mpl()
        kw = {k: v for k, v in styles.get(style, styles["default"]).items() if k in ("levels", "cmap")}
        kw.update({k: v for k, v in kwargs.items() if k in ("levels", "cmap", "alpha")})
        Z = np.asarray(Z).reshape(self.shape)
        if style == "oil":
            Z = 1 - Z
        X, Y = self.mesh
        cc = ax.contourf(X, Y, Z, **{k: v for k, v in kw.items() if v is not None})
        if kwargs.get("title") or styles.get(style, {}).get("title"):
            ax.set_title(kwargs.get("title", styles[style]["title"]))
        if wells:
            ax.plot(*self.inj_xy.T, "v", color="w", mec="k")
            ax.plot(*self.prd_xy.T, "^", color="w", mec="k")
        return cc

    def plt_production(self, ax, production, obs=None, legend_outside=True):
        self._need_mpl()
        hh = ax.plot(1 - np.asarray(production))
        if obs is not None:
            ax.plot(1 - np.asarray(obs), "*")
        ax.set_ylabel("Oil saturation (rel. production)")
        ax.set_xlabel("Time index")
        return hh

    def anim(self, wsats, prod, title="", **kwargs):
        self._need_mpl()
        raise NotImplementedError("animations are out of scope of the B200 hot path")
