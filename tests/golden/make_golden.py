"""Generate golden input/output vectors from the REFERENCE's own code.

Run in the build container only (needs ``/root/reference``; the GPU box never
runs this).  It imports the reference's ``notebooks/tools/{utils,geostat,
localization}.py`` (plotting-only imports stubbed) and AST-extracts the update
functions (``ens_update0``, ``ens_update0_loc``, ``IES``, ``ILES``, ``vect``,
``perm_transf``) from ``notebooks/HistoryMatch.py``, runs them on seeded
inputs and stores inputs + outputs in ``tests/golden/*.npz``.

    python tests/golden/make_golden.py
"""

import ast
import os
import sys
import types

import numpy as np
import scipy.linalg as sla

REF = "/root/reference/notebooks"
OUT = os.path.dirname(os.path.abspath(__file__))


class DotDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def load_reference():
    for name in ("adjustText", "matplotlib", "matplotlib.pyplot", "mpl_tools", "mpl_tools.misc"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["adjustText"].adjust_text = None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_tools.misc"].nRowCol = None
    sys.path.insert(0, REF)
    import tools.geostat as geostat
    import tools.localization as loc
    import tools.utils as utils

    src = open(os.path.join(REF, "HistoryMatch.py")).read()
    wanted = {"ens_update0", "ens_update0_loc", "IES", "ILES", "perm_transf", "rms"}
    ns = dict(np=np, sla=sla, center=utils.center, utils=utils, Dict=DotDict)
    # progress bars off
    _pb = utils.progbar
    utils.progbar = lambda *a, **k: _pb(*a, **{**k, "disable": True})
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            exec(compile(ast.Module([node], []), "HistoryMatch.py", "exec"), ns)
    return utils, geostat, loc, ns


def main():
    utils, geostat, loc, hm = load_reference()
    rng = np.random.RandomState(20261017)

    # ---- primitives --------------------------------------------------------------
    E = rng.randn(7, 5) * 3 + 1
    b = rng.randn(7, 3)
    A = rng.randn(6, 4)
    pts_a, pts_b = rng.rand(9, 2) * [2, 1], rng.rand(4, 2) * [2, 1]
    dist = np.linspace(-1.5, 1.5, 41)
    np.savez(
        os.path.join(OUT, "primitives.npz"),
        E=E, b=b, A=A, pts_a=pts_a, pts_b=pts_b, dist=dist,
        center_X=utils.center(E)[0], center_x=utils.center(E)[1],
        center_X_rescaled=utils.center(E, rescale=True)[0],
        cov=utils.cov(E, b), corr=utils.corr(E, b[:, 0]),
        rinv_tikh=utils.rinv(A, 0.1, tikh=True), rinv_trunc=utils.rinv(A, 0.1, tikh=False),
        pd_ab=loc.pairwise_distances(pts_a, pts_b), pd_aa=loc.pairwise_distances(pts_a),
        pd_periodic=loc.pairwise_distances(pts_a, pts_b, domain=(2, 1)),
        bump1=loc.bump(dist), bump_sharp=loc.bump(dist, 10), bump_soft=loc.bump(dist, 0.1),
        variogram=geostat.variogram_gauss(np.array([0.0, 1.0, 2.0]), 1, n=0.1, a=1),
    )

    # ---- prior: same seed / stream order as HistoryMatch.py:78,167,290 ------------------
    Nx, Ny, Lx, Ly = 20, 20, 2.0, 1.0
    # cell centres exactly as the drop-in Grid2D.mesh / oracle build them: the dense
    # Cholesky of the near-singular covariance amplifies last-bit differences in the
    # coordinates to ~1e-6 in the fields, so "bit-exact" means "given the same points".
    xs = np.linspace(0, Lx, Nx, endpoint=False) + Lx / Nx / 2
    ys = np.linspace(0, Ly, Ny, endpoint=False) + Ly / Ny / 2
    mesh = np.meshgrid(xs, ys, indexing="ij")
    np.random.seed(1)
    truth = geostat.gaussian_fields(mesh, 1, r=0.8)
    prior = geostat.gaussian_fields(mesh, 40, r=0.8)
    np.savez_compressed(os.path.join(OUT, "prior_20x20_seed1.npz"), truth=truth, prior=prior)

    # ---- updates -------------------------------------------------------------------------
    def case(N, M, p, nonlin=False):
        Ep = rng.randn(N, M)
        H = rng.randn(M, p) / np.sqrt(M)
        fwd = (lambda X: np.tanh(X @ H) + 0.1 * (X @ H)) if nonlin else (lambda X: X @ H)
        R12 = np.linalg.cholesky(0.3 * np.eye(p) + 0.05 * np.ones((p, p)))
        return dict(
            prior_ens=Ep, H=H, obs=fwd(rng.randn(1, M))[0] + 0.1 * rng.randn(p),
            perturbs=rng.randn(N, p) @ R12.T, decorr=sla.inv(R12.T),
        ), fwd

    out = {}
    for tag, (N, M, p) in dict(small=(12, 30, 8), wide=(9, 20, 14)).items():
        c, fwd = case(N, M, p, nonlin=True)
        kw = dict(prior_ens=c["prior_ens"], obs=c["obs"], perturbs=c["perturbs"], decorr=c["decorr"])
        # tapers from real geometry: parameters on a line, obs at a few points
        xy_prm = np.stack([np.linspace(0, 2, M), np.full(M, 0.5)], 1)
        xy_obs = np.stack([np.linspace(0.1, 1.9, p), np.full(p, 0.4)], 1)
        taper = loc.bump(loc.pairwise_distances(xy_prm, xy_obs) / 0.9)
        Eo = fwd(c["prior_ens"])
        for k, v in c.items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_taper"] = taper
        out[f"{tag}_obs_ens"] = Eo
        out[f"{tag}_ES"] = hm["ens_update0"](obs_ens=Eo, **kw)
        out[f"{tag}_LES"] = hm["ens_update0_loc"](obs_ens=Eo, taper=taper, **kw)
        Ei, st = hm["IES"](obs_ens=fwd, xStep=0.6, iMax=3, **kw)
        out[f"{tag}_IES"] = Ei
        out[f"{tag}_IES_E"] = np.array(st.E)
        out[f"{tag}_IES_Eo"] = np.array(st.Eo)
        Ei, st = hm["ILES"](obs_ens=fwd, taper=taper, xStep=0.6, iMax=3, **kw)
        out[f"{tag}_ILES"] = Ei
        out[f"{tag}_ILES_E"] = np.array(st.E)
    np.savez_compressed(os.path.join(OUT, "updates.npz"), **out)

    # ---- the notebook's own Gaussian-Gaussian self-checks (HistoryMatch.py:598-612, 811, 949, 1069)
    np.random.seed(7)
    d = 3
    Egg = np.sqrt(4 / 3) * np.random.randn(400, d)
    gg = dict(prior_ens=Egg, obs=4 * np.ones(d), decorr=1 / np.sqrt(4) * np.eye(d),
              perturbs=np.sqrt(4) * np.random.randn(*Egg.shape))
    post = hm["ens_update0"](**gg, obs_ens=Egg)
    post_loc = hm["ens_update0_loc"](**gg, obs_ens=Egg, taper=np.eye(d))
    ies, _ = hm["IES"](**gg, obs_ens=lambda x: x)
    iles, _ = hm["ILES"](**gg, obs_ens=lambda x: x, taper=np.eye(d))
    assert np.allclose(ies, post) and np.allclose(iles, post_loc)
    assert np.allclose(hm["ens_update0_loc"](**gg, obs_ens=Egg, taper=np.ones((d, d))), post)
    np.savez_compressed(os.path.join(OUT, "gauss_gauss.npz"), post=post, post_loc=post_loc, ies=ies, iles=iles, **gg)
    # ---- observation-error model: the notebook's own cell statements (HistoryMatch.py:243-247, 259) --------------
    src = open(os.path.join(REF, "HistoryMatch.py")).read()
    names = {"length_tmp", "corrs1well", "R1well", "R", "R12"}
    cell = dict(np=np, sla=sla, nTime=40, nPrd=4)
    for node in ast.parse(src).body:
        if not isinstance(node, ast.Assign) or len(node.targets) != 1:
            continue
        t = node.targets[0]
        tname = t.id if isinstance(t, ast.Name) else (t.value.id if isinstance(t, ast.Subscript) and isinstance(t.value, ast.Name) else None)
        if tname in names:
            exec(compile(ast.Module([node], []), "HistoryMatch.py", "exec"), cell)
    np.savez_compressed(os.path.join(OUT, "obs_error.npz"), R=cell["R"], R12=cell["R12"], nTime=40, nPrd=4)

    # ---- EnOpt host logic (tools/enopt.py: nabla_ens, backtracker, GD) on a quadratic ----------
    import tools.enopt as enopt

    utils.nCPU = 1  # serial map: the reference's own code path without pathos
    target = np.array([1.0, -2.0])
    obj = lambda u: -np.sum((u - target) ** 2)  # noqa: E731
    eo = {}
    for tag, precond in (("lls", False), ("precond", True)):
        np.random.seed(3)
        path, objs, info = enopt.GD(obj, np.zeros(2), enopt.nabla_ens(0.1, nEns=12, precond=precond), nIter=6, quiet=True)
        eo[f"{tag}_path"], eo[f"{tag}_objs"] = np.asarray(path, float), np.asarray(objs, float)
        eo[f"{tag}_grads"] = np.array([i["grad"] for i in info if "grad" in i])
    np.random.seed(5)
    eo["noise_scalar"] = utils.gaussian_noise(4, 3, 0.5)
    eo["noise_chol"] = utils.gaussian_noise(4, 3, np.linalg.cholesky(np.array([[2.0, 0.3, 0], [0.3, 1, 0.1], [0, 0.1, 0.5]])))
    np.savez_compressed(os.path.join(OUT, "enopt.npz"), **eo)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
