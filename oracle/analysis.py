"""CPU oracle: ensemble-smoother analysis (ES, LES, IES, ILES, ES-MDA) and its
helpers.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

numpy/scipy restatement of the reference algorithms, PINNED against golden
vectors produced by the reference's own code (``tests/golden/make_golden.py``).
Ensemble axis first everywhere ("transposed" convention, ``HistoryMatch.py:574-575``).
"""

from __future__ import annotations

import numpy as np
import scipy.linalg as sla


class Stats(dict):
    """Attribute dict standing in for ``struct_tools.DotDict`` (``HistoryMatch.py:908``)."""

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


# ---- primitives: tools/utils.py -------------------------------------------------
def center(E, axis=0, rescale=False):
    """Anomalies and mean.  Follows ``tools/utils.py:10-28``."""
    mean = np.mean(E, axis=axis, keepdims=True)
    X = E - mean
    if rescale:
        n = E.shape[axis]
        X *= np.sqrt(n / (n - 1))
    return X, mean.squeeze()


def cov(a, b):
    """Sample cross-covariance.  Follows ``tools/utils.py:31-39``."""
    return center(a)[0].T @ center(b)[0] / (len(b) - 1)


def corr(a, b):
    """Sample cross-correlation, clipped to +-999.  Follows ``tools/utils.py:42-55``."""
    C = cov(a, b)
    sa = np.std(a.T, axis=-1, ddof=1)
    sb = np.std(b, axis=0, ddof=1, keepdims=True)
    return (C / sa / sb).clip(-999, 999)


def rinv(A, reg, tikh=True, nMax=None):
    """Tikhonov / truncated-SVD pseudo-inverse.  Follows ``tools/utils.py:70-90``."""
    U, s, VT = sla.svd(A, full_matrices=False)
    thr = reg * s[0]
    if tikh:
        s1 = s / (s**2 + thr**2)
    else:
        s1 = np.where(s >= thr, 1 / np.where(s >= thr, s, 1.0), 0.0)
    if nMax:
        s1[nMax:] = 0
    return (VT.T * s1) @ U.T


# ---- localisation: tools/localization.py ------------------------------------------
def pairwise_distances(A, B=None, domain=None):
    """Euclidean distances ``(mA, mB)``, optionally periodic.  ``tools/localization.py:9-83``."""
    A = np.atleast_2d(A)
    B = A if B is None else np.atleast_2d(B)
    d = A[:, None, :] - B[None, :, :]
    if domain:
        span = np.reshape(domain, (1, 1, -1))
        d = np.abs(d)
        d = np.minimum(d, span - d)
    return np.sqrt((d * d).sum(-1))


def bump(distances, sharpness=1):
    """Compactly supported taper ``exp(1 - 1/(1-x^2))**sharpness``.  ``tools/localization.py:86-92``."""
    distances = np.asarray(distances, float)
    out = np.zeros_like(distances)
    inside = np.abs(distances) < 1
    x = distances[inside]
    out[inside] = np.exp(1 - 1 / (1 - x * x)) ** sharpness
    return out


# ---- prior: tools/geostat.py --------------------------------------------------------
def variogram_gauss(xx, r, n=0, a=1 / 3):
    """``tools/geostat.py:10-31``."""
    xx = np.asarray(xx, float)
    g = (1 - np.exp(-(xx**2) / r**2 / a)) * (1 - n)
    g[xx != 0] += n
    return g


def gaussian_fields(pts, N=1, r=0.2):
    """Dense-covariance Gaussian random fields.  ``tools/geostat.py:86-99``.

    Consumes ``N*M`` normals from the legacy global numpy stream, like the
    reference (``numpy.random.randn``).
    """
    X = np.stack(pts).reshape((len(pts), -1)).T
    diff = X[:, None, :] - X
    dists = np.sqrt(np.sum(diff**2, axis=-1))
    Cov = 1 - variogram_gauss(dists, r)
    C12 = sla.cholesky(Cov + 1e-10 * np.eye(len(Cov)))
    return np.random.randn(N, len(C12)) @ C12


# ---- updates: HistoryMatch.py ---------------------------------------------------------
def ens_update0(prior_ens, obs_ens, obs, perturbs, decorr):
    """Stochastic ES update in whitened obs space.  ``HistoryMatch.py:578-586``."""
    N = len(prior_ens)
    X = center(prior_ens)[0]
    S = center(obs_ens)[0] @ decorr
    D = (obs - obs_ens - perturbs) @ decorr
    C = S.T @ S + (N - 1) * np.eye(len(obs))
    return prior_ens + D @ sla.pinv(C) @ S.T @ X


def ens_update0_loc(prior_ens, obs_ens, obs, perturbs, decorr, taper):
    """Per-parameter local analysis.  ``HistoryMatch.py:774-797``."""
    N, M = prior_ens.shape
    X = center(prior_ens)[0]
    S = center(obs_ens)[0] @ decorr
    D = (obs - obs_ens - perturbs) @ decorr
    out = np.array(prior_ens, dtype=float, copy=True)
    for i in range(M):
        c = np.sqrt(taper[i])
        jj = c > 1e-2
        if not np.any(jj):
            continue
        Si = S[:, jj] * c[jj]
        Di = D[:, jj] * c[jj]
        Ci = Si.T @ Si + (N - 1) * np.eye(int(jj.sum()))
        out[:, i] = prior_ens[:, i] + Di @ sla.pinv(Ci) @ Si.T @ X[:, i]
    return out


def _gn_cov(Y0, N):
    """Gauss-Newton posterior covariance of the weights.  ``HistoryMatch.py:935-938``."""
    excess = Y0.shape[0] - Y0.shape[1]
    V, s, _ = sla.svd(Y0, full_matrices=(excess > 0))
    spec = 1 / (N - 1 + np.pad(s**2, (0, max(0, excess))))
    return (V * spec) @ V.T


def IES(prior_ens, obs_ens, obs, perturbs, decorr, xStep=1.0, iMax=4):
    """Gauss-Newton iterative ES in the ensemble subspace.  ``HistoryMatch.py:906-944``."""
    stats = Stats(E=[], Eo=[])
    N = len(prior_ens)
    y = obs @ decorr
    D = perturbs @ decorr
    I = np.eye(N)
    X0, x0 = center(prior_ens)
    W = I
    for _ in range(iMax):
        E = x0 + W @ X0
        Eo = obs_ens(E)
        stats.E.append(E)
        stats.Eo.append(Eo)
        Eo = Eo @ decorr
        Y0 = center(sla.pinv(W))[0] @ Eo
        grad = (y - D - Eo) @ Y0.T + (N - 1) * (I - W)
        W = W + xStep * (grad @ _gn_cov(Y0, N))
    return x0 + W @ X0, stats


def ILES(prior_ens, obs_ens, obs, perturbs, decorr, taper, xStep=1.0, iMax=4):
    """Localised IES: one weight matrix per parameter.  ``HistoryMatch.py:1007-1064``."""
    stats = Stats(E=[], Eo=[])
    N, M = prior_ens.shape
    I = np.eye(N)
    X0, x0 = center(prior_ens)
    Ws = [I] * M

    def recompose(Ws):
        return x0 + np.array([Ws[i] @ X0[:, i] for i in range(M)]).T

    for _ in range(iMax):
        E = recompose(Ws)
        Eo = obs_ens(E)
        stats.E.append(E)
        stats.Eo.append(Eo)
        S = center(Eo @ decorr)[0]
        D = (obs - Eo - perturbs) @ decorr
        new = []
        for i in range(M):
            c = np.sqrt(taper[i])
            jj = c > 1e-2
            Wi = Ws[i]
            dW = 0
            if np.any(jj):
                Si = S[:, jj] * c[jj]
                Di = D[:, jj] * c[jj]
                Y0 = center(sla.pinv(Wi))[0] @ Si
                grad = Di @ Y0.T + (N - 1) * (I - Wi)
                dW = grad @ _gn_cov(Y0, N)
            new.append(Wi + xStep * dW)
        Ws = new
    return recompose(Ws), stats


def es_mda(prior_ens, obs_ens, obs, R12, alphas, decorr=None, perturbs=None):
    """ES-MDA (Emerick & Reynolds 2013).  NOT IN THE REFERENCE - this repo's
    definition: for each ``alpha_i`` (``sum 1/alpha_i = 1``)
    re-run the forward model and apply ``ens_update0`` with ``R -> alpha_i R``, i.e.
    perturbations ``sqrt(alpha_i) randn(N,p) @ R12.T`` (``HistoryMatch.py:638``) and
    ``decorr/sqrt(alpha_i)`` (``HistoryMatch.py:639``).

    ``obs_ens`` is a callable ``E -> (N,p)``.  ``perturbs`` may be a list of
    pre-drawn ``(N,p)`` standard-normal blocks (one per pass) for reproducible
    parity tests; otherwise the legacy global stream is consumed.
    """
    alphas = np.asarray(alphas, float)
    assert np.isclose(np.sum(1 / alphas), 1.0)
    if decorr is None:
        decorr = sla.inv(R12.T)
    E = np.array(prior_ens, float)
    N, p = len(E), len(obs)
    stats = Stats(E=[], Eo=[])
    for i, a in enumerate(alphas):
        Eo = obs_ens(E)
        stats.E.append(E)
        stats.Eo.append(Eo)
        Z = np.random.randn(N, p) if perturbs is None else perturbs[i]
        E = ens_update0(E, Eo, obs, np.sqrt(a) * (Z @ R12.T), decorr / np.sqrt(a))
    return E, stats


def obs_error_model(nTime, nPrd, length_tmp=2, var=1e-2):
    """``R``, ``R12`` of ``HistoryMatch.py:243-259``."""
    c = np.exp(-np.arange(nTime) / length_tmp)
    c[c < 1e-2] = 0
    R = np.kron(var * sla.toeplitz(c), np.eye(nPrd))
    return R, sla.cholesky(R, lower=True)
