"""Ensemble-smoother analysis: host side of ``hm_es_update`` / ``hm_les_update`` /
``hm_ies_step`` (include/hm_b200.h).

Drop-in equivalents of the notebook's update cells, same names, argument
meaning and return values (ensemble axis first, float64):

* ``ens_update0``      - ``HistoryMatch.py:578-586``
* ``ens_update0_loc``  - ``HistoryMatch.py:774-797``
* ``IES``              - ``HistoryMatch.py:906-944``
* ``ILES``             - ``HistoryMatch.py:1007-1064``
* ``es_mda``           - not in the reference; this repo's definition on top of
  ``ens_update0`` (Emerick & Reynolds 2013).

numpy in -> numpy out (data staged through the GPU); torch CUDA tensors in ->
torch CUDA tensors out.  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class Stats(dict):
    """Attribute dict with the ``stats.E`` / ``stats.Eo`` lists of the reference (``HistoryMatch.py:908``)."""

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise _lib.HmError("no CUDA device visible: historymatching_b200 has no CPU fallback")
    return torch


def _dev(x, like=None):
    """float64 contiguous CUDA tensor of x (copying numpy input to the device)."""
    torch = _torch()
    dev = like.device if like is not None else None  # follow the ensemble's GPU, not the current device
    if isinstance(x, torch.Tensor):
        if dev is None:
            dev = x.device if x.device.type == "cuda" else "cuda"
        return x.to(device=dev, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=dev if dev is not None else "cuda")


def _back(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


def _p(t):
    return C.c_void_p(t.data_ptr())


def _ctx(t):
    ctx = _lib.Context.get(t.device.index if t.device.index is not None else 0)
    ctx.use_torch_stream()
    return ctx


def center(E, axis=0, rescale=False):
    """``utils.center`` (``tools/utils.py:10-28``) on the device, axis 0 only."""
    was_np = not _is_tensor(E)
    if axis != 0:
        raise ValueError("device center() supports axis=0 (the ensemble axis)")
    Ed = _dev(E)
    flat = Ed.reshape(Ed.shape[0], -1)
    torch = _torch()
    X = torch.empty_like(flat)
    mean = torch.empty(flat.shape[1], dtype=torch.float64, device=flat.device)
    ctx = _ctx(flat)
    _lib.check(ctx.lib.hm_center(ctx.handle, flat.shape[0], flat.shape[1], _p(flat), flat.shape[1], _p(X),
                                 flat.shape[1], _p(mean), int(bool(rescale))))
    X = X.reshape(Ed.shape)
    mean = mean.reshape(Ed.shape[1:]).squeeze() if Ed.ndim > 1 else mean.squeeze()
    return _back(X, was_np), _back(mean, was_np)


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


def _cov_corr(a, b, corr):
    was_np = not _is_tensor(a)
    A = _dev(a)
    B = _dev(b, like=A)
    N = A.shape[0]
    A2 = A.reshape(N, -1).contiguous()
    vec = B.ndim == 1
    B2 = (B.reshape(N, 1) if vec else B.reshape(N, -1)).contiguous()
    if B2.shape[0] != N:
        raise ValueError("a and b must have the same ensemble size (shape[0])")
    M, q = A2.shape[1], B2.shape[1]
    torch = _torch()
    out = torch.empty((M, q), dtype=torch.float64, device=A2.device)
    ctx = _ctx(A2)
    _lib.check(ctx.lib.hm_corr(ctx.handle, N, M, q, _p(A2), M, _p(B2), q, _p(out), int(corr)))
    out = out.reshape(A.shape[1:] + (() if vec else (q,)))
    return _back(out, was_np)


def cov(a, b):
    """``utils.cov`` (``tools/utils.py:31-39``) on the device: ``center(a).T @ center(b) / (N-1)``."""
    return _cov_corr(a, b, False)


def corr(a, b):
    """``utils.corr`` (``tools/utils.py:42-55``) on the device, e.g. the correlation field between the
    permeability ensemble ``a (N,M)`` and one well observation ``b (N,)`` (``HistoryMatch.py:738-748, 829-833``):
    ``cov / std(a) / std(b)`` (ddof=1), clipped to [-999, 999].  For a 2-D ``b (N,q)`` the result is ``(M,q)``."""
    return _cov_corr(a, b, True)


def ens_update0(prior_ens, obs_ens, obs, perturbs, decorr):
    """ES analysis update; see ``HistoryMatch.py:578-586``."""
    was_np = not _is_tensor(prior_ens)
    E = _dev(prior_ens).clone()
    Eo, y, pert, dec = _dev(obs_ens, E), _dev(obs, E), _dev(perturbs, E), _dev(decorr, E)
    N, M = E.shape
    p = y.shape[0]
    assert Eo.shape == (N, p) and pert.shape == (N, p) and dec.shape == (p, p)
    ctx = _ctx(E)
    _lib.check(ctx.lib.hm_es_update(ctx.handle, N, M, p, _p(E), M, _p(Eo), _p(y), _p(pert), _p(dec)))
    return _back(E, was_np)


def bump_taper(xy_prm, xy_obs, radius, sharpness=1.0):
    """``loc.bump(loc.pairwise_distances(xy_prm, xy_obs) / radius, sharpness)``
    (``tools/localization.py:9-92``, ``HistoryMatch.py:717,863``) on the device."""
    was_np = not _is_tensor(xy_prm)
    a = _dev(xy_prm)
    b = _dev(xy_obs, a)
    assert a.shape[1] == 2 and b.shape[1] == 2
    torch = _torch()
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float64, device=a.device)
    ctx = _ctx(a)
    _lib.check(ctx.lib.hm_taper_bump(ctx.handle, a.shape[0], b.shape[0], _p(a), _p(b), float(radius),
                                     float(sharpness), _p(out)))
    return _back(out, was_np)


def ens_update0_loc(prior_ens, obs_ens, obs, perturbs, decorr, taper):
    """Localised ES update; see ``HistoryMatch.py:774-797``."""
    was_np = not _is_tensor(prior_ens)
    E = _dev(prior_ens).clone()
    Eo, y, pert, dec, tap = _dev(obs_ens, E), _dev(obs, E), _dev(perturbs, E), _dev(decorr, E), _dev(taper, E)
    N, M = E.shape
    p = y.shape[0]
    assert tap.shape == (M, p)
    ctx = _ctx(E)
    _lib.check(ctx.lib.hm_les_update(ctx.handle, N, M, p, _p(E), M, _p(Eo), _p(y), _p(pert), _p(dec), _p(tap)))
    return _back(E, was_np)


def _recompose(ctx, x0, W, X0):
    """E = x0 + W @ X0 (``HistoryMatch.py:920, 944``) as one DMMA GEMM."""
    torch = _torch()
    N, M = X0.shape
    E = x0.expand(N, M).contiguous()
    _lib.check(ctx.lib.hm_dgemm(ctx.handle, 0, 0, N, M, N, 1.0, _p(W), N, _p(X0), M, 1.0, _p(E), M))
    return E


def IES(prior_ens, obs_ens, obs, perturbs, decorr, xStep=1.0, iMax=4):
    """Iterative ensemble smoother; see ``HistoryMatch.py:906-944``.

    ``obs_ens`` is a callable mapping an ensemble ``(N,M)`` to predicted data
    ``(N,p)``; it receives / may return numpy arrays or CUDA tensors.
    """
    torch = _torch()
    was_np = not _is_tensor(prior_ens)
    stats = Stats(E=[], Eo=[])
    E0 = _dev(prior_ens)
    y, pert, dec = _dev(obs, E0), _dev(perturbs, E0), _dev(decorr, E0)
    N, M = E0.shape
    p = y.shape[0]
    ctx = _ctx(E0)
    X0 = torch.empty_like(E0)
    x0 = torch.empty(M, dtype=torch.float64, device=E0.device)
    _lib.check(ctx.lib.hm_center(ctx.handle, N, M, _p(E0), M, _p(X0), M, _p(x0), 0))
    W = torch.eye(N, dtype=torch.float64, device=E0.device)
    for _ in range(iMax):
        E = _recompose(ctx, x0, W, X0)
        Eo = obs_ens(_back(E, was_np))
        stats.E.append(_back(E, was_np))
        stats.Eo.append(Eo)
        Eo_d = _dev(Eo, E0)
        assert Eo_d.shape == (N, p)
        ctx.use_torch_stream()
        _lib.check(ctx.lib.hm_ies_step(ctx.handle, N, p, _p(W), _p(Eo_d), _p(y), _p(pert), _p(dec), float(xStep)))
    return _back(_recompose(ctx, x0, W, X0), was_np), stats


def ILES(prior_ens, obs_ens, obs, perturbs, decorr, taper, xStep=1.0, iMax=4):
    """Localised iterative ensemble smoother; see ``HistoryMatch.py:1007-1064``.

    One ``N x N`` weight matrix per parameter (``M N^2`` doubles on the device); every
    iteration re-runs ``obs_ens`` and applies the tapered Gauss-Newton step of all
    parameters in one batched kernel (``hm_iles_step``).
    """
    torch = _torch()
    was_np = not _is_tensor(prior_ens)
    stats = Stats(E=[], Eo=[])
    E0 = _dev(prior_ens)
    y, pert, dec, tap = _dev(obs, E0), _dev(perturbs, E0), _dev(decorr, E0), _dev(taper, E0)
    N, M = E0.shape
    p = y.shape[0]
    assert tap.shape == (M, p)
    ctx = _ctx(E0)
    X0 = torch.empty_like(E0)
    x0 = torch.empty(M, dtype=torch.float64, device=E0.device)
    _lib.check(ctx.lib.hm_center(ctx.handle, N, M, _p(E0), M, _p(X0), M, _p(x0), 0))
    Ws = torch.eye(N, dtype=torch.float64, device=E0.device).repeat(M, 1, 1).contiguous()

    def recompose():
        E = torch.empty_like(E0)
        _lib.check(ctx.lib.hm_iles_recompose(ctx.handle, N, M, _p(Ws), _p(X0), _p(x0), _p(E)))
        return E

    for _ in range(iMax):
        E = recompose()
        Eo = obs_ens(_back(E, was_np))
        stats.E.append(_back(E, was_np))
        stats.Eo.append(Eo)
        Eo_d = _dev(Eo, E0)
        assert Eo_d.shape == (N, p)
        ctx.use_torch_stream()
        _lib.check(ctx.lib.hm_iles_step(ctx.handle, N, M, p, _p(Ws), _p(Eo_d), _p(y), _p(pert), _p(dec), _p(tap),
                                        float(xStep)))
    return _back(recompose(), was_np), stats


def es_mda(prior_ens, obs_ens, obs, R12, alphas, decorr=None, perturbs=None):
    """ES-MDA: ``len(alphas)`` forward runs + ES updates with inflated obs-error
    covariance ``alpha_i R`` (``sum 1/alpha_i == 1``).  Not in the reference; defined on
    ``ens_update0`` (``HistoryMatch.py:578-586``) with ``perturbs`` / ``decorr`` built as in
    ``HistoryMatch.py:638-639`` scaled by ``sqrt(alpha_i)`` / ``1/sqrt(alpha_i)``.

    ``perturbs``: optional list of standard-normal ``(N,p)`` blocks (one per pass);
    by default they are drawn from the legacy global numpy stream like the notebook does.
    """
    import scipy.linalg as sla

    alphas = np.asarray(alphas, float)
    if not np.isclose(np.sum(1 / alphas), 1.0):
        raise ValueError("ES-MDA inflation factors must satisfy sum(1/alpha) == 1")
    was_np = not _is_tensor(prior_ens)
    R12 = np.asarray(R12, float)
    if decorr is None:
        decorr = sla.inv(R12.T)
    decorr = np.asarray(decorr, float)
    stats = Stats(E=[], Eo=[])
    E = _dev(prior_ens)
    N, p = E.shape[0], len(obs)
    for i, a in enumerate(alphas):
        Eo = obs_ens(_back(E, was_np))
        stats.E.append(_back(E, was_np))
        stats.Eo.append(Eo)
        Z = np.random.randn(N, p) if perturbs is None else np.asarray(perturbs[i])
        E = ens_update0(E, _dev(Eo, E), obs, np.sqrt(a) * (Z @ R12.T), decorr / np.sqrt(a))
    return _back(E, was_np), stats
