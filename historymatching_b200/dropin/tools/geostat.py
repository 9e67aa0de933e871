"""Prior field generation - drop-in for reference ``notebooks/tools/geostat.py``.

``gaussian_fields`` makes the same numpy/scipy calls as the reference, on the
host, consuming the legacy global ``numpy.random`` stream with identical shapes:
fields are bit-identical for the same seed and points (the dense Cholesky of the
near-singular covariance is LAPACK-path sensitive, so it stays on the host).
It needs the dense ``M x M`` covariance and is limited to small grids, exactly
like the reference; ``gaussian_fields_separable`` is the scalable generator for
128^2 / 512^2 grids (same Gaussian variogram, distribution-matched, not bit-matched).
"""

from __future__ import annotations

import numpy as np
import scipy.linalg as sla
from numpy.random import randn


def variogram_gauss(xx, r, n=0, a=1 / 3):
    """Gaussian variogram with range ``r``, nugget ``n`` (``tools/geostat.py:10-31``).

    >>> variogram_gauss(np.array([0, 1, 2]), 1, n=0.1, a=1)
    array([0.        , 0.6689085 , 0.98351593])
    """
    gamma = 1 - np.exp(-(xx**2) / r**2 / a)
    gamma *= 1 - n
    gamma[xx != 0] += n
    return gamma


def vectorize(*XYZ):
    """``nDim`` coordinate arrays of equal shape -> ``(nPt, nDim)`` (``tools/geostat.py:34-41``)."""
    return np.stack(XYZ).reshape((len(XYZ), -1)).T


def dist_euclid(X):
    """All pairwise distances of the rows of ``X`` (``tools/geostat.py:44-47``)."""
    diff = X[:, None, :] - X
    return np.sqrt(np.sum(diff**2, axis=-1))


def gaussian_fields(pts, N=1, r=0.2):
    """``N`` Gaussian random fields on ``pts`` with a Gaussian variogram (``tools/geostat.py:86-99``)."""
    dists = dist_euclid(vectorize(*pts))
    Cov = 1 - variogram_gauss(dists, r)
    C12 = sla.cholesky(Cov + 1e-10 * np.eye(len(Cov)))
    return randn(N, len(C12)) @ C12


def _factor_1d(x, r, a=1 / 3):
    """Symmetric square root of the 1-D Gaussian covariance exp(-d^2/(r^2 a))."""
    d = x[:, None] - x[None, :]
    C = np.exp(-(d**2) / r**2 / a)
    w, V = np.linalg.eigh(C)
    return V * np.sqrt(np.clip(w, 0, None))


def gaussian_fields_separable(grid, N=1, r=0.2, rng=None, device=None):
    """Scalable version for regular grids: the Gaussian kernel is separable,
    ``Cov = Cx (x) Cy``, so ``field = Fx Z Fy^T`` with 1-D factors ``Fx, Fy``.

    ``grid`` needs ``Nx, Ny, Lx, Ly``.  Returns ``(N, Nx*Ny)``; a torch CUDA tensor
    when ``device`` is given (the two small matrix products then run on the GPU).
    """
    hx, hy = grid.Lx / grid.Nx, grid.Ly / grid.Ny
    Fx = _factor_1d((np.arange(grid.Nx) + 0.5) * hx, r)
    Fy = _factor_1d((np.arange(grid.Ny) + 0.5) * hy, r)
    rng = rng or np.random
    if device is None:
        Z = rng.standard_normal((N, grid.Nx, grid.Ny))
        return np.einsum("ia,nab,jb->nij", Fx, Z, Fy).reshape(N, -1)
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(int(rng.randint(2**31 - 1)))
    Z = torch.randn((N, grid.Nx, grid.Ny), dtype=torch.float64, device=device, generator=gen)
    return separable_apply(Fx, Z, Fy).reshape(N, -1)


def separable_apply(Fx, Z, Fy):
    """``Fx @ Z[n] @ Fy.T`` for every member ``n`` of the CUDA tensor ``Z (N, Nx, Ny)`` with the library's FP64
    tensor-core GEMM (``hm_dgemm``), three launches for the whole ensemble: the right factor on the ``(N Nx, Ny)``
    view, a swap of the two leading axes (``hm_swap01``), the left factor on the ``(Nx, N Ny)`` view, and the swap back."""
    import ctypes as C

    import torch

    from historymatching_b200 import _lib

    N, Nx, Ny = Z.shape
    dev = Z.device
    ctx = _lib.Context.get(dev.index if dev.index is not None else torch.cuda.current_device())
    ctx.use_torch_stream()
    Fx_d = torch.as_tensor(np.ascontiguousarray(Fx), dtype=torch.float64, device=dev)
    Fy_d = torch.as_tensor(np.ascontiguousarray(Fy), dtype=torch.float64, device=dev)
    Z = Z.contiguous()
    A = torch.empty_like(Z)
    B = torch.empty_like(Z)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    # A[n, a, :] = Z[n, a, :] Fy^T            (N Nx, Ny) x (Ny, Ny)^T
    _lib.check(ctx.lib.hm_dgemm(ctx.handle, 0, 1, N * Nx, Ny, Ny, 1.0, p(Z), Ny, p(Fy_d), Ny, 0.0, p(A), Ny))
    # B[a, n, :] = A[n, a, :]
    _lib.check(ctx.lib.hm_swap01(ctx.handle, N, Nx, Ny, p(A), p(B)))
    # A[i, (n, :)] = sum_a Fx[i, a] B[a, (n, :)]   (Nx, Nx) x (Nx, N Ny)
    _lib.check(ctx.lib.hm_dgemm(ctx.handle, 0, 0, Nx, N * Ny, Nx, 1.0, p(Fx_d), Nx, p(B), N * Ny, 0.0, p(A), N * Ny))
    # out[n, i, :] = A[i, n, :]
    _lib.check(ctx.lib.hm_swap01(ctx.handle, Nx, N, Ny, p(A), p(B)))
    return B
