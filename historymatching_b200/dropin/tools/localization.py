"""Localization tools - drop-in for reference ``notebooks/tools/localization.py``.

``pairwise_distances`` and ``bump`` keep the reference behaviour exactly (any
number of dimensions, periodic domains, 1-D input = one point).  For the
history-matching case (2-D points, ``bump(dist / radius)``) the fused device
version is ``historymatching_b200.analysis.bump_taper`` (``hm_taper_bump``).
"""

from __future__ import annotations

import itertools

import numpy as np


def pairwise_distances(A, B=None, domain=None):
    """Euclidean distances between the points of ``A`` and ``B`` -> ``(nA, nB)``.

    ``domain``: edge lengths of a periodic hyper-rectangle (``tools/localization.py:9-83``).
    """
    A = np.atleast_2d(A)
    B = A if B is None else np.atleast_2d(B)
    assert A.shape[1] == B.shape[1], "The last axis of A and B must have equal length."
    d = A[:, None] - B
    if domain:
        L = np.reshape(domain, (1, 1, -1))
        d = abs(d)
        d = np.minimum(d, L - d)
    return np.sqrt((d * d).sum(axis=-1)).reshape(len(A), len(B))


def bump(distances, sharpness=1):
    """``exp(1 - 1/(1-x^2))**sharpness`` on ``|x| < 1``, zero outside (``tools/localization.py:86-92``)."""
    inside = np.abs(distances) < 1
    x = distances[inside]
    coeffs = np.zeros_like(distances)
    coeffs[inside] = np.exp(1 - 1 / (1 - x * x)) ** sharpness
    return coeffs


def rectangular_partitioning(shape, steps, do_ind=True):
    """Rectangular batches of an N-D grid (``tools/localization.py:95-145``; unused by the notebooks)."""
    assert len(shape) == len(steps)
    counts = [round(n / d) for n, d in zip(shape, steps)]
    edges = [np.array_split(np.arange(n), c) for n, c in zip(shape, counts)]
    batches = [[ii.flatten() for ii in np.meshgrid(*e, indexing="ij")] for e in itertools.product(*edges)]
    if do_ind:
        batches = [np.ravel_multi_index(b, shape) for b in batches]
    return batches
