"""Exercise the analysis / small-grid kernels once at representative sizes (for ncu captures):

    ncu --set full -k regex:<kernel> -c 1 -o out python profiles/tools/run_kernels.py <what>

what: dgemm | les | iles | iles200 | small | corr
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from historymatching_b200 import analysis as ha  # noqa: E402
from historymatching_b200.workflow import HistoryMatchCase  # noqa: E402
from oracle import analysis as oa  # noqa: E402

what = sys.argv[1]
rng = np.random.RandomState(0)
dev = "cuda"


def case(N, M, p):
    E = torch.as_tensor(rng.randn(N, M), device=dev)
    H = rng.randn(M, p) / np.sqrt(M)
    R, R12 = oa.obs_error_model(p // 4, 4)
    Eo = torch.tanh(E @ torch.as_tensor(H, device=dev))
    xy_prm = rng.rand(M, 2) * [2, 1]
    xy_obs = np.tile(rng.rand(4, 2) * [2, 1], (p // 4, 1))
    taper = ha.bump_taper(torch.as_tensor(xy_prm, device=dev), torch.as_tensor(xy_obs, device=dev), 1.2)
    kw = dict(obs=torch.as_tensor(rng.randn(p) * 0.1, device=dev), perturbs=torch.as_tensor(rng.randn(N, p) @ R12.T, device=dev),
              decorr=torch.as_tensor(np.linalg.inv(R12.T), device=dev))
    return E, Eo, H, taper, kw


for rep in range(2):  # the second pass is the warm one
    if what == "dgemm":      # IES recomposition W @ X0 at config C: 1024 x 16384 x 1024
        W = torch.as_tensor(rng.randn(1024, 1024), device=dev)
        X0 = torch.as_tensor(rng.randn(1024, 16384), device=dev)
        from historymatching_b200 import _lib

        ha._recompose(_lib.Context.get(0), X0[0].contiguous(), W, X0)
    elif what == "les":      # localised ES at config C size
        E, Eo, H, taper, kw = case(1024, 16384, 160)
        ha.ens_update0_loc(E, Eo, taper=taper, **kw)
    elif what in ("iles", "iles200"):
        N = 40 if what == "iles" else 200
        E, Eo, H, taper, kw = case(N, 400, 160)
        Hd = torch.as_tensor(H, device=dev)
        ha.ILES(E, lambda X: torch.tanh(X @ Hd), taper=taper, xStep=0.4, iMax=1, **kw)
    elif what == "small":    # notebook default: 20 x 20, 40 members, 40 steps, one fused launch
        c = HistoryMatchCase(20, 20, 2.0, 1.0, 0.025, 40)
        c.forward(torch.as_tensor(0.3 * rng.randn(40, 400), device=dev))
    elif what == "corr":     # correlation of every cell of a 128^2 ensemble with 4 well series
        a = torch.as_tensor(rng.randn(1024, 16384), device=dev)
        b = torch.as_tensor(rng.randn(1024, 4), device=dev)
        ha.corr(a, b)
    torch.cuda.synchronize()
print("done", what)
