"""Experiment: two host threads, two contexts / streams, half the ensemble each, against one call."""
import sys, time, threading
import numpy as np, torch
sys.path.insert(0, '.')
from historymatching_b200 import _lib
from historymatching_b200.sim import GridSpec, run_ensemble
from historymatching_b200.workflow import notebook_wells
from historymatching_b200.dropin.tools import geostat
dev = torch.device('cuda')
grid = GridSpec(Nx=128, Ny=128, Lx=2.0, Ly=1.0)
N, nT = 1024, 8
E = geostat.gaussian_fields_separable(grid, N, r=0.8, rng=np.random.RandomState(100), device=dev)
K = 0.1 + torch.exp(5 * E)
wc, wr = notebook_wells(grid)
wc, wr = torch.as_tensor(wc, device=dev), torch.as_tensor(wr, device=dev)
S0 = torch.zeros(grid.M, dtype=torch.float64, device=dev)

def run(Kpart, ctx, stream, out, i):
    with torch.cuda.stream(stream):
        out[i] = run_ensemble(grid, Kpart, wc, wr, S0, 0.025, nT, ctx=ctx)

def timed(nl):
    ctxs = [_lib.Context(0) for _ in range(nl)]
    streams = [torch.cuda.Stream() for _ in range(nl)]
    parts = torch.chunk(K, nl)
    best = 1e9
    for rep in range(3):
        out = [None] * nl
        torch.cuda.synchronize(); t0 = time.perf_counter()
        th = [threading.Thread(target=run, args=(parts[i], ctxs[i], streams[i], out, i)) for i in range(nl)]
        [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    S = torch.cat([o.S_last for o in out])
    return best, S

t1, S1 = timed(1)
for nl in (2, 3, 4):
    t, S = timed(nl)
    print(f"lanes={nl}: {t*1e3:8.1f} ms vs single {t1*1e3:8.1f} ms  ({t1/t:5.3f}x)  max|dS|={float((S-S1).abs().max()):.2e}", flush=True)
