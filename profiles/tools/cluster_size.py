"""Per-sub-step time of k_sat_cluster against the cluster size (tiles per member): isolates the cost of the DSMEM halo exchange."""
import numpy as np, torch, sys
sys.path.insert(0, '.')
from historymatching_b200.sim import GridSpec, run_ensemble
from historymatching_b200.workflow import notebook_wells
dev = torch.device('cuda')
for Nx, N in ((16, 8192), (32, 4096), (64, 2048), (128, 1024), (128, 960)):
    grid = GridSpec(Nx=Nx, Ny=128, Lx=2.0, Ly=1.0)
    wc, wr = notebook_wells(grid)
    g = torch.Generator(device=dev).manual_seed(1)
    x = 0.3 * torch.randn(N, grid.M, dtype=torch.float64, device=dev, generator=g)
    K = 0.1 + torch.exp(5 * x)
    for rep in range(2):
        res = run_ensemble(grid, K, torch.as_tensor(wc, device=dev), torch.as_tensor(wr, device=dev),
                           torch.zeros(grid.M, dtype=torch.float64, device=dev), 0.025, 1, want_substeps=True, sat_block=2)
    nts = int(res.substeps.max()); ms = res.stats['phase_ms']['saturation']
    tiles = Nx // 16
    print(f"Nx={Nx:4d} tiles/member={tiles:2d} members={N:5d} Nts={nts:4d} sat={ms:8.3f} ms  -> {ms*1e3/nts:8.3f} us per sub-step of the whole batch; "
          f"CTAs={N*tiles}", flush=True)
