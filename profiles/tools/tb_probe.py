"""Time the transport phase of short forward runs for several grid / cluster shapes (the A/B table of profiles/README.md).

    python profiles/tools/tb_probe.py 128,128,1024,sat_block=7 512,512,60,sat_block=7,tb_cluster_rows=16,tb_halo=16

Each argument: Nx,Ny,members[,key=value ...] with the integer keywords of run_ensemble (sat_block, tb_cluster_rows, tb_halo).
"""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from historymatching_b200.sim import GridSpec, run_ensemble
from historymatching_b200.workflow import notebook_wells

def run(Nx, Ny, N, nT=2, **kw):
    grid = GridSpec(Nx, Ny, 2.0, 1.0)
    cells, rates = notebook_wells(grid)
    rng = np.random.RandomState(0)
    x = rng.randn(N, Nx // 8 + 2, Ny // 8 + 2)
    x = np.kron(x, np.ones((8, 8)))[:, :Nx, :Ny]
    for _ in range(4):
        x = 0.25 * (np.roll(x, 1, 1) + np.roll(x, -1, 1) + np.roll(x, 1, 2) + np.roll(x, -1, 2))
    K = torch.as_tensor(0.1 + np.exp(2.0 * x.reshape(N, -1) / x.std()), device="cuda")
    S0 = torch.zeros(grid.M, dtype=torch.float64, device="cuda")
    for rep in range(2):
        res = run_ensemble(grid, K, cells, rates, S0, 0.025, nT, want_substeps=True, **kw)
    st = res.stats
    nts = float(res.substeps.double().mean())
    sat = st["phase_ms"]["saturation"] / nT
    upd = st["sat_cell_updates"] / nT if st["sat_cell_updates"] else N * grid.M * nts
    print(json.dumps(dict(grid=[Nx, Ny], N=N, kw=kw, sat_ms_per_step=round(sat, 3), nts=nts,
                          cluster=st["sat_tb_cluster"], strips=st["sat_tb_strips"], halo=st["sat_tb_halo"],
                          resident=st["sat_resident_ctas"], Gcell_updates_per_s=round(upd / sat / 1e6, 2),
                          useful_Gcells_per_s=round(N * grid.M * nts / sat / 1e6, 2))), flush=True)

if __name__ == "__main__":
    for spec in sys.argv[1:]:
        a = spec.split(",")
        kw = dict(x.split("=") for x in a[3:])
        run(int(a[0]), int(a[1]), int(a[2]), **{k: int(v) for k, v in kw.items()})
