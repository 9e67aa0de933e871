"""world_size-2 NCCL test of the member-sharded analysis step (needs 2 GPUs; skipped otherwise).

Every rank simulates its own block of members (no communication), then the ES update runs
member-sharded: all_gather of the predicted data, all_to_all member rows -> parameter columns,
update on the local columns, all_to_all back.  The result must equal the
single-device update of the full ensemble (same CUDA kernels, columns are independent).
"""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, out):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=size, device_id=dev)
    from historymatching_b200 import analysis as ha
    from historymatching_b200 import dist as hd
    from historymatching_b200.workflow import HistoryMatchCase

    case = HistoryMatchCase(20, 20, 2.0, 1.0, 0.025, 6)
    N, M, p = 13, case.grid.M, case.p  # uneven member blocks (7 + 6), uneven column blocks
    rng = np.random.RandomState(3)  # same stream on every rank
    E = torch.as_tensor(np.clip(rng.randn(N, M), -2, 2) * 0.3, device=dev)
    obs = torch.as_tensor(rng.rand(p) * 0.3, device=dev)
    Z = torch.as_tensor(rng.randn(N, p), device=dev)
    pert = Z @ torch.as_tensor(case.R12.T.copy(), device=dev)
    dec = torch.as_tensor(case.decorr, device=dev)
    taper = torch.as_tensor(rng.rand(M, p), device=dev)
    lo, hi = hd.member_slice(N)
    clo, chi = hd.column_slice(M)

    Eo_loc, res = case.forward(E[lo:hi].contiguous())
    assert not res.status.any()
    Eo_full, _ = case.forward(E)  # reference: the whole ensemble on this device
    ok = torch.equal(hd.gather_members(Eo_loc, N), Eo_full)  # members are independent: bit-identical

    post = hd.sharded_update(ha.ens_update0, E[lo:hi].contiguous(), Eo_loc, N, obs=obs, perturbs=pert, decorr=dec)
    ref = ha.ens_update0(E, Eo_full, obs=obs, perturbs=pert, decorr=dec)
    ok = ok and torch.allclose(post, ref[lo:hi], rtol=1e-10, atol=1e-12)
    post_loc = hd.sharded_update(ha.ens_update0_loc, E[lo:hi].contiguous(), Eo_loc, N, obs=obs, perturbs=pert,
                                 decorr=dec, taper=taper[clo:chi].contiguous())
    ref_loc = ha.ens_update0_loc(E, Eo_full, obs=obs, perturbs=pert, decorr=dec, taper=taper)
    ok = ok and torch.allclose(post_loc, ref_loc[lo:hi], rtol=1e-10, atol=1e-12)
    # the sharded cycles (product API) against the single-device functions on the full ensemble
    fwd_loc = lambda X: case.forward(X.contiguous())[0]  # noqa: E731
    Zs = [rng.randn(N, p) for _ in range(2)]
    post_mda, st = hd.es_mda_sharded(fwd_loc, E[lo:hi].contiguous(), N, obs, case.R12, [2.0, 2.0], perturbs=Zs)
    ref_mda, _ = ha.es_mda(E, lambda X: case.forward(X)[0], obs.cpu().numpy(), case.R12, [2.0, 2.0], perturbs=Zs)
    ok_mda = torch.allclose(post_mda, ref_mda[lo:hi], rtol=1e-8, atol=1e-10) and len(st["Eo"]) == 2
    post_ies, st = hd.ies_sharded(fwd_loc, E[lo:hi].contiguous(), N, obs, pert, dec, xStep=0.6, iMax=3)
    ref_ies, _ = ha.IES(E, lambda X: case.forward(X)[0], obs, pert, dec, xStep=0.6, iMax=3)
    ok_ies = torch.allclose(post_ies, ref_ies[lo:hi], rtol=1e-8, atol=1e-10) and len(st["Eo"]) == 3
    post_iles, st = hd.iles_sharded(fwd_loc, E[lo:hi].contiguous(), N, obs, pert, dec, taper[clo:chi].contiguous(),
                                    xStep=0.6, iMax=2)
    ref_iles, _ = ha.ILES(E, lambda X: case.forward(X)[0], obs, pert, dec, taper, xStep=0.6, iMax=2)
    ok_iles = torch.allclose(post_iles, ref_iles[lo:hi], rtol=1e-8, atol=1e-10) and len(st["Eo"]) == 2
    with open(os.path.join(out, f"rank{rank}"), "w") as f:
        f.write(f"{int(ok)}{int(ok_mda)}{int(ok_ies)}{int(ok_iles)}")
    dist.destroy_process_group()


def test_sharded_forward_and_update_world2_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"rank{r}").read() for r in range(2)] == ["1111", "1111"]  # update, ES-MDA, IES, ILES


def test_sharded_cycles_single_process_equal_the_unsharded_functions():
    """Without a process group the sharded cycles are the single-device functions (world size 1: no exchange)."""
    import torch

    from historymatching_b200 import analysis as ha
    from historymatching_b200 import dist as hd
    from historymatching_b200.workflow import HistoryMatchCase

    case = HistoryMatchCase(20, 20, 2.0, 1.0, 0.025, 5)
    N, M, p = 12, case.grid.M, case.p
    rng = np.random.RandomState(5)
    dev = torch.device("cuda")
    E = torch.as_tensor(np.clip(rng.randn(N, M), -2, 2) * 0.3, device=dev)
    obs = torch.as_tensor(rng.rand(p) * 0.3, device=dev)
    pert = torch.as_tensor(rng.randn(N, p) @ case.R12.T, device=dev)
    dec = torch.as_tensor(case.decorr, device=dev)
    taper = torch.as_tensor(rng.rand(M, p), device=dev)
    fwd = lambda X: case.forward(X.contiguous())[0]  # noqa: E731
    post, st = hd.iles_sharded(fwd, E, N, obs, pert, dec, taper, xStep=0.5, iMax=2)
    ref, st_ref = ha.ILES(E, fwd, obs, pert, dec, taper, xStep=0.5, iMax=2)
    assert torch.allclose(post, ref, rtol=1e-10, atol=1e-12) and len(st["Eo"]) == 2
    assert torch.allclose(st["Eo"][1], st_ref.Eo[1], rtol=1e-10, atol=1e-12)
    post, st = hd.ies_sharded(fwd, E, N, obs, pert, dec, xStep=0.5, iMax=2)
    ref, _ = ha.IES(E, fwd, obs, pert, dec, xStep=0.5, iMax=2)
    assert torch.allclose(post, ref, rtol=1e-10, atol=1e-12) and len(st["Eo"]) == 2
    with pytest.raises(ValueError):
        hd.iles_sharded(fwd, E, N, obs, pert, dec, taper[:10], iMax=1)
