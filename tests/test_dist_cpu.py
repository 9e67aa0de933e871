"""world_size-2 gloo tests (CPU) of the member <-> column resharding used by the
multi-GPU analysis step; the update itself is stood in by the oracle."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import analysis as oa


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, N, M, p, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from historymatching_b200 import dist as hd

    rng = np.random.RandomState(0)  # same data on every rank
    E, Eo = rng.randn(N, M), rng.randn(N, p)
    obs, pert, dec = rng.randn(p), rng.randn(N, p), np.linalg.cholesky(np.eye(p) + 0.1 * np.ones((p, p)))
    taper = rng.rand(M, p)
    lo, hi = hd.member_slice(N)
    E_loc, Eo_loc = torch.as_tensor(E[lo:hi]), torch.as_tensor(Eo[lo:hi])

    # round trip
    cols = hd.members_to_columns(E_loc, N)
    clo, chi = hd.column_slice(M)
    assert torch.equal(cols, torch.as_tensor(E[:, clo:chi]))
    assert torch.equal(hd.columns_to_members(cols, N, M), E_loc)
    assert torch.equal(hd.gather_members(Eo_loc, N), torch.as_tensor(Eo))

    def es(Ec, Eof, **kw):
        return torch.as_tensor(oa.ens_update0(Ec.numpy(), Eof.numpy(), **kw))

    def les(Ec, Eof, taper, **kw):
        return torch.as_tensor(oa.ens_update0_loc(Ec.numpy(), Eof.numpy(), taper=taper, **kw))

    post = hd.sharded_update(es, E_loc, Eo_loc, N, obs=obs, perturbs=pert, decorr=dec)
    post_loc = hd.sharded_update(les, E_loc, Eo_loc, N, obs=obs, perturbs=pert, decorr=dec, taper=taper[clo:chi])
    ref = oa.ens_update0(E, Eo, obs, pert, dec)
    ref_loc = oa.ens_update0_loc(E, Eo, obs, pert, dec, taper)
    ok = np.allclose(post.numpy(), ref[lo:hi], rtol=1e-10, atol=1e-12)
    ok_loc = np.allclose(post_loc.numpy(), ref_loc[lo:hi], rtol=1e-10, atol=1e-12)
    # the sharded ES-MDA cycle (host logic + collectives); the CUDA update is stood in by the oracle on CPU
    from historymatching_b200 import analysis as ha

    ha.ens_update0 = lambda Ec, Eof, obs, perturbs, decorr: torch.as_tensor(  # noqa: E731
        oa.ens_update0(Ec.numpy(), Eof.numpy(), np.asarray(obs), np.asarray(perturbs), np.asarray(decorr)))
    H = rng.randn(M, p) / 3
    R12 = np.linalg.cholesky(0.01 * (np.eye(p) + 0.3 * np.ones((p, p))))
    Zs = [rng.randn(N, p) for _ in range(2)]
    post_mda, st = hd.es_mda_sharded(lambda X: torch.tanh(X @ torch.as_tensor(H)), E_loc, N, obs, R12, [2.0, 2.0],
                                     perturbs=Zs)
    Er = E.copy()
    for Z, a in zip(Zs, [2.0, 2.0]):
        Er = oa.ens_update0(Er, np.tanh(Er @ H), obs, np.sqrt(a) * (Z @ R12.T), np.linalg.inv(R12.T) / np.sqrt(a))
    ok_mda = np.allclose(post_mda.numpy(), Er[lo:hi], rtol=1e-10, atol=1e-12) and len(st["Eo"]) == 2
    # the seeded perturbations are the same block on every rank
    zz = hd._normal_block(N, p, 5, "cpu")
    gathered = [torch.empty_like(zz) for _ in range(size)]
    dist.all_gather(gathered, zz)
    ok_mda = ok_mda and all(torch.equal(g, zz) for g in gathered)
    with open(os.path.join(out, f"rank{rank}"), "w") as f:
        f.write(str(int(ok and ok_loc and ok_mda)))
    dist.destroy_process_group()


def test_sharded_update_world2_gloo(tmp_path):
    size = 2
    mp.spawn(_worker, args=(size, _free_port(), 7, 11, 5, str(tmp_path)), nprocs=size, join=True)
    assert [open(tmp_path / f"rank{r}").read() for r in range(size)] == ["1", "1"]


def test_sharded_update_world3_gloo_even_blocks(tmp_path):
    """Three ranks, equal member blocks (the all_gather_into_tensor path) and uneven column blocks."""
    size = 3
    mp.spawn(_worker, args=(size, _free_port(), 9, 10, 4, str(tmp_path)), nprocs=size, join=True)
    assert [open(tmp_path / f"rank{r}").read() for r in range(size)] == ["1", "1", "1"]


def test_single_process_is_identity():
    from historymatching_b200 import dist as hd

    E = torch.randn(4, 6, dtype=torch.float64)
    assert hd.members_to_columns(E, 4) is E and hd.columns_to_members(E, 4, 6) is E
    assert hd.member_slice(10, 1, 3) == (4, 7) and hd.member_slice(10, 2, 3) == (7, 10)


def test_block_partitions_cover_without_overlap():
    """member_slice / column_slice: contiguous blocks in rank order, sizes differing by at most one, also when there are
    fewer items than ranks (empty blocks at the end)."""
    from historymatching_b200 import dist as hd

    for size in range(1, 9):
        for n in range(0, 40):
            blocks = [hd.member_slice(n, r, size) for r in range(size)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            lens = [hi - lo for lo, hi in blocks]
            assert max(lens) - min(lens) <= 1 and lens == sorted(lens, reverse=True)
            assert blocks == [hd.column_slice(n, r, size) for r in range(size)]
