"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: section sharding and the flat
gradient all-reduce that makes `sample_plane` data-parallel."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import multigpu
    torch.manual_seed(0)
    model = multigpu.DataParallelSPHNCA(16, 0.5, normalize_grad=True)
    # rank-dependent gradients; model.2.bias left without a gradient on rank 1 only
    for i, p in enumerate(model.parameters()):
        if not (rank == 1 and i == 3):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    nbytes = multigpu.allreduce_gradients(model.parameters())
    grads = [p.grad.clone() for p in model.parameters()]
    model.normalize_grads()            # all-reduces again (idempotent on equal grads), then normalises
    norms = [float(p.grad.norm()) for p in model.parameters()]
    ret[rank] = (nbytes, [float(g.flatten()[0]) for g in grads], norms)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] == ret[1][0] == 21025 * 4                      # one flat 84,100-byte buffer
    expect = [(1 * (i + 1) + 2 * (i + 1)) / 2 for i in range(3)] + [(1 * 4 + 0) / 2]
    for r in range(world):
        assert ret[r][1] == expect                                  # average over ranks, identical on both
        assert all(abs(n - 1.0) < 1e-5 for n in ret[r][2])          # normalize_grads ran after the reduce


def test_shard_sections():
    import multigpu
    sections = (5, 7, 3, 9, 2)
    got = [multigpu.shard_sections(sections, 2, r) for r in range(2)]
    assert got == [(0, 3, 0, 15), (3, 5, 15, 26)]
    got8 = [multigpu.shard_sections((4,) * 8, 8, r) for r in range(8)]
    assert [g[2] for g in got8] == [4 * r for r in range(8)] and all(g[1] - g[0] == 1 for g in got8)
    # more ranks than sections: trailing ranks get empty shards
    assert multigpu.shard_sections((4, 4), 4, 3) == (2, 2, 8, 8)


def _slab_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import math
    import multigpu
    from oracle import sphnca_oracle as O
    from tests.util import jittered_lattice
    gen = torch.Generator().manual_seed(0)
    n = 40
    x = torch.nn.functional.pad(jittered_lattice(n, 2, gen), (0, 1))
    h = 3.2 * 2 / n
    dims = (math.ceil(2 / h), math.ceil(2 / h), 3)
    A = torch.rand(x.shape[0], 4, generator=gen)
    # global reference: blur of A over the whole cloud (oracle)
    og, pairs, v = O.prepare(x, (x.shape[0],), h, dims)
    ref = O.blur(pairs, v, A)
    # sharded: each rank computes the same operator on [interior | halo] and keeps its interior
    part = multigpu.SlabPartition(x, h, dims, axis=1, rank=rank, world=world)
    x_loc = part.local_from_global(x)
    A_loc = part.local_from_global(A)
    v_loc = part.local_from_global(v[:, None].contiguous())[:, 0]
    lg = O.build_hashgrid(x_loc, (part.n_local,), h, dims)
    lp = O.make_pairs(x_loc, lg)
    out = O.blur(lp, v_loc, A_loc)[:part.n_own]
    err = float((out - ref[part.own_idx]).abs().max() / ref.abs().max())
    ret[rank] = (err, part.n_own, part.n_halo_prev, part.n_halo_next, part.row_lo, part.row_hi)
    dist.destroy_process_group()


def test_slab_partition_halo_world2():
    """One cell layer of halo from the ring neighbours reproduces the global neighbour sums on every interior."""
    world = 2
    port = _free_port()
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_slab_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][1] + ret[1][1] == 40 * 40                        # every particle owned exactly once
    for r in range(world):
        err, n_own, hp, hn, lo, hi = ret[r]
        assert err < 1e-12, err                                      # identical neighbour sets -> identical fp64 sums
        assert hp > 0 and hn > 0 and hi > lo
    assert ret[0][5] == ret[1][4]                                    # contiguous row ranges


def test_counter_uniform_is_a_function_of_the_global_id():
    import multigpu
    ids = torch.arange(0, 100000, dtype=torch.int64)
    u = multigpu.counter_uniform(ids, 3, seed=5)
    assert u.shape == (100000, 3) and u.dtype == torch.float32
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0
    assert abs(float(u.mean()) - 0.5) < 5e-3 and abs(float((u <= 0.5).float().mean()) - 0.5) < 5e-3
    perm = torch.randperm(100000, generator=torch.Generator().manual_seed(0))
    assert torch.equal(multigpu.counter_uniform(ids[perm], 3, seed=5), u[perm])          # any subset, any order, any rank
    assert not torch.equal(multigpu.counter_uniform(ids, 3, seed=6), u)
    # columns and neighbouring ids are uncorrelated
    c = torch.corrcoef(torch.stack([u[:-1, 0], u[1:, 0], u[:-1, 1]]))
    assert float(c[0, 1].abs()) < 0.02 and float(c[0, 2].abs()) < 0.02


def _presharded_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import math
    import multigpu
    n = 64
    h = 3.2 * 2 / n
    d = math.ceil(2 / h)
    dims = (d, d, 3)
    xg, gidg = multigpu.slab_lattice_2d(n, h, d, 0, 1, "cpu", seed=2)                     # the whole cloud (what 1 rank generates)
    glob = multigpu.SlabPartition(xg, h, dims, None, rank, world)                         # axis picked: y (z holds one cell row)
    xl, gidl = multigpu.slab_lattice_2d(n, h, d, rank, world, "cpu", seed=2)              # this rank's rows + halo rows only
    loc = multigpu.SlabPartition.from_local(xl, gidl, h, dims, 1, rank, world)
    same = (glob.axis == 1 and torch.equal(glob.own_idx, loc.own_idx) and torch.equal(glob.send_prev, loc.send_prev)
            and torch.equal(glob.send_next, loc.send_next) and (glob.n_halo_prev, glob.n_halo_next) == (loc.n_halo_prev, loc.n_halo_next)
            and torch.equal(glob.local_from_global(xg), loc.local_positions()))
    ret[rank] = (bool(same), int(xl.shape[0]), int(xg.shape[0]), loc.n_own, int(loc.boundary.sum()))
    dist.destroy_process_group()


def test_presharded_partition_equals_global_partition_world2():
    """slab_lattice_2d(rank, world) + SlabPartition.from_local: same ownership, send sets, halo order and local positions as
    cutting the whole cloud -- without any rank ever holding the whole cloud."""
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_presharded_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0][3] + ret[1][3] == 64 * 64
    for r in range(world):
        same, n_loc, n_glob, n_own, n_bnd = ret[r]
        assert same
        assert n_own < n_loc < n_glob and 0 < n_bnd < n_own
