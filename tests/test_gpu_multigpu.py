"""2-GPU test of the slab-sharded rollout (run with `gpurun --gpus 2`; skipped with fewer devices): the
interior of every slab must equal the single-GPU rollout of the same kernels."""
import math
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import multigpu
    import nca
    from tests.util import forced_fire_masks, jittered_lattice, load_golden
    gen = torch.Generator().manual_seed(0)
    n = 96
    x = torch.nn.functional.pad(jittered_lattice(n, 2, gen), (0, 1)).to(dev)
    N = x.shape[0]
    h = 3.2 * 2 / n
    d = math.ceil(2 / h)
    dims = (d, d, 3)
    w = load_golden("weights_gecko")
    m = nca.SPHNCA(16, 0.5, use_alpha=True, normalize_perception=1.0 / h).to(dev)
    with torch.no_grad():
        for t, k in zip((m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias), ("W1", "b1", "W2", "b2")):
            t.copy_(torch.from_numpy(w[k]))
    d2 = (x ** 2).sum(-1)
    A0 = ((torch.clamp(1 - d2 / 0.6 ** 2, 0, 1) ** 3)[:, None].repeat(1, 16) + 0.05 * torch.rand(N, 16, generator=gen).to(dev))
    K = 3
    masks = (torch.rand(K, N, generator=gen) <= 0.5).to(dev)
    # single-GPU reference with the same kernels
    nca.USE_FUSED_STEP = True
    grid, v = m.init_grid(x, (N,), h, dims)
    A = A0
    with torch.no_grad(), forced_fire_masks(list(masks.cpu())):
        for _ in range(K):
            _, A = m(x, v, A, h, grid)
    sh = multigpu.SlabShardedNCA(m, x, h, dims, axis=1)
    own = sh.part.own_idx
    out = sh.rollout(A0[own], K, fire_masks=masks[:, own])
    err = float((out - A[own]).abs().max() / A.abs().max())
    ret[rank] = (err, int(own.numel()), sh.bytes_per_step)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_slab_sharded_rollout_equals_single_gpu():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0][1] + ret[1][1] == 96 * 96
    for r in range(world):
        assert ret[r][0] <= 2e-5, ret[r]
        assert ret[r][2] > 0


def _model(dev, use_alpha, h):
    import nca
    from tests.util import load_golden
    w = load_golden("weights_gecko" if use_alpha else "weights_zebra")
    m = nca.SPHNCA(16, 0.5, use_alpha=use_alpha, normalize_perception=1.0 / h).to(dev)
    with torch.no_grad():
        for t, k in zip((m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias), ("W1", "b1", "W2", "b2")):
            t.copy_(torch.from_numpy(w[k]))
    return m


def _single_gpu_counter_rollout(m, x, gid, A0, h, dims, K, seed_of_step):
    """Plain single-GPU rollout (SPHNCA.forward) with the slab path's counter-based fire draws."""
    import multigpu
    import nca
    from tests.util import forced_fire_masks
    nca.USE_FUSED_STEP = True
    grid, v = m.init_grid(x, (x.shape[0],), h, dims)
    masks = [(multigpu.counter_uniform(gid, 1, seed_of_step(k))[:, 0] <= 0.5).cpu() for k in range(K)]
    A = A0
    with torch.no_grad(), forced_fire_masks(masks):
        for _ in range(K):
            _, A = m(x, v, A, h, grid)
    return A


def test_slab_boundary_interior_passes_equal_single_pass_on_one_gpu():
    """The no-alpha slab step is split into a boundary pass and an interior pass (the halo exchange hides behind the latter).
    On ONE rank, with the split forced: bit-identical to the unsplit rollout and to SPHNCA.forward with the same draws --
    covers the list-only mode of sphnca_fire_list and the pre-sharded lattice generator on a 1-GPU lease."""
    import multigpu
    dev = torch.device("cuda", 0)
    n = 128
    h = 3.2 * 2 / n
    d = math.ceil(2 / h)
    dims = (d, d, 3)
    x, gid = multigpu.slab_lattice_2d(n, h, d, 0, 1, dev, seed=3)
    assert x.shape[0] == n * n and torch.equal(gid, torch.arange(n * n, device=dev))
    m = _model(dev, False, h)
    A0 = multigpu.counter_uniform(gid, 16, seed=7)
    K = 3
    outs = []
    for split in (False, True):
        sh = multigpu.SlabShardedNCA(m, (x, gid), h, dims, axis=1, presharded=True, split_passes=split)
        assert int(sh.bnd_s.sum()) > 0 and int(sh.int_s.sum()) > 0
        outs.append(sh.rollout(A0, K))
    assert torch.equal(outs[0], outs[1])
    ref = _single_gpu_counter_rollout(m, x, gid, A0, h, dims, K, lambda k: 0 + 1000003 * 0 + k)
    assert torch.equal(outs[1], ref)


def _worker_noalpha(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import multigpu
    n = 160
    h = 3.2 * 2 / n
    d = math.ceil(2 / h)
    dims = (d, d, 3)
    m = _model(dev, False, h)
    K = 4
    # every rank: the whole cloud on one GPU (the reference result), and its own slab of it, generated slab-locally
    xg, gidg = multigpu.slab_lattice_2d(n, h, d, 0, 1, dev, seed=3)
    A0g = multigpu.counter_uniform(gidg, 16, seed=7)
    ref = _single_gpu_counter_rollout(m, xg, gidg, A0g, h, dims, K, lambda k: k)
    res = {}
    for overlap in (True, False):
        x, gid = multigpu.slab_lattice_2d(n, h, d, rank, world, dev, seed=3)
        sh = multigpu.SlabShardedNCA(m, (x, gid), h, dims, axis=1, presharded=True, overlap=overlap)
        own = sh.part.own_global_ids()
        out = sh.rollout(multigpu.counter_uniform(own, 16, seed=7), K)
        res[overlap] = (bool(torch.equal(out, ref[own])), float((out - ref[own]).abs().max()), int(own.numel()), sh.bytes_per_step)
    ret[rank] = res
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_slab_sharded_noalpha_rollout_equals_single_gpu_bitwise():
    """The benched slab path (no alpha channel: pair lists, fire lists, fired-rows perception, boundary / interior passes with the
    exchange on a side stream) on 2 ranks == the 1-GPU rollout of the whole cloud, bit for bit, with the counter-based global
    fire draws; overlapped and blocking exchange."""
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker_noalpha, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert sum(ret[r][True][2] for r in range(world)) == 160 * 160
    for r in range(world):
        for overlap in (True, False):
            same, err, n_own, nbytes = ret[r][overlap]
            assert same, (r, overlap, err)
            assert nbytes > 0


def _worker_dp(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import multigpu
    import nca
    from tests.util import forced_fire_masks, jittered_lattice, load_golden
    gen = torch.Generator().manual_seed(5)
    n, B, K = 72, 4, 3
    xs = [jittered_lattice(n, 2, gen) for _ in range(B)]
    x = torch.cat(xs); N = x.shape[0]
    sections = (n * n,) * B
    h = 3.2 * 2 / n
    dims = math.ceil(2 / h)
    A0 = torch.rand(N, 16, generator=gen)
    masks = torch.rand(K, N, generator=gen) <= 0.5
    w = load_golden("weights_zebra")
    nca.USE_3D = True
    nca.USE_FUSED_STEP = True

    def run(model, xs_, A_, secs, mk):
        with torch.no_grad():
            for t, k in zip((model.model[0].weight, model.model[0].bias, model.model[2].weight, model.model[2].bias), ("W1", "b1", "W2", "b2")):
                t.copy_(torch.from_numpy(w[k]))
        opt = torch.optim.SGD(model.parameters(), lr=0.0)
        grads = {}
        orig = model.normalize_grads

        def hook():
            orig()
            grads["g"] = [p.grad.clone() for p in model.parameters()]
        model.normalize_grads = hook
        import random
        random.seed(0)
        with forced_fire_masks(list(mk)):
            nca.sample_plane(model, xs_.to(dev), A_.to(dev), secs, h, dims, iter_n=K,
                             loss_fn=lambda x_, A, img, A0=None: (A[:, :4] ** 2).mean(), optimizer=opt)
        return grads["g"]

    # 1 rank, the concatenated batch (plain SPHNCA: no reduction)
    full = run(nca.SPHNCA(16, 0.5, use_alpha=False, normalize_perception=1.0 / h, normalize_grad=False).to(dev), x, A0, sections, masks)
    # this rank's shard, gradients all-reduced (mean over equal shards) before normalize_grads
    lo, hi, p_lo, p_hi = multigpu.shard_sections(sections, world, rank)
    dp = run(multigpu.DataParallelSPHNCA(16, 0.5, use_alpha=False, normalize_perception=1.0 / h, normalize_grad=False).to(dev),
             x[p_lo:p_hi], A0[p_lo:p_hi], sections[lo:hi], masks[:, p_lo:p_hi])
    ret[rank] = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(dp, full)]
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_equals_single_gpu():
    """N ranks, each on its block of batch elements, gradients all-reduced before normalize_grads == 1 rank on the concatenated
    batch (same fire masks; a mean loss over equal shards): all four parameter gradients to fp32 summation order."""
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker_dp, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert max(ret[r]) <= 2e-5, ret[r]
