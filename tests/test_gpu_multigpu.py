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
