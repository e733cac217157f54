"""Where the time of a slab-sharded rollout goes (development aid; run under torch.distributed.run on N GPUs):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/slab_profile.py [n_side] [steps]
Variants: overlapped split (default), blocking split, no split (world == 1 only); per-variant rollout time and, from CUDA events on
the main stream, the time of each phase of a step."""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import torch
import torch.distributed as dist


def main():
    n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import bench
    import multigpu
    import nca
    from sphops import fused
    h = 3.2 * 2 / n_side
    d = math.ceil(2 / h)
    dims = (d, d, 3)
    model = nca.SPHNCA(16, 0.5, use_alpha=False, normalize_perception=-1).to(dev)
    with torch.no_grad():
        for t, src in zip((model.model[0].weight, model.model[0].bias, model.model[2].weight, model.model[2].bias), bench.load_weights()):
            t.copy_(src)
    x = multigpu.slab_lattice_2d(n_side, h, d, rank, world, dev, seed=0)
    for name, kw in (("overlapped", dict(overlap=True)), ("blocking", dict(overlap=False))):
        sh = multigpu.SlabShardedNCA(model, x, h, dims, axis=1, presharded=True, **kw)
        A0 = multigpu.counter_uniform(sh.part.own_global_ids(), 16, seed=7, device=dev)
        for _ in range(2):
            sh.rollout(A0, 2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sh.rollout(A0, K); e1.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        fused.PROFILE = {}
        sh.rollout(A0, K)
        torch.cuda.synchronize()
        prof, fused.PROFILE = fused.PROFILE, None
        parts = {k: (len(v), sum(a.elapsed_time(b) for a, b in v) / len(v)) for k, v in prof.items()}
        print(f"[rank {rank}] {name}: rollout {ms:.2f} ms ({ms / K:.3f} ms/step), host enqueue {t_host * 1e3:.1f} ms, own {sh.part.n_own}, "
              f"halo {sh.part.n_halo_prev + sh.part.n_halo_next}, kernels(avg ms): " + ", ".join(f"{k} x{n} {t:.3f}" for k, (n, t) in parts.items()), flush=True)
        del sh
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
