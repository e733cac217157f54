"""Multi-GPU execution of the SPH-NCA path.  The reference is single-GPU; what must
hold is that N-GPU results equal the single-GPU results of the same kernels.

Data parallel over `sections`: batch elements are independent through the whole rollout
(separate particle / cell ranges), so each rank (one process per GPU) runs the rollout on its own
contiguous block of batch elements with NO data-path collective.  The only exchange is one
sum-all-reduce of the flat 21,025-float parameter gradient (84,100 B) per training iteration,
issued before `normalize_grads` (which is non-linear, reference nca.py:119-122) so that every rank
then applies the identical Adam step.  The reduction is a single NCCL all-reduce on one flat
buffer: at 84 KB it is latency-bound over NVLink/NVSwitch, bucketing would only add launches.
"""
import torch
import torch.distributed as dist

import nca


def shard_sections(sections, world_size, rank):
    """Contiguous block of batch elements for `rank`: returns (first_section, last_section_exclusive,
    first_particle, last_particle_exclusive).  Blocks differ by at most one element."""
    B = len(sections)
    base, extra = divmod(B, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    p_lo = sum(sections[:lo])
    return lo, hi, p_lo, p_lo + sum(sections[lo:hi])


def allreduce_gradients(params, group=None, average=True):
    """One flat all-reduce of every parameter gradient.  No-op when not running distributed.
    Parameters without a gradient contribute zeros (and receive the reduced value)."""
    params = list(params)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel() * flat.element_size()


class DataParallelSPHNCA(nca.SPHNCA):
    """SPHNCA whose `normalize_grads` first all-reduces the parameter gradients across ranks.

    The reference's rollout drivers call `loss.backward(); model.normalize_grads(); optimizer.step()`
    (nca.py:151-153), so hooking the reduction here makes `nca.sample_plane` data-parallel without
    changing it."""

    def normalize_grads(self):
        allreduce_gradients(self.parameters())
        super().normalize_grads()
