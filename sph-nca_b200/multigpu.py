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


# =====================================================================================================
# Spatial slab sharding with a halo exchange (one huge cloud, inference) -- SURVEY.md 8e, BASELINE config 4
# =====================================================================================================
class SlabPartition:
    """Ownership and halo bookkeeping of one packed cloud cut into slabs of cell rows.

    The cell hash is x + dims_x*(y + dims_y*z) (reference single_impl.py:34-44), the support radius equals the
    cell size, and neighbour cells wrap modulo `dims` (single_impl.py:54-55).  Cutting along axis `axis` (the
    last populated one) into contiguous ranges of cell indices therefore needs a halo of exactly ONE cell layer
    on each side, taken from the ring neighbours (rank-1, rank+1 modulo world).

    Local particle order on a rank: [interior | halo from the previous rank | halo from the next rank].
    `exchange(rows_of_my_particles)` fills the halo slots of a local tensor from the owners' current values.
    """

    def __init__(self, x, h, dims, axis, rank, world, group=None):
        import math
        self.rank, self.world, self.group, self.axis = rank, world, group, axis
        D = x.shape[1]
        dims = [int(dims)] * D if isinstance(dims, int) else [int(d) for d in dims]
        self.dims = dims
        nrow = dims[axis]
        # cell row of every particle: int(floor(x/h)) mod dims, in float64 like the hash (exact for these inputs)
        row = torch.remainder(torch.floor(x[:, axis].double() / float(h)).long(), nrow)
        base, extra = divmod(nrow, world)
        starts = [r * base + min(r, extra) for r in range(world + 1)]
        self.row_lo, self.row_hi = starts[rank], starts[rank + 1]            # my rows: [row_lo, row_hi)
        owner_mask = (row >= self.row_lo) & (row < self.row_hi)
        self.own_idx = torch.nonzero(owner_mask).flatten()                     # global indices I own (ascending)
        own_rows = row[self.own_idx]
        self.n_own = int(self.own_idx.numel())
        if world > 1 and self.row_hi - self.row_lo < 1:
            raise ValueError("more ranks than cell rows")
        # what I send: my first row goes to the previous rank (it is its upper halo), my last row to the next rank
        self.send_prev = torch.nonzero(own_rows == self.row_lo).flatten()      # positions inside own_idx
        self.send_next = torch.nonzero(own_rows == self.row_hi - 1).flatten()
        self.prev, self.next = (rank - 1) % world, (rank + 1) % world
        if world > 1:
            cnt = torch.tensor([self.send_prev.numel(), self.send_next.numel()], dtype=torch.int64, device=x.device)
            got = [torch.zeros_like(cnt) for _ in range(world)]
            dist.all_gather(got, cnt, group=group)
            self.n_halo_prev = int(got[self.prev][1])      # the previous rank's last row
            self.n_halo_next = int(got[self.next][0])      # the next rank's first row
        else:
            self.n_halo_prev = self.n_halo_next = 0
        self.n_local = self.n_own + self.n_halo_prev + self.n_halo_next
        self.halo_prev = slice(self.n_own, self.n_own + self.n_halo_prev)
        self.halo_next = slice(self.n_own + self.n_halo_prev, self.n_local)

    def exchange(self, own_rows_tensor, out_local):
        """out_local[halo slots] <- the owners' rows.  `own_rows_tensor` is indexed like own_idx (n_own rows);
        `out_local` has n_local rows.  Blocking; one batched NCCL/gloo P2P round with the two ring neighbours."""
        if self.world == 1:
            return 0
        n = self.n_own
        halo_prev = torch.arange(n, n + self.n_halo_prev, device=out_local.device)
        halo_next = torch.arange(n + self.n_halo_prev, self.n_local, device=out_local.device)
        return self.exchange_indexed(own_rows_tensor, self.send_prev, self.send_next, out_local, halo_prev, halo_next)

    def exchange_indexed(self, src, send_prev_idx, send_next_idx, dst, halo_prev_idx, halo_next_idx):
        """Generic form: rows src[send_*_idx] go to the ring neighbours, received rows land in dst[halo_*_idx].
        Only halo-sized gathers / scatters are performed (the tensors may be in any order, e.g. cell-sorted)."""
        if self.world == 1:
            return 0
        to_prev = src[send_prev_idx].contiguous()
        to_next = src[send_next_idx].contiguous()
        shape = tuple(src.shape[1:])
        from_prev = torch.empty((self.n_halo_prev,) + shape, dtype=src.dtype, device=src.device)
        from_next = torch.empty((self.n_halo_next,) + shape, dtype=src.dtype, device=src.device)
        ops = [dist.P2POp(dist.isend, to_prev, self.prev, self.group), dist.P2POp(dist.isend, to_next, self.next, self.group),
               dist.P2POp(dist.irecv, from_next, self.next, self.group), dist.P2POp(dist.irecv, from_prev, self.prev, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        dst[halo_prev_idx] = from_prev
        dst[halo_next_idx] = from_next
        return (to_prev.numel() + to_next.numel()) * to_prev.element_size()

    def local_from_global(self, t_global):
        """[interior | halo_prev | halo_next] rows of a tensor every rank holds in full (setup only)."""
        out = torch.empty((self.n_local,) + tuple(t_global.shape[1:]), dtype=t_global.dtype, device=t_global.device)
        out[:self.n_own] = t_global[self.own_idx]
        self.exchange(out[:self.n_own], out)
        return out


class SlabShardedNCA:
    """K-step inference rollout of one cloud sharded into slabs of cell rows (one process per GPU).

    Per rank and step, in the cell-sorted domain of the LOCAL particle set (interior + halo):
      1. halo exchange of the state rows (64 B / halo particle)
      2. perception sweep for every local particle (from the slab's pair lists when they fit) + tcgen05 MLP/update
         for the fired ones (halo results are discarded)
      3. halo exchange of the alive bytes of the updated state (1 B / halo particle)
      4. life-mask sweep, masking of the new state
    Volume: computed locally (interior exact), halo values fetched once from their owners.
    Results for the interior equal the single-GPU results of the same kernels up to fp32 summation order.
    """

    def __init__(self, model, x, h, dims, axis=None, group=None):
        import ctypes as C
        import sphops
        from sphops import _lib, fused
        self.C, self._lib, self.fused = C, _lib, fused
        self.model, self.h = model, float(h)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        D = x.shape[1]
        dims_l = [int(dims)] * D if isinstance(dims, int) else [int(d) for d in dims]
        if axis is None:
            axis = D - 1
        self.part = SlabPartition(x, h, dims_l, axis, rank, world, group)
        p = self.part
        self.x_loc = p.local_from_global(x.contiguous())
        self.grid = sphops.initialize_hashgrid(self.x_loc, (p.n_local,), h, tuple(dims_l))
        v = sphops.volume(self.x_loc, h, self.grid).clone()
        p.exchange(v[:p.n_own].clone(), v)                       # halo volumes from their owners
        self.v = v
        self.xv4 = self.grid.sorted_xv(self.x_loc, self.v)
        sid = self.grid.sorted_id.long()
        inv = torch.empty_like(sid); inv[sid] = torch.arange(sid.numel(), device=sid.device)
        self.pos_of_local = inv                                   # sorted position of local particle i
        self.own_pos = inv[:p.n_own]
        # sorted positions of what is sent / received each step (halo-sized index sets)
        self.send_prev_pos = self.own_pos[p.send_prev]
        self.send_next_pos = self.own_pos[p.send_next]
        self.halo_prev_pos = inv[p.n_own:p.n_own + p.n_halo_prev]
        self.halo_next_pos = inv[p.n_own + p.n_halo_prev:]
        self.bytes_per_step = 0

    def _exchange_sorted(self, t_s):
        """halo exchange, in place, of a tensor held in cell-sorted order (rows = local particles)."""
        return self.part.exchange_indexed(t_s, self.send_prev_pos, self.send_next_pos, t_s, self.halo_prev_pos, self.halo_next_pos)

    @torch.no_grad()
    def rollout(self, A_global_own, steps, fire_masks=None, fire_rate=None):
        """A_global_own: state rows of the particles this rank owns (ordered like part.own_idx).  Returns the
        final state of the same particles.  fire_masks: optional bool [steps, n_own] (owned particles); by
        default drawn per step with torch.rand like the reference (each rank its own stream)."""
        C, _lib, fused = self.C, self._lib, self.fused
        lib, ptr, check = _lib.lib, _lib.ptr, _lib.check
        p, m = self.part, self.model
        dev = A_global_own.device
        N = p.n_local
        A_loc = torch.zeros((N, 16), dtype=torch.float32, device=dev)
        A_loc[:p.n_own] = A_global_own
        A_s = self.grid.pack_rows(A_loc, 16)
        f32 = dict(dtype=torch.float32, device=dev)
        An_s = torch.empty((N, 16), **f32); Y_s = torch.empty((N, 32), **f32); O_s = torch.empty((N, 36), **f32)
        flags_s = torch.empty((N,), dtype=torch.uint8, device=dev); alive_s = torch.empty((N,), dtype=torch.uint8, device=dev)
        relu_s = torch.empty((N, 8), dtype=torch.int32, device=dev)
        lin1, lin2 = m.model[0], m.model[2]
        fr = m.fire_rate if fire_rate is None else fire_rate
        prm = fused._step_params(lin1.weight.detach(), lin1.bias.detach(), lin2.weight.detach(), lin2.bias.detach(),
                                 m.update_rule, m.use_alpha, m.perception_scale(self.h), m.fire_rate / fr)
        g, st = self.grid.c_grid(), _lib.stream_ptr()
        # the slab's own pair lists (built once per partition) and, per step, the list of the fired rows: the MLP runs
        # over those only (halo rows never fire: their results are overwritten by the next exchange anyway)
        pl = self.grid.pair_lists(self.x_loc, self.v, blur=bool(m.use_alpha))
        lists = fused._fire_lists_on() and N > 0
        if lists:
            rows = torch.empty((N,), dtype=torch.int32, device=dev)
            count = torch.empty((1,), dtype=torch.int32, device=dev)
            ws_fl = fused._fire_list_workspace(N, dev)
        nbytes = 0
        fb = None
        for k in range(steps):
            nbytes += self._exchange_sorted(A_s)
            fire = torch.zeros((N,), dtype=torch.uint8, device=dev)
            fire[:p.n_own] = (fire_masks[k] if fire_masks is not None else (torch.rand(p.n_own, device=dev) <= fr)).to(torch.uint8)
            fire_s = self.grid.pack_bytes(fire)
            if lists and not m.use_alpha:
                # no alpha channel: list, perception sweep over the fired rows, MLP over compact rows, static life mask --
                # the rollout node's forward step; halo rows never fire, so nothing at all is computed for them
                if fb is None:
                    fb = fused._FireBuffers(N, dev, compact_inputs=True)
                fb.flags = flags_s
                fused._step_forward_lists(self.grid, None, None, g, self.xv4, A_s, fire_s, prm, False, Y_s, An_s, alive_s, fb, ws_fl, st, pl=pl)
                A_s, An_s = An_s, A_s
                continue
            if lists:
                check(lib.sphnca_fire_list(N, ptr(fire_s), ptr(A_s), int(bool(m.use_alpha)), ptr(rows), ptr(count), ptr(An_s),
                                           ptr(alive_s), None, None, ptr(ws_fl), ws_fl.numel(), st))
            fused._perceive_pl(self.grid, pl, g, self.xv4, A_s, prm, Y_s, flags_s, st)
            if lists:
                check(lib.sphnca_step_mlp_list(N, ptr(rows), ptr(count), ptr(A_s), ptr(Y_s), C.byref(prm), ptr(An_s), ptr(alive_s),
                                               ptr(O_s), ptr(flags_s), ptr(relu_s), None, None, 0, st))
            else:
                check(lib.sphnca_step_mlp(N, ptr(A_s), ptr(Y_s), ptr(fire_s), C.byref(prm), ptr(An_s), ptr(alive_s), ptr(O_s),
                                          ptr(flags_s), ptr(relu_s), st))
            if m.use_alpha:
                nbytes += self._exchange_sorted(alive_s)
                # the previous-mask bit of a halo particle is never used (its row is overwritten by the next exchange)
                check(lib.sphnca_step_mask(g, ptr(self.xv4), ptr(alive_s), ptr(An_s), ptr(flags_s), st))
            else:
                # no alpha channel: new life mask == previous one (include/sphnca.h), no sweep and no second halo round
                check(lib.sphnca_step_mask_static(N, ptr(An_s), ptr(flags_s), st))
            A_s, An_s = An_s, A_s
        self.bytes_per_step = nbytes // max(steps, 1)
        return A_s[self.own_pos]
