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


def allreduce_gradients(params, group=None, average=True, weight=None):
    """One flat all-reduce of every parameter gradient.  No-op when not running distributed.
    Parameters without a gradient contribute zeros (and receive the reduced value).
    average=True divides the sum by the number of ranks: equal to the single-GPU gradient of a MEAN loss over the
    concatenated batch only when every rank holds the same number of particles.  With unequal shards (B % world != 0,
    `shard_sections`) pass `weight` = this rank's share of the batch (its particles / all particles): each rank's gradient
    is scaled by it before the SUM and nothing is divided afterwards."""
    params = list(params)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    if weight is not None:
        flat = flat * float(weight)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average and weight is None:
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
    changing it.  `shard_weight` (None = equal shards): this rank's share of the global batch, see allreduce_gradients.
    N ranks reproduce the 1-rank gradient of the concatenated batch when the fire masks are the same (each rank draws
    its own torch.rand stream otherwise): tests/test_gpu_multigpu.py::test_data_parallel_equals_single_gpu."""
    shard_weight = None

    def normalize_grads(self):
        allreduce_gradients(self.parameters(), weight=self.shard_weight)
        super().normalize_grads()


# =====================================================================================================
# Spatial slab sharding with a halo exchange (one huge cloud, inference) -- SURVEY.md 8e, BASELINE config 4
# =====================================================================================================
def counter_uniform(ids, ncols, seed, device=None):
    """Counter-based uniforms in [0,1): value (i, c) depends on (ids[i], c, seed) only -- splitmix64 of the counter, top 24
    bits -- so a quantity drawn per GLOBAL particle id (synthetic states, the per-step fire draw of a slab rollout) is the
    same whatever the number of ranks and whichever rank owns the particle.  ids: int64 [n] -> float32 [n, ncols]."""
    ids = ids.to(torch.int64)
    if device is not None:
        ids = ids.to(device)
    M = (1 << 64) - 1

    def s64(v):                                     # two's-complement int64 of an unsigned 64-bit constant
        v &= M
        return v - (1 << 64) if v >= (1 << 63) else v

    z = ids[:, None] * ncols + torch.arange(ncols, device=ids.device, dtype=torch.int64)[None, :]
    z = z + s64(0x9E3779B97F4A7C15 * (int(seed) + 1))
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * s64(0xBF58476D1CE4E5B9)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * s64(0x94D049BB133111EB)
    z = z ^ ((z >> 31) & ((1 << 33) - 1))
    return ((z >> 40) & ((1 << 24) - 1)).to(torch.float32) * (1.0 / (1 << 24))


def _row_ranges(nrow, world):
    base, extra = divmod(nrow, world)
    return [r * base + min(r, extra) for r in range(world + 1)]


def cell_row(x_axis, h, nrow):
    """int(floor(x/h)) mod dims in float64, exactly like the cell hash (single_impl.py:34-44)."""
    return torch.remainder(torch.floor(x_axis.double() / float(h)).long(), nrow)


def slab_lattice_2d(n_side, h, nrow, rank, world, device, seed=0, jitter=1.0):
    """This rank's part of the synthetic n_side^2 jittered lattice (zero-padded to 3-D): the particles whose
    cell row along y lies in the rank's slab or in the one-row halo either side, generated WITHOUT building the whole cloud
    (16 M particles at BASELINE configs[3]).  Jitter comes from counter_uniform(global id), so the cloud is the same for
    every rank count.  Returns (x [m,3], gid [m]) with gid = iy * n_side + ix ... ordered by gid."""
    s = 2.0 / n_side
    starts = _row_ranges(nrow, world)
    lo, hi = starts[rank], starts[rank + 1]
    need = torch.zeros(nrow, dtype=torch.bool, device=device)
    need[lo:hi] = True
    if world > 1:
        need[(lo - 1) % nrow] = True
        need[hi % nrow] = True
    # lattice rows (index iy along y) that can land in a needed cell row for some jitter in [-0.5, 0.5) spacings
    iy = torch.arange(n_side, device=device, dtype=torch.float64)
    r_min = cell_row((iy + 0.5 - 0.5 * jitter) * s - 1.0, h, nrow)
    r_max = cell_row((iy + 0.5 + 0.5 * jitter) * s - 1.0, h, nrow)
    r_mid = cell_row((iy + 0.5) * s - 1.0, h, nrow)
    keep_rows = torch.nonzero(need[r_min] | need[r_max] | need[r_mid]).flatten()
    ix = torch.arange(n_side, device=device, dtype=torch.int64)
    gid = (keep_rows[:, None] * n_side + ix[None, :]).reshape(-1)
    u = counter_uniform(gid, 2, seed, device)
    x = torch.empty((gid.numel(), 3), dtype=torch.float32, device=device)
    # the lattice position in fp32 like tests/util.jittered_lattice: (i + 0.5) * s - 1 + (u - 0.5) * jitter * s
    x[:, 0] = ((gid % n_side).to(torch.float32) + 0.5) * s - 1.0 + (u[:, 0] - 0.5) * (jitter * s)
    x[:, 1] = ((gid // n_side).to(torch.float32) + 0.5) * s - 1.0 + (u[:, 1] - 0.5) * (jitter * s)
    x[:, 2] = 0.0
    sel = need[cell_row(x[:, 1], h, nrow)]
    return x[sel].contiguous(), gid[sel].contiguous()


class SlabPartition:
    """Ownership and halo bookkeeping of one packed cloud cut into slabs of cell rows.

    The cell hash is x + dims_x*(y + dims_y*z) (reference single_impl.py:34-44), the support radius equals the
    cell size, and neighbour cells wrap modulo `dims` (single_impl.py:54-55).  Cutting along axis `axis` into contiguous
    ranges of cell indices therefore needs a halo of exactly ONE cell layer on each side, taken from the ring neighbours
    (rank-1, rank+1 modulo world).  `axis=None` picks the last axis along which the particles span more than one cell
    row (a zero-padded 2-D cloud has all its particles in ONE z row: cutting there would give one rank everything).

    Two ways in: the whole cloud on every rank (`SlabPartition(x, ...)`, small clouds / tests), or only the particles of
    this rank's rows + halo rows with their global ids (`SlabPartition.from_local`, e.g. from `slab_lattice_2d`).
    Local particle order on a rank: [own (ascending global id) | halo from the previous rank | halo from the next rank],
    halo particles in ascending global id too -- the order in which their owner sends them.
    `exchange(rows_of_my_particles)` fills the halo slots of a local tensor from the owners' current values.
    """

    def __init__(self, x, h, dims, axis, rank, world, group=None, gid=None):
        self.rank, self.world, self.group = rank, world, group
        D = x.shape[1]
        dims = [int(dims)] * D if isinstance(dims, int) else [int(d) for d in dims]
        self.dims = dims
        if axis is None:
            axis = self.pick_axis(x, h, dims)
        self.axis = axis
        nrow = dims[axis]
        if world > 1 and nrow < world:
            raise ValueError("more ranks than cell rows")
        row = cell_row(x[:, axis], h, nrow)
        starts = _row_ranges(nrow, world)
        self.row_lo, self.row_hi = starts[rank], starts[rank + 1]            # my rows: [row_lo, row_hi)
        if gid is None:
            gid = torch.arange(x.shape[0], device=x.device)
            self._global = True
        else:                                                                  # candidates: sort by global id once
            order = torch.argsort(gid)
            gid, row, x = gid[order], row[order], x[order]
            self._global = False
        own = (row >= self.row_lo) & (row < self.row_hi)
        sel = torch.nonzero(own).flatten()
        self.own_idx = gid[sel]                                                # global ids I own (ascending)
        own_rows = row[sel]
        self.n_own = int(sel.numel())
        if world > 1 and self.n_own > 0.9 * x.shape[0] and self._global:
            raise ValueError(f"rank {rank} owns {self.n_own} of {x.shape[0]} particles along axis {axis}: wrong slab axis?")
        # what I send: my first row goes to the previous rank (it is its upper halo), my last row to the next rank
        self.send_prev = torch.nonzero(own_rows == self.row_lo).flatten()      # positions inside own_idx
        self.send_next = torch.nonzero(own_rows == self.row_hi - 1).flatten()
        self.boundary = (own_rows == self.row_lo) | (own_rows == self.row_hi - 1)   # own particles whose neighbourhood reaches a halo row
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
        self._x_local = None
        if not self._global:                               # the halo particles are among the candidates: no exchange needed for x
            hp = torch.nonzero(row == (self.row_lo - 1) % nrow).flatten() if world > 1 else sel[:0]
            hn = torch.nonzero(row == self.row_hi % nrow).flatten() if world > 1 else sel[:0]
            if world > 1 and (hp.numel() != self.n_halo_prev or hn.numel() != self.n_halo_next):
                raise ValueError("the local candidates do not hold the complete halo rows")
            self._x_local = torch.cat([x[sel], x[hp], x[hn]]).contiguous()
            self.halo_gid = torch.cat([gid[hp], gid[hn]])

    @classmethod
    def from_local(cls, x_cand, gid_cand, h, dims, axis, rank, world, group=None):
        """Partition from the particles of this rank's rows + halo rows only (any order) and their global ids."""
        return cls(x_cand, h, dims, axis, rank, world, group, gid=gid_cand)

    @staticmethod
    def pick_axis(x, h, dims):
        for a in range(x.shape[1] - 1, -1, -1):
            if dims[a] > 1 and int(torch.unique(cell_row(x[:, a], h, dims[a])).numel()) > 1:
                return a
        raise ValueError("the cloud occupies a single cell row along every axis: nothing to cut")

    def own_global_ids(self):
        return self.own_idx

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
        Only halo-sized gathers / scatters are performed (the tensors may be in any order, e.g. cell-sorted).
        Enqueued on the CURRENT stream (NCCL's stream is ordered after it and the current stream waits for NCCL): the
        host does not block, and a caller may run it on a side stream to overlap it with compute."""
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
        assert self._global, "this partition was built from local candidates: there is no global tensor to cut"
        out = torch.empty((self.n_local,) + tuple(t_global.shape[1:]), dtype=t_global.dtype, device=t_global.device)
        out[:self.n_own] = t_global[self.own_idx]
        self.exchange(out[:self.n_own], out)
        return out

    def local_positions(self, x_global=None):
        return self._x_local if self._x_local is not None else self.local_from_global(x_global.contiguous())


class SlabShardedNCA:
    """K-step inference rollout of one cloud sharded into slabs of cell rows (one process per GPU), everything kept in the
    cell-sorted order of the LOCAL particle set (own + halo), kernels unchanged.

    Without an alpha channel (pair lists + fire lists; the benched path) a step is split so that the halo exchange hides
    behind the interior work:
      pass B  the BOUNDARY particles (own particles in the slab's first and last cell row: the only ones whose neighbourhood
              reaches a halo row, and the only ones a neighbour rank needs) -- fire list, fired-rows perception, tcgen05 MLP, with
              grids sized for the boundary rows and only the boundary rows carried over;
      exchange  of the NEW boundary rows into the neighbours' halo slots, on a side stream (NCCL P2P, 64 B / halo particle),
              issued as soon as pass B is enqueued;
      pass I  the INTERIOR particles (everything else: ~99 % of the slab at BASELINE configs[3]) on the main stream, while the
              exchange is in flight; the next step's pass B waits for the exchange.
    Each pass's list kernel carries over its own not-fired rows; halo rows are written by the exchange only (pass I runs beside it).  With an alpha channel the life mask needs the neighbours' NEW alive bytes, so the step stays sequential:
    exchange state -> full perception sweep -> MLP over the fire list -> exchange alive bytes -> mask sweep.
    Volume: computed locally (own particles exact), halo values fetched once from their owners.
    Fire draws: `fire_masks` (bool [steps, n_own]) or, by default, counter_uniform(global id, seed = fire_seed + step) <= rate:
    the same global draw whatever the number of ranks, so an N-rank rollout equals the 1-rank rollout.
    Results for the own particles equal the single-GPU results of the same kernels bit for bit (row-independent arithmetic,
    identical neighbour order: cell-sorted local order restricted to a particle's neighbourhood is the global one)."""

    def __init__(self, model, x, h, dims, axis=None, group=None, presharded=False, overlap=True, split_passes=None):
        import ctypes as C
        import sphops
        from sphops import _lib, fused
        self.C, self._lib, self.fused = C, _lib, fused
        self.model, self.h = model, float(h)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        if presharded:
            x_cand, gid = x
            D = x_cand.shape[1]
            dims_l = [int(dims)] * D if isinstance(dims, int) else [int(d) for d in dims]
            self.part = SlabPartition.from_local(x_cand, gid, h, dims_l, axis, rank, world, group)
            self.x_loc = self.part.local_positions()
        else:
            D = x.shape[1]
            dims_l = [int(dims)] * D if isinstance(dims, int) else [int(d) for d in dims]
            self.part = SlabPartition(x, h, dims_l, axis, rank, world, group)
            self.x_loc = self.part.local_positions(x)
        p = self.part
        dev = self.x_loc.device
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
        # per sorted position: global id (-1 for halo particles), boundary / interior membership of the own particles
        gid_loc = torch.full((p.n_local,), -1, dtype=torch.int64, device=dev)
        gid_loc[:p.n_own] = p.own_idx
        self.gid_s = gid_loc[sid].contiguous()
        bnd_loc = torch.zeros((p.n_local,), dtype=torch.uint8, device=dev)
        bnd_loc[:p.n_own] = p.boundary.to(torch.uint8)
        own_loc = torch.zeros((p.n_local,), dtype=torch.uint8, device=dev)
        own_loc[:p.n_own] = 1
        self.bnd_s = bnd_loc[sid].contiguous()
        self.int_s = (own_loc[sid] & (1 - self.bnd_s)).contiguous()
        self.own_s = own_loc[sid].contiguous()
        self.n_bnd = int(p.boundary.sum())                        # upper bound of pass B's list length (sizes its grids)
        # boundary / interior passes: whenever there is a neighbour rank (split_passes=True forces them on one rank: tests)
        self.split = world > 1 if split_passes is None else bool(split_passes)
        self.overlap = bool(overlap) and world > 1
        self.comm_stream = torch.cuda.Stream(device=dev) if (self.overlap and dev.type == "cuda") else None
        self.exchange_mode = "none (1 rank)" if world == 1 else (
            "NCCL P2P on a side stream, overlapped with the interior pass" if self.overlap else "NCCL P2P, blocking")
        self.bytes_per_step = 0
        self._work = None
        self._xchg = None
        self._fire_buf = None
        self._events = (torch.cuda.Event(), torch.cuda.Event()) if dev.type == "cuda" else (None, None)
        self.fire_seed = 0
        self._rollouts = 0

    def _exchange_sorted(self, t_s):
        """halo exchange, in place, of a tensor held in cell-sorted order (rows = local particles)."""
        return self.part.exchange_indexed(t_s, self.send_prev_pos, self.send_next_pos, t_s, self.halo_prev_pos, self.halo_next_pos)

    def _fire_bytes(self, k, fire_masks, fr, N, dev, split=False):
        """Fire bytes of step k in sorted order (0 for halo particles); split: also the boundary / interior copies."""
        p = self.part
        if fire_masks is not None:
            fire = torch.zeros((N,), dtype=torch.uint8, device=dev)
            fire[:p.n_own] = fire_masks[k].to(torch.uint8)
            fire_s = self.grid.pack_bytes(fire)
            return (fire_s, fire_s & self.bnd_s, fire_s & self.int_s) if split else fire_s
        # counter_uniform(gid, 1, seed) <= rate in one kernel (csrc/firelist.cu), bit for bit
        _lib = self._lib
        if self._fire_buf is None:
            self._fire_buf = [torch.empty((N,), dtype=torch.uint8, device=dev) for _ in range(3)]
        fire_s = self._fire_buf[0]
        fire_b, fire_i = (self._fire_buf[1], self._fire_buf[2]) if split else (None, None)
        _lib.check(_lib.lib.sphnca_fire_bytes_counter(N, _lib.ptr(self.gid_s), self.fire_seed + 1000003 * self._run + k, float(fr),
                                                      _lib.ptr(self.bnd_s) if split else None, _lib.ptr(fire_s), _lib.ptr(fire_b),
                                                      _lib.ptr(fire_i), _lib.stream_ptr()))
        return (fire_s, fire_b, fire_i) if split else fire_s

    @torch.no_grad()
    def rollout(self, A_global_own, steps, fire_masks=None, fire_rate=None):
        """A_global_own: state rows of the particles this rank owns (ordered like part.own_idx).  Returns the
        final state of the same particles.  fire_masks: optional bool [steps, n_own] (owned particles); by default the
        counter-based global draw (class docstring)."""
        C, _lib, fused = self.C, self._lib, self.fused
        lib, ptr, check = _lib.lib, _lib.ptr, _lib.check
        p, m = self.part, self.model
        dev = A_global_own.device
        N = p.n_local
        f32 = dict(dtype=torch.float32, device=dev)
        if self._work is None:
            # work buffers of a rollout, allocated once per partition (several GB at 16 M particles: re-allocating them per
            # rollout would make the timing depend on the state of torch's caching allocator)
            split_path = fused._fire_lists_on() and not m.use_alpha
            self._work = dict(
                A_loc=torch.zeros((N, 16), **f32), A=torch.empty((N, 16), **f32), An=torch.empty((N, 16), **f32),
                Y=torch.empty((N, 32), **f32), flags=torch.empty((N,), dtype=torch.uint8, device=dev),
                alive=torch.empty((N,), dtype=torch.uint8, device=dev),
                O=None if split_path else torch.empty((N, 36), **f32),
                relu=None if split_path else torch.empty((N, 8), dtype=torch.int32, device=dev),
                fb=fused._FireBuffers(N, dev, compact_inputs=True) if split_path else None)
        w = self._work
        A_loc = w["A_loc"]
        A_loc[:p.n_own] = A_global_own
        A_s, An_s = w["A"], w["An"]
        if N:
            _lib.check(_lib.lib.sphnca_pack_rows(self.grid.c_grid(), _lib.ptr(A_loc), 16, _lib.ptr(A_s), _lib.stream_ptr()))
        Y_s, O_s, flags_s, alive_s, relu_s = w["Y"], w["O"], w["flags"], w["alive"], w["relu"]
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
        self._run, self._rollouts = self._rollouts, self._rollouts + 1     # the draws of rollout r, step k: seed fire_seed + 1000003 r + k
        if lists and not m.use_alpha and pl is not None:
            out = self._rollout_split(A_s, An_s, Y_s, alive_s, flags_s, steps, fire_masks, fr, prm, pl, ws_fl, N, dev)
            return out
        for k in range(steps):
            nbytes += self._exchange_sorted(A_s)
            fire_s = self._fire_bytes(k, fire_masks, fr, N, dev)
            if lists:
                check(lib.sphnca_fire_list(N, ptr(fire_s), ptr(A_s), int(bool(m.use_alpha)), ptr(rows), ptr(count), ptr(An_s),
                                           ptr(alive_s), None, None, 0, ptr(ws_fl), ws_fl.numel(), st))
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

    def _state_exchange(self, dev):
        """Persistent buffers + P2P op list of the per-step state-row exchange (built once: the op objects, the index tensors and the
        send / receive buffers are reused by every step, which keeps the host side of a step short)."""
        if self._xchg is None and self.part.world > 1:
            p = self.part
            f32 = dict(dtype=torch.float32, device=dev)
            to_prev, to_next = torch.empty((self.send_prev_pos.numel(), 16), **f32), torch.empty((self.send_next_pos.numel(), 16), **f32)
            from_prev, from_next = torch.empty((p.n_halo_prev, 16), **f32), torch.empty((p.n_halo_next, 16), **f32)
            ops = [dist.P2POp(dist.isend, to_prev, p.prev, p.group), dist.P2POp(dist.isend, to_next, p.next, p.group),
                   dist.P2POp(dist.irecv, from_next, p.next, p.group), dist.P2POp(dist.irecv, from_prev, p.prev, p.group)]
            self._xchg = dict(to_prev=to_prev, to_next=to_next, from_prev=from_prev, from_next=from_next, ops=ops,
                              nbytes=(to_prev.numel() + to_next.numel()) * 4)
        return self._xchg

    def _exchange_state(self, t_s, dev):
        """halo exchange of the state rows of t_s (cell-sorted), on the current stream, through the persistent buffers."""
        x = self._state_exchange(dev)
        if x is None:
            return 0
        torch.index_select(t_s, 0, self.send_prev_pos, out=x["to_prev"])
        torch.index_select(t_s, 0, self.send_next_pos, out=x["to_next"])
        for w in dist.batch_isend_irecv(x["ops"]):
            w.wait()
        t_s.index_copy_(0, self.halo_prev_pos, x["from_prev"])
        t_s.index_copy_(0, self.halo_next_pos, x["from_next"])
        return x["nbytes"]

    def _rollout_split(self, A_s, An_s, Y_s, alive_s, flags_s, steps, fire_masks, fr, prm, pl, ws_fl, N, dev):
        """The no-alpha rollout: boundary pass -> exchange of the new boundary rows (side stream) || interior pass.  One C-ABI call
        per pass (sphnca_step_fwd_fired), persistent exchange buffers, two reused events: ~15 host-side operations per step."""
        C, _lib = self.C, self._lib
        lib, ptr, check = _lib.lib, _lib.ptr, _lib.check
        g = self.grid.c_grid()
        main = torch.cuda.current_stream(dev)
        st = _lib.stream_ptr()
        fb = self._work["fb"]
        split = self.split
        nbytes = self._exchange_state(A_s, dev)                              # halo rows of the initial state (no-op on one rank)
        ready, done = self._events
        waiting = False
        pargs = (ptr(pl["slots"]), ptr(pl["base"]), ptr(pl["J"]), ptr(pl["WP"]), ptr(pl["FLG"]), C.byref(prm))
        targs = (ptr(fb.rows), ptr(fb.count), ptr(fb.Y_c), ptr(fb.A_c))
        oargs = (ptr(alive_s), ptr(fb.O_c), ptr(flags_s), ptr(fb.relu_c), ptr(ws_fl), ws_fl.numel(), st)
        pA, pAn = ptr(A_s), ptr(An_s)
        timed = self.fused._timed                     # per-phase CUDA events when sphops.fused.PROFILE is set (tools/slab_profile.py)
        for k in range(steps):
            if not split:
                fire_s = self._fire_bytes(k, fire_masks, fr, N, dev)
                check(lib.sphnca_step_fwd_fired(g, ptr(fire_s), pA, *pargs, None, 1, 0, *targs, pAn, *oargs))
                A_s, An_s, pA, pAn = An_s, A_s, pAn, pA
                continue
            _, fire_b, fire_i = self._fire_bytes(k, fire_masks, fr, N, dev, split=True)
            if waiting:
                if self.fused.PROFILE is not None:                          # how long the main stream sits waiting for the exchange
                    a = torch.cuda.Event(enable_timing=True); a.record(main)
                main.wait_event(done)                                       # pass B reads the halo rows of the current state
                if self.fused.PROFILE is not None:
                    b = torch.cuda.Event(enable_timing=True); b.record(main)
                    self.fused.PROFILE.setdefault("wait_exchange", []).append((a, b))
                waiting = False
            # pass B: boundary particles; its list kernel also carries over every not-fired row and writes the static mask bits
            timed("pass_B", lib.sphnca_step_fwd_fired, g, ptr(fire_b), pA, *pargs, ptr(self.bnd_s), 1, self.n_bnd, *targs, pAn, *oargs)
            # exchange of the NEW state's boundary rows into the neighbours' halo slots
            if self.comm_stream is not None:
                ready.record(main)
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ready)
                    nbytes += self._exchange_state(An_s, dev)
                    done.record(self.comm_stream)
                waiting = True
            # pass I: interior particles (a second list over the same step: no carry-over, no flag writes)
            timed("pass_I", lib.sphnca_step_fwd_fired, g, ptr(fire_i), pA, *pargs, ptr(self.int_s), 0, 0, *targs, pAn, *oargs)
            if self.comm_stream is None:
                nbytes += self._exchange_state(An_s, dev)                   # blocking variant (overlap=False) / no-op on one rank
            A_s, An_s, pA, pAn = An_s, A_s, pAn, pA
        if waiting:
            main.wait_event(done)
        self.bytes_per_step = nbytes // max(steps, 1)
        return A_s[self.own_pos]
