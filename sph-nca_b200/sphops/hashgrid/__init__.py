"""Hash-grid container with the reference's field names (code/sphops/hashgrid/single_impl.py:127-171)."""
import ctypes as C

import torch

from .. import _lib

INTTYPE = torch.int32


class Hashgrid:
    """Cell tables of one packed batch.

    Reference-visible fields (same meaning as single_impl.py:127-171 after `merge()`):
      h: float, dims: i64[D] (device), total_cells: int,
      idx: i64[N]  section-local sorted->original index, concatenated over sections,
      cell_start / cell_end: i32[B*total_cells] section-local positions, -1 = empty cell,
      num_points: i32[B], occupied_cells / points_per_cells: list[tensor] (lazy; unused by kernels).
    Internal (consumed by libsphnca): cell_off, sorted_id, sorted_cell and a cache of the
    cell-sorted {x,y,z,v} array.
    """

    def __init__(self, h, dims, total_cells, sections, device):
        self.h = float(h)
        self.dims = dims
        self.total_cells = int(total_cells)
        self.sections = tuple(int(s) for s in sections)
        self.device = device
        self.idx = self.cell_start = self.cell_end = self.num_points = None
        self.cell_off = self.sorted_id = self.sorted_cell = None
        self._c = None
        self._occupied = None
        self._xv_key = None
        self._xv4 = None
        self._xyz_key = None
        self._xyz4 = None
        self._staged_ok = None
        self._pairs_key = None
        self._pairs = None
        self.period = (0.0, 0.0, 0.0)
        self.band = None
        self._vimg = None

    def set_periodic(self, period, band=None):
        """Declare the cloud periodic: period[d] > 0 = the period along axis d, which must equal dims[d] * h (then the modulo
        wrap of the cell hash, reference single_impl.py:54-55, IS the period) -- what `nca.sample_plane_wrapped` emulates with
        ghost particles (reference nca.py:200-299).  Neighbours found through the wrap are taken at x_j +- period by
        `sphops.volume` and by the pair lists (so by every kernel of the fused rollout); the other operators refuse a
        periodic grid.  Call before the first use of the grid.
        band = ((lo_x, lo_y), (hi_x, hi_y)): ghost-volume emulation for periodic PLANES -- the reference builds its grid and
        volumes once on the ghost-padded cloud (nca.py:262-263), so a ghost copy (of a particle with x <= lo / x >= hi, nca.py:182-197)
        carries the volume of its own, outward-truncated neighbourhood; with `band` the pair lists reproduce exactly that
        (include/sphnca.h: sphnca_grid.image_volume).  Without it neighbours through the wrap keep their own volume."""
        per = [float(p) for p in period] + [0.0] * (3 - len(period))
        D = int(self.dims.shape[0])
        for d in range(D):
            if per[d] > 0 and (self._dims_host[d] < 3 or abs(per[d] / self.h - self._dims_host[d]) > 1e-6 * self._dims_host[d]):
                raise ValueError(f"periodic axis {d}: period {per[d]} must equal dims ({self._dims_host[d]}) x h ({self.h}), dims >= 3")
        self.period = tuple(per[:3])
        if band is not None:
            if per[2] > 0 or not (per[0] > 0 and per[1] > 0):
                raise ValueError("ghost-volume emulation is defined for planes periodic in x and y")
            band = (tuple(float(b) for b in band[0]), tuple(float(b) for b in band[1]))
        self.band = band
        self._c = self._pairs_key = self._pairs = self._vimg = None
        return self

    @property
    def periodic(self):
        return any(p > 0 for p in self.period)

    @staticmethod
    def periodic_compatible(h, dims_host, period):
        """True when `set_periodic(period)` would accept this grid (period == dims * h on every periodic axis, dims >= 3)."""
        return all(p <= 0 or (d >= 3 and abs(p / h - d) <= 1e-6 * d) for p, d in zip(period, dims_host))

    # -- reference API -------------------------------------------------------------------
    def get_num_blocks(self, block_size):
        return torch.ceil(self.num_points / block_size).to(INTTYPE)

    def _occupancy(self):
        if self._occupied is None:
            occ, cnt = [], []
            B = len(self.sections)
            cs = self.cell_start.view(B, self.total_cells)
            ce = self.cell_end.view(B, self.total_cells)
            for b in range(B):
                cells = torch.nonzero(cs[b] >= 0).flatten()
                occ.append(cells.to(INTTYPE))
                cnt.append((ce[b][cells] - cs[b][cells]).to(torch.int64))
            self._occupied = (occ, cnt)
        return self._occupied

    @property
    def occupied_cells(self):
        return self._occupancy()[0]

    @property
    def points_per_cells(self):
        return self._occupancy()[1]

    # -- libsphnca side ------------------------------------------------------------------
    @property
    def num_particles(self):
        return sum(self.sections)

    def c_grid(self):
        if self._c is None:
            g = _lib.CGrid()
            D = int(self.dims.shape[0])
            g.dim = D
            g.num_sections = len(self.sections)
            g.num_particles = self.num_particles
            g.cells_per_section = self.total_cells
            for d in range(3):
                g.dims[d] = self._dims_host[d] if d < D else 1
            g.h = self.h
            for d in range(3):
                g.period[d] = self.period[d]
                g.band_lo[d] = self.band[0][d] if (self.band and d < 2) else 0.0
                g.band_hi[d] = self.band[1][d] if (self.band and d < 2) else 0.0
            g.image_volume = self._vimg.data_ptr() if self._vimg is not None else None
            g.cell_off = self.cell_off.data_ptr()
            g.sorted_id = self.sorted_id.data_ptr()
            g.sorted_cell = self.sorted_cell.data_ptr()
            self._c = g
        return C.byref(self._c)

    @staticmethod
    def _key(t):
        return (t.data_ptr(), t._version, tuple(t.shape))

    def sorted_xyz(self, x):
        """float4 {x,y,z,0} in cell-sorted order (cached per position tensor)."""
        key = self._key(x)
        if self._xyz_key != key:
            out = torch.empty((x.shape[0], 4), dtype=torch.float32, device=x.device)
            _lib.check(_lib.lib.sphnca_pack_xv(self.c_grid(), _lib.ptr(x), None, _lib.ptr(out), _lib.stream_ptr()))
            self._xyz_key, self._xyz4, self._xyz_ref = key, out, x   # the ref pins the storage address
        return self._xyz4

    def sorted_xv(self, x, v):
        """float4 {x,y,z,v} in cell-sorted order (cached per (x, v) pair)."""
        key = (self._key(x), self._key(v))
        if self._xv_key != key:
            out = torch.empty((x.shape[0], 4), dtype=torch.float32, device=x.device)
            _lib.check(_lib.lib.sphnca_pack_xv(self.c_grid(), _lib.ptr(x), _lib.ptr(v), _lib.ptr(out), _lib.stream_ptr()))
            self._xv_key, self._xv4, self._xv_ref = key, out, (x, v)
        return self._xv4

    def remember_xv(self, x, v, xv4):
        self._xv_key, self._xv4, self._xv_ref = (self._key(x), self._key(v)), xv4, (x, v)

    # fraction of the free device memory the pair lists of one grid may take (about 1.3 KB per particle)
    PAIR_LIST_MEMORY_FRACTION = 0.4

    def pair_lists(self, x, v, blur=True, z=False):
        """Per-pair neighbour lists and kernel weights of this grid (csrc/pairs.cu), cached per (x, v) pair; None when
        disabled or when they would not fit the memory budget.  Building them reads back ONE integer (the total number
        of list entries) -- the only host synchronisation besides `staged_sweep_ok`.  blur=False leaves out the life-mask
        blur weights WB (8 of the 28 bytes per pair; only models with an alpha channel read them); lists built without
        them are rebuilt when a later caller needs them."""
        if not _lib.USE_PAIR_LISTS or _lib.FORCE_SWEEP_PATH is not None or "sphnca_pairs_fill" in _lib.MISSING_SYMBOLS:
            if self.periodic:
                raise RuntimeError("a periodic grid needs the pair lists (SPHNCA_PAIRS=0 / SPHNCA_SWEEP_PATH are set)")
            return None
        key = (self._key(x), self._key(v))
        if self._pairs_key == key and self._pairs is not None and ((blur and self._pairs["WB"] is None) or (z and self._pairs["WPZ"] is None)):
            blur, z = blur or self._pairs["WB"] is not None, z or self._pairs["WPZ"] is not None
            self._pairs_key = None                                  # built without the blur / z weights: build again with them
        if self._pairs_key != key:
            self._pairs_key, self._pairs, self._pairs_ref = key, None, (x, v)
            if self.num_particles:
                # a periodic grid has no cell-sweep fallback: its lists are built whatever they cost
                self._pairs = self.build_pair_lists(self.sorted_xv(x, v), blur, float("inf") if self.periodic else None, z=z)
        return self._pairs

    # what one list entry costs (J + WP + WA [+ WB]) and the measured entries per particle of the BASELINE workloads
    # (35.9 slots in padded 2-D at h = 3.2 spacings, 39 in 3-D at h = 1.97 spacings): used to plan micro-batches
    PAIR_ENTRY_BYTES = {False: 20, True: 28}
    PAIR_SLOTS_ESTIMATE = 40

    def build_pair_lists(self, xv4, blur=True, budget_bytes=None, z=False):
        """The pair lists of this grid from its cell-sorted {x,y,z,v} array (uncached).  None when they would exceed
        `budget_bytes` (default: PAIR_LIST_MEMORY_FRACTION of the device memory torch does not hold)."""
        N, dev = self.num_particles, xv4.device
        if self.band is not None and self._vimg is None:    # periodic plane with ghost-volume emulation: the copies' volumes first
            vimg = torch.empty((N, 9), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib.sphnca_volume_images(self.c_grid(), _lib.ptr(xv4), _lib.ptr(vimg), _lib.stream_ptr()))
            self._vimg, self._c = vimg, None
        ng = (N + 7) // 8                                   # list groups of 8 particles (csrc/pairs.cu)
        slots = torch.empty((ng,), dtype=torch.int32, device=dev)
        st = _lib.stream_ptr()
        _lib.check(_lib.lib.sphnca_pairs_count(self.c_grid(), _lib.ptr(xv4), _lib.ptr(slots), st))
        cum = torch.cumsum(slots, 0, dtype=torch.int64)
        base = ((cum - slots) * 8).contiguous()
        total = int(cum[-1].item()) * 8
        if budget_bytes is None:
            # capacity minus torch's live tensors (host-side counters; cudaMemGetInfo stalls ~25 ms on a busy device)
            free = torch.cuda.get_device_properties(dev).total_memory - torch.cuda.memory_allocated(dev)
            budget_bytes = self.PAIR_LIST_MEMORY_FRACTION * free
        if not (0 < total and total * (self.PAIR_ENTRY_BYTES[bool(blur)] + (4 if z else 0)) <= budget_bytes):
            return None
        f32 = dict(dtype=torch.float32, device=dev)
        J = torch.empty((total,), dtype=torch.int32, device=dev)
        WP, WA = (torch.empty((total, 2), **f32) for _ in range(2))
        WB = torch.empty((total, 2), **f32) if blur else None
        WPZ = torch.empty((total,), **f32) if z else None            # tangent-frame sweeps only (surface clouds)
        SXY = torch.empty((N, 2), **f32)
        FLG = torch.empty((N,), dtype=torch.uint8, device=dev)
        _lib.check(_lib.lib.sphnca_pairs_fill(self.c_grid(), _lib.ptr(xv4), _lib.ptr(slots), _lib.ptr(base), _lib.ptr(J),
                                              _lib.ptr(WP), _lib.ptr(WB), _lib.ptr(WA), _lib.ptr(SXY), _lib.ptr(FLG), _lib.ptr(WPZ), st))
        return dict(slots=slots, base=base, J=J, WP=WP, WB=WB, WA=WA, SXY=SXY, FLG=FLG, WPZ=WPZ, entries=total)

    # -- micro-batches over sections ---------------------------------------------------------
    def section_ranges(self):
        """[(first sorted position, last + 1)] per section: sections are contiguous, in order, in the sorted arrays."""
        out, off = [], 0
        for n in self.sections:
            out.append((off, off + n))
            off += n
        return out

    def subgrid(self, s0, s1):
        """The grid of sections [s0, s1) alone, as a self-contained problem on SLICES of this grid's sorted arrays.
        Batch elements are independent (reference operators_batch.py:34-42: separate OFFSET / CELL_OFFSET ranges) and
        contiguous in cell-sorted order, so with row p0 = the first sorted position of section s0 re-based to 0 every
        kernel runs unchanged on `array[p0:p1]`: cell offsets and cell ids are made relative here (two small copies),
        `sorted_id` keeps the GLOBAL particle index (it is only used to gather from / scatter to particle-order arrays).
        Used by the rollout node to hold pair lists for a few sections at a time (`sphops.fused`)."""
        ranges = self.section_ranges()
        p0, p1 = ranges[s0][0], ranges[s1 - 1][1]
        C_ = self.total_cells
        sub = Hashgrid(self.h, self.dims, self.total_cells, self.sections[s0:s1], self.device)
        sub._dims_host = self._dims_host
        sub.cell_off = (self.cell_off[s0 * C_: s1 * C_ + 1] - p0).contiguous()
        sub.sorted_cell = (self.sorted_cell[p0:p1] - s0 * C_).contiguous()
        sub.sorted_id = self.sorted_id[p0:p1]
        sub.num_points = self.num_points[s0:s1]
        sub.period, sub.band = self.period, self.band
        sub._vimg = self._vimg[p0:p1] if self._vimg is not None else None
        sub.range = (p0, p1)
        return sub

    def staged_sweep_ok(self):
        """True when (nearly) every 128-particle tile of this grid fits the shared-memory staging buffer of
        sphnca_step_perceive_staged (planar / padded-2-D clouds).  Evaluated once per grid: one small kernel and
        ONE device->host read of two integers (the only synchronisation the fused step ever does)."""
        if self._staged_ok is None:
            if self.num_particles == 0 or "sphnca_stage_fit" in _lib.MISSING_SYMBOLS or _lib.FORCE_SWEEP_PATH is not None:
                self._staged_ok = False
            else:
                counts = torch.empty((2,), dtype=torch.int32, device=self.cell_off.device)
                _lib.check(_lib.lib.sphnca_stage_fit(self.c_grid(), _lib.ptr(counts), _lib.stream_ptr()))
                tiles, unfit = counts.tolist()
                self._staged_ok = unfit * 10 <= tiles
        return self._staged_ok

    def pack_rows(self, t, row):
        out = torch.empty((self.num_particles, row), dtype=torch.float32, device=t.device)
        if self.num_particles:
            _lib.check(_lib.lib.sphnca_pack_rows(self.c_grid(), _lib.ptr(t), row, _lib.ptr(out), _lib.stream_ptr()))
        return out

    def unpack_rows(self, t, row):
        out = torch.empty((self.num_particles, row), dtype=torch.float32, device=t.device)
        if self.num_particles:
            _lib.check(_lib.lib.sphnca_unpack_rows(self.c_grid(), _lib.ptr(t), row, _lib.ptr(out), _lib.stream_ptr()))
        return out

    def pack_bytes(self, t):
        out = torch.empty((self.num_particles,), dtype=torch.uint8, device=t.device)
        if self.num_particles:
            _lib.check(_lib.lib.sphnca_pack_bytes(self.c_grid(), _lib.ptr(t), _lib.ptr(out), _lib.stream_ptr()))
        return out
