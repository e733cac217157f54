"""`initialize_hashgrid` with the reference's signature (code/sphops/preprocess.py:11-43).

The reference loops over batch elements on the host (hash kernel, torch.sort, table kernel,
unique_consecutive; two device syncs each).  Here ONE call of `sphnca_hashgrid_build` handles
every section on the current stream with no synchronisation; `occupied_cells` /
`points_per_cells` (never read by any kernel) are derived lazily on first access.
"""
import ctypes as C
import math

import torch

from . import _lib
from .hashgrid import Hashgrid

__all__ = ["initialize_hashgrid", "Hashgrid"]


def _host_dims(dims, DIM):
    if isinstance(dims, int):
        return [int(dims)] * DIM
    if isinstance(dims, torch.Tensor):
        return [int(d) for d in dims.detach().cpu().tolist()]   # device->host copy, like the reference's .item()
    return [int(d) for d in dims]


@torch.no_grad()
def initialize_hashgrid(packed_points, sections, h, dims):
    """
    packed_points: f[N,D] - particle positions (packed)
    sections: i[B] - packed sections
    h: float / float[1] - hashgrid cell size
    dims: int / tuple / i[D] - hashgrid dimensions
    """
    x = _lib.require_cuda_f32(packed_points, "packed_points")
    N, DIM = x.shape
    dims_host = _host_dims(dims, DIM)
    assert len(dims_host) == DIM
    if DIM not in (2, 3):
        raise NotImplementedError(f"DIM={DIM}: only 2-D and 3-D kernels exist (as in the reference)")
    sections = tuple(int(s) for s in sections)
    assert sum(sections) == N, f"sections {sections} do not sum to N={N}"
    B = len(sections)
    total_cells = math.prod(dims_host)
    h = h.item() if isinstance(h, torch.Tensor) else float(h)
    dev = x.device

    dims_t = dims if isinstance(dims, torch.Tensor) else torch.tensor(dims_host, dtype=torch.int64, device=dev)
    grid = Hashgrid(h, dims_t, total_cells, sections, dev)
    grid._dims_host = dims_host
    i32 = dict(dtype=torch.int32, device=dev)
    grid.idx = torch.empty((N,), dtype=torch.int64, device=dev)
    grid.cell_start = torch.empty((B * total_cells,), **i32)
    grid.cell_end = torch.empty((B * total_cells,), **i32)
    grid.num_points = torch.empty((B,), **i32)
    grid.cell_off = torch.empty((B * total_cells + 1,), **i32)
    grid.sorted_id = torch.empty((N,), **i32)
    grid.sorted_cell = torch.empty((N,), **i32)
    if B == 0:
        return grid

    nbytes = C.c_size_t(0)
    _lib.check(_lib.lib.sphnca_hashgrid_workspace_bytes(N, B, total_cells, C.byref(nbytes)))
    ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=dev)
    sec_arr = (C.c_int64 * B)(*sections)
    dims_arr = (C.c_int64 * DIM)(*dims_host)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.sphnca_hashgrid_build(
            _lib.ptr(x), N, DIM, sec_arr, B, C.c_double(h), dims_arr,
            _lib.ptr(grid.idx), _lib.ptr(grid.cell_start), _lib.ptr(grid.cell_end), _lib.ptr(grid.num_points),
            _lib.ptr(grid.cell_off), _lib.ptr(grid.sorted_id), _lib.ptr(grid.sorted_cell),
            _lib.ptr(ws), nbytes.value, _lib.stream_ptr()))
    return grid
