"""torch.autograd.Function wrappers with the reference's class names and call signatures
(code/sphops/ops/func.py:19-463), launching libsphnca kernels on torch's current stream.

Differences from the reference, all deliberate:
  * no device synchronisation and no `.item()` round trips inside an operator;
  * outputs are fully written by the kernels (torch.empty, not torch.zeros);
  * `Blur.backward` and `Divergence.backward` return the TRUE adjoint w.r.t. A (the reference's
    kernels are not adjoints: operators_batch.py:751,763 and :517,523);
  * dL/dx is implemented for `Volume` only.  The reference's dL/dx branches of gradient / blur /
    divergence read uninitialised locals (kernels_impl.py:168, operators_batch.py:732) and are
    never exercised (positions never require grad in nca.py); requesting them raises.
"""
import torch

from .. import _lib
from .._lib import check, lib, ptr, require_cuda_f32, stream_ptr


def _h(h):
    return h.item() if isinstance(h, torch.Tensor) else float(h)


def _check_grid(grid, x, h):
    N, D = x.shape
    if grid.num_particles != N or int(grid.dims.shape[0]) != D:
        raise ValueError(f"grid was built for N={grid.num_particles}, D={int(grid.dims.shape[0])}; got x {tuple(x.shape)}")
    h = _h(h)
    if abs(h - grid.h) > 1e-9 * max(abs(grid.h), 1e-30):
        raise ValueError(f"h={h} differs from the cell size the grid was built with ({grid.h}); "
                         "the support radius must equal the cell size (3^D-cell stencil)")


def _no_dLdx(name):
    raise NotImplementedError(
        f"{name}: dL/dx is not implemented (the reference's branch reads uninitialised memory and is "
        "never used by nca.py); detach the positions")


class Volume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed_x, h, grid):
        """
        packed_x: f[N,D]
        h: float / f[1]
        grid: HashGrid
        """
        x = require_cuda_f32(packed_x, "packed_x")
        _check_grid(grid, x, h)
        N = x.shape[0]
        ctx.grid = grid
        v = torch.empty((N,), dtype=x.dtype, device=x.device)
        xv4 = torch.empty((N, 4), dtype=x.dtype, device=x.device)
        if N:
            xyz = grid.sorted_xyz(x)
            check(lib.sphnca_volume_fwd(grid.c_grid(), ptr(xyz), ptr(v), ptr(xv4), stream_ptr()))
            grid.remember_xv(x, v, xv4)
        ctx.save_for_backward(x, v)
        return v

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, packed_dLdv):
        x, v = ctx.saved_tensors
        grid = ctx.grid
        if not ctx.needs_input_grad[0]:
            return None, None, None
        N, D = x.shape
        dLdx = torch.empty_like(x)
        if N:
            g_s = grid.pack_rows(require_cuda_f32(packed_dLdv, "dLdv"), 1)
            xv4 = grid.sorted_xv(x, v)
            check(lib.sphnca_volume_bwd(grid.c_grid(), ptr(xv4), ptr(g_s), ptr(dLdx), stream_ptr()))
        return dLdx, None, None


class Gradient(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed_x, packed_v, packed_A, h, grid):
        """
        packed_x: f[N,D]
        packed_v: f[N]
        packed_A: f[N,F]
        h: float / f[1]
        grid: HashGrid
        """
        x = require_cuda_f32(packed_x, "packed_x")
        v = require_cuda_f32(packed_v, "packed_v")
        A = require_cuda_f32(packed_A, "packed_A")
        _check_grid(grid, x, h)
        N, D = x.shape
        assert A.dim() == 2 and A.shape[0] == N, "packed_A must be [N,F]"
        F = A.shape[1]
        ctx.grid = grid
        GA = torch.empty((N, F, D), dtype=A.dtype, device=A.device)
        if N and F:
            xv4 = grid.sorted_xv(x, v)
            A_s = grid.pack_rows(A, F)
            check(lib.sphnca_gradient_fwd(grid.c_grid(), ptr(xv4), ptr(A_s), F, ptr(GA), stream_ptr()))
        ctx.save_for_backward(x, v)
        ctx.F = F
        return GA

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, packed_dLdGA):
        x, v = ctx.saved_tensors
        grid = ctx.grid
        if ctx.needs_input_grad[0]:
            _no_dLdx("sphops.gradient")
        dLdA = None
        if ctx.needs_input_grad[2]:
            N, D = x.shape
            F = ctx.F
            dLdA = torch.empty((N, F), dtype=x.dtype, device=x.device)
            if N and F:
                G = require_cuda_f32(packed_dLdGA, "dLdGA").view(N, F * D)
                G_s = grid.pack_rows(G, F * D)
                xv4 = grid.sorted_xv(x, v)
                check(lib.sphnca_gradient_bwd(grid.c_grid(), ptr(xv4), ptr(G_s), F, ptr(dLdA), stream_ptr()))
        return None, None, dLdA, None, None


class Divergence(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed_x, packed_v, packed_A, h, grid):
        """
        packed_x: f[N,D]
        packed_v: f[N]
        packed_A: f[N,F,D]
        h: float / f[1]
        grid: HashGrid
        """
        x = require_cuda_f32(packed_x, "packed_x")
        v = require_cuda_f32(packed_v, "packed_v")
        A = require_cuda_f32(packed_A, "packed_A")
        _check_grid(grid, x, h)
        N, D = x.shape
        assert A.dim() == 3 and A.shape[0] == N and A.shape[2] == D, "packed_A must be [N,F,D]"
        F = A.shape[1]
        ctx.grid = grid
        DA = torch.empty((N, F), dtype=A.dtype, device=A.device)
        if N and F:
            xv4 = grid.sorted_xv(x, v)
            A_s = grid.pack_rows(A.view(N, F * D), F * D)
            check(lib.sphnca_divergence_fwd(grid.c_grid(), ptr(xv4), ptr(A_s), F, ptr(DA), stream_ptr()))
        ctx.save_for_backward(x, v)
        ctx.F = F
        return DA

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, packed_dLdDA):
        x, v = ctx.saved_tensors
        grid = ctx.grid
        if ctx.needs_input_grad[0]:
            _no_dLdx("sphops.divergence")
        dLdA = None
        if ctx.needs_input_grad[2]:
            N, D = x.shape
            F = ctx.F
            dLdA = torch.empty((N, F, D), dtype=x.dtype, device=x.device)
            if N and F:
                G_s = grid.pack_rows(require_cuda_f32(packed_dLdDA, "dLdDA"), F)
                xv4 = grid.sorted_xv(x, v)
                check(lib.sphnca_divergence_bwd(grid.c_grid(), ptr(xv4), ptr(G_s), F, ptr(dLdA), stream_ptr()))
        return None, None, dLdA, None, None


class Count(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed_x, h, grid):
        """
        packed_x: f[N,D]
        h: float / f[1]
        grid: HashGrid
        """
        x = require_cuda_f32(packed_x, "packed_x")
        _check_grid(grid, x, h)
        N = x.shape[0]
        c = torch.empty((N,), dtype=x.dtype, device=x.device)
        if N:
            check(lib.sphnca_count_fwd(grid.c_grid(), ptr(grid.sorted_xyz(x)), ptr(c), stream_ptr()))
        ctx.mark_non_differentiable(c)
        return c

    @staticmethod
    def backward(ctx, packed_dLdc):
        # a neighbour count is piecewise constant (the reference's backward raises, func.py:353-354)
        return None, None, None


class Blur(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed_x, packed_v, packed_A, h, grid):
        """
        packed_x: f[N,D]
        packed_v: f[N]
        packed_A: f[N,F]
        h: float / f[1]
        grid: HashGrid
        """
        x = require_cuda_f32(packed_x, "packed_x")
        v = require_cuda_f32(packed_v, "packed_v")
        A = require_cuda_f32(packed_A, "packed_A")
        _check_grid(grid, x, h)
        N = x.shape[0]
        assert A.dim() == 2 and A.shape[0] == N, "packed_A must be [N,F]"
        F = A.shape[1]
        ctx.grid = grid
        SA = torch.empty((N, F), dtype=A.dtype, device=A.device)
        if N and F:
            xv4 = grid.sorted_xv(x, v)
            A_s = grid.pack_rows(A, F)
            check(lib.sphnca_blur_fwd(grid.c_grid(), ptr(xv4), ptr(A_s), F, ptr(SA), stream_ptr()))
        ctx.save_for_backward(x, v)
        ctx.F = F
        return SA

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, packed_dLdSA):
        x, v = ctx.saved_tensors
        grid = ctx.grid
        if ctx.needs_input_grad[0]:
            _no_dLdx("sphops.blur")
        dLdA = None
        if ctx.needs_input_grad[2]:
            N = x.shape[0]
            F = ctx.F
            dLdA = torch.empty((N, F), dtype=x.dtype, device=x.device)
            if N and F:
                G_s = grid.pack_rows(require_cuda_f32(packed_dLdSA, "dLdSA"), F)
                xv4 = grid.sorted_xv(x, v)
                check(lib.sphnca_blur_bwd(grid.c_grid(), ptr(xv4), ptr(G_s), F, ptr(dLdA), stream_ptr()))
        return None, None, dLdA, None, None
