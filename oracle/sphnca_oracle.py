"""CPU oracle for the SPH-NCA per-timestep particle update.  TEST INFRASTRUCTURE ONLY.

This file is a plain numpy / torch-CPU restatement of the reference's algorithm for the hot path
(hash-grid build, the five neighbour-sweep operators, the NCA step and the K-step rollout).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs
may import it -- always as the checker or as the timed CPU baseline, never on the product path.
The product (`sph-nca_b200/`) never imports this module and raises if its CUDA library is absent.

Parity status: PINNED.  `tests/golden/*.npz` holds outputs of the unmodified reference
(`/root/reference/code`, run in the build container under Numba's CUDA simulator by
`tests/golden/make_golden.py`); `tests/test_oracle_golden.py` checks every function here
against those vectors.  The reference itself has no tests or golden vectors.

Arithmetic recipe (follows the reference's): pair offsets `r = x_j - x_i` are formed in the
input precision (fp32, reference `operators_batch.py:72,224`), everything after is evaluated in
float64, so this oracle is the "infinitely precise" version of the reference's fp32/fp64 mix.
Backward passes come from torch autograd over these float64 forwards, i.e. they are the TRUE
adjoints.  Where the reference's hand-written backward is a true adjoint (`Volume` dL/dx,
`Gradient` dL/dA) the two agree; where the reference is broken (`Blur`/`Divergence` backward, all
other dL/dx branches: SURVEY.md 8a rows a8,a10,a11) the product follows the true adjoint and the
deviation is documented in DESIGN.md.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

F64 = torch.float64


# --------------------------------------------------------------------------------------------
# Normalisation constants -- reference sphops/ops/kernels_impl.py:20-30 (poly6), :131-141 (spiky)
# --------------------------------------------------------------------------------------------
def poly6_normalization(dim: int, h: float) -> float:
    if dim == 2:
        return 4.0 / (math.pi * h ** 8)
    if dim == 3:
        return 315.0 / (64.0 * math.pi * h ** 9)
    raise NotImplementedError(dim)


def spiky_normalization(dim: int, h: float) -> float:
    if dim == 2:
        return 10.0 / (math.pi * h ** 5)
    if dim == 3:
        return 15.0 / (math.pi * h ** 6)
    raise NotImplementedError(dim)


# --------------------------------------------------------------------------------------------
# Hash grid -- reference sphops/hashgrid/single_impl.py:34-44,62-124 and sphops/preprocess.py:11-43
# --------------------------------------------------------------------------------------------
def as_dims(dims, dim: int) -> np.ndarray:
    """`dims` may be an int, a tuple or a tensor (reference preprocess.py:22-27)."""
    if isinstance(dims, (int, np.integer)):
        return np.full((dim,), int(dims), dtype=np.int64)
    if isinstance(dims, torch.Tensor):
        dims = dims.detach().cpu().numpy()
    out = np.asarray(dims, dtype=np.int64).reshape(-1)
    assert out.shape[0] == dim
    return out


def cell_coords(x: np.ndarray, h: float) -> np.ndarray:
    """int(floor(x/h)) with x promoted to float64 and IEEE division (single_impl.py:40)."""
    return np.floor(x.astype(np.float64) / np.float64(h)).astype(np.int64)


def grid_hash(x: np.ndarray, h: float, dims: np.ndarray) -> np.ndarray:
    """hi[n] = sum_d c_d * (g_d mod dims_d), c_0 = 1 (x fastest); Python modulo.  single_impl.py:34-44."""
    g = cell_coords(x, h)
    out = np.zeros(x.shape[0], dtype=np.int64)
    c = 1
    for d in range(x.shape[1]):
        out += c * np.mod(g[:, d], dims[d])
        c *= int(dims[d])
    return out.astype(np.int32)


@dataclass
class OracleGrid:
    h: float
    dims: np.ndarray            # i64[D]
    total_cells: int
    sections: tuple
    idx: np.ndarray             # i64[N]  batch-local sorted -> original index, concatenated
    cell_start: np.ndarray      # i32[B*cells]  batch-local, -1 = empty
    cell_end: np.ndarray        # i32[B*cells]
    num_points: np.ndarray      # i32[B]
    hi_sorted: list = field(default_factory=list)   # per batch element, sorted cell hashes


def build_hashgrid(packed_x, sections, h, dims) -> OracleGrid:
    """reference preprocess.py:11-43: per batch element hash -> (stable) sort -> run boundaries."""
    x = np.asarray(packed_x.detach().cpu().numpy() if isinstance(packed_x, torch.Tensor) else packed_x)
    dim = x.shape[1]
    dims = as_dims(dims, dim)
    h = float(h.item() if isinstance(h, torch.Tensor) else h)
    total_cells = int(np.prod(dims))
    idxs, starts, ends, his = [], [], [], []
    off = 0
    for n in sections:
        hi = grid_hash(x[off:off + n], h, dims)
        order = np.argsort(hi, kind="stable").astype(np.int64)
        hs = hi[order]
        cs = np.full(total_cells, -1, dtype=np.int32)
        ce = np.full(total_cells, -1, dtype=np.int32)
        if n > 0:
            # single_impl.py:93-109: a run starts where the hash changes
            first = np.flatnonzero(np.r_[True, hs[1:] != hs[:-1]])
            last = np.r_[first[1:], n]
            cs[hs[first]] = first.astype(np.int32)
            ce[hs[first]] = last.astype(np.int32)
        idxs.append(order); starts.append(cs); ends.append(ce); his.append(hs)
        off += n
    return OracleGrid(
        h=h, dims=dims, total_cells=total_cells, sections=tuple(int(s) for s in sections),
        idx=np.concatenate(idxs) if idxs else np.zeros(0, np.int64),
        cell_start=np.concatenate(starts) if starts else np.zeros(0, np.int32),
        cell_end=np.concatenate(ends) if ends else np.zeros(0, np.int32),
        num_points=np.asarray(sections, dtype=np.int32), hi_sorted=his)


def candidate_pairs(packed_x, grid: OracleGrid, chunk: int | None = None):
    """All (i, j) pairs the reference's sweep visits, in global packed indices.

    For every particle i the 3^D neighbouring cells (offsets -1,0,1 per axis, wrapped modulo
    `dims`, reference single_impl.py:46-60 and operators_batch.py:57-76) are enumerated and every
    particle j of those cells is a candidate.  An axis with dims<3 makes the stencil wrap onto
    itself, so a pair can appear 2 or 3 times -- exactly as in the reference.
    Yields (i, j) int64 arrays, chunked over i to bound memory.
    """
    x = np.asarray(packed_x.detach().cpu().numpy() if isinstance(packed_x, torch.Tensor) else packed_x)
    dim = x.shape[1]
    dims = grid.dims
    off, coff = 0, 0
    strides = np.cumprod(np.r_[1, dims[:-1]]).astype(np.int64)
    for b, n in enumerate(grid.sections):
        cs = grid.cell_start[coff:coff + grid.total_cells].astype(np.int64)
        ce = grid.cell_end[coff:coff + grid.total_cells].astype(np.int64)
        cnt = np.where(cs >= 0, ce - cs, 0)
        idx_b = grid.idx[off:off + n]
        step = n if not chunk else chunk
        for lo in range(0, n, max(step, 1)):
            hi_ = min(n, lo + step)
            g = np.mod(cell_coords(x[off + lo:off + hi_], grid.h), dims)      # [m, D]
            ii, jj = [], []
            for k in range(3 ** dim):
                kk = k
                ncell = np.zeros(hi_ - lo, dtype=np.int64)
                for d in range(dim):
                    dg = (kk % 3) - 1
                    kk //= 3
                    ncell += strides[d] * np.mod(g[:, d] + dims[d] + dg, dims[d])
                c = cnt[ncell]
                tot = int(c.sum())
                if tot == 0:
                    continue
                i_rep = np.repeat(np.arange(lo, hi_, dtype=np.int64), c)
                # position inside each run: 0..c-1
                run_start = np.repeat(np.cumsum(c) - c, c)
                within = np.arange(tot, dtype=np.int64) - run_start
                pos = np.repeat(cs[ncell], c) + within
                ii.append(off + i_rep)
                jj.append(off + idx_b[pos])
            if ii:
                yield np.concatenate(ii), np.concatenate(jj)
        off += n
        coff += grid.total_cells


def all_pairs(packed_x, grid: OracleGrid):
    ii, jj = [], []
    for i, j in candidate_pairs(packed_x, grid, chunk=16384):
        ii.append(i); jj.append(j)
    if not ii:
        z = np.zeros(0, np.int64)
        return z, z
    return np.concatenate(ii), np.concatenate(jj)


# --------------------------------------------------------------------------------------------
# Pair geometry
# --------------------------------------------------------------------------------------------
@dataclass
class Pairs:
    """Candidate pairs reduced to those inside the support radius, plus their geometry (float64)."""
    i: torch.Tensor       # i64[P]
    j: torch.Tensor       # i64[P]
    r: torch.Tensor       # f64[P, D]   r = x_j - x_i formed in fp32
    d2: torch.Tensor      # f64[P]
    n: int
    h: float
    dim: int


def make_pairs(packed_x: torch.Tensor, grid: OracleGrid, x_requires_grad: bool = False) -> Pairs:
    """In-radius pair list.  Out-of-radius candidates have exactly zero weight in every operator
    (poly6 `max(...,0)` kernels_impl.py:38, spiky `(h-d>0) and ...` :153, count `d2<h^2` :17)."""
    x = packed_x.detach().cpu()
    i, j = all_pairs(x, grid)
    i = torch.from_numpy(i); j = torch.from_numpy(j)
    r32 = x[j] - x[i]                                    # input precision, like the reference
    d2_32 = (r32 * r32).sum(-1)
    keep = d2_32.to(F64) < grid.h ** 2
    i, j = i[keep], j[keep]
    if x_requires_grad:
        r = (packed_x[j] - packed_x[i]).to(F64)
    else:
        r = r32[keep].to(F64)
    return Pairs(i=i, j=j, r=r, d2=(r * r).sum(-1), n=x.shape[0], h=grid.h, dim=x.shape[1])


def _scatter_sum(values: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.zeros((n,) + values.shape[1:], dtype=values.dtype)
    return out.index_add(0, index, values)


def poly6(p: Pairs) -> torch.Tensor:
    """max((h^2 - d^2)^3, 0)  -- kernels_impl.py:32-39"""
    return torch.clamp(p.h ** 2 - p.d2, min=0.0) ** 3


def spiky(p: Pairs) -> torch.Tensor:
    """3 (h-d)^2 / d * r for 0 < d < h else 0 -- kernels_impl.py:143-156 (points toward the neighbour)."""
    d = torch.sqrt(p.d2)
    ok = (d > 0) & (d < p.h)
    dsafe = torch.where(ok, d, torch.ones_like(d))
    mag = torch.where(ok, 3.0 * (p.h - dsafe) ** 2 / dsafe, torch.zeros_like(d))
    return mag[:, None] * p.r


# --------------------------------------------------------------------------------------------
# The five operators (forward) -- reference sphops/ops/operators_batch.py
# --------------------------------------------------------------------------------------------
def volume(p: Pairs) -> torch.Tensor:
    """v_i = 1 / (Cw * sum_j W_ij), self included.  operators_batch.py:16-77"""
    s = _scatter_sum(poly6(p), p.i, p.n)
    return 1.0 / (poly6_normalization(p.dim, p.h) * s)


def gradient(p: Pairs, v: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """GA_i[f,d] = Cg sum_j (A_j[f]-A_i[f]) v_j g_ij[d].  operators_batch.py:158-235"""
    v = v.to(F64); A = A.to(F64)
    gw = spiky(p) * v[p.j][:, None]                      # [P, D]
    dA = A[p.j] - A[p.i]                                 # [P, F]
    contrib = dA[:, :, None] * gw[:, None, :]            # [P, F, D]
    return spiky_normalization(p.dim, p.h) * _scatter_sum(contrib, p.i, p.n)


def blur(p: Pairs, v: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """SA_i[f] = Cw sum_j A_j[f] W_ij v_j.  operators_batch.py:600-671"""
    v = v.to(F64); A = A.to(F64)
    w = poly6(p) * v[p.j]
    return poly6_normalization(p.dim, p.h) * _scatter_sum(A[p.j] * w[:, None], p.i, p.n)


def divergence(p: Pairs, v: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """DA_i[f] = Cg sum_j v_j (A_j[f,:]-A_i[f,:]) . g_ij.  operators_batch.py:346-423"""
    v = v.to(F64); A = A.to(F64)
    g = spiky(p)
    dA = A[p.j] - A[p.i]                                 # [P, F, D]
    contrib = (dA * g[:, None, :]).sum(-1) * v[p.j][:, None]
    return spiky_normalization(p.dim, p.h) * _scatter_sum(contrib, p.i, p.n)


def count(p: Pairs) -> torch.Tensor:
    """#{j : d2 < h^2}, self included.  operators_batch.py:536-597, kernels_impl.py:10-18"""
    return _scatter_sum(torch.ones_like(p.d2), p.i, p.n)


# Hand-written backward formulas of the reference (only the two that ARE true adjoints).
def gradient_backward_dLdA_reference(p: Pairs, v, dLdGA) -> torch.Tensor:
    """dL/dA_i[f] = -Cg sum_j g_ij . (v_i G_j[f,:] + v_j G_i[f,:]).  operators_batch.py:319-342"""
    v = v.to(F64); G = dLdGA.to(F64)
    g = spiky(p)
    vec = -(v[p.i][:, None, None] * G[p.j] + v[p.j][:, None, None] * G[p.i])
    contrib = (vec * g[:, None, :]).sum(-1)
    return spiky_normalization(p.dim, p.h) * _scatter_sum(contrib, p.i, p.n)


def volume_backward_dLdx_reference(p: Pairs, v, dLdv) -> torch.Tensor:
    """dL/dx_i = Cw sum_j (v_i^2 g_i + v_j^2 g_j) (-6 (h^2-d^2)^2) r.  operators_batch.py:137-150, kernels_impl.py:41-53"""
    v = v.to(F64); g = dLdv.to(F64)
    mult = v[p.i] ** 2 * g[p.i] + v[p.j] ** 2 * g[p.j]
    mag = mult * -6.0 * torch.clamp(p.h ** 2 - p.d2, min=0.0) ** 2
    return poly6_normalization(p.dim, p.h) * _scatter_sum(mag[:, None] * p.r, p.i, p.n)


# --------------------------------------------------------------------------------------------
# NCA step and rollout -- reference code/nca.py:17-31,60-117,124-164
# --------------------------------------------------------------------------------------------
@dataclass
class StepConfig:
    in_features: int = 16
    update_rule: str = "gated"          # nca.py:45-48,102-108
    use_alpha: bool = True              # nca.py:17-21
    normalize_perception: float = -1.0  # nca.py:65-69
    model_fire_rate: float = 0.5        # SPHNCA.fire_rate
    fire_rate: float | None = None      # per-call override, nca.py:98-99


def cell_activity(A: torch.Tensor, use_alpha: bool) -> torch.Tensor:
    alpha = A[..., 3]
    return alpha if use_alpha else torch.ones_like(alpha)


def life_mask(p: Pairs, v, activity: torch.Tensor) -> torch.Tensor:
    """(blur(activity > 0.1) > 0.1)  -- nca.py:71-79.  Returns bool [N,1]."""
    m = (activity > 0.1).to(F64)
    return blur(p, v, m) > 0.1


def perceive(p: Pairs, v, A, cfg: StepConfig) -> torch.Tensor:
    """y = [A, dA/dx, dA/dy] (3F wide); z-derivative dropped.  nca.py:23-31,65-69,95"""
    gA = gradient(p, v, A)
    if cfg.normalize_perception > 0:
        gA = p.h * gA * cfg.normalize_perception
    return torch.cat([A.to(F64), gA[..., 0], gA[..., 1]], dim=-1)


def nca_step(p: Pairs, v, A, params, update_mask, cfg: StepConfig, return_aux=False, relu_bits=None):
    """One SPHNCA.forward (nca.py:87-117) with the stochastic fire mask supplied by the caller.

    params = (W1[Hd,3F], b1[Hd], W2[O,Hd], b2[O]); update_mask bool[N].
    relu_bits: optional bool[N,Hd], or (rows, bool[len(rows),Hd]) for a subset of the rows -- the ReLU decisions of ANOTHER evaluation (the GPU's saved bits) imposed on this
    one: hidden = pre * bit instead of max(pre, 0).  Where a pre-activation is within rounding of 0 the two evaluations
    may decide differently, and a flipped unit changes a gradient row by O(1/Hd); with the bits imposed the comparison
    measures arithmetic error only (the aux output reports every unit that differs and how close to 0 it is).
    """
    W1, b1, W2, b2 = [t.to(F64) for t in params]
    A = A.to(F64)
    nf = cfg.in_features
    prev_mask = life_mask(p, v, cell_activity(A, cfg.use_alpha)[:, None])
    y = perceive(p, v, A, cfg)
    pre = y @ W1.T + b1
    if relu_bits is None:
        bits = None
        hid = torch.relu(pre)
    else:
        if isinstance(relu_bits, tuple):            # (rows i64[n], bits bool[n,Hd]): imposed on those rows only
            rows, sub = relu_bits
            bits = pre.detach() > 0
            bits[rows] = sub
        else:
            bits = relu_bits
        hid = pre * bits.to(F64)
    dA = hid @ W2.T + b2
    fire_rate = cfg.model_fire_rate if cfg.fire_rate is None else cfg.fire_rate
    if cfg.update_rule == "gated":
        gate = torch.sigmoid(dA[..., :nf])
        delta = torch.tanh(dA[..., nf:-1])
        mult = torch.sigmoid(dA[..., [-1]])
        nA = A * gate + delta * mult
    elif cfg.update_rule == "orig":
        nA = A + dA * cfg.model_fire_rate / fire_rate
    else:
        raise ValueError(cfg.update_rule)
    nA = torch.where(update_mask[..., None], nA, A)
    new_mask = life_mask(p, v, cell_activity(nA, cfg.use_alpha)[..., None])
    living = prev_mask & new_mask
    out = nA * living.to(F64)
    if return_aux:
        with torch.no_grad():
            # conditioning of every pre-activation: |pre| / (sum_k |y_k W1_nk| + |b1_n|); fp32 evaluations can only decide
            # a ReLU differently where this is at rounding level
            scale = y.abs() @ W1.abs().T + b1.abs()
            margin = pre.abs() / scale.clamp_min(1e-300)
        return out, dict(prev_mask=prev_mask, new_mask=new_mask, y=y, dA=dA, nA_unmasked=nA, pre=pre.detach(), margin=margin)
    return out


def rollout(p: Pairs, v, A0, params, update_masks, cfg: StepConfig, relu_bits=None, aux_out=None, trajectory=None):
    """K applications of nca_step (nca.py:138-141).  update_masks: bool[K, N].  Returns [A_0..A_K].
    relu_bits: optional list of K bool[N,Hd] tensors imposed per step (see nca_step); aux_out: optional list that
    receives every step's aux dict.
    trajectory: optional list of K+1 states of ANOTHER evaluation (the GPU's fp32 states).  Step k is then evaluated AT
    trajectory[k] (value substituted, gradient passed straight through to this rollout's own state), so out[k+1] is the
    single-step result from identical inputs and the backward is the chain of THIS oracle's Jacobians along that
    trajectory: rounding differences of earlier steps, which the dynamics amplify step by step, are taken out of both
    comparisons (parity is defined per step from identical inputs)."""
    out = [A0.to(F64)]
    A = A0
    for k in range(update_masks.shape[0]):
        bits = None if relu_bits is None else relu_bits[k]
        if trajectory is not None and k > 0:
            A = A + (trajectory[k].to(F64) - A).detach()
        if aux_out is not None:
            A, aux = nca_step(p, v, A, params, update_masks[k], cfg, return_aux=True, relu_bits=bits)
            aux_out.append(aux)
        else:
            A = nca_step(p, v, A, params, update_masks[k], cfg, relu_bits=bits)
        out.append(A)
    return out


def bench_loss(A_final: torch.Tensor) -> torch.Tensor:
    """Scalar loss used by bench/tests in place of the image losses: mean(A_K[:, :4]^2)."""
    return (A_final[:, :4] ** 2).mean()


# --------------------------------------------------------------------------------------------
# Convenience: one call from raw inputs (used by tests and by the CPU-baseline leg of bench.py)
# --------------------------------------------------------------------------------------------
def prepare(packed_x: torch.Tensor, sections, h, dims):
    grid = build_hashgrid(packed_x, sections, h, dims)
    pairs = make_pairs(packed_x, grid)
    return grid, pairs, volume(pairs)


def rollout_fwd_bwd(packed_x, sections, h, dims, A0, params, update_masks, cfg: StepConfig, relu_bits=None, prepared=None,
                    return_all=False, trajectory=None):
    """Grid build + volume + K steps + bench loss + backward.  Returns (A_K, loss, dA0, dparams)
    [+ (states, per-step aux) with return_all].  relu_bits: per-step ReLU decisions to impose (see nca_step);
    prepared: a (grid, pairs, v) tuple from `prepare` to reuse."""
    grid, pairs, v = prepared if prepared is not None else prepare(packed_x, sections, h, dims)
    A0 = A0.detach().to(F64).requires_grad_(True)
    ps = [t.detach().to(F64).requires_grad_(True) for t in params]
    aux = [] if return_all else None
    states = rollout(pairs, v, A0, ps, update_masks, cfg, relu_bits=relu_bits, aux_out=aux, trajectory=trajectory)
    last = states[-1]
    if trajectory is not None:          # the loss (and its gradient) at the other evaluation's final state too
        last = last + (trajectory[-1].to(F64) - last).detach()
    loss = bench_loss(last)
    loss.backward()
    res = (states[-1].detach(), loss.detach(), A0.grad, [t.grad for t in ps])
    if return_all:
        return res + ([s.detach() for s in states], aux)
    return res
