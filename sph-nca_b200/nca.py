"""SPH-NCA model and rollout drivers with the reference's API (code/nca.py), on top of the
B200 `sphops` package in this directory.

Kept from the reference so that its train.py / test.py run unchanged against this module:
`USE_3D`, `cell_activity`, `default_feature_processs`, `SPHNCA` (constructor arguments, the
`model` nn.Sequential and therefore the checkpoint keys `model.0.*`, `model.2.*`, `init_grid`,
`perceive`, `life_mask`, `to_rgba`, `forward`, `normalize_grads`), `sample_plane`.

B200-specific: when the step is the default one (16 channels, hidden 256, default feature
process, no per-step feature hook) `SPHNCA.forward` dispatches to the fused CUDA step
(`sphops.fused.nca_step`), otherwise it composes the general `sphops` operators exactly like
the reference does.  Both paths draw the stochastic fire mask with the same single
`torch.rand(N)` call as the reference (nca.py:100), so a seeded run consumes the RNG identically.
"""
import random

import torch
import torch.nn as nn
import torch.nn.functional as F

import sphops

USE_3D = False
# set to False to force the operator-by-operator path (used by the parity tests)
USE_FUSED_STEP = True
# set to False to run periodic planes (sample_plane_wrapped) on ghost particles like the reference even where the native periodic
# neighbour search applies (no alpha channel, period == dims * h)
USE_NATIVE_PERIODIC = True
# set to False to run rollouts as K separate fused steps (one autograd node per step, states permuted back
# to particle order after every step) instead of one cell-sorted-resident K-step node
USE_FUSED_ROLLOUT = True


def cell_activity(A, use_alpha=True):
    """Per-particle activity: the alpha channel (index 3) or all-ones (reference nca.py:17-21)."""
    alpha = A[..., 3]
    return alpha if use_alpha else torch.ones_like(alpha)


def default_feature_processs(features):
    """[(A [N,F], gA [N,F,3])...] -> [A | dA/dx | dA/dy] concatenated on the channel axis.
    The z derivative is dropped (reference nca.py:23-31); the name's spelling is the reference's."""
    values = [a for a, _ in features]
    ddx = [g[..., 0] for _, g in features]
    ddy = [g[..., 1] for _, g in features]
    return torch.cat(values + ddx + ddy, dim=-1)


class SPHNCA(nn.Module):
    """Neural cellular automaton on particles: SPH-gradient perception -> 2-layer MLP -> gated update.

    Mirrors reference nca.py:33-122.  `model` stays an nn.Sequential(Linear, ReLU, Linear) so
    reference checkpoints load; the fused kernels read its parameters in place.
    """

    def __init__(self, in_features, fire_rate, update_rule='gated', hidden_features=256,
                 base_feature_process=None, normalize_grad=False, use_alpha=True, normalize_perception=-1):
        super().__init__()
        self.in_features = in_features
        self.fire_rate = fire_rate
        self.update_rule = update_rule
        self._default_process = base_feature_process is None
        self.base_feature_process = default_feature_processs if base_feature_process is None else base_feature_process
        self.normalize_grad = normalize_grad
        self.normalize_perception = normalize_perception
        self.use_alpha = use_alpha
        self.hidden_features = hidden_features

        if update_rule == 'gated':
            out_features = 2 * in_features + 1
        elif update_rule == 'orig':
            out_features = in_features
        else:
            raise KeyError(update_rule)
        self.model = nn.Sequential(
            nn.Linear(3 * in_features, hidden_features, bias=True),
            nn.ReLU(),
            nn.Linear(hidden_features, out_features, bias=True),
        )
        if update_rule == 'orig':          # "do nothing" initialisation of the residual rule
            with torch.no_grad():
                self.model[-1].weight.zero_()
                self.model[-1].bias.zero_()

    # ---- building blocks (reference nca.py:60-85) ------------------------------------------
    def init_grid(self, x, sections, h, dims):
        grid = sphops.initialize_hashgrid(x, sections, h, dims)
        return grid, sphops.volume(x, h, grid)

    def perception_scale(self, h):
        h = h.item() if isinstance(h, torch.Tensor) else h
        return h * self.normalize_perception if self.normalize_perception > 0 else 1.0

    def perceive(self, x, v, A, h, grid):
        gA = sphops.gradient(x, v, A, h, grid)
        if self.normalize_perception > 0:
            gA = h * gA * self.normalize_perception
        return [(A, gA)]

    def life_mask(self, x, v, activity, h, grid):
        if activity.shape[-1] > 1:
            activity = cell_activity(activity, self.use_alpha)[:, None]
        alive = (activity > 0.1).float()
        return sphops.blur(x, v, alive, h, grid).detach() > 0.1

    def to_rgba(self, A):
        return torch.cat([A[..., :3], cell_activity(A, self.use_alpha)[..., None]], dim=-1)

    # ---- one step (reference nca.py:87-117) ------------------------------------------------
    def _fusable(self, x, A, feature_process):
        return (USE_FUSED_STEP and feature_process is None and self._default_process
                and self.in_features == 16 and self.hidden_features == 256
                and x.shape[-1] == 3 and A.is_cuda and sphops.fused_available())

    def _tangent_frame(self, x, v, A, grid, feature_process):
        """[N,8] = {T, 0, N x T, 0} when `feature_process` is a `feature_process_tangent(N, T)` hook and this step can run on the
        fused kernels without a tape (inference), else None."""
        NT = getattr(feature_process, "tangent_frame", None)
        if NT is None or not self._fusable(x, A, None):
            return None
        if torch.is_grad_enabled() and (A.requires_grad or any(p.requires_grad for p in self.parameters())):
            return None                                   # training through a surface rollout: the operator-by-operator path has the tape
        if not sphops.fused.frame_step_available(grid, x, v):
            return None
        Nn, T = NT
        frame = torch.zeros((x.shape[0], 8), dtype=torch.float32, device=x.device)
        frame[:, 0:3] = T
        frame[:, 4:7] = torch.linalg.cross(Nn, T)
        return frame

    def forward(self, x, v, A, h, grid, fire_rate=None, feature_process=None):
        if fire_rate is None:
            fire_rate = self.fire_rate
        frame = self._tangent_frame(x, v, A, grid, feature_process)
        if frame is not None:
            # surface clouds: the tangent-frame projection of `feature_process_tangent` runs INSIDE the perception sweep (weights
            # projected per pair), the MLP / update / life mask on the fused kernels -- no operator-by-operator step, no library GEMM
            update_mask = torch.rand(x.shape[0], device=x.device) <= fire_rate
            lin1, lin2 = self.model[0], self.model[2]
            nA = sphops.fused.nca_step(
                x, v, A, update_mask, lin1.weight, lin1.bias, lin2.weight, lin2.bias, h, grid,
                update_rule=self.update_rule, use_alpha=self.use_alpha,
                perception_scale=self.perception_scale(h), orig_rate=self.fire_rate / fire_rate, frame=frame)
            return x, nA
        if self._fusable(x, A, feature_process):
            update_mask = torch.rand(x.shape[0], device=x.device) <= fire_rate
            lin1, lin2 = self.model[0], self.model[2]
            nA = sphops.fused.nca_step(
                x, v, A, update_mask, lin1.weight, lin1.bias, lin2.weight, lin2.bias, h, grid,
                update_rule=self.update_rule, use_alpha=self.use_alpha,
                perception_scale=self.perception_scale(h), orig_rate=self.fire_rate / fire_rate)
            return x, nA

        prev_mask = self.life_mask(x, v, cell_activity(A, self.use_alpha)[:, None], h, grid)
        features = self.perceive(x, v, A, h, grid)
        if feature_process is not None:
            features = feature_process(features)
        dA = self.model(self.base_feature_process(features))

        update_mask = torch.rand(x.shape[0], device=x.device) <= fire_rate
        n = self.in_features
        if self.update_rule == 'gated':
            nA = A * torch.sigmoid(dA[..., :n]) + torch.tanh(dA[..., n:-1]) * torch.sigmoid(dA[..., [-1]])
        else:
            nA = A + dA * self.fire_rate / fire_rate
        nA = torch.where(update_mask[..., None], nA, A)

        new_mask = self.life_mask(x, v, cell_activity(nA, self.use_alpha)[..., None], h, grid)
        return x, nA * (prev_mask & new_mask).float()

    def rollout(self, x, v, A, h, grid, steps, fire_rate=None, keep_all=True, checkpoint=None, microbatch=None, fire_index=None):
        """`steps` chained `forward` calls as one autograd node with the state kept cell-sorted on the
        device (sphops.fused.nca_rollout); returns the sequence of the steps+1 states, or None when the
        configuration is not the fused one (callers then loop over `forward`).  `checkpoint`: True = keep only
        the states + fire bytes and recompute each step in the backward, False = full tape, None = the
        sphops.fused.CHECKPOINT default ("auto": by free device memory).  `microbatch`: sections per group when the
        batch is run group by group (sphops.fused.plan_microbatches), None = sphops.fused.MICROBATCH ("auto")."""
        if not (USE_FUSED_ROLLOUT and self._fusable(x, A, None)):
            return None
        if fire_rate is None:
            fire_rate = self.fire_rate
        lin1, lin2 = self.model[0], self.model[2]
        return sphops.fused.nca_rollout(
            x, v, A, lin1.weight, lin1.bias, lin2.weight, lin2.bias, h, grid, steps, fire_rate,
            update_rule=self.update_rule, use_alpha=self.use_alpha,
            perception_scale=self.perception_scale(h), orig_rate=self.fire_rate / fire_rate, keep_all=keep_all,
            checkpoint=checkpoint, microbatch=microbatch, fire_index=fire_index)

    def normalize_grads(self):
        if not self.normalize_grad:
            return
        for p in self.parameters():
            if p.grad is not None:
                p.grad = p.grad / (p.grad.norm() + 1e-8)


def _draw_steps(iter_n, default_range):
    if iter_n is None:
        lo, hi = default_range
        return torch.randint(lo, hi, (1,)).item()
    if isinstance(iter_n, (tuple, list)):
        lo, hi = iter_n
        return torch.randint(lo, hi, (1,)).item()
    return iter_n


def _train_on(loss_fn, x, A, img, out, model, optimizer, scheduler):
    """Loss on the final state + 0.1 x loss on 4 random recorded states, then one optimiser step
    (reference nca.py:147-158)."""
    loss = loss_fn(x, A, img, A0=out[0])
    for nA in random.choices(out, k=4):
        loss += 0.1 * loss_fn(x, nA, img)
    loss.backward()
    model.normalize_grads()
    optimizer.step()
    optimizer.zero_grad()
    if scheduler:
        scheduler.step()
    return loss.item()


def sample_plane(model, x, A, sections, h, dims, out_steps=False, iter_n=None, loss_fn=None, img=None,
                 optimizer=None, scheduler=None, fire_rate=None):
    """K-step rollout on a (possibly zero-padded 2-D) particle plane; reference nca.py:124-164."""
    iter_n = _draw_steps(iter_n, (32, 48))
    xdim = x.shape[-1]
    if USE_3D and xdim == 2:
        x = F.pad(x, (0, 1))

    grid, v = model.init_grid(x, sections, h, dims)
    out = model.rollout(x, v, A, h, grid, iter_n, fire_rate, keep_all=out_steps or loss_fn is not None) \
        if hasattr(model, "rollout") else None
    if out is not None:
        A = out[-1]
    else:
        out = []
        for _ in range(iter_n):
            out.append(A.clone())
            x, A = model(x, v, A, h, grid, fire_rate=fire_rate, feature_process=None)
        out.append(A.clone())

    loss = 0.0
    if loss_fn is not None:
        if xdim == 2:
            x = x[..., :2]
        loss = _train_on(loss_fn, x, A, img, out, model, optimizer, scheduler)
    if USE_3D and xdim == 2:
        x = x[..., :2]
    if out_steps:
        return loss, x, out
    return loss, x, A


# =====================================================================================================
# Periodic ("wrapped") planes -- reference code/nca.py:167-299
# =====================================================================================================
def get_wrapped_points_index(x, A, sections, gmin, gsize, h):
    """Per section, the 9 ghost index lists of the reference (None at slot 4 = the plane itself): particles
    within h of the left/right (x) and bottom/top (y) edges and the four corner intersections, each list in
    ascending particle order.  Device-side (the reference round-trips through numpy)."""
    gmin = torch.as_tensor(gmin, dtype=x.dtype, device=x.device)
    gmax = gmin + torch.as_tensor(gsize, dtype=x.dtype, device=x.device)
    out, off = [], 0
    for n in sections:
        xs = x[off:off + n]
        lo_x, hi_x = xs[:, 0] <= gmin[0] + h, xs[:, 0] >= gmax[0] - h
        lo_y, hi_y = xs[:, 1] <= gmin[1] + h, xs[:, 1] >= gmax[1] - h
        pick = lambda m: torch.nonzero(m).flatten()
        out.append([pick(lo_x & lo_y), pick(lo_x), pick(lo_x & hi_y), pick(lo_y), None, pick(hi_y),
                    pick(hi_x & lo_y), pick(hi_x), pick(hi_x & hi_y)])
        off += n
    return out


class _WrapPlan:
    """One gather that performs the reference's wrap (and unwrap+wrap) for a whole packed batch.

    wrapped row r of the packed batch = original row src[r], positions shifted by shift[r];
    own[i] = wrapped row of original particle i (the reference's `unwrap_tensors` keeps exactly these)."""

    def __init__(self, sections, idxs, gsize, device):
        src, shift, own, new_sections = [], [], [], []
        off, woff = 0, 0
        gs = torch.as_tensor(gsize, dtype=torch.float32, device=device)
        for n, idx in zip(sections, idxs):
            parts = [torch.arange(n, device=device)]
            shifts = [torch.zeros((n, 2), device=device)]
            for k in range(9):
                if idx[k] is None:
                    continue
                sx, sy = 1 - (k // 3), 1 - (k % 3)        # ghost offset signs (reference nca.py:214-217)
                parts.append(idx[k])
                shifts.append(torch.tensor([sx, sy], dtype=torch.float32, device=device).expand(idx[k].numel(), 2) * gs[:2])
            p = torch.cat(parts)
            src.append(p + off)
            shift.append(torch.cat(shifts))
            own.append(torch.arange(n, device=device) + woff)
            new_sections.append(int(p.numel()))
            off += n
            woff += int(p.numel())
        self.src, self.shift, self.own = torch.cat(src), torch.cat(shift), torch.cat(own)
        self.sections = new_sections
        self.rewrap = self.own[self.src]                   # unwrap followed by wrap, as one gather

    def wrap_positions(self, x):
        xw = x[self.src].clone()
        xw[:, :2] += self.shift.to(x.dtype)
        return xw


def wrap_tensors(tensors, sections, idxs, gsize, x=None):
    """API of reference nca.py:200-227: ghost-pad every tensor (and shift the ghost positions of `x`)."""
    dev = tensors[0].device if tensors else x.device
    plan = _WrapPlan(sections, idxs, gsize, dev)
    new = [t[plan.src] for t in tensors]
    if x is not None:
        return new, plan.sections, plan.wrap_positions(x)
    return new, plan.sections


def unwrap_tensors(tensors, sections, orig_sections):
    """API of reference nca.py:229-242: keep the first `orig_sections[b]` rows of every wrapped section."""
    keep, off = [], 0
    for n_w, n in zip(sections, orig_sections):
        keep.append(torch.arange(n) + off)
        off += n_w
    keep = torch.cat(keep)
    return [t[keep.to(t.device)] for t in tensors]


class _GatheredStates:
    """The reference's `out_wrapped` list (ghost-padded state after every step) as a lazy view: wrapped row r = own row src[r]."""

    def __init__(self, states, src):
        self._s, self._src = states, src

    def __len__(self):
        return len(self._s)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self._s)))]
        return self._s[k][self._src]

    def __iter__(self):
        return (self[k] for k in range(len(self._s)))


def _native_periodic_rollout(model, x, A, sections, h, dims, gmin, gsize, plan, iter_n, fire_rate, keep_all):
    """The K steps of a periodic plane on the N real particles (periodic hash grid, sphops Hashgrid.set_periodic), or None
    when this configuration needs the ghost particles: an alpha channel (the reference evolves ghosts on their own within a
    step and reads THEIR new alpha in the life mask), a plane whose size is not dims x h (the cell hash then does not wrap
    with the plane), or a step the fused node does not take."""
    if not (USE_NATIVE_PERIODIC and USE_FUSED_ROLLOUT and hasattr(model, "rollout") and not model.use_alpha):
        return None
    if not model._fusable(x, A, None):
        return None
    from sphops.hashgrid import Hashgrid
    from sphops.preprocess import _host_dims
    hval = h.item() if isinstance(h, torch.Tensor) else float(h)
    gs = [float(g) for g in torch.as_tensor(gsize).flatten().tolist()][:2]
    period = (gs[0], gs[1], 0.0)
    dims_host = _host_dims(dims, 3)
    if not Hashgrid.periodic_compatible(hval, dims_host, period):
        return None
    # which particles have ghost copies: the very comparisons of get_wrapped_points_index (reference nca.py:182-197)
    gmin_t = torch.as_tensor(gmin, dtype=x.dtype, device=x.device).flatten()[:2]
    gmax_t = gmin_t + torch.as_tensor(gsize, dtype=x.dtype, device=x.device).flatten()[:2]
    band = ((gmin_t + hval).tolist(), (gmax_t - hval).tolist())
    grid = sphops.initialize_hashgrid(x, sections, h, dims).set_periodic(period, band=band)
    v = sphops.volume(x, h, grid)
    # the reference draws torch.rand over the WRAPPED particles every step (nca.py:100 on x_w); the real ones come first in
    # every wrapped section (`plan.own`)
    n_wrapped = int(plan.src.numel())
    return model.rollout(x, v, A, h, grid, iter_n, fire_rate, keep_all=keep_all, fire_index=(n_wrapped, plan.own))


def sample_plane_wrapped(model, x, A, sections, h, dims, gmin, gsize, out_steps=False, iter_n=None, loss_fn=None,
                         img=None, optimizer=None, scheduler=None, out_wrapped=False, fire_rate=None):
    """Rollout on a periodic plane (reference nca.py:245-299): ghost copies of the particles within h of an
    edge are appended, the grid/volume are built on the wrapped set once, and after every step the ghosts are
    refreshed from their sources.  Here wrap / unwrap+wrap are single device-side gathers."""
    iter_n = _draw_steps(iter_n, (32, 48))
    xdim = x.shape[-1]
    if USE_3D and xdim == 2:
        x = F.pad(x, (0, 1))
    idxs = get_wrapped_points_index(x, A, sections, gmin, gsize, h)
    plan = _WrapPlan(sections, idxs, gsize, x.device)
    x_w = plan.wrap_positions(x)

    native = _native_periodic_rollout(model, x, A, sections, h, dims, gmin, gsize, plan, iter_n, fire_rate,
                                      keep_all=out_steps or loss_fn is not None)
    if native is not None:
        # Periodic neighbour search instead of ghost particles: N particles, ONE autograd node, fire lists, no per-step
        # gathers.  Own-particle results are the reference's (same pairs at the same fp32 offsets; the wrapped states are
        # the own states gathered, which is what the reference's per-step ghost refresh makes them).
        out = native
        A = out[-1]
        loss = 0.0
        if loss_fn is not None:
            if xdim == 2:
                x = x[..., :2]
            loss = _train_on(loss_fn, x, A, img, out, model, optimizer, scheduler)
        if USE_3D and xdim == 2:
            x = x[..., :2]
        if out_steps:
            return loss, x, out, x_w, _GatheredStates(out, plan.src)
        return loss, x, A

    A_w = A[plan.src]
    grid, v_w = model.init_grid(x_w, plan.sections, h, dims)

    out, out_w = [], []
    for _ in range(iter_n):
        out.append(A.clone())
        out_w.append(A_w.clone())
        x_w, A_w = model(x_w, v_w, A_w, h, grid, fire_rate=fire_rate, feature_process=None)
        A = A_w[plan.own]
        A_w = A_w[plan.rewrap]
    out.append(A.clone())
    out_w.append(A_w.clone())

    loss = 0.0
    if loss_fn is not None:
        if xdim == 2:
            x = x[..., :2]
        loss = _train_on(loss_fn, x, A, img, out, model, optimizer, scheduler)
    if USE_3D and xdim == 2:
        x = x[..., :2]
    if out_steps:
        if out_w:       # the reference's `out_wrapped` flag is shadowed by its list of the same name (always truthy)
            return loss, x, out, x_w, out_w
        return loss, x, out
    return loss, x, A


# =====================================================================================================
# Surface meshes: tangent frames -- reference code/nca.py:303-398
# =====================================================================================================
def normalize(v):
    return v / (1e-8 + torch.linalg.norm(v, dim=-1, keepdims=True))


def orthogonalize(N, T):
    """Remove the component of T along N and renormalise."""
    return normalize(T - N * (N * T).sum(-1, keepdims=True))


_DIFFUSE_GRIDS = {}


def diffuse(N, T, x, A, sections, h, dims, v=None, grid=None, lerp_multiplier=1.0, w_multiplier=1.0):
    """Spread the tangent field over the surface with an activity-weighted SPH blur (reference nca.py:312-323).
    The reference rebuilds the hash grid and volumes on every call; positions do not change during a rollout,
    so they are cached per position tensor (same data pointer / version / h / dims)."""
    if v is None or grid is None:
        key = (x.data_ptr(), x._version, tuple(x.shape), float(h), str(dims), tuple(sections))
        hit = _DIFFUSE_GRIDS.get("last")
        if hit is None or hit[0] != key:
            g = sphops.initialize_hashgrid(x, sections, h, dims)
            hit = (key, g, sphops.volume(x, h, g), x)       # keeping x alive pins its address
            _DIFFUSE_GRIDS["last"] = hit
        grid, v = hit[1], hit[2]
    w = torch.clip(cell_activity(A)[..., None], 0, 1)
    m = torch.lerp(torch.ones_like(w), w, w_multiplier)
    blurred = sphops.blur(x, v, torch.cat([m, m * T], dim=-1), h, grid)
    T2 = blurred[..., 1:] / (1e-8 + blurred[..., [0]])
    return orthogonalize(N, torch.lerp(T2, T, w * lerp_multiplier))


def project_tangent_space(Av, N, T):
    """Av [n,C,3] expressed in the local frame (T, N x T, N)."""
    frame = torch.stack([T, torch.linalg.cross(N, T), N], dim=-1)
    return torch.matmul(Av, frame)


def feature_process_tangent(N, T):
    def func(features):
        return [(As, project_tangent_space(Av, N, T)) for As, Av in features]
    func.tangent_frame = (N, T)          # lets SPHNCA.forward run the projection inside the perception sweep
    return func


def sample_mesh(model, x, A, N, T, sections, h, dims, out_steps=False, iter_n=None, loss_fn=None, img=None,
                optimizer=None, scheduler=None, lerp_multiplier=1.0, w_multiplier=1.0, out_np=False, iter_wrapper=None,
                fire_rate=None):
    """Rollout on a surface point cloud (reference nca.py:338-381): perception is projected onto the tangent
    frame, and the tangent field T is diffused (in place) after every step."""
    iter_n = _draw_steps(iter_n, (32, 48))
    xdim = x.shape[-1]
    if USE_3D and xdim == 2:
        x = F.pad(x, (0, 1))
    grid, v = model.init_grid(x, sections, h, dims)
    steps = range(iter_n)
    if iter_wrapper is not None:
        steps = iter_wrapper(steps)
    out = []
    for _ in steps:
        out.append(A.clone())
        x, A = model(x, v, A, h, grid, fire_rate=fire_rate, feature_process=feature_process_tangent(N, T))
        T[:, :] = diffuse(N, T, x, A, sections, 0.1, 20, lerp_multiplier=lerp_multiplier, w_multiplier=w_multiplier).detach()
    out.append(A.clone())

    loss = 0.0
    if loss_fn is not None:
        if xdim == 2:
            x = x[..., :2]
        loss = _train_on(loss_fn, x, A, img, out, model, optimizer, scheduler)
    if out_np:
        out = [a.cpu().numpy() for a in out]
    if out_steps:
        return loss, x, out
    return loss, x, A


def add_radial_seed(x3, A3, N=None, T=None, R=0.2, idx=None, center=None):
    """Add a smooth bump (1 - d^2/R^2)^3 around a particle or a given centre to every channel, in place
    (reference nca.py:383-398); optionally seed the tangent of the closest particle."""
    if idx is None:
        import numpy as np
        idx = int(np.random.randint(0, x3.shape[0]))           # the reference's draw (np.random, seeded by train.py:252)
    c = x3[idx, :] if center is None else center
    d2 = ((x3 - c) ** 2).sum(dim=-1)
    A3 += (torch.clamp(1 - d2 / R ** 2, 0, 1) ** 3)[..., None]
    if N is not None:
        i = torch.argmin(d2)
        T[i, :] = orthogonalize(N[i, :], torch.randn((x3.shape[-1],), device=T.device))


def BNC2BCHW(A, H, W, B=None):
    """[B*N, C] (with B given) or [B, N, C] particle rows of a lattice, particle n = w*H + h (the order of the reference's
    `geometry.grange`) -> image batch [B, C, H, W] (reference nca.py:555-561; used by the image-space losses of losses.py)."""
    C_ = A.shape[-1]
    if B is not None:
        A = A.reshape(B, -1, C_)
    return A.reshape(A.shape[0], W, H, C_).permute(0, 3, 2, 1)


# Downstream of the per-timestep path (SURVEY.md 2 rows 9-19, 8 "out of scope"): the VGG / optimal-transport / CLIP image losses
# of reference nca.py:470-715.  `train.py --loss mse_simple` needs none of them; `--loss ot|clip_multiscale` needs the
# reference's own implementations (pretrained VGG19 / open_clip weights), which this module does not re-implement.
_NOT_ON_THE_PATH = ("get_style_loss", "eval_style_loss", "OptimalTransportLoss", "StyleLoss", "Normalization", "gram_matrix",
                    "get_clip_loss")


def __getattr__(name):
    if name in _NOT_ON_THE_PATH:
        raise AttributeError(
            f"nca.{name} is an image-space loss helper of the reference (code/nca.py:470-715), downstream of the particle update "
            "this package replaces; it is not re-implemented here.  Take it from the reference's nca.py (e.g. "
            "`import importlib.util` on that file) -- `train.py --loss mse_simple` does not need it.")
    raise AttributeError(f"module 'nca' has no attribute {name!r}")


def get_rgba(A, margin=None):
    rgba = A[..., :4].clone()
    if margin is not None:
        rgba = rgba + (torch.clamp(rgba, -margin, 1 + margin) - rgba).detach()
    return rgba


class Pool:
    """Sample pool of NCA states (reference nca.py:400-460): `sample` draws a batch (optionally sorted by loss,
    worst one re-seeded), `update` writes the evolved states back.  By default the states stay on the CPU like the
    reference; `device=` keeps the pool RESIDENT on that device: `sample` / `update` then move no
    state over PCIe (a 1024 x 65,536 x 16 pool is 4.3 GB of the 180 GB), the draws (`randperm`, re-seed, damage) are the
    same calls in the same order."""

    def __init__(self, seed, total_size, randomized_feat=False, feature_process=None, device=None):
        self.total_size, self.seed = total_size, seed
        self.num_points = seed[0].shape[0]
        self.randomized_feat, self.feature_process = randomized_feat, feature_process
        self.device = torch.device("cpu") if device is None else torch.device(device)
        self.x = seed[0].to(self.device).repeat((total_size, 1, 1))
        self.A = seed[1].to(self.device).repeat((total_size, 1, 1))
        self.num_features = self.A.shape[-1]
        if feature_process:
            x, A = self.get_initial_feature()
            self.x[:] = x.to(self.device)
            self.A[:] = A.to(self.device)
        else:
            for i in range(total_size):
                x, A = self.get_initial_feature()
                self.x[i], self.A[i] = x.to(self.device), A.to(self.device)

    def get_initial_feature(self):
        x, A = self.seed
        if self.randomized_feat:
            A = torch.rand_like(A)
        elif self.feature_process:
            x, A = self.feature_process(x, A)
        return x, A

    def sample(self, batch_size, device, replace_worst=False, loss_fn=None, img=None, degrade_prob=0.0, erase_R=0.0):
        idx = torch.randperm(self.total_size)[:batch_size]
        sel = idx.to(self.device)
        x, A = self.x[sel].to(device), self.A[sel].to(device)
        if replace_worst:
            import numpy as np
            order = torch.from_numpy(np.array(loss_fn(x, A, img)).argsort()[::-1].copy())
            x, A, idx = x[order], A[order], idx[order]
            x[0], A[0] = self.get_initial_feature()
        if degrade_prob > 0.0:
            hit = torch.rand_like(A[..., 0]) < degrade_prob
            A[hit, :] = torch.rand_like(A[hit, :])
        if erase_R > 0.0:
            for b in range(x.shape[0]):
                c = x[b, int(torch.randint(0, x.shape[1], (1,)))]
                A[b, ((x[b] - c) ** 2).sum(-1) < erase_R ** 2, :] = 0.0
        return x.view(-1, 2), A.view(-1, self.num_features), (self.num_points,) * batch_size, idx

    def update(self, x, A, idx):
        b = idx.shape[0]
        sel = idx.to(self.device)
        self.x[sel] = x.view(b, -1, 2).detach().to(self.device)
        self.A[sel] = A.view(b, -1, self.num_features).detach().to(self.device)
