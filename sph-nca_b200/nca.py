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

    def forward(self, x, v, A, h, grid, fire_rate=None, feature_process=None):
        if fire_rate is None:
            fire_rate = self.fire_rate
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
