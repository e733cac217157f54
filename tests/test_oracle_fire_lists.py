"""CPU check, on the oracle (the restatement of reference code/nca.py:87-117), of the two facts the fire-list path of the
product relies on (csrc/firelist.cu): for a particle whose fire draw is off
  (1) the new state is the old one (times the life mask) -- the MLP output is discarded by `torch.where(update_mask, nA, A)`;
  (2) no gradient reaches its MLP output or its perception vector: dL/d(dA) and dL/dy are EXACTLY zero there, so it adds
      nothing to the parameter gradients and its dL/dY row in the adjoint sweep is zero.
The GPU tests then show that the product, which skips those particles, gives the oracle's / the reference's results."""
import math

import numpy as np
import pytest
import torch

from oracle import sphnca_oracle as O
from tests.util import jittered_lattice, load_golden


@pytest.mark.parametrize("use_alpha,rule", [(False, "gated"), (True, "gated"), (False, "orig")])
def test_not_fired_particles_neither_change_nor_receive_gradient(use_alpha, rule):
    gen = torch.Generator().manual_seed(11)
    n = 24
    x = torch.nn.functional.pad(jittered_lattice(n, 2, gen), (0, 1))
    N = x.shape[0]
    h = 3.2 * 2 / n
    dims = math.ceil(2 / h)
    w = load_golden("weights_gecko" if use_alpha else "weights_zebra")
    params = [torch.from_numpy(w[k]).double() for k in ("W1", "b1", "W2", "b2")]
    if rule == "orig":
        params[2] = params[2][:16].clone(); params[3] = params[3][:16].clone()
    params = [p.requires_grad_(True) for p in params]
    if use_alpha:
        A = (torch.clamp(1 - (x ** 2).sum(-1) / 0.5 ** 2, 0, 1) ** 3)[:, None].repeat(1, 16) + 0.05 * torch.rand(N, 16, generator=gen)
    else:
        A = torch.rand(N, 16, generator=gen)
    A = A.double().requires_grad_(True)
    mask = torch.rand(N, generator=gen) <= 0.5
    cfg = O.StepConfig(use_alpha=use_alpha, update_rule=rule, normalize_perception=1.0 / h)
    grid, pairs, v = O.prepare(x, (N,), h, dims)
    out, aux = O.nca_step(pairs, v, A, params, mask, cfg, return_aux=True)
    aux["dA"].retain_grad(); aux["y"].retain_grad()
    ((out[:, :5] ** 2).sum() + out.sum()).backward()
    off = ~mask
    assert int(off.sum()) > N // 4 and int(mask.sum()) > N // 4
    living = (aux["prev_mask"] & aux["new_mask"]).double()
    assert torch.equal(out[off].detach(), (A.detach() * living)[off]), "a not-fired particle keeps its state (times the life mask)"
    assert float(aux["dA"].grad[off].abs().max()) == 0.0, "no gradient reaches the MLP output of a not-fired particle"
    assert float(aux["y"].grad[off].abs().max()) == 0.0, "no gradient reaches the perception vector of a not-fired particle"
    assert float(aux["dA"].grad[mask].abs().max()) > 0.0 and float(aux["y"].grad[mask].abs().max()) > 0.0
    # the direct part of dL/dA of a not-fired particle is the gradient of its new state, masked -- what the adjoint takes from the flags
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in params)
