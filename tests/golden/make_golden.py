"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's Numba-CUDA kernels execute on the CPU under Numba's CUDA simulator through the
harness shim in `ref_shim.py`.  Every array written here is an input to, or an output of, the
reference's own `sphops` / `nca.SPHNCA` code; nothing is produced by this repo's oracle or kernels.
Sizes are small because the simulator costs ~5 ms per particle per operator.
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

sphops, nca = ref_shim.load_reference()

import numpy as np  # noqa: E402
import torch  # noqa: E402

WEIGHTS_DIR = "/root/reference/docs/weights"


def np32(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {name}.npz ({os.path.getsize(path)/1024:.1f} KiB)")


def grid_arrays(grid):
    return dict(
        idx=np32(grid.idx).astype(np.int64), cell_start=np32(grid.cell_start), cell_end=np32(grid.cell_end),
        num_points=np32(grid.num_points), grid_dims=np32(grid.dims).astype(np.int64),
        total_cells=np.int64(grid.total_cells))


def jittered_lattice(n_side, dim, gen, jitter=1.0):
    """(i+0.5)*s - 1 in [-1,1]^dim plus uniform jitter (U-0.5)*jitter*s  (SURVEY.md 8d)."""
    s = 2.0 / n_side
    axes = [(torch.arange(n_side, dtype=torch.float32) + 0.5) * s - 1.0 for _ in range(dim)]
    # x fastest like commons/geometry.py grange order is irrelevant for the kernels
    mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, dim)
    return mesh + (torch.rand(mesh.shape, generator=gen) - 0.5) * jitter * s


def load_weights(name):
    w = json.load(open(os.path.join(WEIGHTS_DIR, name + ".json")))
    L0, L2 = w["layers"]
    return (torch.tensor(L0["weight"], dtype=torch.float32), torch.tensor(L0["bias"], dtype=torch.float32),
            torch.tensor(L2["weight"], dtype=torch.float32), torch.tensor(L2["bias"], dtype=torch.float32))


def ops_case(name, x, sections, h, dims, F_grad=16, F_blur=(1, 4), F_div=3, seed=1):
    """All five operators + the two sound backwards of the reference on one particle set."""
    t0 = time.time()
    gen = torch.Generator().manual_seed(seed)
    N, D = x.shape
    grid = sphops.initialize_hashgrid(x, sections, h, dims)
    out = dict(x=np32(x), sections=np.asarray(sections, np.int64), h=np.float64(h),
               dims=np.asarray(dims if not isinstance(dims, int) else [dims] * D, np.int64), **grid_arrays(grid))

    xg = x.clone().requires_grad_(True)
    v = sphops.volume(xg, h, grid)
    dLdv = torch.rand(N, generator=gen)
    v.backward(dLdv)
    out.update(v=np32(v), dLdv=np32(dLdv), volume_dLdx=np32(xg.grad))
    v = v.detach()

    out["count"] = np32(sphops.count(x, h, grid))

    A = torch.rand(N, F_grad, generator=gen).requires_grad_(True)
    GA = sphops.gradient(x, v, A, h, grid)
    dLdGA = torch.randn(N, F_grad, D, generator=gen)
    GA.backward(dLdGA)
    out.update(A=np32(A), GA=np32(GA), dLdGA=np32(dLdGA), gradient_dLdA=np32(A.grad))

    for F in F_blur:
        Ab = torch.rand(N, F, generator=gen)
        out[f"blur_A{F}"] = np32(Ab)
        out[f"blur_SA{F}"] = np32(sphops.blur(x, v, Ab, h, grid))

    Ad = torch.randn(N, F_div, D, generator=gen)
    out.update(div_A=np32(Ad), div_DA=np32(sphops.divergence(x, v, Ad, h, grid)))
    save(name, **out)
    print(f"  {name}: N={N} D={D} took {time.time()-t0:.1f}s")


def radial_seed(x, F, R):
    """A = (clamp(1 - |x-c|^2/R^2, 0, 1))^3 on every channel, like nca.add_radial_seed (nca.py:383-392)."""
    d2 = (x[:, :2] ** 2).sum(-1)
    w = torch.clamp(1 - d2 / R ** 2, 0, 1) ** 3
    return w[:, None].repeat(1, F).contiguous()


def step_case(name, weights, n_side, h, dims, sections_split, use_alpha, normalize_perception, update_rule,
              fire_rate_call, A_kind, K, seed):
    """K calls of the reference's SPHNCA.forward (nca.py:87-117) + backward of the bench loss."""
    t0 = time.time()
    gen = torch.Generator().manual_seed(seed)
    x2 = jittered_lattice(n_side, 2, gen)
    N = x2.shape[0]
    x = torch.nn.functional.pad(x2, (0, 1))          # USE_3D padding, nca.py:132-134
    sections = sections_split(N)
    assert sum(sections) == N
    model = nca.SPHNCA(16, 0.5, update_rule=update_rule, use_alpha=use_alpha,
                       normalize_perception=normalize_perception)
    if weights is not None:
        W1, b1, W2, b2 = load_weights(weights)
        with torch.no_grad():
            model.model[0].weight.copy_(W1); model.model[0].bias.copy_(b1)
            model.model[2].weight.copy_(W2); model.model[2].bias.copy_(b2)
    else:
        torch.manual_seed(seed)
        with torch.no_grad():
            for p in model.parameters():
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    if A_kind == "rand":
        A0 = torch.rand(N, 16, generator=gen)
    else:
        A0 = radial_seed(x, 16, 0.45) + 0.05 * torch.rand(N, 16, generator=gen)
    A0 = A0.requires_grad_(True)

    grid, v = model.init_grid(x, sections, h, dims)
    fire = model.fire_rate if fire_rate_call is None else fire_rate_call
    masks, states = [], [A0]
    A = A0
    for k in range(K):
        torch.manual_seed(1000 + seed * 10 + k)
        masks.append(torch.rand(N) <= fire)           # what nca.py:100 will draw
        torch.manual_seed(1000 + seed * 10 + k)
        _, A = model(x, v, A, h, grid, fire_rate=fire_rate_call)
        states.append(A)
    loss = (states[-1][:, :4] ** 2).mean()
    if K > 1:
        loss = loss + 0.1 * (states[1][:, :4] ** 2).mean()   # an intermediate state, like nca.py:149-150
    loss.backward()
    P = dict(model.named_parameters())
    save(name,
         x=np32(x), sections=np.asarray(sections, np.int64), h=np.float64(h), dims=np.int64(dims),
         use_alpha=np.bool_(use_alpha), normalize_perception=np.float64(normalize_perception),
         update_rule=np.str_(update_rule), model_fire_rate=np.float64(model.fire_rate),
         fire_rate_call=np.float64(-1.0 if fire_rate_call is None else fire_rate_call),
         W1=np32(P["model.0.weight"]), b1=np32(P["model.0.bias"]),
         W2=np32(P["model.2.weight"]), b2=np32(P["model.2.bias"]),
         v=np32(v), masks=np.stack([np32(m) for m in masks]),
         states=np.stack([np32(s) for s in states]), loss=np.float64(loss.item()),
         dA0=np32(A0.grad), dW1=np32(P["model.0.weight"].grad), db1=np32(P["model.0.bias"].grad),
         dW2=np32(P["model.2.weight"].grad), db2=np32(P["model.2.bias"].grad), **grid_arrays(grid))
    print(f"  {name}: N={N} K={K} took {time.time()-t0:.1f}s")


def sample_plane_case(name, seed):
    """The reference's own rollout driver nca.sample_plane (nca.py:124-164), inference mode."""
    t0 = time.time()
    gen = torch.Generator().manual_seed(seed)
    nca.USE_3D = True
    x2 = jittered_lattice(12, 2, gen)
    N = x2.shape[0]
    sections = (N // 2, N - N // 2)
    h, dims = 0.4, 5
    model = nca.SPHNCA(16, 0.5, update_rule="gated", use_alpha=False, normalize_perception=-1)
    W1, b1, W2, b2 = load_weights("zebra")
    with torch.no_grad():
        model.model[0].weight.copy_(W1); model.model[0].bias.copy_(b1)
        model.model[2].weight.copy_(W2); model.model[2].bias.copy_(b2)
    A0 = torch.rand(N, 16, generator=gen)
    torch.manual_seed(77)
    with torch.no_grad():
        loss, x_out, out = nca.sample_plane(model, x2, A0, sections, h, dims, out_steps=True, iter_n=3)
    save(name, x=np32(x2), sections=np.asarray(sections, np.int64), h=np.float64(h), dims=np.int64(dims),
         W1=np32(W1), b1=np32(b1), W2=np32(W2), b2=np32(b2), A0=np32(A0), torch_seed=np.int64(77),
         x_out=np32(x_out), states=np.stack([np32(s) for s in out]))
    print(f"  {name}: N={N} took {time.time()-t0:.1f}s")


def main():
    torch.set_num_threads(1)
    gen = torch.Generator().manual_seed(0)

    print("ops cases")
    # 3-D uniform cloud, ragged batch, coordinates spilling outside [-1,1] so the modulo wrap is hit
    x = (torch.rand(260, 3, generator=gen) * 2 - 1) * 1.15
    ops_case("ops_3d_ragged", x, (150, 110), 0.45, 5)
    # true 2-D specialisation (USE_3D=False path), non-square dims as a tuple
    x = torch.rand(220, 2, generator=gen) * 2 - 1
    ops_case("ops_2d", x, (220,), 0.3, (7, 8), F_grad=5, F_div=2)
    # zero-padded 2-D lattice as the shipped scripts run it (nca.USE_3D=True)
    x = torch.nn.functional.pad(jittered_lattice(15, 2, gen), (0, 1))
    ops_case("ops_padded2d", x, (100, 125), 0.4, 5)
    # an axis with dims<3: the stencil wraps onto itself and pairs are counted 2x / 3x
    x = torch.cat([torch.rand(160, 2, generator=gen) * 2 - 1, torch.rand(160, 1, generator=gen) * 0.6], dim=-1)
    ops_case("ops_thin_axis", x, (160,), 0.5, (4, 4, 2))
    x = torch.nn.functional.pad(jittered_lattice(12, 2, gen), (0, 1))
    ops_case("ops_single_layer", x, (144,), 0.4, (5, 5, 1), F_grad=4)

    print("step cases")
    step_case("step_zebra_noalpha", "zebra", 15, 0.4, 5, lambda n: (n,), use_alpha=False,
              normalize_perception=-1.0, update_rule="gated", fire_rate_call=None, A_kind="rand", K=1, seed=3)
    step_case("step_gecko_alpha", "gecko", 16, 0.3, 7, lambda n: (100, n - 100), use_alpha=True,
              normalize_perception=1.0 / 0.3, update_rule="gated", fire_rate_call=None, A_kind="radial", K=1, seed=4)
    step_case("step_orig_rule", None, 13, 0.4, 5, lambda n: (n,), use_alpha=True,
              normalize_perception=2.5, update_rule="orig", fire_rate_call=0.8, A_kind="rand", K=1, seed=5)
    step_case("rollout_gecko_alpha_k3", "gecko", 14, 0.35, 6, lambda n: (96, n - 96), use_alpha=True,
              normalize_perception=1.0 / 0.35, update_rule="gated", fire_rate_call=None, A_kind="radial", K=3, seed=6)
    step_case("rollout_zebra_k3", "zebra", 13, 0.4, 5, lambda n: (n,), use_alpha=False,
              normalize_perception=-1.0, update_rule="gated", fire_rate_call=None, A_kind="rand", K=3, seed=7)
    sample_plane_case("sample_plane_zebra", 8)

    # trained MLP weights as a compact fixture for the GPU-side tests/bench (no /root/reference there)
    for nm in ("zebra", "gecko"):
        W1, b1, W2, b2 = load_weights(nm)
        save("weights_" + nm, W1=np32(W1), b1=np32(b1), W2=np32(W2), b2=np32(b2))


if __name__ == "__main__":
    main()
