"""Rollout drivers of SURVEY.md 8f (sample_plane_wrapped, sample_mesh/diffuse, Pool) against golden vectors
produced by the unmodified reference (tests/golden/make_golden_drivers.py)."""
import numpy as np
import pytest
import torch

from tests.util import assert_close, forced_fire_masks, load_golden

DEV = "cuda"


def test_wrap_plan_matches_reference_ghosts():
    """CPU: ghost selection, order and position shifts equal the reference's wrap_tensors output bit for bit."""
    import nca
    g = load_golden("wrapped_zebra_k3")
    x = torch.nn.functional.pad(torch.from_numpy(g["x"]), (0, 1))
    sections = tuple(int(s) for s in g["sections"])
    gmin, gsize = torch.tensor([-1.0, -1.0]), torch.tensor([2.0, 2.0])
    idxs = nca.get_wrapped_points_index(x, None, sections, gmin, gsize, float(g["h"]))
    [Aw], new_sections, xw = nca.wrap_tensors([torch.arange(x.shape[0], dtype=torch.float32)[:, None]], sections, idxs, gsize, x=x)
    assert np.array_equal(xw.numpy(), g["x_wrapped"])
    assert sum(new_sections) == g["x_wrapped"].shape[0] and len(new_sections) == 2
    [back] = nca.unwrap_tensors([Aw], new_sections, sections)
    assert torch.equal(back[:, 0], torch.arange(x.shape[0], dtype=torch.float32))


def test_pool_api():
    import nca
    x = torch.rand(50, 2); A = torch.zeros(50, 16)
    pool = nca.Pool((x, A), 8, randomized_feat=True)
    xs, As, sections, idx = pool.sample(4, "cpu", replace_worst=True, loss_fn=lambda x, A, img: [float(a.sum()) for a in A])
    assert xs.shape == (200, 2) and As.shape == (200, 16) and sections == (50,) * 4 and idx.shape == (4,)
    pool.update(xs, As + 1, idx)
    assert torch.allclose(pool.A[idx[1]], As.view(4, 50, 16)[1] + 1)


def _pool_run(device):
    import nca
    torch.manual_seed(7)
    x = torch.rand(40, 2); A = torch.rand(40, 16)
    pool = nca.Pool((x, A), 6, randomized_feat=False, device=device)
    out = []
    for it in range(3):
        xs, As, sections, idx = pool.sample(3, pool.device if device else "cpu", replace_worst=True,
                                            loss_fn=lambda x, A, img: [float(a.sum()) for a in A], degrade_prob=0.0)
        pool.update(xs, As * 0.5 + it, idx)
        out.append((xs.cpu(), As.cpu(), idx.clone()))
    return out, pool.A.cpu()


def test_pool_resident_on_a_device_equals_cpu_pool():
    """`device=` keeps the pool on that device; the samples, the order and the stored states equal the CPU pool's."""
    a, Aa = _pool_run(None)
    b, Ab = _pool_run("cpu")
    for (x0, A0, i0), (x1, A1, i1) in zip(a, b):
        assert torch.equal(x0, x1) and torch.equal(A0, A1) and torch.equal(i0, i1)
    assert torch.equal(Aa, Ab)


@pytest.mark.gpu
def test_pool_resident_on_the_gpu_equals_cpu_pool():
    a, Aa = _pool_run(None)
    b, Ab = _pool_run(DEV)
    for (x0, A0, i0), (x1, A1, i1) in zip(a, b):
        assert torch.equal(x0, x1) and torch.equal(A0, A1) and torch.equal(i0, i1)
    assert torch.equal(Aa, Ab)


def _model(g, **kw):
    import nca
    m = nca.SPHNCA(16, 0.5, **kw).to(DEV)
    with torch.no_grad():
        m.model[0].weight.copy_(torch.from_numpy(g["W1"])); m.model[0].bias.copy_(torch.from_numpy(g["b1"]))
        m.model[2].weight.copy_(torch.from_numpy(g["W2"])); m.model[2].bias.copy_(torch.from_numpy(g["b2"]))
    return m


@pytest.mark.gpu
def test_sample_plane_wrapped_matches_reference():
    import nca
    g = load_golden("wrapped_zebra_k3")
    nca.USE_3D, nca.USE_FUSED_STEP = True, True
    m = _model(g, update_rule="gated", use_alpha=False, normalize_perception=float(g["normalize_perception"]))
    x = torch.from_numpy(g["x"]).to(DEV)
    gmin, gsize = torch.tensor([-1.0, -1.0], device=DEV), torch.tensor([2.0, 2.0], device=DEV)
    with torch.no_grad(), forced_fire_masks([torch.from_numpy(mk) for mk in g["masks"]]):
        loss, x_out, out, x_w, out_w = nca.sample_plane_wrapped(m, x, torch.from_numpy(g["A0"]).to(DEV), tuple(g["sections"]),
                                                                float(g["h"]), int(g["dims"]), gmin, gsize, out_steps=True, iter_n=3)
    assert x_out.shape == x.shape and np.array_equal(x_w.cpu().numpy(), g["x_wrapped"])
    for k in range(4):
        assert_close(out[k], g["states"][k], what=f"state {k}")
    assert_close(out_w[-1], g["states_wrapped_last"], what="wrapped final state")


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False], ids=["frame_sweep", "operators"])
def test_sample_mesh_matches_reference(fused, monkeypatch):
    """Surface rollout (reference nca.py:338-381) against the reference golden: with the tangent-frame projection inside the
    perception sweep + fused MLP / update / mask kernels (default for inference), and operator by operator."""
    import nca
    from sphops import fused as F
    g = load_golden("mesh_gecko_k2")
    nca.USE_3D = True
    calls = []
    orig = F.nca_step
    monkeypatch.setattr(F, "nca_step", lambda *a, **k: calls.append(k.get("frame") is not None) or orig(*a, **k))
    if not fused:
        monkeypatch.setattr(F, "frame_step_available", lambda *a: False)
    m = _model(g, update_rule="gated", use_alpha=True, normalize_perception=float(g["normalize_perception"]))
    x = torch.from_numpy(g["x"]).to(DEV)
    N = torch.from_numpy(g["N"]).to(DEV)
    T = torch.from_numpy(g["T0"]).to(DEV).clone()
    n = x.shape[0]
    with torch.no_grad(), forced_fire_masks([torch.from_numpy(mk) for mk in g["masks"]]):
        loss, x_out, out = nca.sample_mesh(m, x, torch.from_numpy(g["A0"]).to(DEV), N, T, (n,), float(g["h"]), int(g["dims"]),
                                           out_steps=True, iter_n=2)
    assert calls == ([True, True] if fused else []), calls
    for k in range(3):
        assert_close(out[k], g["states"][k], what=f"state {k}")
    assert_close(T, g["T_final"], rtol=2e-5, what="diffused tangents")


@pytest.mark.gpu
@pytest.mark.parametrize("native", [True, False], ids=["periodic_grid", "ghost_particles"])
def test_sample_plane_wrapped_periodic_matches_reference_with_gradients(native, monkeypatch):
    """The periodic plane as train.py / test.py set it up (dims = ceil(2 / h), so dims x h is the plane size), forward states AND
    the gradients through the K steps, against the unmodified reference (tests/golden/make_golden_drivers.py::
    wrapped_periodic_case): the NATIVE periodic path (periodic hash grid: N particles, one rollout node with fire lists, no
    ghosts, no per-step gathers) and the reference-shaped ghost path."""
    import nca
    g = load_golden("wrapped_periodic_k3")
    nca.USE_3D, nca.USE_FUSED_STEP = True, True
    monkeypatch.setattr(nca, "USE_NATIVE_PERIODIC", native)
    taken = []
    orig = nca._native_periodic_rollout
    monkeypatch.setattr(nca, "_native_periodic_rollout", lambda *a, **k: taken.append(orig(*a, **k)) or taken[-1])
    m = _model(g, update_rule="gated", use_alpha=False, normalize_perception=float(g["normalize_perception"]))
    x = torch.from_numpy(g["x"]).to(DEV)
    A0 = torch.from_numpy(g["A0"]).to(DEV).requires_grad_(True)
    gmin, gsize = torch.tensor([-1.0, -1.0], device=DEV), torch.tensor([2.0, 2.0], device=DEV)
    with forced_fire_masks([torch.from_numpy(mk) for mk in g["masks"]]):
        loss, x_out, out, x_w, out_w = nca.sample_plane_wrapped(m, x, A0, tuple(g["sections"]), float(g["h"]), int(g["dims"]), gmin, gsize,
                                                                out_steps=True, iter_n=3)
    assert (taken[0] is not None) == native
    assert np.array_equal(x_w.cpu().numpy(), g["x_wrapped"])
    for k in range(4):
        assert_close(out[k], g["states"][k], what=f"state {k}")
    assert_close(out_w[-1], g["states_wrapped_last"], what="wrapped final state")
    L = (out[-1][:, :4] ** 2).mean() + 0.1 * (out[1][:, :4] ** 2).mean()
    L.backward()
    assert abs(float(L) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert_close(A0.grad, g["dA0"], rtol=2e-5, what="dL/dA0")
    P = dict(m.named_parameters())
    for key, pn in (("dW1", "model.0.weight"), ("db1", "model.0.bias"), ("dW2", "model.2.weight"), ("db2", "model.2.bias")):
        assert_close(P[pn].grad, g[key], rtol=2e-5, what=key)


@pytest.mark.gpu
def test_periodic_grid_is_refused_where_it_is_not_implemented():
    """A plane that is not dims x h wide keeps the ghost particles; operators without a periodic neighbour search refuse a
    periodic grid instead of silently ignoring the wrap."""
    import nca
    import sphops
    from sphops._lib import SphncaError
    g = load_golden("wrapped_zebra_k3")                     # h = 0.4, dims = 7: 2.8 != 2
    m = _model(g, update_rule="gated", use_alpha=False, normalize_perception=2.5)
    x = torch.nn.functional.pad(torch.from_numpy(g["x"]), (0, 1)).to(DEV)
    sections = tuple(int(s) for s in g["sections"])
    plan = nca._WrapPlan(sections, nca.get_wrapped_points_index(x, None, sections, torch.tensor([-1.0, -1.0]), torch.tensor([2.0, 2.0]), 0.4),
                         torch.tensor([2.0, 2.0]), x.device)
    assert nca._native_periodic_rollout(m, x, torch.from_numpy(g["A0"]).to(DEV), sections, 0.4, 7, torch.tensor([-1.0, -1.0]),
                                        torch.tensor([2.0, 2.0]), plan, 2, None, True) is None
    grid = sphops.initialize_hashgrid(x, sections, 0.4, 5)
    with pytest.raises(ValueError):
        grid.set_periodic((2.0, 2.0, 2.0))                  # z: dims 5 x 0.4 == 2 is fine, but (2.8, ...) is not
        grid.set_periodic((2.8, 2.0, 0.0))
    grid.set_periodic((2.0, 2.0, 0.0))
    v = sphops.volume(x, 0.4, grid)                          # implemented
    assert torch.isfinite(v).all()
    with pytest.raises(SphncaError):
        sphops.count(x, 0.4, grid)                           # not implemented for periodic grids


@pytest.mark.gpu
def test_periodic_rollout_microbatched_over_sections_equals_whole_batch():
    """A periodic grid cut into micro-batches of sections (Hashgrid.subgrid carries the period, the ghost-copy band and its slice of
    the ghost volumes): states bit-identical to the whole-batch periodic rollout."""
    import nca
    import sphops
    g = load_golden("wrapped_periodic_k3")
    m = _model(g, update_rule="gated", use_alpha=False, normalize_perception=float(g["normalize_perception"]))
    x = torch.nn.functional.pad(torch.from_numpy(g["x"]), (0, 1)).to(DEV)
    A0 = torch.from_numpy(g["A0"]).to(DEV)
    sections = tuple(int(s) for s in g["sections"])
    h = float(g["h"])
    band = ((-1.0 + h, -1.0 + h), (1.0 - h, 1.0 - h))
    outs = []
    for mb in (None, 1):
        grid = sphops.initialize_hashgrid(x, sections, h, int(g["dims"])).set_periodic((2.0, 2.0, 0.0), band=band)
        v = sphops.volume(x, h, grid)
        torch.manual_seed(3)
        with torch.no_grad():
            out = m.rollout(x, v, A0, h, grid, 3, microbatch=mb)
        outs.append([out[k].clone() for k in range(4)])
    for a, b in zip(*outs):
        assert torch.equal(a, b)
