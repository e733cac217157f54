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
def test_sample_mesh_matches_reference():
    import nca
    g = load_golden("mesh_gecko_k2")
    nca.USE_3D = True
    m = _model(g, update_rule="gated", use_alpha=True, normalize_perception=float(g["normalize_perception"]))
    x = torch.from_numpy(g["x"]).to(DEV)
    N = torch.from_numpy(g["N"]).to(DEV)
    T = torch.from_numpy(g["T0"]).to(DEV).clone()
    n = x.shape[0]
    with torch.no_grad(), forced_fire_masks([torch.from_numpy(mk) for mk in g["masks"]]):
        loss, x_out, out = nca.sample_mesh(m, x, torch.from_numpy(g["A0"]).to(DEV), N, T, (n,), float(g["h"]), int(g["dims"]),
                                           out_steps=True, iter_n=2)
    for k in range(3):
        assert_close(out[k], g["states"][k], what=f"state {k}")
    assert_close(T, g["T_final"], rtol=2e-5, what="diffused tangents")
