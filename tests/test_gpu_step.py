"""GPU parity tests of the NCA step and K-step rollouts: both the
operator-by-operator path (general sphops ops + torch MLP, like the reference composes them) and
the fused CUDA step, against the reference's golden vectors and the fp64 oracle.

Bar: states, gradients |new-ref| <= 1e-5*max|ref| (fp32, BASELINE north_star); the two `> 0.1`
life-mask thresholds make a K-step state chaotic, so rows whose mask bit flips are counted and
must be zero on these inputs (fraction reported in the assertion message)."""
import math

import numpy as np
import pytest
import torch

from tests.util import RTOL, assert_close, forced_fire_masks, jittered_lattice, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
STEP_CASES = ["step_zebra_noalpha", "step_gecko_alpha", "step_orig_rule", "rollout_gecko_alpha_k3", "rollout_zebra_k3"]
GRAD_RTOL = 2e-5   # gradients pass through K chained fp32 steps (the reference itself is fp32)


def build_model(g):
    import nca
    m = nca.SPHNCA(16, float(g["model_fire_rate"]), update_rule=str(g["update_rule"]), use_alpha=bool(g["use_alpha"]),
                   normalize_perception=float(g["normalize_perception"])).to(DEV)
    with torch.no_grad():
        m.model[0].weight.copy_(torch.from_numpy(g["W1"])); m.model[0].bias.copy_(torch.from_numpy(g["b1"]))
        m.model[2].weight.copy_(torch.from_numpy(g["W2"])); m.model[2].bias.copy_(torch.from_numpy(g["b2"]))
    return m


def run_case(g, fused):
    import nca
    nca.USE_FUSED_STEP = fused
    model = build_model(g)
    x = torch.from_numpy(g["x"]).to(DEV)
    h, dims = float(g["h"]), int(g["dims"])
    fr = float(g["fire_rate_call"])
    fr = None if fr < 0 else fr
    grid, v = model.init_grid(x, tuple(g["sections"]), h, dims)
    assert_close(v, g["v"], what="volume")
    A0 = torch.from_numpy(g["states"][0]).to(DEV).requires_grad_(True)
    K = g["masks"].shape[0]
    states = [A0]
    A = A0
    with forced_fire_masks(g["masks"]):
        for _ in range(K):
            _, A = model(x, v, A, h, grid, fire_rate=fr)
            states.append(A)
    loss = (states[-1][:, :4] ** 2).mean()
    if K > 1:
        loss = loss + 0.1 * (states[1][:, :4] ** 2).mean()
    loss.backward()
    return model, A0, states, loss


@pytest.fixture(autouse=True)
def _default_mlp_path():
    """every test starts (and leaves) with the default tcgen05 MLP path"""
    from sphops import _lib, fused
    _lib.check(_lib.lib.sphnca_set_mlp_path(1))
    fused.SIMT_BACKWARD = False
    fused.SPLIT_FORWARD = True
    yield
    _lib.check(_lib.lib.sphnca_set_mlp_path(1))
    fused.SIMT_BACKWARD = False
    fused.SPLIT_FORWARD = True


@pytest.mark.parametrize("path", ["ops", "fused_simt", "fused_tc", "fused_tc_onekernel"])
@pytest.mark.parametrize("name", STEP_CASES)
def test_step_matches_reference(name, path):
    """ops = general sphops operators + torch MLP; fused_simt = fused step with the fp32 CUDA-core MLP
    (forward, data gradient, weight gradient); fused_tc = tcgen05 (3xTF32) MLP kernels for all three with the
    forward split into sweep + MLP kernels (the default); fused_tc_onekernel = same with the forward in one
    persistent kernel."""
    import sphops
    from sphops import _lib, fused as fused_mod
    fused = path != "ops"
    if fused:
        assert sphops.fused_available(), "libsphnca.so lacks the fused step"
        _lib.check(_lib.lib.sphnca_set_mlp_path({"fused_tc": 1, "fused_tc_onekernel": 2, "fused_simt": 0}[path]))
        fused_mod.SIMT_BACKWARD = path == "fused_simt"
        fused_mod.SPLIT_FORWARD = path == "fused_tc"
    g = load_golden(name)
    model, A0, states, loss = run_case(g, fused)
    K = g["masks"].shape[0]
    for k in range(1, K + 1):
        ref = g["states"][k]
        diff = np.abs(states[k].detach().cpu().numpy().astype(np.float64) - ref)
        bad = (diff.max(-1) > RTOL * np.abs(ref).max()).mean()
        assert bad == 0.0, f"step {k}: {bad:.3%} of particles differ (max rel {diff.max()/np.abs(ref).max():.2e})"
    assert abs(float(loss) - float(g["loss"])) <= RTOL * abs(float(g["loss"]))
    assert_close(A0.grad, g["dA0"], rtol=GRAD_RTOL, what="dL/dA0")
    P = dict(model.named_parameters())
    for key, pn in (("dW1", "model.0.weight"), ("db1", "model.0.bias"), ("dW2", "model.2.weight"), ("db2", "model.2.bias")):
        assert_close(P[pn].grad, g[key], rtol=GRAD_RTOL, what=key)


@pytest.mark.parametrize("checkpoint", [False, True], ids=["tape", "checkpointed"])
@pytest.mark.parametrize("name", STEP_CASES)
def test_rollout_node_matches_reference(name, checkpoint):
    """The DEFAULT product path -- `SPHNCA.rollout`: one autograd node, fire lists, fired-rows perception, tcgen05 kernels,
    with and without activation checkpointing -- against the reference's golden vectors: every state, the loss,
    dL/dA0 and the four parameter gradients (same loss as run_case: final state + 0.1 x state 1)."""
    import nca
    g = load_golden(name)
    nca.USE_FUSED_STEP = True
    model = build_model(g)
    x = torch.from_numpy(g["x"]).to(DEV)
    h, dims = float(g["h"]), int(g["dims"])
    fr = float(g["fire_rate_call"])
    fr = None if fr < 0 else fr
    grid, v = model.init_grid(x, tuple(g["sections"]), h, dims)
    A0 = torch.from_numpy(g["states"][0]).to(DEV).requires_grad_(True)
    K = g["masks"].shape[0]
    with forced_fire_masks(g["masks"]):
        out = model.rollout(x, v, A0, h, grid, K, fire_rate=fr, checkpoint=checkpoint)
    assert out is not None and len(out) == K + 1
    loss = (out[-1][:, :4] ** 2).mean()
    if K > 1:
        loss = loss + 0.1 * (out[1][:, :4] ** 2).mean()
    loss.backward()
    for k in range(1, K + 1):
        ref = g["states"][k]
        diff = np.abs(out[k].detach().cpu().numpy().astype(np.float64) - ref)
        bad = (diff.max(-1) > RTOL * np.abs(ref).max()).mean()
        assert bad == 0.0, f"step {k}: {bad:.3%} of particles differ (max rel {diff.max()/np.abs(ref).max():.2e})"
    assert abs(float(loss.detach()) - float(g["loss"])) <= RTOL * abs(float(g["loss"]))
    assert_close(A0.grad, g["dA0"], rtol=GRAD_RTOL, what="dL/dA0")
    P = dict(model.named_parameters())
    for key, pn in (("dW1", "model.0.weight"), ("db1", "model.0.bias"), ("dW2", "model.2.weight"), ("db2", "model.2.bias")):
        assert_close(P[pn].grad, g[key], rtol=GRAD_RTOL, what=key)


def test_rollout_node_matches_oracle_midsize_and_is_reproducible():
    """The default rollout node vs the fp64 oracle at 14 k 3-D particles with the BASELINE neighbour count (K=4, fwd+bwd) -- with
    the GPU's ReLU decisions imposed on the oracle and every step compared from identical inputs (tests/
    test_gpu_parity_config_sizes.py: states, loss, dL/dA0 and the four parameter gradients at the fp32 bars, no allowance for
    flipped kinks) -- and a second run of it: bit-identical states and gradients (ordered fire lists, fixed-order reductions,
    gather adjoints)."""
    import nca
    from tests.test_gpu_parity_config_sizes import compare_with_imposed_bits
    K = 4
    x, sections, h, dims, A0, params, masks, cfg = oracle_case(24, True, False, K, seed=21)
    compare_with_imposed_bits(x, sections, h, dims, A0, params, masks, cfg, "3-D 14k")
    nca.USE_FUSED_STEP = True
    m = nca.SPHNCA(16, 0.5, use_alpha=False, normalize_perception=cfg.normalize_perception).to(DEV)
    with torch.no_grad():
        for t, src in zip((m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias), params):
            t.copy_(src)
    xd = x.to(DEV)
    runs = []
    for _ in range(2):
        m.zero_grad()
        grid, v = m.init_grid(xd, sections, h, dims)
        A = A0.to(DEV).requires_grad_(True)
        with forced_fire_masks(list(masks)):
            out = m.rollout(xd, v, A, h, grid, K)
        loss = (out[-1][:, :4] ** 2).mean()
        loss.backward()
        runs.append([out[-1].detach().clone(), A.grad.clone()] + [p.grad.clone() for p in m.parameters()])
    for a, b in zip(*runs):
        assert torch.equal(a, b)


def test_sample_plane_driver_matches_reference_call_pattern():
    """nca.sample_plane on the fused path: pads 2-D to 3-D, K steps, returns x in 2-D; masks forced to the
    reference run's draws (the reference drew them on the CPU generator)."""
    import nca
    g = load_golden("sample_plane_zebra")
    nca.USE_3D = True
    nca.USE_FUSED_STEP = True
    m = nca.SPHNCA(16, 0.5, update_rule="gated", use_alpha=False, normalize_perception=-1).to(DEV)
    with torch.no_grad():
        m.model[0].weight.copy_(torch.from_numpy(g["W1"])); m.model[0].bias.copy_(torch.from_numpy(g["b1"]))
        m.model[2].weight.copy_(torch.from_numpy(g["W2"])); m.model[2].bias.copy_(torch.from_numpy(g["b2"]))
    x = torch.from_numpy(g["x"]).to(DEV)
    N = x.shape[0]
    K = g["states"].shape[0] - 1
    torch.manual_seed(int(g["torch_seed"]))
    masks = [torch.rand(N) <= 0.5 for _ in range(K)]
    with torch.no_grad(), forced_fire_masks(masks):
        loss, x_out, out = nca.sample_plane(m, x, torch.from_numpy(g["A0"]).to(DEV), tuple(g["sections"]),
                                            float(g["h"]), int(g["dims"]), out_steps=True, iter_n=K)
    assert loss == 0.0 and x_out.shape == (N, 2) and len(out) == K + 1
    for k in range(K + 1):
        assert_close(out[k], g["states"][k], what=f"state {k}")


def oracle_case(n_side, dim3, use_alpha, K, seed):
    from oracle import sphnca_oracle as O
    gen = torch.Generator().manual_seed(seed)
    if dim3:
        x = jittered_lattice(n_side, 3, gen); h = 1.97 * 2 / n_side
        sections = (n_side ** 3,)
    else:
        x = torch.nn.functional.pad(jittered_lattice(n_side, 2, gen), (0, 1)); h = 3.2 * 2 / n_side
        sections = (n_side * n_side // 2, n_side * n_side - n_side * n_side // 2)
    N = x.shape[0]
    dims = math.ceil(2 / h)
    w = load_golden("weights_gecko" if use_alpha else "weights_zebra")
    params = [torch.from_numpy(w[k]) for k in ("W1", "b1", "W2", "b2")]
    if use_alpha:
        d2 = (x ** 2).sum(-1)
        A0 = (torch.clamp(1 - d2 / 0.5 ** 2, 0, 1) ** 3)[:, None].repeat(1, 16) + 0.05 * torch.rand(N, 16, generator=gen)
    else:
        A0 = torch.rand(N, 16, generator=gen)
    masks = torch.rand(K, N, generator=gen) <= 0.5
    cfg = O.StepConfig(use_alpha=use_alpha, normalize_perception=(1.0 / h if use_alpha else -1.0))
    return x, sections, h, dims, A0, params, masks, cfg


@pytest.mark.parametrize("dim3,use_alpha,n_side", [(True, False, 24), (False, True, 100)], ids=["3d_14k", "2d_10k_alpha"])
def test_fused_rollout_matches_oracle_midsize(dim3, use_alpha, n_side):
    """K=4 rollout fwd+bwd vs the fp64 oracle at the BASELINE neighbour counts (two sections in 2-D; the alpha-channel model
    with its life-mask sweeps): the GPU's ReLU decisions imposed, every step from identical inputs, fp32 bars on every state
    and gradient (tests/test_gpu_parity_config_sizes.py).  The single-step autograd nodes (`SPHNCA.forward`) are tied to
    this path bit for bit by test_rollout_node_equals_stepwise_tape."""
    from tests.test_gpu_parity_config_sizes import compare_with_imposed_bits
    x, sections, h, dims, A0, params, masks, cfg = oracle_case(n_side, dim3, use_alpha, 4, seed=21)
    compare_with_imposed_bits(x, sections, h, dims, A0, params, masks, cfg, "mid-size " + ("alpha 2-D" if use_alpha else "3-D"))


def test_fused_equals_ops_path_bitwise_masks():
    """Fused and operator-by-operator paths agree on a 40k-particle step, including every life-mask bit."""
    import nca
    gen = torch.Generator().manual_seed(5)
    n = 200
    x = torch.nn.functional.pad(jittered_lattice(n, 2, gen), (0, 1)).to(DEV)
    N = x.shape[0]
    h = 3.2 * 2 / n
    w = load_golden("weights_gecko")
    m = nca.SPHNCA(16, 0.5, use_alpha=True, normalize_perception=1.0 / h).to(DEV)
    with torch.no_grad():
        for t, k in zip((m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias), ("W1", "b1", "W2", "b2")):
            t.copy_(torch.from_numpy(w[k]))
    d2 = (x ** 2).sum(-1)
    A0 = ((torch.clamp(1 - d2 / 0.6 ** 2, 0, 1) ** 3)[:, None].repeat(1, 16) + 0.05 * torch.rand(N, 16, generator=gen).to(DEV))
    grid, v = m.init_grid(x, (N,), h, math.ceil(2 / h))
    mask = torch.rand(N, generator=gen) <= 0.5
    outs = []
    for fused in (False, True):
        nca.USE_FUSED_STEP = fused
        A = A0.clone().requires_grad_(True)
        with forced_fire_masks([mask]):
            _, nA = m(x, v, A, h, grid)
        (nA[:, :4] ** 2).mean().backward()
        outs.append((nA.detach(), A.grad.clone(), [p.grad.clone() for p in m.parameters()]))
        m.zero_grad()
    zero_rows = [(o[0].abs().sum(-1) == 0) for o in outs]
    assert torch.equal(zero_rows[0], zero_rows[1])
    assert 0.05 < float(zero_rows[0].float().mean()) < 0.95      # the mask actually cuts something
    assert_close(outs[1][0], outs[0][0], rtol=2e-5, what="nA")
    assert_close(outs[1][1], outs[0][1], rtol=5e-5, what="dA")
    for a, b in zip(outs[1][2], outs[0][2]):
        assert_close(a, b, rtol=5e-5, what="param grad")


@pytest.mark.parametrize("n_particles", [1, 127, 128, 129, 5000, 148 * 128 * 2 + 77])
def test_tc_forward_equals_simt_forward(n_particles):
    """The tcgen05 step_fwd (persistent, 128-particle tiles) against the CUDA-core step_fwd on ragged tile
    counts (partial tiles, fewer tiles than SMs, more than 2 tiles per SM): every saved tensor must agree.

    Tolerance: the state here is white noise at a fine lattice (perception values ~50, hidden sums of
    thousands) while the outputs are O(1), so both kernels sit at the fp32 noise floor of the dot
    products: fp32 FMA chains ~3e-5 of max|A_new| vs fp64, the tensor core's truncating fp32
    accumulator ~1.5e-4 (measured).  The reference-scale parity bar (1e-5) is enforced by the golden and
    oracle tests above on trained-weight, BASELINE-scaled inputs; this test pins tiling, not rounding."""
    import ctypes as C
    import sphops
    from sphops import _lib, fused
    gen = torch.Generator().manual_seed(n_particles)
    N = n_particles
    side = max(int(math.ceil(N ** 0.5)), 1)
    x = torch.nn.functional.pad(jittered_lattice(side, 2, gen)[:N], (0, 1))     # first N points of a jittered lattice
    h = max(3.2 * 2 / side, 0.02)
    dims = max(math.ceil(2 / h), 1)
    xd = x.to(DEV)
    grid = sphops.initialize_hashgrid(xd, (N,), h, (dims, dims, 3))
    v = sphops.volume(xd, h, grid)
    w = load_golden("weights_gecko")
    W = [torch.from_numpy(w[k]).to(DEV) for k in ("W1", "b1", "W2", "b2")]
    A = torch.rand(N, 16, generator=gen).to(DEV)
    fire = (torch.rand(N, generator=gen) <= 0.5).to(DEV)
    xv4 = grid.sorted_xv(xd, v)
    A_s = grid.pack_rows(A, 16)
    fire_s = grid.pack_bytes(fire.to(torch.uint8))
    res = []
    for path in (0, 1, 2):
        _lib.check(_lib.lib.sphnca_set_mlp_path(path))
        An = torch.zeros(N, 16, device=DEV); Y = torch.zeros(N, 32, device=DEV); O = torch.zeros(N, 36, device=DEV)
        fl = torch.zeros(N, dtype=torch.uint8, device=DEV); al = torch.zeros(N, dtype=torch.uint8, device=DEV)
        rl = torch.zeros(N, 8, dtype=torch.int32, device=DEV)
        p = fused._step_params(*W, "gated", True, 1.0 / h, 1.0)
        _lib.check(_lib.lib.sphnca_step_fwd(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(A_s), _lib.ptr(fire_s), C.byref(p),
                                            _lib.ptr(An), _lib.ptr(al), _lib.ptr(Y), _lib.ptr(O), _lib.ptr(fl), _lib.ptr(rl), _lib.stream_ptr()))
        torch.cuda.synchronize()
        res.append((An, Y, O, fl, al, rl.view(torch.uint8)))
    # ReLU bits: a pre-activation within fp32 noise of 0 may land on either side in the two kernels
    for tc in (1, 2):
        nbits = int(torch.tensor([bin(v).count("1") for v in (res[0][5] ^ res[tc][5]).flatten().tolist()]).sum()) if N <= 6000 else 0
        assert nbits <= max(4, N // 200), f"relu masks differ in {nbits} bits (path {tc})"
        for a, b, nm in zip(res[tc][:5], res[0][:5], ("A_new", "Y", "O", "flags", "alive")):
            if a.dtype == torch.uint8:
                # a `> 0.1` decision on a value within fp32 noise of 0.1 may differ between the two kernels
                assert int((a != b).sum()) <= max(2, N // 10000), f"{nm}: {int((a != b).sum())} bytes differ (path {tc})"
            else:
                assert_close(a, b, rtol=2e-5 if nm == "Y" else 5e-4, what=f"{nm} path {tc}")
    # the two tensor-core forwards run the same MMA sequence on the same operands
    assert_close(res[1][0], res[2][0], rtol=2e-4, what="A_new split vs fused")   # same noise floor (gate sums ordered differently)


def test_backward_chunking_is_equivalent(monkeypatch):
    """The backward processes particles in chunks (bounded [256][chunk] scratch); odd chunk sizes
    (not multiples of 4/32, so the scalar tail paths run) must give the same gradients."""
    import nca
    from sphops import fused
    g = load_golden("rollout_gecko_alpha_k3")
    res = []
    monkeypatch.setattr(fused, "SIMT_BACKWARD", True)
    for chunk in (4 * 1024 * 1024, 61):
        monkeypatch.setattr(fused, "WGRAD_CHUNK", chunk)
        model, A0, states, loss = run_case(g, True)
        res.append([A0.grad.clone()] + [p.grad.clone() for p in model.parameters()])
    for a, b in zip(res[1], res[0]):
        assert_close(a, b, rtol=2e-6, what="chunked vs unchunked")
    for t, key in zip(res[1][1:], ("dW1", "db1", "dW2", "db2")):
        assert_close(t, g[key], rtol=GRAD_RTOL, what=key)


@pytest.mark.parametrize("n_particles", [1, 63, 64, 65, 129, 5000, 148 * 128 * 2 + 77])
def test_tc_backward_equals_simt_backward(n_particles):
    """tcgen05 data-gradient (step_bwd_tc_kernel) and weight-gradient (mlp_wgrad_tc_kernel) kernels against the
    CUDA-core kernels and an fp64 torch evaluation of the same formulas, on ragged tile / chunk counts."""
    import ctypes as C
    import sphops
    from sphops import _lib, fused
    gen = torch.Generator().manual_seed(100 + n_particles)
    N = n_particles
    side = max(int(math.ceil(N ** 0.5)), 1)
    x = torch.nn.functional.pad(jittered_lattice(side, 2, gen)[:N], (0, 1))
    h = max(3.2 * 2 / side, 0.02)
    dims = max(math.ceil(2 / h), 1)
    xd = x.to(DEV)
    grid = sphops.initialize_hashgrid(xd, (N,), h, (dims, dims, 3))
    v = sphops.volume(xd, h, grid)
    w = load_golden("weights_gecko")
    W = [torch.from_numpy(w[k]).to(DEV) for k in ("W1", "b1", "W2", "b2")]
    d2 = (x ** 2).sum(-1)
    A = ((torch.clamp(1 - d2 / 0.6 ** 2, 0, 1) ** 3)[:, None].repeat(1, 16) + 0.05 * torch.rand(N, 16, generator=gen)).to(DEV)
    fire = (torch.rand(N, generator=gen) <= 0.5).to(DEV)
    xv4 = grid.sorted_xv(xd, v); A_s = grid.pack_rows(A, 16); fire_s = grid.pack_bytes(fire.to(torch.uint8))
    f32 = dict(dtype=torch.float32, device=DEV)
    An = torch.zeros(N, 16, **f32); Y = torch.zeros(N, 32, **f32); O = torch.zeros(N, 36, **f32)
    fl = torch.zeros(N, dtype=torch.uint8, device=DEV); al = torch.zeros(N, dtype=torch.uint8, device=DEV)
    rl = torch.zeros(N, 8, dtype=torch.int32, device=DEV)
    p = fused._step_params(*W, "gated", True, 1.0 / h, 1.0)
    st = _lib.stream_ptr()
    _lib.check(_lib.lib.sphnca_set_mlp_path(0))
    _lib.check(_lib.lib.sphnca_step_fwd(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(A_s), _lib.ptr(fire_s), C.byref(p), _lib.ptr(An),
                                        _lib.ptr(al), _lib.ptr(Y), _lib.ptr(O), _lib.ptr(fl), _lib.ptr(rl), st))
    _lib.check(_lib.lib.sphnca_step_mask(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(al), _lib.ptr(An), _lib.ptr(fl), st))
    dAn = torch.randn(N, 16, generator=gen).to(DEV)
    nb = C.c_size_t(0); _lib.check(_lib.lib.sphnca_mlp_wgrad_workspace_bytes(N, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=DEV)
    tiles = (N + 31) // 32
    out = {}
    for path in (0, 1):
        _lib.check(_lib.lib.sphnca_set_mlp_path(path))
        dO = torch.zeros(N, 36, **f32); dY = torch.zeros(N, 32, **f32); dAd = torch.zeros(N, 16, **f32)
        Ht = torch.empty(tiles, 256, 32, **f32) if path == 0 else None
        dHt = torch.empty(tiles, 256, 32, **f32) if path == 0 else None
        _lib.check(_lib.lib.sphnca_step_bwd(N, _lib.ptr(A_s), _lib.ptr(Y), _lib.ptr(O), _lib.ptr(fl), _lib.ptr(rl), _lib.ptr(dAn),
                                            C.byref(p), _lib.ptr(dO), _lib.ptr(dY), _lib.ptr(dAd), _lib.ptr(Ht), _lib.ptr(dHt), st))
        dW = [torch.zeros_like(t) for t in W]
        _lib.check(_lib.lib.sphnca_mlp_wgrad(N, _lib.ptr(A_s), _lib.ptr(Y), _lib.ptr(dO), _lib.ptr(Ht), _lib.ptr(dHt), C.byref(p),
                                             *[_lib.ptr(t) for t in dW], _lib.ptr(ws), nb.value, st))
        torch.cuda.synchronize()
        out[path] = (dO, dY, dAd, dW)
    # fp64 evaluation of the same formulas from the saved tensors
    Yf = torch.cat([A_s, Y], 1).double(); W1, b1, W2, b2 = [t.double() for t in W]
    pre = Yf @ W1.t() + b1
    dOd = out[0][0][:, :33].double()
    dH = (dOd @ W2) * (pre > 0)
    ref = dict(dY=(dH @ W1)[:, 16:], dW1=dH.t() @ Yf, db1=dH.sum(0), dW2=dOd.t() @ torch.relu(pre), db2=dOd.sum(0))
    for path in (0, 1):
        dO, dY, dAd, dW = out[path]
        tol = 2e-5
        assert_close(dO, out[0][0], rtol=1e-6, what=f"dO path {path}")
        assert_close(dY, ref["dY"], rtol=tol, what=f"dY path {path}")
        for t, nm in zip(dW, ("dW1", "db1", "dW2", "db2")):
            assert_close(t, ref[nm], rtol=tol, what=f"{nm} path {path}")
    assert_close(out[1][2], out[0][2], rtol=2e-5, what="dAdir tc vs simt")


def _sweep_inputs(kind):
    gen = torch.Generator().manual_seed(7)
    if kind.startswith("golden:"):
        g = load_golden(kind.split(":")[1])
        return torch.from_numpy(g["x"]), tuple(int(s) for s in g["sections"]), float(g["h"]), tuple(int(d) for d in g["dims"])
    if kind == "one_cell_dense":           # 400 particles in ONE cell: chunking (CAP=128) and hit-list flushes (SLOTS=48)
        return torch.rand(400, 3, generator=gen) * 0.2 + 0.1, (400,), 0.5, (4, 4, 4)
    if kind == "lattice3d":
        x = jittered_lattice(30, 3, gen)
        return x, (9000, 27000 - 9000), 1.97 * 2 / 30, (31, 31, 31)
    if kind == "lattice2d_padded":
        x = torch.nn.functional.pad(jittered_lattice(150, 2, gen), (0, 1))
        return x, (150 * 150,), 3.2 * 2 / 150, (47, 47, 47)
    raise KeyError(kind)


@pytest.mark.parametrize("kind", ["golden:ops_3d_ragged", "golden:ops_padded2d", "golden:ops_thin_axis", "golden:ops_single_layer",
                                  "golden:ops_2d", "one_cell_dense", "lattice3d", "lattice2d_padded"])
def test_tile_sweeps_equal_gather_sweeps(kind):
    """Shared-memory tile engine (segments per cell row, wrapped-cell runs, chunked staging, hit lists) against the
    one-thread-per-particle gather sweeps: perception + previous life mask, and the gather adjoint.  Covers ragged
    sections, coordinates outside the box (modulo wrap), axes with dims < 3 (double / triple counting), true 2-D,
    a cell holding more particles than a chunk, and BASELINE-like lattices."""
    import ctypes as C
    import sphops
    from sphops import _lib, fused
    x, sections, h, dims = _sweep_inputs(kind)
    N, D = x.shape
    xd = x.to(DEV)
    grid = sphops.initialize_hashgrid(xd, sections, h, dims)
    v = sphops.volume(xd, h, grid)
    gen = torch.Generator().manual_seed(3)
    A = torch.rand(N, 16, generator=gen).to(DEV)
    dY = torch.randn(N, 32, generator=gen).to(DEV)
    dAd = torch.randn(N, 16, generator=gen).to(DEV)
    xv4 = grid.sorted_xv(xd, v)
    A_s = grid.pack_rows(A, 16); dY_s = grid.pack_rows(dY, 32); dAd_s = grid.pack_rows(dAd, 16)
    w = load_golden("weights_gecko")
    W = [torch.from_numpy(w[k]).to(DEV) for k in ("W1", "b1", "W2", "b2")]
    p = fused._step_params(*W, "gated", True, 1.7, 1.0)
    st = _lib.stream_ptr()
    res = []
    for path in (0, 1, 2):
        _lib.check(_lib.lib.sphnca_set_sweep_path(path))
        Y = torch.zeros(N, 32, device=DEV); fl = torch.zeros(N, dtype=torch.uint8, device=DEV); dA = torch.zeros(N, 16, device=DEV)
        _lib.check(_lib.lib.sphnca_step_perceive(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(A_s), C.byref(p), _lib.ptr(Y), _lib.ptr(fl), st))
        _lib.check(_lib.lib.sphnca_step_adjoint(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(dY_s), _lib.ptr(dAd_s), C.byref(p), _lib.ptr(dA), st))
        torch.cuda.synchronize()
        res.append((Y, fl, dA))
    _lib.check(_lib.lib.sphnca_set_sweep_path(0))
    assert_close(res[1][0], res[0][0], rtol=2e-5 if kind != "one_cell_dense" else 2e-4, what="perception")
    assert int((res[1][1] != res[0][1]).sum()) <= max(1, N // 5000), "previous life mask"
    assert_close(res[1][2], res[0][2], rtol=2e-5 if kind != "one_cell_dense" else 2e-4, what="adjoint")
    # TMA-staged tiles (path 2): same visiting order and arithmetic as the gathers -> bit-identical, including the
    # tiles that fall back to global loads (more rows or candidates than the staging buffer holds)
    assert torch.equal(res[2][0], res[0][0]) and torch.equal(res[2][1], res[0][1]) and torch.equal(res[2][2], res[0][2])
    # pair lists (csrc/pairs.cu): geometry, in-radius test and weights precomputed once per grid; same visiting order and
    # the same floats -> bit-identical perception, life mask and adjoint
    pl = grid.pair_lists(xd, v)
    if _lib.USE_PAIR_LISTS and _lib.FORCE_SWEEP_PATH is None:
        assert pl is not None and pl["entries"] % 32 == 0      # groups of 8 particles, slot counts in multiples of 4
        Y = torch.zeros(N, 32, device=DEV); fl = torch.zeros(N, dtype=torch.uint8, device=DEV); dA = torch.zeros(N, 16, device=DEV)
        _lib.check(_lib.lib.sphnca_step_perceive_pairs(grid.c_grid(), _lib.ptr(pl["slots"]), _lib.ptr(pl["base"]), _lib.ptr(pl["J"]),
                                                       _lib.ptr(pl["WP"]), _lib.ptr(pl["WB"]), _lib.ptr(pl["FLG"]), _lib.ptr(A_s), C.byref(p), _lib.ptr(Y), _lib.ptr(fl), st))
        _lib.check(_lib.lib.sphnca_step_adjoint_pairs(grid.c_grid(), _lib.ptr(pl["slots"]), _lib.ptr(pl["base"]), _lib.ptr(pl["J"]),
                                                      _lib.ptr(pl["WA"]), _lib.ptr(pl["SXY"]), _lib.ptr(xv4), _lib.ptr(dY_s), _lib.ptr(dAd_s),
                                                      C.byref(p), _lib.ptr(dA), st))
        torch.cuda.synchronize()
        assert torch.equal(Y, res[0][0]), "perception from pair lists"
        assert torch.equal(fl, res[0][1]), "life mask from pair lists"
        assert torch.equal(dA, res[0][2]), "adjoint from pair lists"
        # dL/dY rows as [x quarter | y quarter] pairs (sphnca_step_params.dy_interleaved: what sphnca_step_bwd_list writes when the
        # flag is set): one 32-byte gather per lane instead of two 16-byte ones, same arithmetic
        p_ilv = fused._step_params(*W, "gated", True, 1.7, 1.0); p_ilv.dy_interleaved = 1
        dY_ilv = dY_s.view(N, 2, 4, 4).transpose(1, 2).contiguous().view(N, 32)
        dA2 = torch.zeros(N, 16, device=DEV)
        _lib.check(_lib.lib.sphnca_step_adjoint_pairs(grid.c_grid(), _lib.ptr(pl["slots"]), _lib.ptr(pl["base"]), _lib.ptr(pl["J"]),
                                                      _lib.ptr(pl["WA"]), _lib.ptr(pl["SXY"]), _lib.ptr(xv4), _lib.ptr(dY_ilv), _lib.ptr(dAd_s),
                                                      C.byref(p_ilv), _lib.ptr(dA2), st))
        torch.cuda.synchronize()
        assert torch.equal(dA2, res[0][2]), "adjoint from pair lists, interleaved dL/dY rows"
        # no alpha channel: the life mask comes from the flag the fill kernel evaluated once (the blur weights are not read)
        p0 = fused._step_params(*W, "gated", False, 1.7, 1.0)
        outs = []
        for use_lists in (False, True):
            Y = torch.zeros(N, 32, device=DEV); fl = torch.zeros(N, dtype=torch.uint8, device=DEV)
            if use_lists:
                _lib.check(_lib.lib.sphnca_step_perceive_pairs(grid.c_grid(), _lib.ptr(pl["slots"]), _lib.ptr(pl["base"]), _lib.ptr(pl["J"]),
                                                               _lib.ptr(pl["WP"]), _lib.ptr(pl["WB"]), _lib.ptr(pl["FLG"]), _lib.ptr(A_s), C.byref(p0),
                                                               _lib.ptr(Y), _lib.ptr(fl), st))
            else:
                _lib.check(_lib.lib.sphnca_step_perceive(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(A_s), C.byref(p0), _lib.ptr(Y), _lib.ptr(fl), st))
            torch.cuda.synchronize()
            outs.append((Y, fl))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), "no-alpha perception from pair lists"


def test_rollout_is_bit_reproducible_at_config_size():
    """No float atomics anywhere (gather adjoints, fixed-order weight-gradient reduction): two runs of the same
    8 x 65,536-particle rollout (BASELINE configs[1] size, 3 steps fwd+bwd) give bit-identical states and gradients."""
    import nca
    from bench import make_inputs, WORKLOADS, load_weights
    wl = WORKLOADS["train2d_b8x65k_k64"]
    x, A0, sections, h, dims = make_inputs(wl, 0)
    nca.USE_3D, nca.USE_FUSED_STEP = True, True
    m = nca.SPHNCA(16, 0.5, use_alpha=False, normalize_perception=1.0 / h).to(DEV)
    with torch.no_grad():
        for t, src in zip((m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias), load_weights()):
            t.copy_(src)
    xd = torch.nn.functional.pad(x, (0, 1)).to(DEV)
    runs = []
    for _ in range(2):
        m.zero_grad()
        torch.manual_seed(99)
        grid, v = m.init_grid(xd, sections, h, dims)
        A = A0.to(DEV).requires_grad_(True)
        A_in = A
        for _ in range(3):
            _, A = m(xd, v, A, h, grid)
        (A[:, :4] ** 2).mean().backward()
        runs.append([A.detach().clone(), A_in.grad.clone()] + [p.grad.clone() for p in m.parameters()])
    for a, b in zip(*runs):
        assert torch.equal(a, b)
    # and the size-independent sanity of the result: finite, masks not degenerate
    assert bool(torch.isfinite(runs[0][0]).all()) and bool(torch.isfinite(runs[0][1]).all())


# ---- K-step rollout as one autograd node ---------------------------------------
def _rollout_model(use_alpha, h):
    import nca
    from bench import load_weights
    nca.USE_3D, nca.USE_FUSED_STEP = True, True
    w = load_golden("weights_gecko" if use_alpha else "weights_zebra")
    m = nca.SPHNCA(16, 0.5, use_alpha=use_alpha, normalize_perception=1.0 / h).to(DEV)
    with torch.no_grad():
        for t, k in zip((m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias), ("W1", "b1", "W2", "b2")):
            t.copy_(torch.from_numpy(w[k]))
    return m


def _multi_state_loss(out, picks):
    loss = (out[-1][:, :4] ** 2).mean() + 0.3 * (out[0][:, 4:8] ** 2).mean()
    for k in picks:
        loss = loss + 0.1 * (out[k][:, :4] - 0.25).abs().mean()
    return loss


@pytest.mark.parametrize("use_alpha", [False, True], ids=["noalpha", "alpha"])
@pytest.mark.parametrize("picks", [(1, 3, 3), ()], ids=["multi_state_loss", "final_only"])
def test_rollout_node_equals_stepwise_tape(use_alpha, picks):
    """One K-step node (sorted-resident state, in-place accumulated parameter gradients) == K fused single-step
    nodes: states bit-identical, gradients to fp32 summation order."""
    import nca
    K = 5
    x, sections, h, dims, A0, params, masks, cfg = oracle_case(64, False, use_alpha, K, seed=5)
    m = _rollout_model(use_alpha, h)
    xd = x.to(DEV)
    res = []
    for fused_rollout in (False, True):
        m.zero_grad()
        grid, v = m.init_grid(xd, sections, h, dims)
        A = A0.to(DEV).requires_grad_(True)
        with forced_fire_masks(list(masks)):
            if fused_rollout:
                out = m.rollout(xd, v, A, h, grid, K)
                assert out is not None and len(out) == K + 1
            else:
                out, cur = [], A
                for _ in range(K):
                    out.append(cur.clone())
                    _, cur = m(xd, v, cur, h, grid)
                out.append(cur.clone())
        _multi_state_loss(out, picks).backward()
        res.append(([o.detach().clone() for o in out], A.grad.clone(), [p.grad.clone() for p in m.parameters()]))
    for k, (a, b) in enumerate(zip(res[0][0], res[1][0])):
        assert torch.equal(a, b), f"state {k}"
    assert_close(res[1][1], res[0][1], rtol=1e-6, what="dL/dA0")
    for a, b, nm in zip(res[0][2], res[1][2], ("dW1", "db1", "dW2", "db2")):
        assert_close(b, a, rtol=1e-6, what=nm)


def test_rollout_node_loss_on_an_intermediate_state_only():
    """A loss that reads only state 2 of 5: steps 3..5 carry no gradient and are skipped in the backward."""
    K = 5
    x, sections, h, dims, A0, params, masks, cfg = oracle_case(48, False, False, K, seed=8)
    m = _rollout_model(False, h)
    xd = x.to(DEV)
    res = []
    for fused_rollout in (False, True):
        m.zero_grad()
        grid, v = m.init_grid(xd, sections, h, dims)
        A = A0.to(DEV).requires_grad_(True)
        with forced_fire_masks(list(masks)):
            if fused_rollout:
                s2 = m.rollout(xd, v, A, h, grid, K)[2]
            else:
                cur = A
                for k in range(K):
                    if k == 2:
                        s2 = cur
                    _, cur = m(xd, v, cur, h, grid)
        (s2[:, :6] ** 2).mean().backward()
        res.append((A.grad.clone(), [p.grad.clone() for p in m.parameters()]))
    assert_close(res[1][0], res[0][0], rtol=1e-6, what="dL/dA0")
    for a, b, nm in zip(res[0][1], res[1][1], ("dW1", "db1", "dW2", "db2")):
        assert_close(b, a, rtol=1e-6, what=nm)


@pytest.mark.parametrize("use_alpha", [False, True], ids=["noalpha", "alpha"])
@pytest.mark.parametrize("pairs", [True, False], ids=["pairlists", "cellsweep"])
def test_rollout_node_checkpointing_is_bit_identical(use_alpha, pairs, monkeypatch):
    """Activation checkpointing: keeping only the states + fire bytes and re-running each step's
    forward in the backward gives the SAME bits as the full tape (deterministic kernels), for a multi-state loss and
    with steps past the last read state skipped; and it holds ~4.7x less memory per particle-step."""
    from sphops import _lib, fused
    if not pairs:
        monkeypatch.setattr(_lib, "USE_PAIR_LISTS", False)
    K = 6
    x, sections, h, dims, A0, params, masks, cfg = oracle_case(72, False, use_alpha, K, seed=11)
    m = _rollout_model(use_alpha, h)
    xd = x.to(DEV)
    res, held = [], []
    for ckpt in (False, True):
        m.zero_grad()
        grid, v = m.init_grid(xd, sections, h, dims)
        A = A0.to(DEV).requires_grad_(True)
        torch.cuda.synchronize()
        before = torch.cuda.memory_allocated()
        with forced_fire_masks(list(masks)):
            out = m.rollout(xd, v, A, h, grid, K, checkpoint=ckpt)
        held.append(torch.cuda.memory_allocated() - before)
        loss = (out[4][:, :4] ** 2).mean() + 0.1 * (out[2][:, :4] - 0.25).abs().mean()     # steps 5, 6 carry no gradient
        loss.backward()
        res.append([o.detach().clone() for o in out] + [A.grad.clone()] + [p.grad.clone() for p in m.parameters()])
    for k, (a, b) in enumerate(zip(*res)):
        assert torch.equal(a, b), f"tensor {k} differs between full tape and checkpointed rollout"
    N = x.shape[0]
    assert held[0] - held[1] >= 0.9 * N * (K - 1) * fused.TAPE_BYTES, held     # the tape that is not kept


def test_rollout_node_auto_checkpoint_follows_free_memory(monkeypatch):
    from sphops import fused
    free = torch.cuda.get_device_properties(0).total_memory
    n_small, n_huge = 1000, int(free / fused.TAPE_BYTES)
    assert fused._tape_mode(n_small, 8, torch.device(DEV), None) == 1
    assert fused._tape_mode(n_huge, 8, torch.device(DEV), None) == 2
    assert fused._tape_mode(n_small, 8, torch.device(DEV), True) == 2 and fused._tape_mode(n_huge, 8, torch.device(DEV), False) == 1
    monkeypatch.setattr(fused, "CHECKPOINT", True)
    assert fused._tape_mode(n_small, 8, torch.device(DEV), None) == 2


@pytest.mark.parametrize("n,rate", [(1, 1.0), (5, 0.0), (1023, 0.5), (1024, 0.5), (1025, 0.3), (300_000, 0.5), (2_100_000, 0.9), (6_000_000, 0.5)])
def test_fire_list_is_the_ordered_compaction(n, rate):
    """sphnca_fire_list: rows[0..count) == nonzero(fire) in increasing order, count on the device; not-fired rows carry
    their state over and get their alive byte; fired rows of A_new are left to the MLP kernel."""
    import ctypes as C
    from sphops import _lib
    gen = torch.Generator().manual_seed(n)
    fire = (torch.rand(n, generator=gen) < rate).to(torch.uint8).to(DEV)
    A = torch.rand(n, 16, generator=gen).to(DEV)
    rows = torch.full((n,), -1, dtype=torch.int32, device=DEV)
    count = torch.full((1,), -7, dtype=torch.int32, device=DEV)
    An = torch.full((n, 16), -5.0, device=DEV)
    alive = torch.full((n,), 9, dtype=torch.uint8, device=DEV)
    nb = C.c_size_t(0)
    _lib.check(_lib.lib.sphnca_fire_list_workspace_bytes(n, C.byref(nb)))
    ws = torch.zeros((nb.value,), dtype=torch.uint8, device=DEV)           # zero before the first use (include/sphnca.h)
    # three calls on the same workspace: the launch leaves its look-back state zeroed for the next one
    for rep in range(3):
        if rep:
            fire = (torch.rand(n, generator=gen) < rate).to(torch.uint8).to(DEV)
            rows.fill_(-1); count.fill_(-7); An.fill_(-5.0); alive.fill_(9)
        _lib.check(_lib.lib.sphnca_fire_list(n, _lib.ptr(fire), _lib.ptr(A), 1, _lib.ptr(rows), _lib.ptr(count), _lib.ptr(An), _lib.ptr(alive),
                                             None, None, 0, _lib.ptr(ws), nb.value, _lib.stream_ptr()))
        want = torch.nonzero(fire)[:, 0].to(torch.int32)
        assert int(count.item()) == want.numel()
        assert torch.equal(rows[:want.numel()], want)
        off = fire == 0
        assert torch.equal(An[off], A[off]) and bool((An[~off] == -5.0).all())
        assert torch.equal(alive[off], (A[off][:, 3] > 0.1).to(torch.uint8)) and bool((alive[~off] == 9).all())
        assert not bool(ws.any())


@pytest.mark.parametrize("use_alpha", [False, True], ids=["noalpha", "alpha"])
@pytest.mark.parametrize("rate", [0.5, 1.0, 0.0], ids=["half", "all_fire", "none_fire"])
def test_rollout_fire_lists_equal_full_mlp(use_alpha, rate, monkeypatch):
    """The MLP half over the fired rows only (csrc/firelist.cu) == the MLP over every row followed by the reference's
    `where(update_mask, nA, A)`: states bit-identical (row-independent arithmetic), gradients to fp32 summation order
    (the weight gradient sums the same per-particle terms in another chunking)."""
    from sphops import fused
    K = 4
    x, sections, h, dims, A0, params, masks, cfg = oracle_case(72, False, use_alpha, K, seed=21)
    gen = torch.Generator().manual_seed(3)
    masks = torch.rand(K, x.shape[0], generator=gen) < rate
    m = _rollout_model(use_alpha, h)
    xd = x.to(DEV)
    res = []
    for lists in (False, True):
        monkeypatch.setattr(fused, "FIRE_LISTS", lists)
        m.zero_grad()
        grid, v = m.init_grid(xd, sections, h, dims)
        A = A0.to(DEV).requires_grad_(True)
        with forced_fire_masks(list(masks)):
            out = m.rollout(xd, v, A, h, grid, K, checkpoint=False)
        _multi_state_loss(out, (1, 2)).backward()
        res.append(([o.detach().clone() for o in out], A.grad.clone(), [p.grad.clone() for p in m.parameters()]))
    for k, (a, b) in enumerate(zip(res[0][0], res[1][0])):
        assert torch.equal(a, b), f"state {k}"
    assert_close(res[1][1], res[0][1], rtol=1e-6, what="dL/dA0")
    for a, b, nm in zip(res[0][2], res[1][2], ("dW1", "db1", "dW2", "db2")):
        if float(a.abs().max()) == 0.0:
            assert float(b.abs().max()) == 0.0, nm
        else:
            assert_close(b, a, rtol=2e-6, what=nm)


def test_rollout_fire_lists_orig_rule_and_call_fire_rate(monkeypatch):
    """Same equality for the `orig` update rule (16 MLP outputs, rate-scaled residual) with a per-call fire rate."""
    import nca
    from sphops import fused
    K = 3
    x, sections, h, dims, A0, params, masks, cfg = oracle_case(56, False, False, K, seed=31)
    nca.USE_3D, nca.USE_FUSED_STEP = True, True
    torch.manual_seed(5)
    m = nca.SPHNCA(16, 0.5, update_rule="orig", use_alpha=False, normalize_perception=1.0 / h).to(DEV)
    with torch.no_grad():
        m.model[2].weight.copy_(0.05 * torch.randn_like(m.model[2].weight))
    xd = x.to(DEV)
    res = []
    for lists in (False, True):
        monkeypatch.setattr(fused, "FIRE_LISTS", lists)
        m.zero_grad()
        grid, v = m.init_grid(xd, sections, h, dims)
        A = A0.to(DEV).requires_grad_(True)
        with forced_fire_masks(list(masks)):
            out = m.rollout(xd, v, A, h, grid, K, fire_rate=0.25, checkpoint=False)
        (out[-1][:, :5] ** 2).mean().backward()
        res.append((out[-1].detach().clone(), A.grad.clone(), [p.grad.clone() for p in m.parameters()]))
    assert torch.equal(res[0][0], res[1][0])
    assert_close(res[1][1], res[0][1], rtol=1e-6, what="dL/dA0")
    for a, b, nm in zip(res[0][2], res[1][2], ("dW1", "db1", "dW2", "db2")):
        assert_close(b, a, rtol=2e-6, what=nm)


def test_rollout_node_inference_keeps_only_what_is_asked():
    """No tape, keep_all=False: two states (first, last); same final state as the stepwise path; sequence protocol."""
    import random
    K = 4
    x, sections, h, dims, A0, params, masks, cfg = oracle_case(48, False, True, K, seed=9)
    m = _rollout_model(True, h)
    xd = x.to(DEV)
    grid, v = m.init_grid(xd, sections, h, dims)
    with torch.no_grad(), forced_fire_masks(list(masks)):
        cur = A0.to(DEV)
        for _ in range(K):
            _, cur = m(xd, v, cur, h, grid)
    with torch.no_grad(), forced_fire_masks(list(masks)):
        short = m.rollout(xd, v, A0.to(DEV), h, grid, K, keep_all=False)
    with torch.no_grad(), forced_fire_masks(list(masks)):
        full = m.rollout(xd, v, A0.to(DEV), h, grid, K)
    assert len(short) == 2 and len(full) == K + 1
    assert torch.equal(short[-1], cur) and torch.equal(full[-1], cur) and torch.equal(full[0], A0.to(DEV))
    assert len(full[1:3]) == 2 and len(list(full)) == K + 1 and full[-1] is full[K]
    assert len(random.choices(full, k=4)) == 4
    with pytest.raises(IndexError):
        full[K + 1]


def test_static_life_mask_equals_sweep_when_no_alpha():
    """use_alpha=False: the elementwise sphnca_step_mask_static gives bit-identical flags / state to the neighbour-sweep
    sphnca_step_mask fed with the all-ones activity, including particles whose life mask is 0 (sparse outliers)."""
    import ctypes as C
    import sphops
    from sphops import _lib, fused
    gen = torch.Generator().manual_seed(11)
    h = 3.2 * 2 / 90
    x = torch.nn.functional.pad(jittered_lattice(90, 2, gen), (0, 1))
    # particles whose life mask IS 0 (blur of the all-ones activity <= 0.1): a ring of 16 isolated particles at 0.8 h
    # around a very dense, very tight blob, in a region cleared of lattice particles.  Their own volume is tiny (the blob
    # sits inside their support) and the blob's volumes are tiny too: blur = 0.0625 (fp64 oracle), far from the threshold.
    c = torch.tensor([0.5, 0.5, 0.0])
    x = x[((x - c) ** 2).sum(-1) > (2.5 * h) ** 2]
    blob = (torch.rand(4000, 3, generator=gen) - 0.5) * torch.tensor([0.002, 0.002, 0.0]) + c
    ang = torch.arange(16) * (2 * math.pi / 16)
    ring = c + 0.8 * h * torch.stack([torch.cos(ang), torch.sin(ang), torch.zeros(16)], -1)
    x = torch.cat([x, blob, ring])
    N = x.shape[0]
    xd = x.to(DEV)
    grid = sphops.initialize_hashgrid(xd, (N,), h, math.ceil(2 / h))
    v = sphops.volume(xd, h, grid)
    xv4 = grid.sorted_xv(xd, v)
    A_s = grid.pack_rows(torch.rand(N, 16, generator=gen).to(DEV), 16)
    w = load_golden("weights_zebra")
    W = [torch.from_numpy(w[k]).to(DEV) for k in ("W1", "b1", "W2", "b2")]
    p = fused._step_params(*W, "gated", False, 1.0 / h, 1.0)
    st = _lib.stream_ptr()
    fire_s = (torch.rand(N, generator=gen) <= 0.5).to(torch.uint8).to(DEV)
    res = []
    for static in (False, True):
        An = torch.empty(N, 16, device=DEV); Y = torch.empty(N, 32, device=DEV); O = torch.empty(N, 36, device=DEV)
        fl = torch.empty(N, dtype=torch.uint8, device=DEV); al = torch.empty(N, dtype=torch.uint8, device=DEV)
        rl = torch.empty(N, 8, dtype=torch.int32, device=DEV)
        _lib.check(_lib.lib.sphnca_step_perceive(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(A_s), C.byref(p), _lib.ptr(Y), _lib.ptr(fl), st))
        _lib.check(_lib.lib.sphnca_step_mlp(N, _lib.ptr(A_s), _lib.ptr(Y), _lib.ptr(fire_s), C.byref(p), _lib.ptr(An), _lib.ptr(al),
                                            _lib.ptr(O), _lib.ptr(fl), _lib.ptr(rl), st))
        if static:
            _lib.check(_lib.lib.sphnca_step_mask_static(N, _lib.ptr(An), _lib.ptr(fl), st))
        else:
            assert bool((al == 1).all())
            _lib.check(_lib.lib.sphnca_step_mask(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(al), _lib.ptr(An), _lib.ptr(fl), st))
        torch.cuda.synchronize()
        res.append((An, fl))
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][0], res[1][0])
    dead = (res[0][1] & 1) == 0
    fired = fire_s.bool()
    # the 16 ring particles are dead; both the fired-and-dead and the not-fired-and-dead (carry-over) branches are exercised
    assert int(dead.sum()) == 16, int(dead.sum())
    assert int((dead & fired).sum()) > 0 and int((dead & ~fired).sum()) > 0
    assert bool((res[1][0][dead] == 0).all()), "a dead particle kept a non-zero state"


def test_staged_sweep_is_chosen_per_grid():
    """sphnca_stage_fit: planar clouds fit the staging buffer of the TMA-staged perception sweep, 3-D clouds do not;
    the fused step picks its kernel from that (results are bit-identical either way, see the sweep test above)."""
    import sphops
    gen = torch.Generator().manual_seed(2)
    x2 = torch.nn.functional.pad(jittered_lattice(160, 2, gen), (0, 1)).to(DEV)
    g2 = sphops.initialize_hashgrid(x2, (160 * 160,), 3.2 * 2 / 160, 50)
    x3 = jittered_lattice(30, 3, gen).to(DEV)
    g3 = sphops.initialize_hashgrid(x3, (27000,), 1.97 * 2 / 30, 31)
    from sphops import _lib
    if _lib.FORCE_SWEEP_PATH is None:
        assert g2.staged_sweep_ok() is True and g3.staged_sweep_ok() is False
    else:
        assert g2.staged_sweep_ok() is False


@pytest.mark.parametrize("use_alpha", [False, True], ids=["noalpha", "alpha"])
@pytest.mark.parametrize("per_group", [1, 2], ids=["one_section_per_group", "two_sections_per_group"])
def test_rollout_node_microbatched_over_sections_equals_whole_batch(use_alpha, per_group):
    """Micro-batching over sections (sphops.fused.plan_microbatches: groups of consecutive batch elements run one after the
    other with their own pair lists; BASELINE configs[4] needs it) against the whole-batch node: every state and dL/dA0
    bit-identical, parameter gradients equal up to the summation order over groups; same fire draws (one torch.rand(N)
    per step).  Ragged sections, none of them a multiple of 4 or 8 (slices of the sorted arrays start anywhere)."""
    import nca
    from sphops import fused
    gen = torch.Generator().manual_seed(31)
    n_side = 110
    x = torch.nn.functional.pad(jittered_lattice(n_side, 2, gen), (0, 1))
    N = x.shape[0]
    sections = (3001, 2047, 777, 4099, N - 3001 - 2047 - 777 - 4099)
    h = 3.2 * 2 / n_side
    dims = math.ceil(2 / h)
    w = load_golden("weights_gecko" if use_alpha else "weights_zebra")
    params = [torch.from_numpy(w[k]) for k in ("W1", "b1", "W2", "b2")]
    if use_alpha:
        d2 = (x ** 2).sum(-1)
        A0 = (torch.clamp(1 - d2 / 0.6 ** 2, 0, 1) ** 3)[:, None].repeat(1, 16) + 0.05 * torch.rand(N, 16, generator=gen)
    else:
        A0 = torch.rand(N, 16, generator=gen)
    K = 3
    masks = torch.rand(K, N, generator=gen) <= 0.5
    nca.USE_FUSED_STEP = True
    m = nca.SPHNCA(16, 0.5, use_alpha=use_alpha, normalize_perception=1.0 / h).to(DEV)
    with torch.no_grad():
        for t, src in zip((m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias), params):
            t.copy_(src)
    xd = x.to(DEV)
    res = {}
    for mb in (None, per_group):
        m.zero_grad()
        grid, v = m.init_grid(xd, sections, h, dims)
        A = A0.to(DEV).requires_grad_(True)
        prev, fused.MICROBATCH = fused.MICROBATCH, None          # "whole batch" must really be the whole batch
        try:
            with forced_fire_masks(list(masks)):
                out = m.rollout(xd, v, A, h, grid, K, checkpoint=True, microbatch=mb)
        finally:
            fused.MICROBATCH = prev
        loss = (out[-1][:, :4] ** 2).mean() + 0.1 * (out[1][:, :4] ** 2).mean()
        loss.backward()
        res[mb] = ([out[k].detach().clone() for k in range(K + 1)], A.grad.clone(), [p.grad.clone() for p in m.parameters()])
        # inference through the same groups: only [first, last] are kept
        with torch.no_grad(), forced_fire_masks(list(masks)):
            o2 = m.rollout(xd, v, A0.to(DEV), h, grid, K, keep_all=False, microbatch=mb)
        assert len(o2) == 2 and torch.equal(o2[-1], res[mb][0][-1])
    for a, b in zip(res[None][0], res[per_group][0]):
        assert torch.equal(a, b)
    assert torch.equal(res[None][1], res[per_group][1])
    for a, b, nm in zip(res[None][2], res[per_group][2], ("dW1", "db1", "dW2", "db2")):
        assert_close(b, a, rtol=GRAD_RTOL, what=nm)


def test_microbatch_plan_follows_memory(monkeypatch):
    """plan_microbatches: whole batch while lists + work buffers fit the budget, else groups of consecutive sections that do."""
    import sphops
    from sphops import fused
    gen = torch.Generator().manual_seed(1)
    x = torch.nn.functional.pad(jittered_lattice(64, 2, gen), (0, 1)).to(DEV)
    sections = (1024,) * 4
    grid = sphops.initialize_hashgrid(x, sections, 0.1, 20)
    assert fused.plan_microbatches(grid, 8, False, x.device, 2, "auto") is None
    per = grid.PAIR_SLOTS_ESTIMATE * grid.PAIR_ENTRY_BYTES[False] + fused.MICROBATCH_WORK_BYTES

    class Props:
        total_memory = int(torch.cuda.memory_allocated(x.device) + 4096 * (64 * 9 + 8) + 2.5 * 1024 * per / fused.MICROBATCH_FRACTION)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda d: Props)
    assert fused.plan_microbatches(grid, 8, False, x.device, 2, "auto") == [(0, 2), (2, 4)]
    assert fused.plan_microbatches(grid, 8, False, x.device, 2, 3) == [(0, 3), (3, 4)]
    monkeypatch.setattr(fused, "MICROBATCH", None)
    assert fused.plan_microbatches(grid, 8, False, x.device, 2, None) is None


@pytest.mark.parametrize("use_alpha", [False, True], ids=["noalpha", "alpha"])
def test_microbatched_rollout_with_empty_and_tiny_sections(use_alpha):
    """Micro-batches whose groups hold an empty section, a 1-particle section and a section that is not a multiple of the list
    group size: forward states and gradients equal the whole-batch node (bitwise states / dL/dA0)."""
    import nca
    gen = torch.Generator().manual_seed(8)
    x = torch.nn.functional.pad(jittered_lattice(40, 2, gen), (0, 1))
    N = x.shape[0]
    sections = (0, 1, 613, 0, N - 614)
    h = 3.2 * 2 / 40
    dims = math.ceil(2 / h)
    w = load_golden("weights_gecko" if use_alpha else "weights_zebra")
    nca.USE_FUSED_STEP = True
    m = nca.SPHNCA(16, 0.5, use_alpha=use_alpha, normalize_perception=1.0 / h).to(DEV)
    with torch.no_grad():
        for t, k in zip((m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias), ("W1", "b1", "W2", "b2")):
            t.copy_(torch.from_numpy(w[k]))
    A0 = (0.3 + 0.7 * torch.rand(N, 16, generator=gen))
    masks = torch.rand(3, N, generator=gen) <= 0.5
    xd = x.to(DEV)
    res = []
    for mb in (None, 2):
        m.zero_grad()
        grid, v = m.init_grid(xd, sections, h, dims)
        A = A0.to(DEV).requires_grad_(True)
        with forced_fire_masks(list(masks)):
            out = m.rollout(xd, v, A, h, grid, 3, checkpoint=True, microbatch=mb)
        (out[-1][:, :4] ** 2).mean().backward()
        res.append(([out[k].detach().clone() for k in range(4)], A.grad.clone(), [p.grad.clone() for p in m.parameters()]))
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    assert torch.equal(res[0][1], res[1][1])
    for a, b, nm in zip(res[0][2], res[1][2], ("dW1", "db1", "dW2", "db2")):
        assert_close(b, a, rtol=GRAD_RTOL, what=nm)
