"""Gradient parity of the DEFAULT product path at BASELINE.json sizes, without a ReLU-kink allowance.

The rollout node (one autograd node, fire lists, fired-rows perception, tcgen05 MLP / data-gradient / weight-gradient
kernels, fire-aware adjoint) runs one element of configs[1] (65,536 padded-2-D particles) and configs[2] (250,047 3-D
particles); the GPU's saved ReLU decisions are read back (`sphops.fused.TAPE_PROBE`) and the fp64 oracle is evaluated
WITH THOSE DECISIONS IMPOSED (`oracle.nca_step(relu_bits=...)`) ALONG THE GPU's TRAJECTORY (`trajectory=`: step k is
evaluated at the GPU's state k, so every step is compared from identical inputs -- the parity definition of SURVEY.md 8c --
and the backward is the chain of the oracle's own fp64 Jacobians at those states).  What is compared is then arithmetic only:

* the stages of every step on their own: perception rows <= 1e-6 of max|perception| (measured 3-4e-7), MLP output rows
  <= 1.5e-6 of max|output| (measured 0.7-1.1e-6 = the 3xTF32 split; torch fp32 / cuBLAS sgemm on the same rows, printed next
  to it: 3.3-4.4e-7); every single-step state <= 1e-5 of max|state| (measured <= 8.2e-6 with the no-alpha texture model on
  random states, where the gated update sees |output| up to 215, <= 2.7e-6 with the alpha model, <= 4.4e-6 in 3-D); loss
  <= 1e-5; dL/dA0 and all four parameter gradients <= 2e-5 (max-norm relative to the tensor's maximum, no row-fraction
  allowance);
* printed for the record, not asserted: the error against the FREE-RUNNING fp64 oracle, next to the divergence of that
  oracle from ITSELF when A0 is perturbed by one fp32 rounding -- the dynamics amplify any perturbation by 2-5x per step
  at these sizes, which no fp32 implementation (the reference's included) can escape;
* the units whose GPU decision differs from the oracle's own are counted, and each must be ill-conditioned:
  |pre| / (sum_k |y_k W1_nk| + |b1_n|) < 2e-6, i.e. inside the rounding of ANY fp32 evaluation (the reference's cuBLAS
  sgemm included) -- that separates "a kink fell the other way" from "the kernel is wrong".

Reference: code/nca.py:87-117 (step), sphops/ops/func.py:145-200 (gradient backward), SURVEY.md 8c.
"""
import math

import numpy as np
import pytest
import torch

from tests.util import jittered_lattice, load_golden, forced_fire_masks, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
STATE_RTOL = 1e-5         # single step, relative to max|state|: the fp32 bar of BASELINE.json's north star
GRAD_RTOL = 2e-5
KINK_MARGIN = 2e-6
PERCEPTION_RTOL = 1e-6    # perception rows relative to max|perception|   (measured 3-4e-7)
MLP_OUT_RTOL = 1.5e-6     # MLP output rows relative to max|output|       (measured 0.7-1.1e-6: the 3xTF32 split; cuBLAS
                          # sgemm on the same rows: 3.3-4.4e-7)


def unpack_bits(words):
    """int32 [n, 8] -> bool [n, 256]; bit b of word w = hidden unit 32 w + b (csrc/step.cu:109-116)."""
    w = words.to(torch.int64) & 0xFFFFFFFF
    sh = torch.arange(32, device=words.device, dtype=torch.int64)
    return ((w[:, :, None] >> sh) & 1).bool().reshape(words.shape[0], 256)


def gpu_relu_decisions(probe, grid):
    """[(original particle indices of the fired rows, their ReLU decisions bool[n,256])] per step, on the CPU."""
    sid = grid.sorted_id.long()
    out = []
    for rows, count, relu_c, Y_c, O_c in probe:
        n = int(count.item())
        out.append((sid[rows[:n].long()].cpu(), unpack_bits(relu_c[:n]).cpu(), Y_c[:n].cpu(), O_c[:n].cpu()))
    return out


def run_product(x, sections, h, dims, A0, params, masks, use_alpha, norm_perc, loss_states=()):
    import nca
    from sphops import fused
    nca.USE_FUSED_STEP = True
    m = nca.SPHNCA(16, 0.5, use_alpha=use_alpha, normalize_perception=norm_perc).to(DEV)
    with torch.no_grad():
        for t, src in zip((m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias), params):
            t.copy_(src)
    xd = x.to(DEV)
    grid, v = m.init_grid(xd, sections, h, dims)
    A = A0.to(DEV).requires_grad_(True)
    fused.TAPE_PROBE = []
    try:
        with forced_fire_masks(list(masks)):
            out = m.rollout(xd, v, A, h, grid, masks.shape[0], checkpoint=False)
        probe = fused.TAPE_PROBE
    finally:
        fused.TAPE_PROBE = None
    assert out is not None and len(probe) == masks.shape[0], "the fire-list rollout node did not run"
    loss = (out[-1][:, :4] ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    states = [out[k].detach().cpu() for k in range(len(out))]
    grads = [A.grad.cpu()] + [p.grad.cpu() for p in (m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias)]
    return states, float(loss.detach()), grads, gpu_relu_decisions(probe, grid), v.cpu()


def compare_with_imposed_bits(x, sections, h, dims, A0, params, masks, cfg, tag):
    from oracle import sphnca_oracle as O
    states, loss, grads, bits, v_gpu = run_product(x, sections, h, dims, A0, params, masks, cfg.use_alpha, cfg.normalize_perception)
    stage = bits
    bits = [(rows, b) for rows, b, _, _ in stage]
    for k, (rows, _) in enumerate(bits):       # the list IS the set of fired particles, in every step
        fired = torch.zeros(x.shape[0], dtype=torch.bool); fired[rows] = True
        assert torch.equal(fired, masks[k]), f"step {k}: fire list differs from the fire mask"
    prepared = O.prepare(x, sections, h, dims)
    assert rel_err(v_gpu, prepared[2]) <= STATE_RTOL
    # The oracle along the GPU's trajectory, with the GPU's ReLU decisions: every step from identical inputs,
    # the backward = the chain of the oracle's own fp64 Jacobians at those states.
    A_ref, loss_ref, dA0_ref, dP_ref, st_ref, aux = O.rollout_fwd_bwd(x, sections, h, dims, A0, params, masks, cfg, relu_bits=bits,
                                                                     prepared=prepared, return_all=True, trajectory=states)
    # 1. how many ReLU decisions differ from the oracle's own, and how ill-conditioned those units are
    n_diff, worst_margin, n_units = 0, 0.0, 0
    for (rows, b), a in zip(bits, aux):
        nat = a["pre"][rows] > 0
        d = nat != b
        n_units += d.numel()
        n_diff += int(d.sum())
        if d.any():
            worst_margin = max(worst_margin, float(a["margin"][rows][d].max()))
    # 2. arithmetic parity
    errs = {f"S{k}": rel_err(states[k], st_ref[k]) for k in range(1, len(states))}
    errs["loss"] = abs(loss - float(loss_ref)) / abs(float(loss_ref))
    for nm, g, r in zip(("dA0", "dW1", "db1", "dW2", "db2"), grads, [dA0_ref] + list(dP_ref)):
        errs[nm] = rel_err(g, r)
    # 3. for the record (not asserted): the free-running fp64 oracle (its own trajectory, its own ReLU decisions), and how fast
    #    the dynamics themselves amplify a perturbation of one fp32 rounding of A0 (fp64 vs fp64)
    _, pairs, v = prepared
    with torch.no_grad():
        free = O.rollout(pairs, v, A0, params, masks, cfg)
        bump = O.rollout(pairs, v, A0.double() * (1.0 + 6e-8 * torch.sign(torch.randn(A0.shape, generator=torch.Generator().manual_seed(1)))),
                         params, masks, cfg)
    # the two stages of every step on their own (same inputs: the GPU's state k): perception rows, MLP output rows
    n_out = params[2].shape[0]
    perc_err = [rel_err(Y, a["y"][rows][:, 16:]) for (rows, _, Y, _), a in zip(stage, aux)]
    mlp_err = [rel_err(Oc[:, :n_out], a["dA"][rows]) for (rows, _, _, Oc), a in zip(stage, aux)]
    mlp_mag = [float(a["dA"][rows].detach().abs().max()) for (rows, _, _, _), a in zip(stage, aux)]
    # for scale: the same MLP rows evaluated the reference's way -- torch fp32 on the GPU (cuBLAS sgemm, TF32 off) -- against the oracle
    prev_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        Wd = [t.to(DEV) for t in params]
        sgemm_err = []
        for (rows, _, _, _), a in zip(stage, aux):
            y32 = a["y"][rows].to(torch.float32).to(DEV)
            o32 = torch.relu(y32 @ Wd[0].t() + Wd[1]) @ Wd[2].t() + Wd[3]
            sgemm_err.append(rel_err(o32.cpu(), a["dA"][rows].detach()))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev_tf32
    free_err = [rel_err(states[k], free[k]) for k in range(1, len(states))]
    growth = [rel_err(bump[k], free[k]) for k in range(1, len(states))]
    print(f"\n[{tag}] N={x.shape[0]} K={masks.shape[0]}: ReLU decisions differing from the fp64 oracle's own: {n_diff} of {n_units} "
          f"({n_diff / max(n_units, 1):.2e}), worst |pre|/sum|terms| among them {worst_margin:.2e}\n    per-step / gradient rel errs (oracle "
          "along the GPU trajectory): " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items())
          + "\n    stages per step: perception " + ", ".join(f"{e:.2e}" for e in perc_err) + "; MLP output " + ", ".join(f"{e:.2e}" for e in mlp_err)
          + " (max |output| " + ", ".join(f"{e:.1f}" for e in mlp_mag) + "); torch fp32 / cuBLAS sgemm on the same rows: "
          + ", ".join(f"{e:.2e}" for e in sgemm_err)
          + "\n    free-running fp64 oracle, states: " + ", ".join(f"{e:.2e}" for e in free_err)
          + "\n    fp64 oracle vs itself from A0*(1 +- 6e-8) (amplification by the dynamics): " + ", ".join(f"{e:.2e}" for e in growth))
    assert worst_margin < KINK_MARGIN, f"a well-conditioned ReLU unit was decided differently (margin {worst_margin:.2e})"
    assert max(perc_err) <= PERCEPTION_RTOL and max(mlp_err) <= MLP_OUT_RTOL, (perc_err, mlp_err)
    for k, e in errs.items():
        bar = 1e-5 if k == "loss" else (STATE_RTOL if k.startswith("S") else GRAD_RTOL)
        assert e <= bar, f"{tag}: {k} rel err {e:.3e} > {bar:g}   (all: {errs})"
    return errs


def _weights(name):
    w = load_golden(name)
    return [torch.from_numpy(w[k]) for k in ("W1", "b1", "W2", "b2")]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("use_alpha", [False, True], ids=["zebra_noalpha", "gecko_alpha"])
def test_configs1_element_gradients_with_gpu_relu_bits(use_alpha):
    """One batch element of BASELINE configs[1]: 256^2 = 65,536 jittered particles, zero-padded to 3-D, h = 3.2 spacings,
    normalised perception, K = 4 steps fwd+bwd; the no-alpha texture model (bench.py's) and the alpha-channel model."""
    from oracle import sphnca_oracle as O
    gen = torch.Generator().manual_seed(3)
    n = 256
    x = torch.nn.functional.pad(jittered_lattice(n, 2, gen), (0, 1))
    h = 3.2 * 2 / n
    dims = math.ceil(2 / h)
    N = x.shape[0]
    if use_alpha:
        d2 = (x ** 2).sum(-1)
        A0 = (torch.clamp(1 - d2 / 0.5 ** 2, 0, 1) ** 3)[:, None].repeat(1, 16) + 0.05 * torch.rand(N, 16, generator=gen)
    else:
        A0 = torch.rand(N, 16, generator=gen)
    masks = torch.rand(4, N, generator=gen) <= 0.5
    cfg = O.StepConfig(use_alpha=use_alpha, normalize_perception=1.0 / h)
    compare_with_imposed_bits(x, (N,), h, dims, A0, _weights("weights_gecko" if use_alpha else "weights_zebra"), masks, cfg,
                              "configs[1] element, " + ("alpha" if use_alpha else "no alpha"))


@pytest.mark.timeout(1800)
def test_configs2_gradients_with_gpu_relu_bits():
    """BASELINE configs[2]: 63^3 = 250,047 jittered 3-D particles, h = 1.97 spacings, K = 2 steps fwd+bwd."""
    from oracle import sphnca_oracle as O
    gen = torch.Generator().manual_seed(5)
    n = 63
    x = jittered_lattice(n, 3, gen)
    h = 1.97 * 2 / n
    dims = math.ceil(2 / h)
    N = x.shape[0]
    A0 = torch.rand(N, 16, generator=gen)
    masks = torch.rand(2, N, generator=gen) <= 0.5
    cfg = O.StepConfig(use_alpha=False, normalize_perception=1.0 / h)
    compare_with_imposed_bits(x, (N,), h, dims, A0, _weights("weights_zebra"), masks, cfg, "configs[2]")


@pytest.mark.timeout(600)
def test_weight_gradient_accumulation_262144_rows_vs_fp64():
    """The weight-gradient kernel accumulates over ROWS in a truncating fp32 TMEM accumulator: at the
    262,144 fired rows of a configs[1] launch, against an fp64 evaluation of the same formulas from the same inputs.
    The kernel recomputes the hidden layer and takes its own ReLU decisions, so rows that hold an ill-conditioned unit
    (|pre| / sum|terms| < 1e-5: 0.3 % of the rows) get a zero output gradient -- what is measured is the accumulation over
    rows, not kink flips (a first version of this test used zero-mean random gradients, whose sums cancel to 1/512 of
    sum|terms|: ONE flipped unit then shows up as 2e-3 of the result)."""
    import ctypes as C
    from sphops import _lib, fused
    R = 262144
    gen = torch.Generator().manual_seed(9)
    W = [t.to(DEV) for t in _weights("weights_zebra")]
    A_c = torch.rand(R, 16, generator=gen).to(DEV)
    Y_c = (0.3 * torch.randn(R, 32, generator=gen)).to(DEV)
    W1d, b1d = W[0].double().cpu(), W[1].double().cpu()
    Yf_cpu = torch.cat([A_c.cpu(), Y_c.cpu()], 1).double()
    pre_cpu = Yf_cpu @ W1d.t() + b1d
    cond = pre_cpu.abs() / (Yf_cpu.abs() @ W1d.abs().t() + b1d.abs())
    keep = (cond >= 1e-5).all(1)
    dO = torch.zeros(R, 36)
    dO[:, :33] = 1e-6 * (0.25 + torch.rand(R, 33, generator=gen))     # gradient magnitudes of a mean loss over 2^18 rows
    dO[~keep] = 0.0
    print(f"\nrows with an ill-conditioned hidden unit (zero gradient): {int((~keep).sum())} of {R}")
    dO = dO.to(DEV)
    count = torch.tensor([R], dtype=torch.int32, device=DEV)
    p = fused._step_params(*W, "gated", False, 1.0, 1.0)
    nb = C.c_size_t(0)
    _lib.check(_lib.lib.sphnca_mlp_wgrad_workspace_bytes(R, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=DEV)
    dW = [torch.zeros_like(t) for t in W]
    _lib.check(_lib.lib.sphnca_mlp_wgrad_list(R, _lib.ptr(count), _lib.ptr(A_c), _lib.ptr(Y_c), _lib.ptr(dO), C.byref(p),
                                              *[_lib.ptr(t) for t in dW], _lib.ptr(ws), nb.value, _lib.stream_ptr()))
    torch.cuda.synchronize()
    Yf = torch.cat([A_c, Y_c], 1).double()
    W1, b1, W2, b2 = [t.double() for t in W]
    pre = Yf @ W1.t() + b1
    dOd = dO[:, :33].double()
    dH = (dOd @ W2) * (pre > 0)
    ref = dict(dW1=dH.t() @ Yf, db1=dH.sum(0), dW2=dOd.t() @ torch.relu(pre), db2=dOd.sum(0))
    errs = {nm: rel_err(t, ref[nm]) for t, nm in zip(dW, ("dW1", "db1", "dW2", "db2"))}
    print("\nweight gradient over 262,144 rows vs fp64:", {k: f"{v:.2e}" for k, v in errs.items()})
    for nm, e in errs.items():
        assert e <= GRAD_RTOL, f"{nm}: {e:.3e} (all {errs})"


def test_weight_gradient_partials_over_a_series_of_calls_vs_fp64():
    """sphnca_mlp_wgrad_list_partial x 3 (accumulate = 0, 1, 1: the steps of a rollout's backward, different row counts) +
    ONE sphnca_mlp_wgrad_reduce == the fp64 sum of the three weight gradients, and == three sphnca_mlp_wgrad_list calls up to
    fp32 summation order; a stream without rows in one of the calls keeps what the others gave it."""
    import ctypes as C
    from sphops import _lib, fused
    gen = torch.Generator().manual_seed(7)
    W = [t.to(DEV) for t in _weights("weights_zebra")]
    R = 40_000
    p = fused._step_params(*W, "gated", False, 1.0, 1.0)
    nb = C.c_size_t(0)
    _lib.check(_lib.lib.sphnca_mlp_wgrad_workspace_bytes(R, C.byref(nb)))
    ws = torch.full((nb.value,), 0xA5, dtype=torch.uint8, device=DEV)      # garbage: the first call of a series must not read it
    ws2 = torch.empty(nb.value, dtype=torch.uint8, device=DEV)
    dW = [torch.zeros_like(t) for t in W]
    dW_each = [torch.zeros_like(t) for t in W]
    ref = None
    W1, b1, W2, b2 = [t.double() for t in W]
    for call, rows in enumerate((R, 37, 25_001)):                          # 37 rows: one 64-row chunk, every other stream idle
        A_c = torch.rand(R, 16, generator=gen).to(DEV)
        Y_c = (0.3 * torch.randn(R, 32, generator=gen)).to(DEV)
        # same-sign gradients, zero on rows that hold an ill-conditioned hidden unit (see the test above: kink flips are not the subject)
        Yc = torch.cat([A_c.cpu(), Y_c.cpu()], 1).double()
        keep = ((Yc @ W1.cpu().t() + b1.cpu()).abs() / (Yc.abs() @ W1.cpu().abs().t() + b1.cpu().abs()) >= 1e-5).all(1)
        dO = torch.zeros(R, 36)
        dO[:, :33] = 1e-3 * (0.25 + torch.rand(R, 33, generator=gen))
        dO[~keep] = 0.0
        dO = dO.to(DEV)
        count = torch.tensor([rows], dtype=torch.int32, device=DEV)
        _lib.check(_lib.lib.sphnca_mlp_wgrad_list_partial(R, _lib.ptr(count), _lib.ptr(A_c), _lib.ptr(Y_c), _lib.ptr(dO), C.byref(p),
                                                          1 if call else 0, _lib.ptr(ws), nb.value, _lib.stream_ptr()))
        _lib.check(_lib.lib.sphnca_mlp_wgrad_list(R, _lib.ptr(count), _lib.ptr(A_c), _lib.ptr(Y_c), _lib.ptr(dO), C.byref(p),
                                                  *[_lib.ptr(t) for t in dW_each], _lib.ptr(ws2), nb.value, _lib.stream_ptr()))
        Yf = torch.cat([A_c[:rows], Y_c[:rows]], 1).double()
        pre = Yf @ W1.t() + b1
        dOd = dO[:rows, :33].double()
        dH = (dOd @ W2) * (pre > 0)
        part = [dH.t() @ Yf, dH.sum(0), dOd.t() @ torch.relu(pre), dOd.sum(0)]
        ref = part if ref is None else [a + b for a, b in zip(ref, part)]
    _lib.check(_lib.lib.sphnca_mlp_wgrad_reduce(R, C.byref(p), *[_lib.ptr(t) for t in dW], _lib.ptr(ws), nb.value, _lib.stream_ptr()))
    torch.cuda.synchronize()
    for t, e, r, nm in zip(dW, dW_each, ref, ("dW1", "db1", "dW2", "db2")):
        assert rel_err(t, r) <= GRAD_RTOL, f"{nm}: series vs fp64 {rel_err(t, r):.3e}"
        assert rel_err(t, e.double()) <= 2e-6, f"{nm}: series vs per-call reduction {rel_err(t, e.double()):.3e}"
