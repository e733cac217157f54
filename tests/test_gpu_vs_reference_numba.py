"""Parity against the REFERENCE ITSELF on the GPU box, at BASELINE size: the unmodified reference (staged `oracle/_ref/code`: its
Numba-CUDA sphops kernels + eager torch, fp32) runs one element of configs[1] -- 65,536 padded-2-D particles -- in a subprocess
(`oracle/ref_gpu_baseline.py --dump`; also configs[2], 250,047 3-D particles), and this repo's default path is compared with it on
the SAME inputs and fire masks
(BASELINE.json north_star: "correctness is checked against the reference's own sphops path on the same inputs").

* volumes <= 1e-5; every step compared FROM THE REFERENCE'S OWN STATE (the dynamics amplify one fp32 rounding 2-5x per step, so
  two fp32 implementations can only be compared step by step): states <= 1e-5 of max|state|;
* gradients of a single step (no chaos): the four parameter gradients <= 1e-4; dL/dA0 <= 2e-5 on all rows but the few (<= 0.05 %)
  whose ReLU decisions fall differently in cuBLAS fp32 and in the 3xTF32 kernels (counted; tests/test_gpu_parity_config_sizes.py
  pins the arithmetic at 2e-5 against the oracle with the decisions imposed).
Skipped when the reference is not staged or numba cannot drive this GPU."""
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.util import forced_fire_masks, load_golden, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "code")
DEV = "cuda"


def reference_dump(tmp_path, K, workload):
    out = tmp_path / f"ref_k{K}.npz"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_gpu_baseline.py"), "--workload", workload, "--batch", "1",
                        "--nca-steps", str(K), "--steps", "1", "--warmup", "0", "--dump", str(out)], capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    info = json.loads(line[-1]) if line else {"ok": False, "error": r.stderr[-500:]}
    if not info.get("ok"):
        pytest.skip(f"the reference's Numba-CUDA path did not run here: {info.get('error')}")
    return dict(np.load(out))


def build_model(h):
    import nca
    w = load_golden("weights_zebra")
    m = nca.SPHNCA(16, 0.5, use_alpha=False, normalize_perception=1.0 / h, normalize_grad=True).to(DEV)
    with torch.no_grad():
        for t, k in zip((m.model[0].weight, m.model[0].bias, m.model[2].weight, m.model[2].bias), ("W1", "b1", "W2", "b2")):
            t.copy_(torch.from_numpy(w[k]))
    return m


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "sphops")), reason="oracle/_ref/code is not staged")
@pytest.mark.timeout(1800)
@pytest.mark.parametrize("workload,n_particles", [("train2d_b8x65k_k64", 65536), ("train3d_250k_k96", 250047)],
                         ids=["configs1_element_65k_2d", "configs2_250k_3d"])
def test_states_and_gradients_against_the_reference_numba_path_at_config_size(tmp_path, workload, n_particles):
    import nca
    nca.USE_FUSED_STEP = True
    # ---- states, K = 3, every step from the reference's own state
    g = reference_dump(tmp_path, 3, workload)
    h, dims = float(g["h"]), int(g["dims"])
    x = torch.from_numpy(g["x"]).to(DEV)
    N = x.shape[0]
    assert N == n_particles
    m = build_model(h)
    grid, v = m.init_grid(x, (N,), h, dims)
    assert rel_err(v, g["v"]) <= 1e-5
    prev = torch.from_numpy(g["A0"]).to(DEV)
    errs = []
    with torch.no_grad():
        for k in range(3):
            with forced_fire_masks([torch.from_numpy(g["masks"][k])]):
                out = m.rollout(x, v, prev, h, grid, 1, keep_all=False)
            errs.append(rel_err(out[-1], g["states"][k]))
            prev = torch.from_numpy(g["states"][k]).to(DEV)
    # where a difference sits: the fp64 oracle on the same inputs as the third party
    from oracle import sphnca_oracle as O
    xc, A0c = torch.from_numpy(g["x"]), torch.from_numpy(g["A0"])
    w = load_golden("weights_zebra")
    params = [torch.from_numpy(w[k]) for k in ("W1", "b1", "W2", "b2")]
    og, pairs, ov = O.prepare(xc, (N,), h, dims)
    with torch.no_grad():
        o1 = O.nca_step(pairs, ov, A0c, params, torch.from_numpy(g["masks"][0]), O.StepConfig(use_alpha=False, normalize_perception=1.0 / h))
        with forced_fire_masks([torch.from_numpy(g["masks"][0])]):
            mine1 = m.rollout(x, v, torch.from_numpy(g["A0"]).to(DEV), h, grid, 1, keep_all=False)[-1].cpu()
    with torch.no_grad():
        oga = O.gradient(pairs, ov, A0c)                                # [N,16,3] fp64
        oy = O.perceive(pairs, ov, A0c, O.StepConfig(use_alpha=False, normalize_perception=1.0 / h))
        W1, b1, W2, b2 = [t.double() for t in params]
        omlp = torch.relu(oy @ W1.T + b1) @ W2.T + b2
        rmlp_on_ref_y = torch.relu(torch.from_numpy(g["y0"]).double() @ W1.T + b1) @ W2.T + b2
    print(f"\nreference stages vs the oracle: sphops.gradient {rel_err(g['gA0'], oga):.2e}; perception rows y {rel_err(g['y0'], oy):.2e}; "
          f"MLP output {rel_err(g['mlp0'], omlp):.2e}; its torch MLP vs fp64 on ITS OWN y {rel_err(g['mlp0'], rmlp_on_ref_y):.2e}")
    ref1 = torch.from_numpy(g["states"][0])
    d_ref = (ref1.double() - o1).abs().max(-1).values / o1.abs().max()
    d_mine = (mine1.double() - o1).abs().max(-1).values / o1.abs().max()
    print(f"\nstep 1 vs the fp64 oracle: this repo {float(d_mine.max()):.2e}; reference Numba-CUDA {float(d_ref.max()):.2e} "
          f"(rows above 1e-5: {int((d_ref > 1e-5).sum())} of {N}, above 1e-3: {int((d_ref > 1e-3).sum())}; fired rows among them: "
          f"{int((torch.from_numpy(g['masks'][0]) & (d_ref > 1e-5)).sum())})")
    # ---- gradients of one step
    g1 = reference_dump(tmp_path, 1, workload)
    m.zero_grad()
    A0 = torch.from_numpy(g1["A0"]).to(DEV).requires_grad_(True)
    with forced_fire_masks([torch.from_numpy(g1["masks"][0])]):
        out = m.rollout(x, v, A0, h, grid, 1)
    loss = (out[-1][:, :4] ** 2).mean()
    loss.backward()
    P = dict(m.named_parameters())
    gerr = {nm: rel_err(P[pn].grad, g1[nm]) for nm, pn in (("dW1", "model.0.weight"), ("db1", "model.0.bias"), ("dW2", "model.2.weight"), ("db2", "model.2.bias"))}
    ref = g1["dA0"].astype(np.float64)
    row_err = np.abs(A0.grad.cpu().numpy().astype(np.float64) - ref).max(-1) / np.abs(ref).max()
    off = int((row_err > 2e-5).sum())
    print(f"\\n[vs reference Numba-CUDA, {N} particles] per-step states {', '.join(f'{e:.2e}' for e in errs)}; loss "
          f"{abs(float(loss) - float(g1['loss'])) / abs(float(g1['loss'])):.2e}; parameter gradients "
          + ", ".join(f"{k} {e:.2e}" for k, e in gerr.items()) + f"; dL/dA0 rows above 2e-5: {off} of {N} (worst {row_err.max():.2e})")
    assert float(d_mine.max()) <= 1e-5
    assert max(errs) <= 1e-5, errs
    assert abs(float(loss) - float(g1["loss"])) <= 1e-5 * abs(float(g1["loss"]))
    # Gradients against ANOTHER fp32 evaluation carry its ReLU decisions: a pre-activation within rounding of 0 falls one way in
    # cuBLAS sgemm and the other way in the 3xTF32 kernels (measured: 0 rows of 65,536 in 2-D, 55 of 250,047 in 3-D), which moves
    # that row of dL/dA0 by ~1e-3 and one term of the weight-gradient sums.  Bars: parameter gradients 1e-4, dL/dA0 2e-5 on all
    # but 0.05 % of the rows; the arithmetic itself is pinned at 2e-5 against the oracle with the decisions imposed
    # (tests/test_gpu_parity_config_sizes.py).
    for k, e in gerr.items():
        assert e <= 1e-4, gerr
    assert off <= 5e-4 * N and row_err.max() <= 1e-2, (off, row_err.max())
