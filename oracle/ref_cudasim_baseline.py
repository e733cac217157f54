#!/usr/bin/env python
"""Baseline #1 of BASELINE.json's north star, "the reference's CPU path": the UNMODIFIED reference (`oracle/_ref/code`,
staged by `oracle/stage_reference.py`) under Numba's CUDA simulator (`NUMBA_ENABLE_CUDASIM=1`, harness shim
`tests/golden/ref_shim.py`), through its own `nca.sample_plane` -- forward + loss + backward + optimiser step -- timed on
the host.  BASELINE INFRASTRUCTURE ONLY (started by `bench.py` as a subprocess).

The simulator interprets every CUDA thread as a Python thread under the GIL: ONE effective core, ~5 ms per particle per
operator, so the sample is a few hundred particles of BASELINE configs[0]'s cloud and 1 NCA step.
One JSON line: {"ok": true, "value": particle-steps/s, "particles": N, "nca_steps": K, "seconds": t, "cores": 1}.
"""
import argparse
import json
import os
import sys
import time
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-side", type=int, default=16)
    ap.add_argument("--nca-steps", type=int, default=1)
    args = ap.parse_args()
    out = {"ok": False, "impl": "reference under NUMBA_ENABLE_CUDASIM=1 (oracle/_ref/code, unmodified)", "cores": 1}
    try:
        os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
        os.environ.setdefault("SPHNCA_REFERENCE_CODE", os.path.join(HERE, "_ref", "code"))
        if not os.path.isdir(os.path.join(os.environ["SPHNCA_REFERENCE_CODE"], "sphops")):
            raise RuntimeError("oracle/_ref/code is not staged (run oracle/stage_reference.py in the build container)")
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import numpy as np
        import torch
        import ref_shim
        sphops, nca = ref_shim.load_reference()
        n = args.n_side
        gen = torch.Generator().manual_seed(0)
        s = 2.0 / n
        ax = (torch.arange(n, dtype=torch.float32) + 0.5) * s - 1.0
        x = torch.stack(torch.meshgrid(ax, ax, indexing="ij"), dim=-1).reshape(-1, 2)
        x = x + (torch.rand(x.shape, generator=gen) - 0.5) * s
        A = torch.rand(x.shape[0], 16, generator=gen)
        h = 3.2 * s
        dims = max(int(np.ceil(2.0 / h)), 3)
        w = np.load(os.path.join(ROOT, "tests", "golden", "weights_zebra.npz"))
        nca.USE_3D = True
        model = nca.SPHNCA(16, 0.5, update_rule="gated", normalize_grad=True, use_alpha=False, normalize_perception=1.0 / h)
        with torch.no_grad():
            for t, k in zip((model.model[0].weight, model.model[0].bias, model.model[2].weight, model.model[2].bias),
                            ("W1", "b1", "W2", "b2")):
                t.copy_(torch.from_numpy(w[k]))
        opt = torch.optim.Adam(model.parameters(), lr=1e-5)
        loss_fn = lambda x_, A_, img, A0=None: (A_[:, :4] ** 2).mean()      # noqa: E731  (bench.py's loss)
        t0 = time.perf_counter()
        nca.sample_plane(model, x, A, (x.shape[0],), h, dims, iter_n=args.nca_steps, loss_fn=loss_fn, img=None, optimizer=opt)
        dt = time.perf_counter() - t0
        out.update(ok=True, value=x.shape[0] * args.nca_steps / dt, unit="particle-steps/s", particles=int(x.shape[0]),
                   nca_steps=args.nca_steps, seconds=dt,
                   sample=f"{x.shape[0]} jittered 2-D particles (configs[0] spacing rule, h = 3.2 spacings) x {args.nca_steps} NCA "
                          "step(s), grid build + volume + fwd + loss + bwd + Adam through the reference's nca.sample_plane")
    except BaseException as e:  # noqa: BLE001
        out["error"] = f"{type(e).__name__}: {e}"
        out["traceback"] = traceback.format_exc()[-2000:]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
