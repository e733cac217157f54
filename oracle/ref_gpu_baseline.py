#!/usr/bin/env python
"""Baseline #2 of BASELINE.json's north star: the REFERENCE's Numba-CUDA kernels + eager torch, through its
unchanged `nca.sample_plane`, timed on one B200.  BASELINE INFRASTRUCTURE ONLY (started by `bench.py`, as a
subprocess so that Numba's CUDA context and the JIT never share a process with the product).

Runs the staged copy `oracle/_ref/code` (made by `oracle/stage_reference.py` in the build container; the GPU box has no
`/root/reference`).  Harness-only accommodation, no reference file touched: `import numba.cuda.libdevice` BEFORE
`import sphops` -- numba 0.65 (image) exposes `cuda.libdevice` only after that import, and reference
`sphops/ops/kernels_impl.py:7` dereferences it at import time (the reference pins numba 0.61.2 + numba-cuda 0.16.0).

Same synthetic inputs, weights, loss and optimiser as `bench.py` (imported from it).  One JSON line on stdout:
    {"ok": true, "value": particle-NCA-steps/s, "ms_per_step": ..., "nca_steps": K, "particles": N, ...}
or  {"ok": false, "error": "<text>"}.
The reference keeps ~3-4 KB of autograd residue per particle-step, so K is bounded (`--nca-steps`); throughput per
particle-step does not depend on K beyond the once-per-rollout grid build (reported separately as `grid_ms`).
"""
import argparse
import json
import os
import sys
import time
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref", "code")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="train2d_b8x65k_k64")
    ap.add_argument("--nca-steps", type=int, default=8)
    ap.add_argument("--batch", type=int, default=0, help="override the number of clouds (0 = the workload's)")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--use-alpha", choices=["on", "off"], default="off")
    ap.add_argument("--dump", default="", help="write x, A0, fire masks, states and gradients of ONE rollout to this .npz "
                                               "(GPU oracle for the parity tests)")
    args = ap.parse_args()
    out = {"ok": False, "impl": "reference numba-cuda + eager torch (oracle/_ref/code, unmodified)"}
    try:
        if not os.path.isdir(os.path.join(REF, "sphops")):
            raise RuntimeError(f"{REF} is not staged (run oracle/stage_reference.py in the build container)")
        sys.path.insert(0, ROOT)
        import bench                                   # make_inputs, WORKLOADS, weights, loss (no product import at module level)
        for m in ("sphops", "nca"):
            assert m not in sys.modules
        # the reference's modules must win over this repo's drop-in package of the same names
        sys.path = [REF] + [p for p in sys.path if not p.rstrip("/").endswith("sph-nca_b200")]
        import torch
        import numba
        import numba.cuda.libdevice                    # noqa: F401  (see the module docstring)
        from numba import cuda
        if not hasattr(cuda, "libdevice"):
            cuda.libdevice = numba.cuda.libdevice
        out["numba"] = numba.__version__
        out["torch"] = torch.__version__
        import sphops                                   # the reference's
        import nca                                      # the reference's
        assert os.path.realpath(sphops.__file__).startswith(os.path.realpath(REF)), sphops.__file__
        assert os.path.realpath(nca.__file__).startswith(os.path.realpath(REF)), nca.__file__

        wl = dict(bench.WORKLOADS[args.workload])
        if args.batch > 0:
            wl["batch"] = args.batch
        K = args.nca_steps
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(0)
        torch.manual_seed(1234)
        x_cpu, A_cpu, sections, h, dims = bench.make_inputs(wl, seed=0)
        N = x_cpu.shape[0]
        use_alpha = args.use_alpha == "on"
        nca.USE_3D = True
        model = nca.SPHNCA(16, 0.5, update_rule="gated", normalize_grad=True, use_alpha=use_alpha,
                           normalize_perception=1.0 / h).to(dev)
        with torch.no_grad():
            for t, src in zip((model.model[0].weight, model.model[0].bias, model.model[2].weight, model.model[2].bias),
                              bench.load_weights("gecko" if use_alpha else "zebra")):
                t.copy_(src)
        opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=0)
        x_dev, A_dev = x_cpu.to(dev), A_cpu.to(dev)
        dims_arg = dims                                   # an int: every axis, the zero-padded z axis included (like bench.py)

        def one_step():
            return nca.sample_plane(model, x_dev, A_dev, sections, h, dims_arg, iter_n=K, loss_fn=bench.bench_loss, img=None,
                                    optimizer=opt)

        t0 = time.perf_counter()
        one_step()                                      # JIT compilation of every kernel happens here
        torch.cuda.synchronize()
        out["jit_first_call_s"] = time.perf_counter() - t0
        for _ in range(args.warmup):
            one_step()
        torch.cuda.synchronize()
        # the once-per-rollout part on its own
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xp = torch.nn.functional.pad(x_dev, (0, 1)) if x_dev.shape[1] == 2 else x_dev
        e0.record()
        model.init_grid(xp, sections, h, dims_arg)
        e1.record()
        torch.cuda.synchronize()
        out["grid_ms"] = e0.elapsed_time(e1)
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            one_step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out.update(ok=True, value=N * K / (ms * 1e-3), unit="particle-steps/s", ms_per_step=ms, nca_steps=K, particles=N,
                   workload=args.workload, use_alpha=use_alpha, steps=args.steps, warmup=args.warmup,
                   peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9,
                   gpu=torch.cuda.get_device_name(0), cc=list(torch.cuda.get_device_capability(0)))

        if args.dump:
            # one more rollout with recorded fire draws: states + gradients of the reference's own GPU path (fp32)
            import numpy as np
            draws = []
            orig_rand = torch.rand

            def rec(*a, **k):
                u = orig_rand(*a, **k)
                draws.append(u.detach().cpu())
                return u
            model.zero_grad(set_to_none=True)
            with torch.no_grad():       # the timed iterations above were TRAINING iterations (Adam updated the weights): start over
                for t, src in zip((model.model[0].weight, model.model[0].bias, model.model[2].weight, model.model[2].bias),
                                  bench.load_weights("gecko" if use_alpha else "zebra")):
                    t.copy_(src)
            A0 = A_dev.clone().requires_grad_(True)
            torch.rand = rec
            try:
                grid, v = model.init_grid(xp, sections, h, dims_arg)
                A, states = A0, []
                for _ in range(K):
                    _, A = model(xp, v, A, h, grid)
                    states.append(A.detach().cpu().numpy())
            finally:
                torch.rand = orig_rand
            loss = bench.bench_loss(xp, A, None)
            loss.backward()
            with torch.no_grad():       # the two stages of the first step on their own: the sphops gradient op, the torch MLP on top of it
                gA0 = sphops.gradient(xp, v, A_dev, h, grid)
                y0 = model.base_feature_process(model.perceive(xp, v, A_dev, h, grid))
                mlp0 = model.model(y0)
            P = dict(model.named_parameters())
            np.savez_compressed(args.dump, x=xp.cpu().numpy(), A0=A_cpu.numpy(), v=v.cpu().numpy(), h=h, dims=dims,
                                sections=np.asarray(sections), masks=np.stack([(u <= 0.5).numpy() for u in draws]),
                                states=np.stack(states), loss=float(loss), dA0=A0.grad.cpu().numpy(),
                                gA0=gA0.cpu().numpy(), y0=y0.cpu().numpy(), mlp0=mlp0.cpu().numpy(),
                                dW1=P["model.0.weight"].grad.cpu().numpy(), db1=P["model.0.bias"].grad.cpu().numpy(),
                                dW2=P["model.2.weight"].grad.cpu().numpy(), db2=P["model.2.bias"].grad.cpu().numpy())
            out["dump"] = args.dump
    except BaseException as e:  # noqa: BLE001  (the error text IS the result when the reference cannot run here)
        out["error"] = f"{type(e).__name__}: {e}"
        out["traceback"] = traceback.format_exc()[-3000:]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
