#!/usr/bin/env python
"""bench.py -- particle-NCA-steps/s (fwd+bwd) of the SPH-NCA per-timestep particle update on B200.

One bench "step" = one training iteration of the reference's public API on one synthetic batch:
`nca.sample_plane(model, x0, A0, sections, h, dims, iter_n=K_nca, loss_fn, optimizer)`, i.e.
hash-grid build + volume + K_nca NCA steps forward + loss + backward + normalize_grads + Adam step
(reference code/nca.py:124-164, code/train.py:295-305).  Units processed per step =
particles x K_nca ("particle-NCA-steps").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--use-alpha on|off]

N>1 is launched by the driver with torch.distributed.run (one rank per GPU, NCCL).

The ONE JSON line (rank 0):
* headline (`value`, `e2e`, `roofline`, `kernels`, `perception`): BASELINE.json configs[1] -- 8 x 65,536 padded-2-D particles, 64 NCA
  steps fwd+bwd per GPU, data parallel over batch elements (weak scaling), the 84 KB gradient all-reduced before
  `normalize_grads`;
* `workloads`: bounded sub-records measured in the SAME run at the SAME N -- configs[2] (`train3d_250k_k96`, one cloud:
  replicas at N>1), configs[4] (`train3d_b64x1m_k16`: 64 clouds x 1,000,000 3-D particles split over the ranks = STRONG
  scaling; at N=1 the rollout node micro-batches over sections), configs[3] (`infer2d_16m_slab`: one 16 M cloud cut into
  slabs of cell rows with a halo exchange = STRONG scaling) and the alpha-channel model on configs[1] (`train2d_alpha`);
* `roofline`: dominant kernel; `achieved` = SURVEY.md 8(d) ALGORITHMIC bytes (perception 272 B, gradient adjoint 336 B per row) x
  rows per launch / CUDA-event launch time, `frac` = achieved / MEASURED_PEAKS.json; `traffic` = ncu DRAM bytes per launch
  (profiles/ncu_traffic.json), `traffic_frac` = traffic / time / peak, `traffic_ratio` = traffic / algorithmic;
* `gpu_baseline` (N=1): the REFERENCE's Numba-CUDA kernels + eager torch through its unchanged nca.sample_plane on this GPU
  (oracle/ref_gpu_baseline.py on the staged oracle/_ref/code) -- baseline #2 of the north star;
* `cpu_baseline` (N=1): the CPU oracle port on all host threads (bounded sample); `cpu_baseline_ref`: the reference itself under
  Numba's CUDA simulator (1 effective core) -- baseline #1.
`--impl reference` times the CPU oracle port with the thread count pinned (the launcher's OMP_NUM_THREADS is overridden).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

# NCCL writes its version / debug lines to stdout by default; stdout carries exactly one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

# workload name -> description, lattice dimension, particles per side, clouds per GPU (or in total), NCA steps, h / spacing
WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N=1
    "train2d_b8x65k_k64": dict(desc="2D texture training, batch 8 x 65,536 jittered particles (256^2, zero-padded to 3-D), "
                                    "16 channels, 64 NCA steps fwd+bwd", lattice_dim=2, n_side=256, batch=8, k_nca=64, hf=3.2),
    # the same with the alpha-channel model (the reference's train.py default and its gecko example, train-example.sh:8-13)
    "train2d_alpha": dict(desc="2D training with an alpha channel (gecko weights, radial seed), batch 8 x 65,536 particles, "
                               "64 NCA steps fwd+bwd", lattice_dim=2, n_side=256, batch=8, k_nca=64, hf=3.2, use_alpha=True),
    # BASELINE.json configs[2]
    "train3d_250k_k96": dict(desc="3D point cloud, 250,047 jittered particles (63^3), 96 NCA steps fwd+bwd",
                             lattice_dim=3, n_side=63, batch=1, k_nca=96, hf=1.97),
    # BASELINE.json configs[4]: 64 clouds of 1,000,000 3-D particles split data-parallel over the GPUs (strong scaling)
    "train3d_b64x1m_k16": dict(desc="3D training, batch 64 x 1,000,000 jittered particles (100^3), 16 NCA steps fwd+bwd, "
                                    "data-parallel over batch elements + NCCL gradient all-reduce; states + fire bytes kept, "
                                    "steps recomputed in the backward; micro-batched over sections where the pair lists of "
                                    "all clouds of a rank do not fit", lattice_dim=3, n_side=100, batch=64, batch_total=64,
                               k_nca=16, hf=1.97, device_inputs=True),
    # BASELINE.json configs[3]: one cloud, forward only, slab shard + NVLink halo (strong scaling over GPUs)
    "infer2d_16m_slab": dict(desc="discretization-agnostic inference on 16,000,000 jittered 2D particles (4000^2, zero-padded to 3-D, "
                                  "dims (1250,1250,3)), 16 NCA steps forward, spatial slab shard + halo exchange",
                             lattice_dim=2, n_side=4000, batch=1, k_nca=16, hf=3.2, slab=True),
    # reduced sizes for quick checks (not a bench line)
    "tiny": dict(desc="tiny 2D, batch 2 x 4,096, 8 steps", lattice_dim=2, n_side=64, batch=2, k_nca=8, hf=3.2),
    "tiny3d_b4": dict(desc="tiny 3D, batch 4 x 8,000, 4 steps", lattice_dim=3, n_side=20, batch=4, batch_total=4, k_nca=4, hf=1.97,
                      device_inputs=True),
}
DEFAULT_WORKLOAD = "train2d_b8x65k_k64"
SUBRECORDS = ("train2d_alpha", "train3d_250k_k96", "train3d_b64x1m_k16", "infer2d_16m_slab")
METRIC = "particle-NCA-steps/sec (fwd+bwd)"
UNIT = "particle-steps/s"
# Per-kernel roofline bookkeeping.  bytes = SURVEY.md 8(d) ALGORITHMIC HBM bytes per row a launch processes (compulsory traffic of
# the operator, index / list reads excluded), flops = algorithmic FLOPs per row (FMA = 2), bound = the roof that applies.
#   perception  = `gradient` forward:  read x,v,A 80 + write GA 192                  = 272 B   (8d row "perception")
#   adjoint     = `gradient` backward: read x,v,A 80 + dL/dGA 192 + write dL/dA 64   = 336 B   (8d row "gradient bwd")
#   MLP kernels: tensor-bound, 41,472 FLOP forward / data gradient, 82,944 weight gradient (recompute + accumulate);
#   the bytes of their own interfaces are listed for the hbm_gbs column only.
KERNELS = {
    "step_perceive": dict(bytes=272.0, flops=6700.0, bound="hbm", rows="fired_if_no_alpha"),
    "step_adjoint": dict(bytes=336.0, flops=6400.0, bound="hbm"),
    "fire_list": dict(bytes=68.5, flops=0.0, bound="hbm"),
    "step_mlp": dict(bytes=152.0, flops=41472.0, bound="tensor", rows="fired"),
    "step_fwd": dict(bytes=152.0, flops=41472.0 + 6700.0, bound="tensor"),
    "step_mask": dict(bytes=24.0, flops=1300.0, bound="hbm"),
    "step_mask_static": dict(bytes=2.0, flops=0.0, bound="hbm"),
    "step_bwd": dict(bytes=208.0, flops=41472.0, bound="tensor", rows="fired"),
    "step_bwd_unfired": dict(bytes=129.0, flops=0.0, bound="hbm"),
    "mlp_wgrad": dict(bytes=336.0, flops=82944.0, bound="tensor", rows="fired"),
}
STEP_BYTES = 360.0            # 8(d): fused step fwd+bwd, A_t only kept


def jittered_lattice(n_side, dim, gen, device="cpu"):
    s = 2.0 / n_side
    axes = [(torch.arange(n_side, dtype=torch.float32, device=device) + 0.5) * s - 1.0 for _ in range(dim)]
    mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, dim)
    return mesh + (torch.rand(mesh.shape, generator=gen, device=device) - 0.5) * s


def make_inputs(wl, seed, device="cpu"):
    """Synthetic batch: jittered lattices in [-1,1]^d; A ~ U[0,1), or (alpha-channel model) the reference's
    radial seed bump (nca.py:383-398) + noise.  CPU generator by default, `device` generator for the very large workloads."""
    gen = torch.Generator(device=device).manual_seed(seed)
    xs = [jittered_lattice(wl["n_side"], wl["lattice_dim"], gen, device) for _ in range(wl["batch"])]
    x = torch.cat(xs)
    del xs
    sections = tuple([wl["n_side"] ** wl["lattice_dim"]] * wl["batch"])
    if wl.get("use_alpha"):
        d2 = (x ** 2).sum(-1)
        A = (torch.clamp(1 - d2 / 0.5 ** 2, 0, 1) ** 3)[:, None].repeat(1, 16) \
            + 0.05 * torch.rand(x.shape[0], 16, generator=gen, device=device)
    else:
        A = torch.rand(x.shape[0], 16, generator=gen, device=device)
    h = wl["hf"] * 2.0 / wl["n_side"]
    return x, A, sections, h, math.ceil(2.0 / h)


def load_weights(name="zebra"):
    """Trained MLP weights of the reference's demo (docs/weights/<name>.json, committed as tests/golden/weights_<name>.npz):
    zebra = texture model without an alpha channel, gecko = the alpha-channel model of train-example.sh:8-13."""
    w = np.load(os.path.join(ROOT, "tests", "golden", f"weights_{name}.npz"))
    return [torch.from_numpy(w[k]) for k in ("W1", "b1", "W2", "b2")]


def bench_loss(x, A, img, A0=None):
    return (A[:, :4] ** 2).mean()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def host_threads():
    """All host cores, whatever the launcher put into OMP_NUM_THREADS (torch.distributed.run sets it to 1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(wl, steps, warmup, sample_sections=1, sample_k=4):
    """The CPU oracle port on a bounded sample of the workload: `sample_sections` batch elements,
    `sample_k` NCA steps, grid build + volume + fwd + loss + bwd, all host threads."""
    from oracle import sphnca_oracle as O
    cores = host_threads()
    sub = dict(wl, batch=sample_sections, k_nca=sample_k)
    x, A, sections, h, dims = make_inputs(sub, 0)
    if x.shape[1] == 2:
        x = torch.nn.functional.pad(x, (0, 1))
    params = load_weights()
    cfg = O.StepConfig(use_alpha=False, normalize_perception=1.0 / h)
    gen = torch.Generator().manual_seed(1)
    times = []
    for it in range(warmup + steps):
        masks = torch.rand(sample_k, x.shape[0], generator=gen) <= 0.5
        t0 = time.perf_counter()
        O.rollout_fwd_bwd(x, sections, h, dims, A, params, masks, cfg)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    units = x.shape[0] * sample_k
    sample = (f"{sample_sections} of {wl['batch']} batch elements ({x.shape[0]} particles) x {sample_k} of "
              f"{wl['k_nca']} NCA steps, grid build + volume + fwd + bwd, fp64 oracle (oracle/sphnca_oracle.py)")
    return units / (sum(times) / len(times)), sum(times) / len(times), sample, cores


def run_subprocess_json(cmd, timeout):
    """Run a baseline script (oracle/ref_*.py) and parse the JSON line it prints last."""
    try:
        r = subprocess.run([sys.executable] + cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"ok": False, "error": f"no JSON line (rc {r.returncode}): {r.stderr[-400:]}"}
    except subprocess.TimeoutExpired:
        return {"ok": False, "error": f"timed out after {timeout} s"}
    except Exception as e:  # noqa: BLE001
        return {"ok": False, "error": f"{type(e).__name__}: {e}"}


def gpu_baseline(use_alpha):
    """Baseline #2: the reference's Numba-CUDA path on this GPU (separate processes; the product is idle meanwhile)."""
    alpha = ["--use-alpha", "on" if use_alpha else "off"]
    out = {"what": "reference sphops Numba-CUDA kernels + eager torch MLP through the unchanged nca.sample_plane "
                   "(oracle/_ref/code staged from the reference, numba of this image), same inputs / weights / loss / optimiser, "
                   "CUDA-event timed after the JIT warm-up; K bounded by the reference's ~4 KB/particle-step autograd residue"}
    out["train2d_b8x65k"] = run_subprocess_json([os.path.join("oracle", "ref_gpu_baseline.py"), "--workload", DEFAULT_WORKLOAD,
                                                 "--nca-steps", "8", "--steps", "2", "--warmup", "1"] + alpha, 600)
    out["train3d_1m"] = run_subprocess_json([os.path.join("oracle", "ref_gpu_baseline.py"), "--workload", "train3d_b64x1m_k16",
                                             "--batch", "1", "--nca-steps", "4", "--steps", "2", "--warmup", "1"], 600)
    for v in out.values():
        if isinstance(v, dict):
            v.pop("traceback", None)
    return out


# ------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.distributed.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """`steps` calls between barrier + synchronize, CUDA events on the launching stream, MAX over ranks (ms)."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    def close(self):
        if self.world > 1:
            torch.distributed.destroy_process_group()


def release_memory():
    """Between workloads: drop autograd garbage and hand torch's cached blocks back, so that the next workload does not
    start inside an allocator full of the previous one's 100+ GB."""
    import gc
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def run_slab(ctx, args, name, wl, steps, warmup):
    """Forward-only rollout of ONE cloud sharded into slabs of cell rows (multigpu.SlabShardedNCA): strong scaling."""
    import nca
    import multigpu
    from sphops import _lib
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    n_side, K = wl["n_side"], wl["k_nca"]
    s_ = 2.0 / n_side
    h = wl["hf"] * s_
    d = math.ceil(2.0 / h)
    dims = (d, d, 3)
    model = nca.SPHNCA(16, 0.5, update_rule="gated", use_alpha=False, normalize_perception=-1).to(dev)
    with torch.no_grad():
        for t, src in zip((model.model[0].weight, model.model[0].bias, model.model[2].weight, model.model[2].bias), load_weights()):
            t.copy_(src)
    x = multigpu.slab_lattice_2d(n_side, h, d, rank, world, dev, seed=0)     # only this rank's rows (+ halo rows) are generated
    sh = multigpu.SlabShardedNCA(model, x, h, dims, axis=1, presharded=True)
    N = n_side * n_side
    n_own = sh.part.n_own
    del x
    A0 = multigpu.counter_uniform(sh.part.own_global_ids(), 16, seed=7, device=dev)   # same global state whatever the rank count
    for _ in range(max(warmup, 1)):
        sh.rollout(A0, 2)
    clk = ClockSampler(ctx.local_rank).start() if rank == 0 else None
    l0 = _lib.lib.sphnca_launch_count()
    ms = ctx.timed(lambda: sh.rollout(A0, K), steps)
    launches = int(_lib.lib.sphnca_launch_count() - l0)
    rec = {"metric": "particle-NCA-steps/sec (fwd)", "value": N * K * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
           "steps": steps, "warmup": max(warmup, 1), "ms_per_step": ms / steps, "scaling": "strong", "dtype": "f32",
           "config": {"workload": name, "description": wl["desc"], "particles_total": N, "particles_rank0": n_own,
                      "halo_rank0": sh.part.n_halo_prev + sh.part.n_halo_next, "nca_steps": K, "h": h, "dims": list(dims),
                      "parallelism": f"slab{world}", "halo_bytes_per_nca_step_rank0": sh.bytes_per_step,
                      "exchange": sh.exchange_mode, "l2": "state per rank >> 126 MB L2"},
           "gpu_launches": launches, "clocks": clk.stop() if clk else None}
    del sh, A0
    release_memory()
    return rec


def run_train(ctx, args, name, wl, steps, warmup, detail):
    """Training iterations of one workload, data parallel over batch elements.  detail=True (the headline): second pass with
    per-kernel CUDA events, the end-to-end pass with host buffers, launch count."""
    import nca
    from sphops import _lib, fused
    import multigpu
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    wl = dict(wl)
    strong = bool(wl.get("batch_total"))
    if strong:                                      # the batch is split over the ranks
        if wl["batch_total"] % world:
            raise SystemExit("batch_total must be divisible by the number of GPUs")
        wl["batch"] = wl["batch_total"] // world
    use_alpha = bool(wl.get("use_alpha")) or args.use_alpha == "on"
    wl["use_alpha"] = use_alpha
    nca.USE_3D = True
    nca.USE_FUSED_STEP = not args.no_fused
    fused.CHECKPOINT = {"auto": "auto", "on": True, "off": False}[args.checkpoint]
    torch.manual_seed(1234 + rank)
    on_device = bool(wl.get("device_inputs"))
    x_in, A_in, sections, h, dims = make_inputs(wl, seed=rank, device=dev if on_device else "cpu")
    N, K_nca = x_in.shape[0], wl["k_nca"]
    model = multigpu.DataParallelSPHNCA(16, 0.5, update_rule="gated", normalize_grad=True, use_alpha=use_alpha,
                                        normalize_perception=1.0 / h).to(dev)
    wname = "gecko" if use_alpha else "zebra"
    with torch.no_grad():
        for t, src in zip((model.model[0].weight, model.model[0].bias, model.model[2].weight, model.model[2].bias), load_weights(wname)):
            t.copy_(src)
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=0)
    if on_device:
        x_dev, A_dev = x_in, A_in
    else:
        x_pin, A_pin = x_in.pin_memory(), A_in.pin_memory()
        x_dev, A_dev = x_pin.to(dev), A_pin.to(dev)

    def one_step(x0=None, A0=None):
        return nca.sample_plane(model, x_dev if x0 is None else x0, A_dev if A0 is None else A0, sections, h, dims, iter_n=K_nca,
                                loss_fn=bench_loss, img=None, optimizer=opt)

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize()
    grid_probe = model.init_grid(torch.nn.functional.pad(x_dev, (0, 1)) if x_dev.shape[1] == 2 else x_dev, sections, h, dims)[0]
    taped = fused._tape_mode(N, K_nca, dev, None)
    groups = fused.plan_microbatches(grid_probe, K_nca, use_alpha, dev, taped)
    del grid_probe
    tape_mode = {1: "full tape", 2: "checkpointed (states + fire bytes kept, forward recomputed in backward)"}[2 if groups else taped]
    torch.cuda.reset_peak_memory_stats(dev)
    clk = ClockSampler(ctx.local_rank).start() if rank == 0 else None
    launches0 = _lib.lib.sphnca_launch_count()
    ms = ctx.timed(one_step, steps)
    launches = int(_lib.lib.sphnca_launch_count() - launches0)
    units = N * K_nca * world
    rec = {"metric": METRIC, "value": units * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms / steps, "scaling": "strong" if strong else ("weak" if wl["batch"] > 1 else "replicas"), "dtype": "f32",
           "config": {"workload": name, "description": wl["desc"], "particles_per_gpu": N, "clouds_per_gpu": wl["batch"],
                      "nca_steps": K_nca, "h": h, "dims": dims, "use_alpha": use_alpha,
                      "weights": f"{wname} (trained, tests/golden/weights_{wname}.npz)", "optimizer": "Adam lr=1e-5",
                      "path": "fused" if nca.USE_FUSED_STEP else "ops", "parallelism": f"dp{world}", "rollout_tape": tape_mode,
                      "microbatches": [list(g) for g in groups] if groups else None,
                      "l2": f"no flush needed: >= 65 B/particle/NCA-step of saved state => working set "
                            f"{65 * N * K_nca / 1e9:.1f}+ GB per step >> 126 MB L2"},
           "gpu_launches": launches, "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    if not detail:
        rec["clocks"] = clk.stop() if clk else None
        del x_dev, A_dev, model, opt
        release_memory()
        return rec

    # per-kernel durations: the same K steps once more with a CUDA-event pair (on the launching stream) around
    # every hot kernel launch.  ~770 event pairs per step cost ~4 % of a step, so they stay out of `value`'s pass.
    fused.PROFILE = {}
    ms_instr = ctx.timed(one_step, steps)
    prof, fused.PROFILE = fused.PROFILE, None
    rec["clocks"] = clk.stop() if clk else None

    # End to end: every step's inputs come from pinned host memory and its final state + loss go back to the host, all inside the
    # timed region.  The copies run on a copy stream, double buffered like an input pipeline would: the H2D of step i+1 and the D2H
    # of step i-1's result overlap step i's kernels; every step still consumes a copy made for it and the region ends with a
    # synchronize that covers the last D2H.
    out_pin = torch.empty_like(A_pin).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    def fetch():
        with torch.cuda.stream(copy_stream):
            x0 = x_pin.to(dev, non_blocking=True)
            A0 = A_pin.to(dev, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream)
        return x0, A0, ev

    def e2e_steps(n):
        nxt = fetch()
        for i in range(n):
            x0, A0, ev = nxt
            main_stream.wait_event(ev)
            x0.record_stream(main_stream); A0.record_stream(main_stream)
            if i + 1 < n:
                nxt = fetch()
            loss, _, A = one_step(x0, A0)                       # ends with loss.item(): the loss is on the host
            done = torch.cuda.Event(); done.record(main_stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                A = A.detach()
                A.record_stream(copy_stream)
                out_pin.copy_(A, non_blocking=True)
        copy_stream.synchronize()

    e2e_steps(2)
    ms_e2e = ctx.timed(lambda: e2e_steps(steps), 1)
    rec["e2e"] = {"value": units * steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / steps,
                  "h2d_bytes_per_step": x_pin.numel() * 4 + A_pin.numel() * 4, "d2h_bytes_per_step": out_pin.numel() * 4 + 4,
                  "copies": "pinned host buffers; H2D / D2H on a copy stream, double buffered (overlap the neighbouring steps' kernels)"}
    if rank != 0:
        return rec

    hbm, bf16_burst, bf16_sust, which = peaks()
    lists = nca.USE_FUSED_STEP and fused._fire_lists_on()
    fire_fraction = model.fire_rate if lists else 1.0
    tj = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
    kernels, dominant = {}, None
    for kname, evs in prof.items():
        tot = sum(a.elapsed_time(b) for a, b in evs)
        k = {"launches": len(evs), "avg_ms": tot / len(evs), "share_of_step": tot / ms_instr}
        info = KERNELS.get(kname)
        if info:
            sec = k["avg_ms"] * 1e-3
            fired = info.get("rows") == "fired" or (info.get("rows") == "fired_if_no_alpha" and not use_alpha)
            rows = N * (fire_fraction if fired else 1.0)                     # rows one launch processes
            k["rows_per_launch"] = rows
            k["algorithmic_bytes_per_row"] = info["bytes"]
            k["hbm_gbs"] = info["bytes"] * rows / sec / 1e9
            k["hbm_frac"] = k["hbm_gbs"] / hbm
            k["tflops"] = info["flops"] * rows / sec / 1e12
            k["bound"] = info["bound"]
            t = tj.get("bytes_per_launch", {}).get(kname)
            if t is not None and not use_alpha:
                t = t * N / tj["particles"]                                   # measured once under `ncu --set full`, scaled to this launch
                k["traffic_bytes_per_launch"] = t
                k["traffic_gbs"] = t / sec / 1e9
                k["traffic_frac"] = k["traffic_gbs"] / hbm
                k["traffic_ratio"] = t / (info["bytes"] * rows)
        kernels[kname] = k
        if dominant is None or tot > kernels[dominant]["avg_ms"] * kernels[dominant]["launches"]:
            dominant = kname
    roofline = None
    if dominant and dominant in KERNELS:
        info, k = KERNELS[dominant], kernels[dominant]
        if info["bound"] == "tensor":
            # kernels timed inside a long step -> sustained bf16 figure; kind::tf32 runs at half the bf16 rate and the
            # 3xTF32 split issues 3 MMAs per product, so 1/6 of this peak is the ceiling of an fp32-accurate kernel
            roofline = {"kernel": dominant, "bound": "tensor", "achieved": k["tflops"], "peak": bf16_sust, "unit": "TFLOP/s",
                        "frac": k["tflops"] / bf16_sust, "traffic": k.get("traffic_bytes_per_launch"),
                        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({which})",
                        "algorithmic_flops_per_row": info["flops"], "rows_per_launch": k["rows_per_launch"], "avg_launch_ms": k["avg_ms"],
                        "note": "fp32-accurate MLP on tcgen05 with a 3-pass split: ceiling = peak/6 of algorithmic FLOPs",
                        "frac_of_3pass_ceiling": k["tflops"] / (bf16_sust / 6)}
        else:
            roofline = {"kernel": dominant, "bound": "hbm", "achieved": k["hbm_gbs"], "peak": hbm, "unit": "GB/s",
                        "frac": k["hbm_frac"], "traffic": k.get("traffic_bytes_per_launch"),
                        "traffic_frac": k.get("traffic_frac"), "traffic_ratio": k.get("traffic_ratio"),
                        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({which})",
                        "algorithmic_bytes_per_row": info["bytes"], "rows_per_launch": k["rows_per_launch"], "avg_launch_ms": k["avg_ms"],
                        "note": "achieved = SURVEY.md 8(d) algorithmic bytes x rows / launch time; the pair-list sweeps also stream "
                                "~36 list slots x 12 B per row (an implementation cost, in `traffic`, not in `achieved`)"}
        roofline["traffic_source"] = tj.get("source")
    if "step_perceive" in kernels:                          # the metric's second half: perception HBM GB/s vs peak
        k = kernels["step_perceive"]
        rec["perception"] = {"kernel": "step_perceive", "achieved_gbs": k["hbm_gbs"], "peak_gbs": hbm, "frac": k["hbm_frac"],
                             "algorithmic_bytes_per_row": 272.0, "rows_per_launch": k["rows_per_launch"], "avg_launch_ms": k["avg_ms"],
                             "traffic_gbs": k.get("traffic_gbs"), "traffic_frac": k.get("traffic_frac"), "traffic_ratio": k.get("traffic_ratio")}
    rec["step_roofline"] = {"algorithmic_bytes_per_particle_step": STEP_BYTES, "achieved_gbs": STEP_BYTES * rec["value"] / world / 1e9,
                            "frac_of_hbm_peak": STEP_BYTES * rec["value"] / world / 1e9 / hbm}
    rec["config"]["mlp_rows"] = "fired particles only (fire lists)" if fire_fraction < 1.0 else "all particles"
    rec["roofline"], rec["kernels"] = roofline, kernels
    rec["kernel_timing"] = {"how": "second timed pass of the same steps with a CUDA-event pair around every hot kernel launch "
                                   "(kept out of the pass `value` is measured on)", "ms_per_step_instrumented": ms_instr / steps}
    del x_dev, A_dev, model, opt
    release_memory()
    return rec


def run_reference(args, wl, rank):
    """--impl reference: the reference's path on the host cores = the CPU oracle port (kind 'port'), all host threads."""
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), min(args.warmup, 1)
    rate, sec, sample, cores = cpu_oracle_rate(wl, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": args.workload, "description": wl["desc"]},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--nca-steps", type=int, default=0, help="override the workload's NCA steps per rollout (profiling runs)")
    ap.add_argument("--use-alpha", choices=["on", "off"], default="off",
                    help="headline model: off = texture model without an alpha channel (zebra weights; static life mask, fired-rows "
                         "perception), on = alpha-channel model (gecko weights; the reference's train.py default).  The other one is "
                         "always reported as a sub-record")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-subrecords", action="store_true", help="headline only (profiling runs)")
    ap.add_argument("--no-fused", action="store_true", help="operator-by-operator path (sphops ops + torch MLP)")
    ap.add_argument("--checkpoint", choices=["auto", "on", "off"], default="auto",
                    help="activation checkpointing inside the rollout node: keep states + fire bytes only (65 B per "
                         "particle-step) and recompute each step's forward in the backward; auto = by free HBM")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.nca_steps > 0:
        wl["k_nca"] = args.nca_steps
        wl["desc"] += f" [NCA steps overridden to {args.nca_steps}]"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if world != args.gpus and args.gpus > 1 and world == 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU oracle)")

    ctx = Ctx()
    warmup = max(args.warmup, 3)
    if wl.get("slab"):
        line = run_slab(ctx, args, args.workload, wl, args.steps, warmup)
    else:
        line = run_train(ctx, args, args.workload, wl, args.steps, warmup, detail=True)
    subs = {}
    if args.workload == DEFAULT_WORKLOAD and not args.no_subrecords and args.nca_steps == 0 and not args.no_fused:
        for name in SUBRECORDS:
            sw = dict(WORKLOADS[name])
            if name == "train2d_alpha" and args.use_alpha == "on":          # the headline is the alpha model: report the other one
                name, sw = "train2d_noalpha", dict(WORKLOADS[DEFAULT_WORKLOAD], use_alpha=False)
            n_steps = min(args.steps, 2 if name.startswith("train3d_b64") else 3)
            try:
                if sw.get("slab"):
                    subs[name] = run_slab(ctx, args, name, sw, n_steps, 1)
                else:
                    saved = args.use_alpha
                    args.use_alpha = "off"
                    subs[name] = run_train(ctx, args, name, sw, n_steps, 1, detail=False)
                    args.use_alpha = saved
            except Exception as e:  # noqa: BLE001  (a sub-record must not take the headline down; the error is the record)
                subs[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
                release_memory()
    ctx.barrier()
    if ctx.rank != 0:
        ctx.close()
        return
    line.update({"higher_is_better": True, "vs_baseline": None, "data": "synthetic"})
    if subs:
        line["workloads"] = subs
    if ctx.world == 1 and args.workload == DEFAULT_WORKLOAD:
        if not args.no_gpu_baseline:
            gb = gpu_baseline(args.use_alpha == "on")
            ours_3d = subs.get("train3d_b64x1m_k16", {}).get("value")
            b2, b3 = gb.get("train2d_b8x65k", {}), gb.get("train3d_1m", {})
            if b2.get("ok"):
                gb["speedup_train2d_b8x65k"] = line["value"] / b2["value"]
            if b3.get("ok") and ours_3d:
                gb["speedup_train3d_1m"] = ours_3d / b3["value"]
                gb["target"] = ">= 10x on the 1M-particle 3-D clouds (BASELINE.json north_star)"
            line["gpu_baseline"] = gb
        if not args.no_cpu_baseline:
            rate, sec, sample, cores = cpu_oracle_rate(wl, 1, 0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            sim = run_subprocess_json([os.path.join("oracle", "ref_cudasim_baseline.py"), "--n-side", "16", "--nca-steps", "1"], 300)
            sim.pop("traceback", None)
            line["cpu_baseline_ref"] = dict(sim, kind="reference", cores=1,
                                            note="the reference itself under NUMBA_ENABLE_CUDASIM=1: every CUDA thread is a Python thread "
                                                 "under the GIL, 1 effective core")
    print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
