#!/usr/bin/env python
"""bench.py -- particle-NCA-steps/s (fwd+bwd) of the SPH-NCA per-timestep particle update on B200.

One bench "step" = one training iteration of the reference's public API on one synthetic batch:
`nca.sample_plane(model, x0, A0, sections, h, dims, iter_n=K_nca, loss_fn, optimizer)`, i.e.
hash-grid build + volume + K_nca NCA steps forward + loss + backward + normalize_grads + Adam step
(reference code/nca.py:124-164, code/train.py:295-305).  Units processed per step =
particles x K_nca ("particle-NCA-steps").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

N>1 is launched by the driver with torch.distributed.run (one rank per GPU, NCCL): batches shard
data-parallel over sections (each rank its own equally sized batch = weak scaling) and the 84 KB
parameter gradient is all-reduced before `normalize_grads`.

JSON keys beyond the base contract: `roofline` (dominant kernel, CUDA-event timed inside the timed
region), `kernels` (per-kernel share of the step), `cpu_baseline` (the CPU oracle port on a bounded
sample, rank 0, N=1), `e2e` (same metric with host buffers: pinned H2D of the inputs and D2H of
the final state + loss inside the timed region).  `--impl reference` times the CPU oracle port
(the reference itself is Numba-CUDA Python with no CPU build).
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

# workload name -> (description, dim of the lattice, n_side, batch, K_nca, h factor)
WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N=1
    "train2d_b8x65k_k64": dict(desc="2D texture training, batch 8 x 65,536 jittered particles (256^2, zero-padded to 3-D), "
                                    "16 channels, 64 NCA steps fwd+bwd", lattice_dim=2, n_side=256, batch=8, k_nca=64, hf=3.2),
    # BASELINE.json configs[2]
    "train3d_250k_k96": dict(desc="3D point cloud, 250,047 jittered particles (63^3), 96 NCA steps fwd+bwd",
                             lattice_dim=3, n_side=63, batch=1, k_nca=96, hf=1.97),
    # BASELINE.json configs[4]: 64 clouds of 1,000,000 3-D particles, split data-parallel over the GPUs
    # (batch_total / n_gpus clouds per GPU = strong scaling; 3 NCA steps keep the saved state of 64 M particles in HBM)
    "train3d_b64x1m_k3": dict(desc="3D training, batch 64 x 1,000,000 jittered particles (100^3), 3 NCA steps fwd+bwd, "
                                   "data-parallel over batch elements + NCCL gradient all-reduce",
                              lattice_dim=3, n_side=100, batch=64, batch_total=64, k_nca=3, hf=1.97),
    # BASELINE.json configs[3]: one cloud, forward only, slab shard + NVLink halo (strong scaling over GPUs)
    "infer2d_16m_slab": dict(desc="discretization-agnostic inference on 16,000,000 jittered 2D particles (4000^2, zero-padded to 3-D, "
                                  "dims (1250,1250,3)), 16 NCA steps forward, spatial slab shard + halo exchange",
                             lattice_dim=2, n_side=4000, batch=1, k_nca=16, hf=3.2, slab=True),
    # reduced sizes for quick checks (not a bench line)
    "tiny": dict(desc="tiny 2D, batch 2 x 4,096, 8 steps", lattice_dim=2, n_side=64, batch=2, k_nca=8, hf=3.2),
}
DEFAULT_WORKLOAD = "train2d_b8x65k_k64"
METRIC = "particle-NCA-steps/sec (fwd+bwd)"
UNIT = "particle-steps/s"
# Per-kernel roofline bookkeeping.  bytes = algorithmic HBM bytes per particle per launch
# (compulsory traffic of that kernel's interface), flops = algorithmic FLOPs per particle (FMA = 2),
# bound = which roof applies: "hbm" (sweeps: memory/latency side) or "tensor" (tcgen05 MLP kernels).
KERNELS = {
    # perception from the pair lists (csrc/pairs.cu), over the FIRED rows only (no alpha channel): ~36 list slots x 12 B (J, WP)
    # + A row (64) + list entry (4) + write compact Y_c (128) and A_c (64); 24 packed FMA/ADD pairs per slot
    "step_perceive": dict(bytes=692.0, flops=36 * 64.0, bound="hbm", rows="fired"),
    # ordered compaction of the fire mask + carry-over of the not-fired rows: fire bytes twice (2), list entry (4 x 0.5),
    # not-fired row copy (129 x 0.5)
    "fire_list": dict(bytes=68.5, flops=0.0, bound="hbm"),
    # The three MLP kernels run over the FIRED rows only (csrc/firelist.cu); their figures are per fired row and are scaled
    # by the fire rate (0.5) below ("rows": "fired").
    # MLP + update: read A,Y (192) + list entry, flag (5) ; write A' (64), O (144), relu bits (32), flags/alive (2), compact A,Y
    # copies for the backward (192); 2*(48*256+256*33) FLOP
    "step_mlp": dict(bytes=631.0, flops=41472.0, bound="tensor", rows="fired"),
    "step_fwd": dict(bytes=152.0, flops=41472.0 + 5500.0, bound="tensor"),   # fused forward (A/B path)
    "step_mask": dict(bytes=24.0, flops=1300.0, bound="hbm"),               # neighbour-sweep life mask (alpha channel only)
    "step_mask_static": dict(bytes=2.0, flops=0.0, bound="hbm"),            # no alpha channel: elementwise
    # data gradient: read A,O,dA',relu,flags,list entry (64+144+64+32+1+4) ; write dO (144), dY (128), dAdir (64); 2*(33*256+256*48) FLOP
    "step_bwd": dict(bytes=645.0, flops=41472.0, bound="tensor", rows="fired"),
    # backward of the not-fired rows: flag (1) + (read dA' 64 + write dAdir 64, dY 128) x 0.5
    "step_bwd_unfired": dict(bytes=129.0, flops=0.0, bound="hbm"),
    # weight gradient with transposed recompute: read A,Y,dO (336); recompute 2*(48*256+33*256) + accumulate the same
    "mlp_wgrad": dict(bytes=336.0, flops=82944.0, bound="tensor", rows="fired"),
    # adjoint from the pair lists: ~36 slots x 12 B (J, WA) + x,v (16) + SXY (8) + dY (128) + dAdir (64) ; write dA (64)
    "step_adjoint": dict(bytes=712.0, flops=36 * 64.0, bound="hbm"),
}


def jittered_lattice(n_side, dim, gen):
    s = 2.0 / n_side
    axes = [(torch.arange(n_side, dtype=torch.float32) + 0.5) * s - 1.0 for _ in range(dim)]
    mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, dim)
    return mesh + (torch.rand(mesh.shape, generator=gen) - 0.5) * s


def make_inputs(wl, seed):
    """Synthetic batch: jittered lattices in [-1,1]^d, A ~ U[0,1), CPU generator."""
    gen = torch.Generator().manual_seed(seed)
    xs = [jittered_lattice(wl["n_side"], wl["lattice_dim"], gen) for _ in range(wl["batch"])]
    x = torch.cat(xs)
    sections = tuple(t.shape[0] for t in xs)
    A = torch.rand(x.shape[0], 16, generator=gen)
    h = wl["hf"] * 2.0 / wl["n_side"]
    return x, A, sections, h, math.ceil(2.0 / h)


def load_weights():
    w = np.load(os.path.join(ROOT, "tests", "golden", "weights_zebra.npz"))
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


# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(wl, steps, warmup, sample_sections=1, sample_k=4):
    """The CPU oracle port on a bounded sample of the workload: `sample_sections` batch elements,
    `sample_k` NCA steps, grid build + volume + fwd + loss + bwd, all host threads torch uses."""
    from oracle import sphnca_oracle as O
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
    return units / (sum(times) / len(times)), sum(times) / len(times), sample


def run_slab(args, wl, rank, world, local_rank):
    """Forward-only rollout of ONE cloud sharded into slabs of cell rows (multigpu.SlabShardedNCA)."""
    import nca
    import multigpu
    from sphops import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    n_side, K = wl["n_side"], wl["k_nca"]
    s_ = 2.0 / n_side
    h = wl["hf"] * s_
    d = math.ceil(2.0 / h)
    dims = (d, d, 3)
    gen = torch.Generator(device=dev).manual_seed(0)               # every rank builds the same cloud on its GPU
    ax = (torch.arange(n_side, dtype=torch.float32, device=dev) + 0.5) * s_ - 1.0
    x = torch.stack(torch.meshgrid(ax, ax, indexing="ij"), dim=-1).reshape(-1, 2)
    x = x + (torch.rand(x.shape, generator=gen, device=dev) - 0.5) * s_
    x = torch.nn.functional.pad(x, (0, 1))
    N = x.shape[0]
    model = nca.SPHNCA(16, 0.5, update_rule="gated", use_alpha=False, normalize_perception=-1).to(dev)
    with torch.no_grad():
        for t, src in zip((model.model[0].weight, model.model[0].bias, model.model[2].weight, model.model[2].bias), load_weights()):
            t.copy_(src)
    sh = multigpu.SlabShardedNCA(model, x, h, dims, axis=1)
    n_own = sh.part.n_own
    del x
    torch.manual_seed(1234 + rank)
    A0 = torch.rand(n_own, 16, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        sh.rollout(A0, 2)
    barrier()
    l0 = _lib.lib.sphnca_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = sh.rollout(A0, K)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(ms.item())
    launches = int(_lib.lib.sphnca_launch_count() - l0)
    if rank == 0:
        line = {"metric": "particle-NCA-steps/sec (fwd)", "value": N * K * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "description": wl["desc"], "particles_total": N, "particles_rank0": n_own,
                           "halo_rank0": sh.part.n_halo_prev + sh.part.n_halo_next, "nca_steps": K, "h": h, "dims": list(dims),
                           "parallelism": f"slab{world}", "halo_bytes_per_nca_step_rank0": sh.bytes_per_step,
                           "l2": "state per rank >> 126 MB L2"},
                "gpu_launches": launches}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_reference(args, wl, rank):
    """--impl reference: the reference's path on the host cores = the CPU oracle port (kind 'port')."""
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), min(args.warmup, 1)
    rate, sec, sample = cpu_oracle_rate(wl, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": args.workload, "description": wl["desc"]},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU oracle)")
    if wl.get("slab"):
        run_slab(args, wl, rank, world, local_rank)
        return

    import nca
    import sphops
    from sphops import _lib, fused
    import multigpu

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)

    if wl.get("batch_total"):                     # strong scaling: the batch is split over the ranks
        if wl["batch_total"] % world:
            raise SystemExit("batch_total must be divisible by the number of GPUs")
        wl["batch"] = wl["batch_total"] // world
    nca.USE_3D = True
    nca.USE_FUSED_STEP = not args.no_fused
    fused.CHECKPOINT = {"auto": "auto", "on": True, "off": False}[args.checkpoint]
    torch.manual_seed(1234 + rank)
    x_cpu, A_cpu, sections, h, dims = make_inputs(wl, seed=rank)
    N, K_nca = x_cpu.shape[0], wl["k_nca"]
    model = multigpu.DataParallelSPHNCA(16, 0.5, update_rule="gated", normalize_grad=True, use_alpha=False,
                                        normalize_perception=1.0 / h).to(dev)
    with torch.no_grad():
        for t, src in zip((model.model[0].weight, model.model[0].bias, model.model[2].weight, model.model[2].bias), load_weights()):
            t.copy_(src)
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=0)

    x_pin, A_pin = x_cpu.pin_memory(), A_cpu.pin_memory()
    out_pin = torch.empty_like(A_pin).pin_memory()
    x_dev, A_dev = x_pin.to(dev), A_pin.to(dev)
    tape_mode = {1: "full tape", 2: "checkpointed (states + fire bytes kept, forward recomputed in backward)"}[
        fused._tape_mode(N, K_nca, dev, None)]

    def one_step(x0, A0):
        return nca.sample_plane(model, x0, A0, sections, h, dims, iter_n=K_nca, loss_fn=bench_loss, img=None, optimizer=opt)

    def one_step_e2e():
        x0 = x_pin.to(dev, non_blocking=True)
        A0 = A_pin.to(dev, non_blocking=True)
        loss, _, A = one_step(x0, A0)
        out_pin.copy_(A.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        one_step(x_dev, A_dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.lib.sphnca_launch_count()
    ms = timed(lambda: one_step(x_dev, A_dev), args.steps)
    launches = int(_lib.lib.sphnca_launch_count() - launches0)
    # per-kernel durations: the same K steps once more with a CUDA-event pair (on the launching stream) around
    # every hot kernel launch.  ~770 event pairs per step cost ~4 % of a step, so they stay out of `value`'s pass.
    fused.PROFILE = {}
    ms_instr = timed(lambda: one_step(x_dev, A_dev), args.steps)
    prof, fused.PROFILE = fused.PROFILE, None
    clocks = sampler.stop() if rank == 0 else None

    for _ in range(2):
        one_step_e2e()
    ms_e2e = timed(one_step_e2e, args.steps)

    units_per_step = N * K_nca * world
    value = units_per_step * args.steps / (ms * 1e-3)
    e2e_value = units_per_step * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    hbm, bf16_burst, bf16_sust, which = peaks()
    # rows the MLP kernels process per launch: the fired ones when the rollout runs on fire lists (model fire rate 0.5)
    fire_fraction = model.fire_rate if (nca.USE_FUSED_STEP and fused._fire_lists_on()) else 1.0
    kernels, dominant = {}, None
    for name, evs in prof.items():
        tot = sum(a.elapsed_time(b) for a, b in evs)
        k = {"launches": len(evs), "avg_ms": tot / len(evs), "share_of_step": tot / ms}
        info = KERNELS.get(name)
        if info:
            sec = k["avg_ms"] * 1e-3
            rows = N * (fire_fraction if info.get("rows") == "fired" else 1.0)     # rows one launch processes
            k["rows_per_launch"] = rows
            k["hbm_gbs"] = info["bytes"] * rows / sec / 1e9
            k["hbm_frac"] = k["hbm_gbs"] / hbm
            k["tflops"] = info["flops"] * rows / sec / 1e12
            k["bound"] = info["bound"]
        kernels[name] = k
        if dominant is None or tot > kernels[dominant]["avg_ms"] * kernels[dominant]["launches"]:
            dominant = name
    roofline = None
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if dominant and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if dominant in tj.get("bytes_per_launch", {}):
            # measured once under `ncu --set full` (never during a timed run); scaled to this launch size
            traffic = tj["bytes_per_launch"][dominant] * N / tj["particles"]
            traffic_src = tj["source"]
    if dominant and dominant in KERNELS:
        info, k = KERNELS[dominant], kernels[dominant]
        if info["bound"] == "tensor":
            # kernels timed inside a long step -> sustained bf16 figure; kind::tf32 runs at half the bf16 rate and the
            # 3xTF32 split issues 3 MMAs per product, so 1/6 of this peak is the ceiling of an fp32-accurate kernel
            roofline = {"kernel": dominant, "bound": "tensor", "achieved": k["tflops"], "peak": bf16_sust, "unit": "TFLOP/s",
                        "frac": k["tflops"] / bf16_sust, "traffic": None,
                        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({which})",
                        "algorithmic_flops_per_row": info["flops"], "rows_per_launch": k["rows_per_launch"], "avg_launch_ms": k["avg_ms"],
                        "note": "fp32-accurate MLP on tcgen05 kind::tf32 with the 3xTF32 split: ceiling = peak/6 "
                                f"= {bf16_sust / 6:.0f} TFLOP/s of algorithmic FLOPs; frac_of_3xtf32_ceiling below",
                        "frac_of_3xtf32_ceiling": k["tflops"] / (bf16_sust / 6),
                        "algorithmic_bytes_per_launch": info["bytes"] * k["rows_per_launch"]}
        else:
            roofline = {"kernel": dominant, "bound": "hbm", "achieved": k["hbm_gbs"], "peak": hbm, "unit": "GB/s",
                        "frac": k["hbm_frac"], "traffic": None, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({which})",
                        "algorithmic_bytes_per_particle": info["bytes"], "particles_per_launch": N, "avg_launch_ms": k["avg_ms"],
                        "note": "pair-list sweep: list stream from HBM + ~36 gathered rows per particle out of L1/L2; bound by the L1 "
                                "data pipe (DESIGN.md 4)"}

    if roofline is not None:
        roofline["traffic"] = traffic
        roofline["traffic_source"] = traffic_src
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if wl.get("batch_total") else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"], "particles_per_gpu": N, "nca_steps": K_nca,
                   "h": h, "dims": dims, "weights": "zebra (trained, tests/golden/weights_zebra.npz)", "optimizer": "Adam lr=1e-5",
                   "path": "fused" if nca.USE_FUSED_STEP else "ops", "parallelism": f"dp{world}", "rollout_tape": tape_mode,
                   "mlp_rows": "fired particles only (fire lists)" if fire_fraction < 1.0 else "all particles",
                   "l2": f"no flush needed: ~{fused.TAPE_BYTES + 64} B/particle/NCA-step of saved state => working set "
                         f"{(fused.TAPE_BYTES + 64) * N * K_nca / 1e9:.1f} GB per step >> 126 MB L2"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": x_pin.numel() * 4 + A_pin.numel() * 4, "d2h_bytes_per_step": out_pin.numel() * 4 + 4},
        "roofline": roofline, "kernels": kernels,
        "kernel_timing": {"how": "second timed pass of the same steps with a CUDA-event pair around every hot kernel launch "
                                 "(kept out of the pass `value` is measured on)", "ms_per_step_instrumented": ms_instr / args.steps},
    }
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, sample = cpu_oracle_rate(wl, 1, 0)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
