"""Kernel micro-benchmark through the C ABI (development aid): times sphnca_step_fwd variants with CUDA events."""
import ctypes as C, math, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import torch, numpy as np
import sphops
from sphops import _lib, fused
from bench import make_inputs, WORKLOADS, load_weights
dev = "cuda"
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "train2d_b8x65k_k64"]
x, A, sections, h, dims = make_inputs(wl, 0)
if x.shape[1] == 2: x = torch.nn.functional.pad(x, (0, 1))
x = x.to(dev); A = A.to(dev); N = x.shape[0]
grid = sphops.initialize_hashgrid(x, sections, h, dims)
v = sphops.volume(x, h, grid)
W = [t.to(dev) for t in load_weights()]
xv4 = grid.sorted_xv(x, v); A_s = grid.pack_rows(A, 16)
fire_s = grid.pack_bytes((torch.rand(N, device=dev) <= 0.5).to(torch.uint8))
An = torch.zeros(N, 16, device=dev); Y = torch.zeros(N, 32, device=dev); O = torch.zeros(N, 36, device=dev)
fl = torch.zeros(N, dtype=torch.uint8, device=dev); al = torch.zeros(N, dtype=torch.uint8, device=dev)
rl = torch.zeros(N, 8, dtype=torch.int32, device=dev)
p = fused._step_params(*W, "gated", False, 1.0 / h, 1.0)
def run():
    _lib.check(_lib.lib.sphnca_step_fwd(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(A_s), _lib.ptr(fire_s), C.byref(p),
               _lib.ptr(An), _lib.ptr(al), _lib.ptr(Y), _lib.ptr(O), _lib.ptr(fl), _lib.ptr(rl), _lib.stream_ptr()))
for path, name in ((0, "simt fused"), (2, "tcgen05 fused"), (1, "tcgen05 split (perceive + mlp)")):
    _lib.check(_lib.lib.sphnca_set_mlp_path(path))
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"step_fwd [{name}]: {ms:.4f} ms  ({N/ms/1e6:.2f} G particles/s, {ms*1e-3*1.965e9*148/N:.0f} clk*SM/particle)")
_lib.check(_lib.lib.sphnca_set_mlp_path(1))
# perception sweep alone, per sweep path (0 gathers, 1 smem tile engine, 2 TMA-staged tiles)
for sp in (0, 1, 2):
    _lib.check(_lib.lib.sphnca_set_sweep_path(sp))
    def perc():
        _lib.check(_lib.lib.sphnca_step_perceive(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(A_s), C.byref(p), _lib.ptr(Y), _lib.ptr(fl), _lib.stream_ptr()))
    for _ in range(3): perc()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): perc()
    b.record(); torch.cuda.synchronize()
    print(f"step_perceive [sweep path {sp}]: {a.elapsed_time(b) / 10:.4f} ms")
_lib.check(_lib.lib.sphnca_set_sweep_path(0))
# pair-list sweeps
import time
torch.cuda.synchronize(); t0 = time.time()
pl = grid.pair_lists(x, v)
torch.cuda.synchronize(); print(f"pair lists: {pl['entries']} entries ({pl['entries'] * 28 / 1e6:.0f} MB, {pl['entries'] / N:.1f} slots/particle), built in {(time.time() - t0) * 1e3:.1f} ms")
dY = torch.randn(N, 32, device=dev); dAd = torch.randn(N, 16, device=dev); dA = torch.zeros(N, 16, device=dev)
def percp():
    _lib.check(_lib.lib.sphnca_step_perceive_pairs(grid.c_grid(), _lib.ptr(pl["slots"]), _lib.ptr(pl["base"]), _lib.ptr(pl["J"]), _lib.ptr(pl["WP"]),
               _lib.ptr(pl["WB"]), _lib.ptr(pl["FLG"]), _lib.ptr(A_s), C.byref(p), _lib.ptr(Y), _lib.ptr(fl), _lib.stream_ptr()))
def adjp():
    _lib.check(_lib.lib.sphnca_step_adjoint_pairs(grid.c_grid(), _lib.ptr(pl["slots"]), _lib.ptr(pl["base"]), _lib.ptr(pl["J"]), _lib.ptr(pl["WA"]),
               _lib.ptr(pl["SXY"]), _lib.ptr(xv4), _lib.ptr(dY), _lib.ptr(dAd), C.byref(p), _lib.ptr(dA), _lib.stream_ptr()))
def adjg():
    _lib.check(_lib.lib.sphnca_step_adjoint(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(dY), _lib.ptr(dAd), C.byref(p), _lib.ptr(dA), _lib.stream_ptr()))
for name, fn in (("perceive from pair lists", percp), ("adjoint from pair lists", adjp), ("adjoint gather", adjg)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / 10:.4f} ms")
