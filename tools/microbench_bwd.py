"""Micro-benchmark of the backward kernels through the C ABI (development aid)."""
import ctypes as C, math, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import torch
import sphops
from sphops import _lib, fused
from bench import make_inputs, WORKLOADS, load_weights
dev = "cuda"
wl = WORKLOADS["train2d_b8x65k_k64"]
x, A, sections, h, dims = make_inputs(wl, 0)
x = torch.nn.functional.pad(x, (0, 1)).to(dev); A = A.to(dev); N = x.shape[0]
grid = sphops.initialize_hashgrid(x, sections, h, dims)
v = sphops.volume(x, h, grid)
W = [t.to(dev) for t in load_weights()]
xv4 = grid.sorted_xv(x, v); A_s = grid.pack_rows(A, 16)
fire_s = grid.pack_bytes((torch.rand(N, device=dev) <= 0.5).to(torch.uint8))
f32 = dict(dtype=torch.float32, device=dev)
An = torch.zeros(N, 16, **f32); Y = torch.zeros(N, 32, **f32); O = torch.zeros(N, 36, **f32)
fl = torch.zeros(N, dtype=torch.uint8, device=dev); al = torch.zeros(N, dtype=torch.uint8, device=dev)
rl = torch.zeros(N, 8, dtype=torch.int32, device=dev)
p = fused._step_params(*W, "gated", False, 1.0 / h, 1.0)
st = _lib.stream_ptr()
_lib.check(_lib.lib.sphnca_step_fwd(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(A_s), _lib.ptr(fire_s), C.byref(p), _lib.ptr(An), _lib.ptr(al), _lib.ptr(Y), _lib.ptr(O), _lib.ptr(fl), _lib.ptr(rl), st))
_lib.check(_lib.lib.sphnca_step_mask(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(al), _lib.ptr(An), _lib.ptr(fl), st))
dAn = torch.randn(N, 16, **f32); dO = torch.empty(N, 36, **f32); dY = torch.empty(N, 32, **f32); dAd = torch.empty(N, 16, **f32); dA = torch.empty(N, 16, **f32)
tiles = (N + 31) // 32
Ht = torch.empty(tiles, 256, 32, **f32); dHt = torch.empty(tiles, 256, 32, **f32)
nb = C.c_size_t(0); _lib.check(_lib.lib.sphnca_mlp_wgrad_workspace_bytes(N, C.byref(nb)))
ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
dW = [torch.zeros_like(t) for t in W]
def bwd(): _lib.check(_lib.lib.sphnca_step_bwd(N, _lib.ptr(A_s), _lib.ptr(Y), _lib.ptr(O), _lib.ptr(fl), _lib.ptr(rl), _lib.ptr(dAn), C.byref(p), _lib.ptr(dO), _lib.ptr(dY), _lib.ptr(dAd), _lib.ptr(Ht), _lib.ptr(dHt), st))
def bwd_nostore(): _lib.check(_lib.lib.sphnca_step_bwd(N, _lib.ptr(A_s), _lib.ptr(Y), _lib.ptr(O), _lib.ptr(fl), _lib.ptr(rl), _lib.ptr(dAn), C.byref(p), _lib.ptr(dO), _lib.ptr(dY), _lib.ptr(dAd), None, None, st))
def wgrad(): _lib.check(_lib.lib.sphnca_mlp_wgrad(N, _lib.ptr(A_s), _lib.ptr(Y), _lib.ptr(dO), _lib.ptr(Ht), _lib.ptr(dHt), C.byref(p), *[_lib.ptr(t) for t in dW], _lib.ptr(ws), nb.value, st))
def wgrad_tc(): _lib.check(_lib.lib.sphnca_mlp_wgrad(N, _lib.ptr(A_s), _lib.ptr(Y), _lib.ptr(dO), None, None, C.byref(p), *[_lib.ptr(t) for t in dW], _lib.ptr(ws), nb.value, st))
def adj(): _lib.check(_lib.lib.sphnca_step_adjoint(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(dY), _lib.ptr(dAd), C.byref(p), _lib.ptr(dA), st))
def mask(): _lib.check(_lib.lib.sphnca_step_mask(grid.c_grid(), _lib.ptr(xv4), _lib.ptr(al), _lib.ptr(An), _lib.ptr(fl), st))
for name, fn in (("step_bwd simt (+H^T,dH^T stores)", bwd), ("step_bwd tcgen05", bwd_nostore), ("mlp_wgrad simt", wgrad), ("mlp_wgrad tcgen05", wgrad_tc), ("step_adjoint", adj), ("step_mask", mask)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"{name:32s}: {ms:.4f} ms  ({ms*1e-3*1.965e9*148/N:.0f} clk*SM/particle)")
