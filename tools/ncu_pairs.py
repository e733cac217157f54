"""ncu target: pair-list sweeps on configs[1] inputs (development aid)."""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import torch
import sphops
from sphops import _lib, fused
from bench import make_inputs, WORKLOADS, load_weights
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "train2d_b8x65k_k64"]
x, A, sections, h, dims = make_inputs(wl, 0)
if x.shape[1] == 2: x = torch.nn.functional.pad(x, (0, 1))
x = x.cuda(); A = A.cuda(); N = x.shape[0]
grid = sphops.initialize_hashgrid(x, sections, h, dims)
v = sphops.volume(x, h, grid)
W = [t.cuda() for t in load_weights()]
xv4 = grid.sorted_xv(x, v); A_s = grid.pack_rows(A, 16)
Y = torch.zeros(N, 32, device="cuda"); fl = torch.zeros(N, dtype=torch.uint8, device="cuda")
dY = torch.randn(N, 32, device="cuda"); dAd = torch.randn(N, 16, device="cuda"); dA = torch.zeros(N, 16, device="cuda")
p = fused._step_params(*W, "gated", False, 1.0 / h, 1.0)
pl = grid.pair_lists(x, v)
for _ in range(3):
    _lib.check(_lib.lib.sphnca_step_perceive_pairs(grid.c_grid(), _lib.ptr(pl["slots"]), _lib.ptr(pl["base"]), _lib.ptr(pl["J"]), _lib.ptr(pl["WP"]),
               _lib.ptr(pl["WB"]), _lib.ptr(pl["FLG"]), _lib.ptr(A_s), C.byref(p), _lib.ptr(Y), _lib.ptr(fl), _lib.stream_ptr()))
    _lib.check(_lib.lib.sphnca_step_adjoint_pairs(grid.c_grid(), _lib.ptr(pl["slots"]), _lib.ptr(pl["base"]), _lib.ptr(pl["J"]), _lib.ptr(pl["WA"]),
               _lib.ptr(pl["SXY"]), _lib.ptr(xv4), _lib.ptr(dY), _lib.ptr(dAd), C.byref(p), _lib.ptr(dA), _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok")
