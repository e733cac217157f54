"""Launch time of the fire-list entry points (development aid)."""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import torch
import sphops
from sphops import _lib, fused
from bench import make_inputs, WORKLOADS
dev = "cuda"; lib, ptr, check = _lib.lib, _lib.ptr, _lib.check
st = _lib.stream_ptr()
x, A, sections, h, dims = make_inputs(WORKLOADS["train2d_b8x65k_k64"], 0)
x = torch.nn.functional.pad(x, (0, 1)).to(dev); A = A.to(dev); N = x.shape[0]
grid = sphops.initialize_hashgrid(x, sections, h, dims)
A_s = grid.pack_rows(A, 16)
u = torch.rand(N, device=dev)
fire = torch.empty(N, dtype=torch.uint8, device=dev); rows = torch.empty(N, dtype=torch.int32, device=dev); cnt = torch.zeros(1, dtype=torch.int32, device=dev)
An = torch.empty(N, 16, device=dev); al = torch.empty(N, dtype=torch.uint8, device=dev)
FLG = torch.ones(N, dtype=torch.uint8, device=dev); flags = torch.empty(N, dtype=torch.uint8, device=dev)
ws = fused._fire_list_workspace(N, dev)
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); [fn() for _ in range(reps)]; b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
g = grid.c_grid()
t_draw = timeit(lambda: check(lib.sphnca_fire_list_draw(g, ptr(u), 0.5, ptr(fire), ptr(A_s), 0, ptr(rows), ptr(cnt), ptr(An), ptr(al), ptr(FLG), ptr(flags), 1, ptr(ws), ws.numel(), st)))
t_list = timeit(lambda: check(lib.sphnca_fire_list(N, ptr(fire), ptr(A_s), 0, ptr(rows), ptr(cnt), ptr(An), ptr(al), ptr(FLG), ptr(flags), 1, ptr(ws), ws.numel(), st)))
t_only = timeit(lambda: check(lib.sphnca_fire_list(N, ptr(fire), ptr(A_s), 0, ptr(rows), ptr(cnt), None, None, None, None, 0, ptr(ws), ws.numel(), st)))
t_rand = timeit(lambda: torch.rand(N, device=dev))
t_copy = timeit(lambda: An.copy_(A_s))
print(f"N={N}: fire_list_draw {t_draw:.1f} us   fire_list (bytes given) {t_list:.1f} us   list only {t_only:.1f} us   torch.rand {t_rand:.1f} us   plain copy of all rows {t_copy:.1f} us   count {int(cnt.item())}")
