"""Launch time of the three tcgen05 kernels vs the number of rows (development aid): separates the per-launch fixed cost
(weight staging, pipeline fill, final dump) from the per-row cost."""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import torch
from sphops import _lib, fused
from bench import load_weights
dev = "cuda"
lib, ptr, check = _lib.lib, _lib.ptr, _lib.check
W = [t.to(dev) for t in load_weights()]
p = fused._step_params(*W, "gated", False, 1.0, 1.0)
st = _lib.stream_ptr()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for N in (128, 1024, 18944, 32768, 131072, 262144, 524288):
    f32 = dict(dtype=torch.float32, device=dev)
    A = torch.rand(N, 16, **f32); Y = torch.randn(N, 32, **f32); O = torch.randn(N, 36, **f32); dO = torch.randn(N, 36, **f32) * 1e-3
    An = torch.empty(N, 16, **f32); dAn = torch.randn(N, 16, **f32); dY = torch.empty(N, 32, **f32); dAd = torch.empty(N, 16, **f32)
    fire = torch.ones(N, dtype=torch.uint8, device=dev); fl = torch.full((N,), 7, dtype=torch.uint8, device=dev); al = torch.empty(N, dtype=torch.uint8, device=dev)
    rl = torch.randint(-2**31, 2**31 - 1, (N, 8), dtype=torch.int32, device=dev)
    nb = C.c_size_t(0); check(lib.sphnca_mlp_wgrad_workspace_bytes(N, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    dW = [torch.zeros_like(t) for t in W]
    t_mlp = timeit(lambda: check(lib.sphnca_step_mlp(N, ptr(A), ptr(Y), ptr(fire), C.byref(p), ptr(An), ptr(al), ptr(O), ptr(fl), ptr(rl), st)))
    fl.fill_(7)
    t_bwd = timeit(lambda: check(lib.sphnca_step_bwd(N, ptr(A), ptr(Y), ptr(O), ptr(fl), ptr(rl), ptr(dAn), C.byref(p), ptr(dO), ptr(dY), ptr(dAd), None, None, st)))
    t_wg = timeit(lambda: check(lib.sphnca_mlp_wgrad(N, ptr(A), ptr(Y), ptr(dO), None, None, C.byref(p), *[ptr(t) for t in dW], ptr(ws), nb.value, st)))
    print(f"N={N:7d}: step_mlp {t_mlp*1e3:7.1f} us   step_bwd {t_bwd*1e3:7.1f} us   mlp_wgrad(+reduce) {t_wg*1e3:7.1f} us")
