"""Data-gradient kernel: all rows of a 262,144-row array vs the same number of LISTED rows of a 524,288-row array (development aid)."""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import torch
from sphops import _lib, fused
from bench import load_weights
dev = "cuda"; lib, ptr, check = _lib.lib, _lib.ptr, _lib.check
W = [t.to(dev) for t in load_weights()]
p = fused._step_params(*W, "gated", False, 1.0, 1.0); st = _lib.stream_ptr()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); [fn() for _ in range(reps)]; b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
N = 524288; f32 = dict(dtype=torch.float32, device=dev)
for label, rows in (("every other row", torch.arange(0, N, 2, dtype=torch.int32, device=dev)),
                    ("random half", torch.sort(torch.randperm(N, device=dev)[:N // 2])[0].to(torch.int32)),
                    ("first half (contiguous)", torch.arange(0, N // 2, dtype=torch.int32, device=dev))):
    n = rows.numel(); cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    A_c = torch.rand(n, 16, **f32); O_c = torch.randn(n, 36, **f32); relu = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device=dev)
    flags = torch.full((N,), 7, dtype=torch.uint8, device=dev); dAn = torch.randn(N, 16, **f32)
    dO = torch.empty(n, 36, **f32); dY = torch.empty(N, 32, **f32); dAd = torch.empty(N, 16, **f32)
    t = timeit(lambda: check(lib.sphnca_step_bwd_list(N, ptr(rows), ptr(cnt), ptr(A_c), ptr(O_c), ptr(flags), ptr(relu), ptr(dAn), C.byref(p), ptr(dO), ptr(dY), ptr(dAd), st)))
    # forward list mode too (inputs by list)
    Y_c = torch.randn(n, 32, **f32); An = torch.empty(N, 16, **f32); al = torch.empty(N, dtype=torch.uint8, device=dev)
    tf = timeit(lambda: check(lib.sphnca_step_mlp_list(N, ptr(rows), ptr(cnt), ptr(A_c), ptr(Y_c), C.byref(p), ptr(An), ptr(al), ptr(O_c), ptr(flags), ptr(relu), None, None, 3, st)))
    print(f"{label:26s}: step_bwd_list {t:6.1f} us   step_mlp_list {tf:6.1f} us   ({n} rows)")
