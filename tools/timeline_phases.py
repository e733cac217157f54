"""Launch phases of step_mlp_tc_kernel's CTA 0 (development aid; needs `make EXTRA=-DSPHNCA_TIMELINE`)."""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import torch
from sphops import _lib, fused
from bench import load_weights
dev = "cuda"; lib, ptr, check = _lib.lib, _lib.ptr, _lib.check
W = [t.to(dev) for t in load_weights()]
p = fused._step_params(*W, "gated", False, 1.0, 1.0); st = _lib.stream_ptr()
lib.sphnca_debug_timeline_mlp.argtypes = [C.POINTER(C.c_longlong), C.c_int32]
for N in (128, 18944, 262144):
    f32 = dict(dtype=torch.float32, device=dev)
    A = torch.rand(N, 16, **f32); Y = torch.randn(N, 32, **f32); O = torch.empty(N, 36, **f32); An = torch.empty(N, 16, **f32)
    fire = torch.ones(N, dtype=torch.uint8, device=dev); fl = torch.full((N,), 7, dtype=torch.uint8, device=dev); al = torch.empty(N, dtype=torch.uint8, device=dev)
    rl = torch.empty(N, 8, dtype=torch.int32, device=dev)
    run = lambda: check(lib.sphnca_step_mlp(N, ptr(A), ptr(Y), ptr(fire), C.byref(p), ptr(An), ptr(al), ptr(O), ptr(fl), ptr(rl), st))
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); [run() for _ in range(20)]; b.record(); torch.cuda.synchronize()
    buf = (C.c_longlong * (64 * 16))(); lib.sphnca_debug_timeline_mlp(buf, 64 * 16)
    t = [buf[63 * 16 + k] for k in range(4)]
    first = {k: buf[0 * 16 + k] - t[0] for k in (12, 0, 6, 7, 9, 10, 11)}
    print(f"N={N}: {a.elapsed_time(b) / 20 * 1e3:.1f} us per launch; CTA 0 cycles from entry: weights staged {t[1]-t[0]}, ready {t[2]-t[0]}, done {t[3]-t[0]}"
          f" ({(t[3]-t[0]) / 1.9e3:.1f} us at 1.9 GHz); tile 0: loaded {first[12]}, Y_FULL {first[0]}, D1a ok {first[6]}, epi a done {first[7]}, epi b done {first[9]}, D2 ok {first[10]}, final done {first[11]}")
