# needs a library built with the in-kernel time stamps: make -C sph-nca_b200/csrc clean && make -C sph-nca_b200/csrc EXTRA=-DSPHNCA_TIMELINE
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import runpy
g = runpy.run_path(os.path.join(ROOT, "tools", "microbench_bwd.py"))
from sphops import _lib
import torch
g["wgrad_tc"](); torch.cuda.synchronize()
buf = (C.c_longlong * (64 * 16))()
for fn in ("sphnca_debug_timeline_wgrad", "sphnca_debug_timeline_fwd"):
    getattr(_lib.lib, fn).argtypes = [C.POINTER(C.c_longlong), C.c_int32]
print("rc", _lib.lib.sphnca_debug_timeline_wgrad(buf, 64 * 16))
names = ["L K_EMPTY ok", "L K_FULL arr", "L T_EMPTY ok", "L T_FULL arr", "M K_FULL ok", "M recomp issued", "M HT_FULL ok", "M T_FULL ok", "M accum issued", "E R_FULL ok", "E HT arr", "L cpasync landed", "L barsync", "L iter top", "L copy issued"]
t0 = buf[4]
for it in range(2, 8):
    row = [buf[it * 16 + k] - t0 for k in range(15)]
    print(f"chunk {it}: " + "  ".join(f"{n}={v}" for n, v in sorted(zip(names, row), key=lambda t: t[1])))

sys.exit(0)
print("---- step_fwd_tc timeline (cycles, CTA 0)")
import runpy as _r
g2 = _r.run_path(os.path.join(ROOT, "tools", "microbench_fwd.py"))
g2["run"](); torch.cuda.synchronize()
_lib.check(_lib.lib.sphnca_set_mlp_path(1)); g2["run"](); torch.cuda.synchronize()
_lib.lib.sphnca_debug_timeline_fwd(buf, 64 * 16)
nm = ["S start", "S loop done", "S Y_EMPTY ok", "S Y_FULL arr", "M Y_FULL ok", "M G1a issued", "M D1_EMPTY ok", "M G1b issued", "M H_FULL a", "M G2a issued", "M H_FULL b", "M G2b issued", "E D1_FULL a", "E D1_FULL b", "E D2_FULL", "E done"]
t0 = buf[0]
for it in range(3, 7):
    row = [buf[it * 16 + k] - t0 for k in range(16)]
    print(f"tile {it}: " + "  ".join(f"{n}={v}" for n, v in sorted(zip(nm, row), key=lambda t: t[1])))
