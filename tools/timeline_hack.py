"""In-kernel clock64 timeline of step_mlp_tc_kernel (CTA 0), development aid."""
# needs a library built with the in-kernel time stamps: make -C sph-nca_b200/csrc clean && make -C sph-nca_b200/csrc EXTRA=-DSPHNCA_TIMELINE
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import runpy, torch
g = runpy.run_path(os.path.join(ROOT, "tools", "microbench_fwd.py"))
from sphops import _lib
_lib.check(_lib.lib.sphnca_set_mlp_path(1)); g["run"](); torch.cuda.synchronize()
buf = (C.c_longlong * (64 * 16))()
_lib.lib.sphnca_debug_timeline_mlp.argtypes = [C.POINTER(C.c_longlong), C.c_int32]
print("rc", _lib.lib.sphnca_debug_timeline_mlp(buf, 64 * 16))
nm = ["L Y_FULL arr", "M Y_FULL ok", "M G1a issued", "M G1b issued", "M G2a issued", "M G2b issued", "E D1a ok", "E epi a done", "E D1b ok", "E epi b done",
      "g1a ks0", "g1a ks1", "g1a ks2", "g1a ks3", "g1a ks4", "g1a commit"]
t0 = buf[2 * 16 + 0]
for it in range(3, 8):
    row = [buf[it * 16 + k] - t0 for k in range(16)]
    print(f"tile {it}: " + "  ".join(f"{n}={v}" for n, v in sorted(zip(nm, row), key=lambda t: t[1])))
