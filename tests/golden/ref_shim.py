"""Harness-only shim that lets the UNMODIFIED reference (`/root/reference/code`) run on the CPU
under Numba's CUDA simulator, so that golden vectors can be generated in the build container.

Nothing here is product code and nothing here ships to the GPU box; `make_golden.py` is the only
importer.  No reference file is touched -- three run-time patches are applied:

1. `cuda.libdevice.{sqrtf,rsqrtf,frcp_rn}` are provided (the simulator has no libdevice and
   reference `sphops/ops/kernels_impl.py:7` dereferences it at import time).
2. `FakeCUDAKernel.__call__` converts `torch.Tensor` arguments to numpy views: with raw tensors a
   scalar read such as `idx_current = cell_start[c]` (reference `operators_batch.py:65`) is a 0-d
   *view*, and the following `idx_current += 1` (`:75`) corrupts the table in place.
3. the module-global `cuda` of the reference's kernel modules is replaced by a proxy that resolves
   attributes against the simulator's per-launch context.  The simulator swaps that global around
   every nested device-function call from many Python threads at once, which races.
4. the simulator's `cuda.local.array` (np.empty) is made zero-filling.  Reference
   `operators_batch.py:122` accumulates into a local array it never clears (`volume_backward`,
   SURVEY.md 8a row a6), which is undefined on a GPU and garbage under np.empty; zero-fill is the
   evident intent and makes that kernel the true adjoint it is meant to be.
"""
import math
import os
import sys

os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")

import numpy as np
import torch
from numba import cuda
from numba.cuda.simulator import kernel as _simkernel
from numba.cuda.simulator import kernelapi as _kernelapi

REFERENCE_CODE = os.environ.get("SPHNCA_REFERENCE_CODE", "/root/reference/code")


class _Libdevice:
    """fp32 stand-ins for the three libdevice entry points the reference calls."""

    @staticmethod
    def sqrtf(a):
        return np.float32(math.sqrt(float(a)))

    @staticmethod
    def rsqrtf(a):
        a = float(a)
        return np.float32(1.0 / math.sqrt(a)) if a > 0.0 else np.float32(np.inf)

    @staticmethod
    def frcp_rn(a):
        return np.float32(1.0) / np.float32(a)


class _CudaProxy:
    """Stands in for the `cuda` module global inside the reference's kernel modules."""

    libdevice = _Libdevice

    def __getattr__(self, name):
        ctx = _simkernel._get_kernel_context()
        if ctx is not None and hasattr(ctx, name):
            return getattr(ctx, name)
        return getattr(cuda, name)


def _patch_simulator():
    cuda.libdevice = _Libdevice
    _kernelapi.FakeCUDAModule.libdevice = _Libdevice

    def zero_local_array(self, shape, dtype):
        from numba.core import types
        from numba.np import numpy_support
        if isinstance(dtype, types.Type):
            dtype = numpy_support.as_dtype(dtype)
        return np.zeros(shape, dtype)

    _kernelapi.FakeCUDALocal.array = zero_local_array

    orig_call = _simkernel.FakeCUDAKernel.__call__
    if getattr(orig_call, "_sphnca_patched", False):
        return

    def call(self, *args):
        conv = []
        for a in args:
            if isinstance(a, torch.Tensor):
                a = a.detach()
                a = a.numpy() if a.dim() > 0 else a.item()
            conv.append(a)
        return orig_call(self, *conv)

    call._sphnca_patched = True
    _simkernel.FakeCUDAKernel.__call__ = call


def load_reference():
    """Import the reference's `sphops` and `nca` modules (CPU, simulator) and return them."""
    _patch_simulator()
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)
    import sphops  # noqa: the reference's package
    from sphops.hashgrid import single_impl
    from sphops.ops import kernels_impl, operators_batch, func
    from sphops import utils as sph_utils

    proxy = _CudaProxy()
    for mod in (single_impl, kernels_impl, operators_batch, func, sph_utils):
        if getattr(mod, "cuda", None) is cuda:
            mod.cuda = proxy
    kernels_impl.sqrtf = _Libdevice.sqrtf
    import nca  # noqa: the reference's model file
    return sphops, nca
