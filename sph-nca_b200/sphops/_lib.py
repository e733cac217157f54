"""ctypes binding of libsphnca.so (C ABI declared in include/sphnca.h).

The library is built in-tree by `make -C sph-nca_b200/csrc` (or `__graft_entry__.build()`).
There is NO fallback: if the shared library is missing, importing `sphops` fails, and every
operator refuses non-CUDA tensors.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPHNCA_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libsphnca.so"))


class SphncaError(RuntimeError):
    pass


class CGrid(C.Structure):
    """struct sphnca_grid (include/sphnca.h)."""
    _fields_ = [
        ("dim", C.c_int32), ("num_sections", C.c_int32), ("num_particles", C.c_int64),
        ("cells_per_section", C.c_int64), ("dims", C.c_int64 * 3), ("h", C.c_double),
        ("cell_off", C.c_void_p), ("sorted_id", C.c_void_p), ("sorted_cell", C.c_void_p),
        ("period", C.c_double * 3),
        ("image_volume", C.c_void_p), ("band_lo", C.c_double * 3), ("band_hi", C.c_double * 3),
    ]


class CStepParams(C.Structure):
    """struct sphnca_step_params (include/sphnca.h)."""
    _fields_ = [
        ("W1", C.c_void_p), ("b1", C.c_void_p), ("W2", C.c_void_p), ("b2", C.c_void_p),
        ("update_rule", C.c_int32), ("use_alpha", C.c_int32),
        ("perception_scale", C.c_float), ("orig_rate", C.c_float),
        ("dy_interleaved", C.c_int32),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"libsphnca.so not found at {LIB_PATH}: build it with `make -C sph-nca_b200/csrc` "
            "(or __graft_entry__.build()).  sphops has no CPU / eager fallback.")
    lib = C.CDLL(LIB_PATH)
    P, I32, I64, D, G = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.POINTER(CGrid)
    SP = C.POINTER(CStepParams)
    sigs = {
        "sphnca_abi_version": ([], I32),
        "sphnca_last_error": ([], C.c_char_p),
        "sphnca_hashgrid_workspace_bytes": ([I64, I32, I64, C.POINTER(C.c_size_t)], I32),
        "sphnca_hashgrid_build": ([P, I64, I32, C.POINTER(I64), I32, D, C.POINTER(I64),
                                   P, P, P, P, P, P, P, P, C.c_size_t, P], I32),
        "sphnca_pack_xv": ([G, P, P, P, P], I32),
        "sphnca_pack_rows": ([G, P, I32, P, P], I32),
        "sphnca_unpack_rows": ([G, P, I32, P, P], I32),
        "sphnca_pack_bytes": ([G, P, P, P], I32),
        "sphnca_volume_fwd": ([G, P, P, P, P], I32),
        "sphnca_volume_bwd": ([G, P, P, P, P], I32),
        "sphnca_volume_images": ([G, P, P, P], I32),
        "sphnca_count_fwd": ([G, P, P, P], I32),
        "sphnca_gradient_fwd": ([G, P, P, I32, P, P], I32),
        "sphnca_gradient_bwd": ([G, P, P, I32, P, P], I32),
        "sphnca_blur_fwd": ([G, P, P, I32, P, P], I32),
        "sphnca_blur_bwd": ([G, P, P, I32, P, P], I32),
        "sphnca_divergence_fwd": ([G, P, P, I32, P, P], I32),
        "sphnca_divergence_bwd": ([G, P, P, I32, P, P], I32),
        "sphnca_step_fwd": ([G, P, P, P, SP, P, P, P, P, P, P, P], I32),
        "sphnca_step_perceive_staged": ([G, P, P, SP, P, P, P], I32),
        "sphnca_stage_fit": ([G, P, P], I32),
        "sphnca_pairs_count": ([G, P, P, P], I32),
        "sphnca_pairs_fill": ([G, P, P, P, P, P, P, P, P, P, P, P], I32),
        "sphnca_step_mask_pairs": ([G, P, P, P, P, P, I32, P, P, P, P], I32),
        "sphnca_step_perceive_pairs_frame": ([G, P, P, P, P, P, P, P, P, P, SP, P, P, P], I32),
        "sphnca_step_perceive_pairs": ([G, P, P, P, P, P, P, P, SP, P, P, P], I32),
        "sphnca_step_adjoint_pairs": ([G, P, P, P, P, P, P, P, P, SP, P, P], I32),
        "sphnca_step_adjoint_pairs_fired": ([G, P, P, P, P, P, P, P, P, P, P, SP, P, P], I32),
        "sphnca_step_mask": ([G, P, P, P, P, P], I32),
        "sphnca_step_mask_static": ([I64, P, P, P], I32),
        "sphnca_step_perceive": ([G, P, P, SP, P, P, P], I32),
        "sphnca_step_mlp": ([I64, P, P, P, SP, P, P, P, P, P, P], I32),
        "sphnca_step_bwd": ([I64, P, P, P, P, P, P, SP, P, P, P, P, P, P], I32),
        "sphnca_mlp_wgrad_workspace_bytes": ([I64, C.POINTER(C.c_size_t)], I32),
        "sphnca_mlp_wgrad": ([I64, P, P, P, P, P, SP, P, P, P, P, P, C.c_size_t, P], I32),
        "sphnca_step_adjoint": ([G, P, P, P, SP, P, P], I32),
        "sphnca_fire_list_workspace_bytes": ([I64, C.POINTER(C.c_size_t)], I32),
        "sphnca_fire_list": ([I64, P, P, I32, P, P, P, P, P, P, I32, P, C.c_size_t, P], I32),
        "sphnca_fire_list_draw": ([G, P, C.c_float, P, P, I32, P, P, P, P, P, P, I32, P, C.c_size_t, P], I32),
        "sphnca_step_perceive_pairs_fired": ([G, P, P, P, P, P, P, P, SP, P, P, P], I32),
        "sphnca_step_mlp_list": ([I64, P, P, P, P, SP, P, P, P, P, P, P, P, I32, P], I32),
        "sphnca_step_bwd_list": ([I64, P, P, P, P, P, P, P, SP, P, P, P, P], I32),
        "sphnca_step_bwd_unfired": ([I64, P, P, P, P, P], I32),
        "sphnca_fire_bytes_counter": ([I64, P, I64, C.c_float, P, P, P, P, P], I32),
        "sphnca_step_fwd_fired": ([G, P, P, P, P, P, P, P, SP, P, I32, I64, P, P, P, P, P, P, P, P, P, P, C.c_size_t, P], I32),
        "sphnca_mlp_wgrad_list": ([I64, P, P, P, P, SP, P, P, P, P, P, C.c_size_t, P], I32),
        "sphnca_mlp_wgrad_list_partial": ([I64, P, P, P, P, SP, I32, P, C.c_size_t, P], I32),
        "sphnca_mlp_wgrad_reduce": ([I64, SP, P, P, P, P, P, C.c_size_t, P], I32),
        "sphnca_launch_count": ([], I64),
        "sphnca_set_mlp_path": ([I32], I32),
        "sphnca_get_mlp_path": ([], I32),
        "sphnca_set_sweep_path": ([I32], I32),
    }
    missing = []
    for name, (args, res) in sigs.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.argtypes = args
        fn.restype = res
    return lib, tuple(sigs), tuple(missing)


lib, DECLARED_SYMBOLS, MISSING_SYMBOLS = _load()

# optional implementation switches for A/B runs (see include/sphnca.h)
# SPHNCA_SWEEP_PATH set: the fused step always calls sphnca_step_perceive (no per-grid choice of the staged kernel)
FORCE_SWEEP_PATH = os.environ.get("SPHNCA_SWEEP_PATH")
# SPHNCA_PAIRS=0: never build pair lists (the fused step then sweeps the cells every step)
USE_PAIR_LISTS = os.environ.get("SPHNCA_PAIRS", "1") != "0"
for _env, _fn in (("SPHNCA_MLP_PATH", "sphnca_set_mlp_path"), ("SPHNCA_SWEEP_PATH", "sphnca_set_sweep_path")):
    if os.environ.get(_env) is not None and _fn not in MISSING_SYMBOLS:
        getattr(lib, _fn)(int(os.environ[_env]))


def check(rc):
    if rc != 0:
        raise SphncaError(f"libsphnca error {rc}: {lib.sphnca_last_error().decode()}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: sphops runs only on the GPU (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype}); the reference's dtype map "
                        "(func.py:12-16) also lists fp16/fp64 but only fp32 is implemented")
    if t.device.index != torch.cuda.current_device():
        # kernels are launched on the CURRENT device's current stream (stream_ptr): a tensor of another device would be
        # dereferenced in the wrong context (an illegal address at best)
        raise RuntimeError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                           "call torch.cuda.set_device(...) / use `with torch.cuda.device(t.device)` around sphops calls")
    return t.detach().contiguous()
