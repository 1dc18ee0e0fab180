"""ctypes binding of librnerf_b200.so (the C ABI declared in include/rnerf_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
Tensors are validated here (device, dtype, contiguity) before their raw pointers cross the ABI;
outputs are allocated by torch so ownership stays with torch's caching allocator.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# RN_B200_LIB: load another build of the same library (A/B timing of kernel variants on one GPU box)
LIB_PATH = os.environ.get("RN_B200_LIB") or os.path.join(_HERE, "librnerf_b200.so")

NUM_PARAM_TENSORS = 24
NUM_PARAMS = 595844
MLP_FLOP_PER_POINT = 1186816

_lib = None

# name -> (restype, argtypes); mirrors include/rnerf_b200.h one to one
_P = c_void_p
_SIGS = {
    "rn_version": (c_int, []),
    "rn_status_string": (c_char_p, [c_int]),
    "rn_last_cuda_error": (c_int, []),
    "rn_device_sm_count": (c_int, [POINTER(c_int)]),
    "rn_launch_count": (ctypes.c_ulonglong, []),
    "rn_set_flag": (c_int, [c_int, c_int]),
    "rn_get_flag": (c_int, [c_int, POINTER(c_int)]),
    "rn_debug_stream_lag": (c_int, [POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(c_int)]),
    "rn_debug_stream_busy": (c_int, [POINTER(ctypes.c_uint), c_int]),
    "rn_debug_stream_plan": (c_int, [c_int, POINTER(c_int), POINTER(c_int)]),
    "rn_prof_enable": (c_int, [c_int]),
    "rn_prof_collect": (c_int, [POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(c_int)]),
    "rn_se3_poses_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "rn_se3_poses_bwd": (c_int, [_P, _P, _P, c_int, c_int, _P, _P, _P, _P]),
    "rn_ray_directions": (c_int, [c_int, c_int, c_float, c_float, c_float, _P, _P]),
    "rn_get_rays": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "rn_get_rays_bwd": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P]),
    "rn_raygen_fwd": (c_int, [_P, _P, c_int64, _P, c_int, c_int, c_int, c_float, c_float, c_float, _P, _P, _P]),
    "rn_raygen_bwd": (c_int, [_P, _P, c_int64, _P, c_int, c_int, c_int, c_float, c_float, c_float, _P, _P, _P, _P]),
    "rn_raygen_se3_fwd": (c_int, [_P, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                  c_float, _P, _P, _P]),
    "rn_raygen_se3_bwd": (c_int, [_P, _P, c_int64, _P, _P, c_int, c_int, c_int, c_int, c_float, c_float, c_float,
                                  _P, _P, _P, _P, _P]),
    "rn_pixel_gather": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P, _P]),
    "rn_image_metrics_scratch_bytes": (ctypes.c_size_t, [c_int, c_int, c_int]),
    "rn_image_metrics": (c_int, [_P, _P, c_int, c_int, c_int, ctypes.c_float, ctypes.c_float, _P, _P, _P, _P]),
    "rn_pose_noise": (c_int, [_P, c_int, _P, _P, _P, ctypes.c_float, ctypes.c_float, ctypes.c_double, _P, _P, _P]),
    "rn_pose_errors": (c_int, [_P, _P, c_int, _P, _P]),
    "rn_pixel_gather_u8": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P, _P]),
    "rn_stratified_fwd": (c_int, [_P, _P, c_int64, _P, c_int, _P, _P, _P, _P]),
    "rn_points_fwd": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P]),
    "rn_points_bwd": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P]),
    "rn_sample_pdf_fwd": (c_int, [_P, _P, c_int64, c_int, _P, c_int64, c_int, _P, _P, _P]),
    "rn_sample_hierarchical_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P, c_int64, c_int, _P, _P, _P, _P]),
    "rn_composite_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P, _P, _P]),
    "rn_composite_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rn_mse_loss_fwd_bwd": (c_int, [_P, _P, c_int64, c_float, _P, _P, _P]),
    "rn_mse2_loss_fwd_bwd": (c_int, [_P, _P, _P, c_int64, _P, _P, _P, _P]),
    "rn_mlp_packed_weight_bytes": (c_size_t, []),
    "rn_mlp_pack_weights": (c_int, [POINTER(c_void_p), _P, _P]),
    "rn_mlp_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "rn_mlp_infer_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "rn_view_dirs": (c_int, [_P, c_int64, _P, _P]),
    "rn_render_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "rn_render_view": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_float, c_float, c_float, c_int64, c_int64, c_int64, c_int, c_int,
                               _P, c_int, _P, c_int, c_int, _P, _P, _P, _P, POINTER(c_int64), _P]),
    "rn_mlp_fwd": (c_int, [_P, _P, _P, c_int64, c_int, _P, c_int, _P, _P]),
    "rn_mlp_bwd": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P]),
    "rn_head_act_fwd": (c_int, [_P, c_int64, _P, _P, _P]),
    "rn_head_act_bwd": (c_int, [_P, c_int64, _P, _P, _P, _P]),
    "rn_posenc_fwd": (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    "rn_posenc_bwd": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P]),
    "rn_gemm_bf16": (c_int, [c_int, _P, c_int64, _P, c_int64, _P, c_int64, c_int64, c_int, c_int64, _P, c_int, _P,
                             _P, _P, _P, c_size_t, _P]),
    "rn_gemm_scratch_bytes": (c_size_t, []),
    "rn_clip_adam_step": (c_int, [_P, _P, _P, _P, c_int64, POINTER(c_int64), POINTER(c_float), c_int, c_float, c_float,
                                  c_float, c_float, c_int, _P, _P, c_float, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def lib():
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python robust-nerf_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)      # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = handle
        # RN_FLAGS="4=0,5=1": kernel-variant flags applied at load time (A/B timing and running the test suite under a
        # variant; see rn_set_flag in csrc/mlp.cu)
        for kv in filter(None, os.environ.get("RN_FLAGS", "").split(",")):
            k, v = kv.split("=")
            check(handle.rn_set_flag(int(k), int(v)), f"rn_set_flag({kv})")
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        l = lib()
        msg = l.rn_status_string(status).decode()
        extra = f" (cudaError {l.rn_last_cuda_error()})" if status == 2 else ""
        raise RuntimeError(f"rnerf_b200 {what}: {msg}{extra}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    """Validate a tensor that is about to cross the ABI; returns a contiguous version."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the B200 path has no CPU fallback, move it to a CUDA device")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"{name} lives on {t.device} but the current CUDA device is {torch.cuda.current_device()}: "
                           "kernels launch on the current device's stream (use torch.cuda.device(...) / set_device)")
    return t.contiguous()


def call(name: str, *args) -> None:
    """Invoke a C-ABI entry point.  The kernels launch on the CURRENT device's stream (`stream_ptr()`), and every
    wrapper validates its tensors with `require_cuda`, which refuses tensors that live on another device."""
    check(getattr(lib(), name)(*args), name)
