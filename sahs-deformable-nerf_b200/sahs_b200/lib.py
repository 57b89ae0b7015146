"""ctypes binding of libsahs_b200.so (include/sahs_b200.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, a RuntimeError is raised.
PyTorch is used only for device memory, streams and torch.distributed; every tensor crosses the ABI as a raw
device pointer.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SAHS_B200_LIB") or os.path.join(os.path.dirname(_PKG), "lib", "libsahs_b200.so")

_lib: Optional[C.CDLL] = None


class ModelSpecC(C.Structure):
    """struct sahs_model_spec"""
    _fields_ = [(n, C.c_int32) for n in (
        "xyz_L", "xyz_inc", "dir_L", "dir_inc", "use_ambient", "amb_dim", "amb_L", "amb_inc",
        "use_warp", "warp_layers", "warp_hidden", "warp_skip", "hyper_layers", "hyper_hidden", "hyper_skip",
        "trunk_layers", "trunk_hidden", "trunk_skip", "trunk_driving", "trunk_pose", "use_grid")]


class RngC(C.Structure):
    """struct sahs_rng"""
    _fields_ = [("seed", C.c_uint64), ("counter_dev", C.c_void_p), ("stream", C.c_uint32), ("reserved", C.c_uint32)]


class ConvDescC(C.Structure):
    """struct sahs_conv_desc"""
    _fields_ = [("in_", C.c_void_p), ("in_h", C.c_int), ("in_w", C.c_int), ("in_cs", C.c_int), ("cin", C.c_int),
                ("out_h", C.c_int), ("out_w", C.c_int), ("mode", C.c_int), ("up_shift", C.c_int), ("down_shift", C.c_int),
                ("packed_w", C.c_void_p), ("bias", C.c_void_p), ("ntile", C.c_int), ("ntiles", C.c_int),
                ("epilogue", C.c_int), ("aux", C.c_void_p), ("aux_cs", C.c_int), ("aux_shift", C.c_int),
                ("mean", C.c_void_p), ("rstd", C.c_void_p), ("out", C.c_void_p), ("out_cs", C.c_int), ("cout", C.c_int),
                ("t2_class", C.c_int)]


EXPORTS = (
    "sahs_abi_version", "sahs_last_error", "sahs_launch_count", "sahs_param_count", "sahs_get_ray_bundle",
    "sahs_coarse_z", "sahs_positional_encoding", "sahs_field_sizes", "sahs_pack_params", "sahs_fold_frame",
    "sahs_field_fwd", "sahs_composite_fwd", "sahs_composite_bwd", "sahs_sample_pdf_merge", "sahs_sample_pdf",
    "sahs_field_status", "sahs_debug_plan", "sahs_train_layout", "sahs_pack_params_train", "sahs_pack_params_bwd",
    "sahs_field_fwd_train", "sahs_field_bwd", "sahs_field_wgrad", "sahs_frame_postprocess", "sahs_weighted_sample", "sahs_stage1_loss",
    "sahs_adam_step", "sahs_adam_step_dev", "sahs_adam_advance", "sahs_counter_add", "sahs_weighted_sample_dev", "sahs_normal_map",
    "sahs_spade_conv", "sahs_spade_conv_status", "sahs_instnorm_stats", "sahs_avgpool2",
    "sahs_operand_format", "sahs_rng_fill", "sahs_coarse_z_rng", "sahs_composite_fwd_rng", "sahs_composite_bwd_rng", "sahs_sample_pdf_merge_rng",
)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"libsahs_b200.so not found at {LIB_PATH}; build it with `make -C sahs-deformable-nerf_b200` "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    spec_p = C.POINTER(ModelSpecC)
    sigs = {
        "sahs_abi_version": (C.c_int, []),
        "sahs_operand_format": (C.c_int, []),
        "sahs_spade_conv": (C.c_int, [C.POINTER(ConvDescC), vp]),
        "sahs_spade_conv_status": (C.c_int, [C.POINTER(C.c_int)]),
        "sahs_instnorm_stats": (C.c_int, [vp, i64, i32, i32, f32, vp, vp, vp, vp]),
        "sahs_avgpool2": (C.c_int, [vp, i32, i32, i32, vp, vp]),
        "sahs_last_error": (C.c_char_p, []),
        "sahs_launch_count": (C.c_uint64, []),
        "sahs_param_count": (C.c_int, [spec_p]),
        "sahs_get_ray_bundle": (C.c_int, [i32, i32, f32, f32, C.c_double, C.c_double, vp, vp, vp, vp]),
        "sahs_coarse_z": (C.c_int, [i32, i32, f32, f32, i32, vp, vp, vp, vp]),
        "sahs_positional_encoding": (C.c_int, [vp, i64, i32, i32, i32, vp, vp]),
        "sahs_field_sizes": (C.c_int, [spec_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "sahs_pack_params": (C.c_int, [spec_p, i32, C.POINTER(vp), vp, vp, vp]),
        "sahs_fold_frame": (C.c_int, [spec_p, i32, C.POINTER(vp), vp, vp, vp, vp]),
        "sahs_field_fwd": (C.c_int, [spec_p, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, i32, vp]),
        "sahs_composite_fwd": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
        "sahs_composite_bwd": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
        "sahs_sample_pdf_merge": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
        "sahs_sample_pdf": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]),
        "sahs_debug_plan": (C.c_int, [spec_p, i32, C.POINTER(C.c_int32), i32, C.POINTER(C.c_int32), i32,
                                      C.POINTER(C.c_int32), i32, C.POINTER(C.c_int32)]),
        "sahs_train_layout": (C.c_int, [spec_p, C.POINTER(C.c_int32), i32]),
        "sahs_pack_params_train": (C.c_int, [spec_p, i32, C.POINTER(vp), vp, vp]),
        "sahs_pack_params_bwd": (C.c_int, [spec_p, i32, C.POINTER(vp), vp, vp]),
        "sahs_field_fwd_train": (C.c_int, [spec_p, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]),
        "sahs_field_bwd": (C.c_int, [spec_p, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
        "sahs_field_wgrad": (C.c_int, [spec_p, i32, C.POINTER(vp), vp, vp, i32, vp, C.c_size_t, C.POINTER(C.c_uint64), vp]),
        "sahs_frame_postprocess": (C.c_int, [vp, i64, vp, vp, vp, vp]),
        "sahs_weighted_sample": (C.c_int, [vp, vp, i64, i32, i32, C.c_uint64, vp, vp, C.c_size_t, vp]),
        "sahs_adam_step": (C.c_int, [vp, vp, vp, vp, i64, C.c_double, C.c_double, C.c_double, C.c_double, i32, C.c_float,
                           vp]),
        "sahs_adam_step_dev": (C.c_int, [vp, vp, vp, vp, i64, vp, C.c_float, vp]),
        "sahs_adam_advance": (C.c_int, [vp, vp]),
        "sahs_counter_add": (C.c_int, [vp, C.c_uint64, vp]),
        "sahs_weighted_sample_dev": (C.c_int, [vp, vp, i64, i32, i32, C.c_uint64, vp, vp, vp, C.c_size_t, vp]),
        "sahs_normal_map": (C.c_int, [vp, i32, f32, f32, f32, f32, vp, i32, vp, vp]),
        "sahs_rng_fill": (C.c_int, [vp, i64, C.POINTER(RngC), i32, f32, vp]),
        "sahs_coarse_z_rng": (C.c_int, [i32, i32, f32, f32, i32, vp, C.POINTER(RngC), vp, vp]),
        "sahs_composite_fwd_rng": (C.c_int, [vp, vp, vp, f32, C.POINTER(RngC), vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp,
                                             vp]),
        "sahs_composite_bwd_rng": (C.c_int, [vp, vp, vp, f32, C.POINTER(RngC), vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp,
                                             vp, vp]),
        "sahs_sample_pdf_merge_rng": (C.c_int, [vp, vp, C.POINTER(RngC), i32, i32, i32, vp, vp, vp, vp]),
        "sahs_stage1_loss": (C.c_int, [vp, vp, vp, vp, i32, i32, C.c_float, C.c_float, i32, i32, vp, vp, vp, vp, vp, vp]),
        "sahs_field_status": (C.c_int, [C.POINTER(C.c_int)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sahs_last_error()
        raise RuntimeError(f"libsahs_b200 {what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("sahs_b200 ops need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("internal error: non-contiguous tensor handed to the C ABI")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous view/copy (inputs may be non-contiguous views)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def param_ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = ptr(t) if t is not None else None
    return arr
