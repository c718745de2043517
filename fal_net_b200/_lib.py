"""ctypes binding of libfalnet_sm100.so (the C ABI declared in include/falnet_b200.h).

There is NO CPU fallback: if the shared library is missing, or a tensor is not on a CUDA device,
the call fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libfalnet_sm100.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "falnet_b200.h")

_lib = None

_C = ctypes
_p, _i, _ll, _f, _u, _d = _C.c_void_p, _C.c_int, _C.c_longlong, _C.c_float, _C.c_uint, _C.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "faln_version": [],
    "faln_last_error": [],
    "faln_launch_count": [],
    "faln_med_fwd": [_p] * 11 + [_i] * 4 + [_ll, _u, _p],
    "faln_med_bwd": [_p] * 12 + [_i] * 4 + [_ll, _ll, _u, _p],
    "faln_med_disp": [_p] * 3 + [_i] * 4 + [_ll, _p],
    "faln_loss_partials_len": [],
    "faln_loss_rec_l1": [_p] * 6 + [_i] * 4 + [_p],
    "faln_loss_rec_l1_bwd": [_p] * 4 + [_f, _p, _p] + [_i] * 4 + [_p],
    "faln_loss_smooth": [_p, _p, _f, _p, _p] + [_i] * 6 + [_p],
    "faln_loss_smooth_bwd": [_p, _p, _f, _f, _p, _p] + [_i] * 7 + [_p],
    "faln_loss_mirror": [_p] * 6 + [_i] * 6 + [_p],
    "faln_loss_mirror_bwd": [_p] * 4 + [_f, _p, _p] + [_i] * 7 + [_p],
    "faln_mse_bf16": [_p, _p, _ll, _p, _p, _p],
    "faln_mse_bf16_bwd": [_p, _p, _ll, _f, _p, _p, _p],
    "faln_inv_rowmax": [_p, _p, _i, _ll, _p],
    "faln_occ_mask": [_p, _p, _p] + [_i] * 7 + [_p],
    "faln_adam": [_p] * 5 + [_ll] + [_f] * 5 + [_i, _f, _p],
    "faln_adam_dev": [_p] * 5 + [_ll, _p] + [_f] * 5 + [_p],
    "faln_adam_dev_range": [_p] * 5 + [_ll, _p] + [_f] * 5 + [_i, _p],
    "faln_nchw_to_nhwc_bf16": [_p, _p] + [_i] * 6 + [_p],
    "faln_nhwc_bf16_to_planar": [_p, _p] + [_i] * 5 + [_ll, _p],
    "faln_planar_to_nhwc_bf16": [_p, _p] + [_i] * 5 + [_ll, _p],
    "faln_conv3x3_fwd": [_p] * 8 + [_i] * 10 + [_ll, _i, _p],
    "faln_conv3x3_logits_disp": [_p] * 6 + [_i] * 7 + [_p],
    "faln_conv3x3_dgrad": [_p] * 5 + [_i] * 12 + [_p],
    "faln_conv3x3_wgrad": [_p] * 3 + [_i] * 10 + [_u, _p],
    "faln_conv3x3_wgrad_up2": [_p] * 3 + [_i] * 9 + [_p],
    "faln_conv3x3_wgrad_bias": [_p] * 4 + [_i] * 10 + [_u, _p],
    "faln_conv3x3_wgrad_multi": [_p, _i, _p],
    "faln_conv3x3_wgrad_up2_multi": [_p, _i, _p],
    "faln_f32_to_bf16": [_p, _p, _ll, _p],
    "faln_pack_dgrad_batched": [_p, _p, _p, _i, _i, _p],
    "faln_pack_dgrad_flat": [_p] * 3 + [_i, _i, _p],
    "faln_border_sum_nhwc": [_p, _p] + [_i] * 5 + [_p],
    "faln_upsample_nearest_bwd_nhwc": [_p] * 3 + [_i] * 8 + [_p],
    "faln_maxpool2_bwd_nhwc": [_p] * 3 + [_i] * 5 + [_p],
    "faln_channel_sum_nhwc": [_p, _p, _ll, _i, _i, _p],
    "faln_stem_conv": [_p] * 4 + [_i] * 6 + [_p],
    "faln_stem_conv_mma": [_p] * 4 + [_i] * 6 + [_p],
    "faln_stem_wgrad": [_p] * 4 + [_i] * 3 + [_p],
    "faln_upsample_nearest_nhwc": [_p, _p] + [_i] * 6 + [_p],
    "faln_maxpool2_nhwc": [_p, _p] + [_i] * 4 + [_p],
    "faln_stem_conv_tc": [_p] * 6 + [_i] * 6 + [_p],
    "faln_conv3x3_up2_fwd": [_p] * 4 + [_i] * 8 + [_p],
    "faln_conv3x3_up2_dgrad": [_p] * 4 + [_i] * 8 + [_p],
    "faln_pack_up2_weights": [_p] + [_ll] * 4 + [_p, _p] + [_i] * 4 + [_p],
    "faln_pack_up2_weights_multi": [_p, _i, _p],
    "faln_level_tables": [_p] * 4 + [_i] * 3 + [_p],
    "faln_fold_logit_conv": [_p, _p] + [_ll] * 4 + [_p, _p] + [_i] * 4 + [_p],
    "faln_fold_logit_conv_bwd": [_p] * 3 + [_ll] * 4 + [_p] + [_ll] * 4 + [_p, _i, _i, _p],
    "faln_const_channel_table": [_p] + [_ll] * 4 + [_i, _p, _i, _p],
    "faln_const_channel_wgrad": [_p, _p, _p] + [_ll] * 4 + [_i] * 5 + [_p],
    "faln_scalar_combine": [_p, _p, _i, _p, _p],
    "faln_scalar_scale": [_p, _p, _i, _p, _p],
    "faln_flip_resize_bilinear": [_p, _p] + [_i] * 6 + [_p],
    "faln_percentile_rows": [_p, _i, _ll, _ll, _d, _d, _p, _p],
    "faln_mspp_blend": [_p] * 4 + [_i] * 5 + [_f, _p],
    "faln_kitti_errors": [_p] * 3 + [_i] * 8 + [_d] * 4 + [_p],
    "faln_real_epe": [_p] * 3 + [_i] * 6 + [_p],
    "faln_rmse255": [_p] * 3 + [_i] * 3 + [_f] * 3 + [_p],
    "faln_maskr_noalign": [_p] * 6 + [_i] * 4 + [_ll, _p],
    "faln_pil_bicubic_ksize": [_i, _i],
    "faln_pil_bicubic_coeffs": [_i, _i, _i, _i, _p, _p],
    "faln_augment_crops_u8": [_p, _i, _p, _p, _p, _ll, _i, _p, _i, _i, _p],
}
_RESTYPES = {"faln_last_error": _C.c_char_p, "faln_launch_count": _C.c_longlong}


def declared_symbols() -> list[str]:
    """Every function name include/falnet_b200.h declares."""
    with open(HEADER_PATH) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(faln_[a-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing. Build it with `python -m fal_net_b200.build` (needs nvcc). "
                "fal_net_b200 has no CPU / eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _C.c_int)
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().faln_last_error().decode(errors="replace")
        raise RuntimeError(f"libfalnet_sm100 {what} failed (rc={rc}): {msg}")


def ptr(t: torch.Tensor | None):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("fal_net_b200 kernels need CUDA tensors (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(lib().faln_launch_count())


def f32c(t: torch.Tensor, name: str = "tensor") -> torch.Tensor:
    """Contiguous fp32 CUDA view of t (copies only if it has to)."""
    if not t.is_cuda:
        raise RuntimeError(f"{name}: fal_net_b200 kernels need CUDA tensors (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
