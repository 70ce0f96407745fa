"""ctypes binding of the C-ABI library ``libtactilesr_b200.so`` (declared in include/tactilesr_b200.h).

There is no CPU fallback and no torch/cuDNN fallback: if the library is missing or a call fails the
error is raised.  Pointers are passed as integers (``tensor.data_ptr()``), streams as the raw
``cudaStream_t`` of ``torch.cuda.current_stream()``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtactilesr_b200.so")

_P, _I, _L, _F, _Z = c_void_p, c_int, c_longlong, c_float, c_size_t

# name -> (restype, argtypes).  Every entry here must be declared in include/tactilesr_b200.h
# (tests/test_abi.py checks both directions).
SIGNATURES = {
    "tsr_last_error": (c_char_p, []),
    "tsr_version": (_I, []),
    "tsr_check_device": (_I, []),
    "tsr_set_f16_overflow_flag": (None, [_P]),
    "tsr_launch_count": (_L, []),
    "tsr_launch_count_reset": (None, []),
    "tsr_launch_count_add": (None, [_L]),
    # fp32 convolutions
    "tsr_pack_conv_weight_f32": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "tsr_conv2d_f32": (_I, [_P, _I, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "tsr_conv2d_wgrad_f32_workspace": (_Z, [_I, _I, _I, _I, _I, _I]),
    "tsr_conv2d_wgrad_f32": (_I, [_P, _I, _P, _I, _P, _P, _Z, _I, _I, _I, _I, _I, _I, _I, _P]),
    "tsr_colsum_workspace": (_Z, [_L, _I]),
    "tsr_colsum": (_I, [_P, _I, _I, _L, _I, _P, _P, _Z, _I, _P]),
    "tsr_head_fwd": (_I, [_P, _L, _P, _P, _I, _I, _I, _I, _I, _P]),
    "tsr_head_wgrad_workspace": (_Z, [_I]),
    "tsr_head_wgrad": (_I, [_P, _L, _P, _I, _I, _P, _P, _Z, _I, _I, _I, _P]),
    "tsr_tail_fwd": (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "tsr_tail_dgrad": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "tsr_tail_dgrad_masked": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    "tsr_tail_wgrad_workspace": (_Z, [_I, _I, _I, _I]),
    "tsr_tail_wgrad": (_I, [_P, _I, _I, _P, _P, _P, _P, _Z, _I, _I, _I, _I, _I, _I, _P]),
    # elementwise / reductions
    "tsr_bn_workspace": (_Z, [_L, _I]),
    "tsr_bn_train_stats": (_I, [_P, _I, _I, _L, _I, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _Z, _P]),
    "tsr_bn_finalize_partials": (_I, [_P, _I, _I, _L, _I, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P]),
    "tsr_bn_bwd_finalize_partials": (_I, [_P, _I, _I, _L, _I, _P, _P, _P, _P, _I, _P, _P, _I, _P]),
    "tsr_bn_backward_apply": (_I, [_P, _I, _P, _I, _P, _I, _I, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P]),
    "tsr_bn_eval_coeffs": (_I, [_I, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P]),
    "tsr_bn_apply": (_I, [_P, _I, _I, _P, _P, _P, _I, _I, _L, _I, _I, _P, _I, _P]),
    "tsr_bn_backward_workspace": (_Z, [_L, _I]),
    "tsr_bn_backward": (_I, [_P, _I, _P, _I, _P, _I, _I, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _I, _P, _Z, _P]),
    "tsr_relu_backward": (_I, [_P, _I, _P, _I, _P, _I, _I, _L, _I, _P]),
    "tsr_copy_channels": (_I, [_P, _I, _I, _P, _I, _I, _L, _I, _P]),
    "tsr_nchw_to_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "tsr_nhwc_to_nchw": (_I, [_P, _I, _I, _P, _I, _I, _I, _P]),
    "tsr_mse_hr_workspace": (_Z, []),
    "tsr_mse_hr_loss": (_I, [_P, _P, _F, _I, _I, _I, _I, _I, _P, _P, _F, _P, _Z, _P]),
    "tsr_eval_metrics": (_I, [_P, _P, _F, _I, _I, _I, _I, _I, _F, _F, _F, _F, _P, _P, _P, _P]),
    "tsr_adam_step": (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _L, _F, _P]),
    "tsr_adam_step_dev": (_I, [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _P]),
    # tPSFNet
    "tsr_sgemm_strided": (_I, [_P, _L, _L, _P, _L, _L, _P, _L, _I, _I, _I, _P, _I, _I, _P]),
    "tsr_linear_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "tsr_linear_bwd_workspace": (_Z, [_I, _I, _I]),
    "tsr_linear_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "tsr_psf_forward": (_I, [_P, _P, _P, _P, _P, _I, _P]),
    "tsr_psf_forward_tc": (_I, [_P, _P, _P, _P, _P, _P, _I, _P]),
    "tsr_psf_backward_tc": (_I, [_P, _P, _P, _P, _P, _I, _P]),
    "tsr_psf_forward_tc_f16": (_I, [_P, _P, _P, _P, _P, _P, _I, _P]),
    "tsr_psf_backward_tc_f16": (_I, [_P, _P, _P, _P, _P, _I, _P]),
    "tsr_psf_aux_floats": (_Z, []),
    "tsr_psf_forward_ffma": (_I, [_P, _P, _P, _P, _P, _I, _P]),
    "tsr_set_psf_mode": (None, [_I]),
    "tsr_get_psf_mode": (_I, []),
    "tsr_psf_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _P]),
    # tensor-core (tcgen05) convolutions
    "tsr_pack_conv_weight_bf16": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "tsr_pack_conv_weight_f16": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "tsr_pack_conv_weights_multi": (_I, [_P, _I, _L, _P]),
    "tsr_pack_conv_weight_folded": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "tsr_conv2d_tc": (_I, [_P, _I, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _Z, _P, _P, _I, _P]),
    "tsr_conv2d_tc_stat_rows": (_I, []),
    "tsr_conv2d_tc_workspace": (_Z, [_I, _I, _I, _I, _I, _I]),
    "tsr_conv2d_wgrad_tc": (_I, [_P, _I, _P, _I, _P, _P, _Z, _I, _I, _I, _I, _I, _I, _I, _P]),
    "tsr_conv2d_wgrad_tc_workspace": (_Z, [_I, _I, _I, _I, _I, _I]),
    "tsr_conv2d_wgrad_tc_x": (_I, [_P, _I, _I, _P, _I, _P, _P, _Z, _I, _I, _I, _I, _I, _I, _I, _P]),
    "tsr_conv2d_tc2": (_I, [_P, _P]),
    "tsr_conv2d_tc2_stat_rows": (_I, []),
    "tsr_conv2d_tc2_debug": (None, [_P]),
    "tsr_pack_conv_weight_dual_elems": (_Z, [_I]),
    "tsr_pack_conv_weight_dual": (_I, [_P, _P, _P, _I, _I, _P]),
    "tsr_set_tc_desc_mode": (None, [_I]),
    "tsr_get_tc_desc_mode": (_I, []),
}



class ConvSrc(ctypes.Structure):
    """TsrConvSrc of include/tactilesr_b200.h."""
    _fields_ = [("inp", c_void_p), ("w_packed", c_void_p), ("in_ld", c_int), ("Cin", c_int), ("KS", c_int), ("pad_", c_int)]


class ConvTc2(ctypes.Structure):
    """TsrConvTc2 of include/tactilesr_b200.h (argument block of tsr_conv2d_tc2)."""
    _fields_ = [("src", ConvSrc * 2), ("bias", c_void_p), ("residual", c_void_p), ("out", c_void_p), ("out2_bf16", c_void_p),
                ("stat", c_void_p), ("aux", c_void_p), ("aux_scale", c_void_p), ("aux_shift", c_void_p),
                ("nsrc", c_int), ("dual_fwd", c_int), ("res_ld", c_int), ("out_ld", c_int), ("out2_ld", c_int),
                ("stat_ld", c_int), ("aux_ld", c_int), ("B", c_int), ("H", c_int), ("W", c_int), ("Cout", c_int),
                ("flags", c_int)]


TC2_RELU, TC2_F16, TC2_MASK, TC2_BNB, TC2_BNB_RELU, TC2_AUX_F16, TC2_STAT_PRECLEARED = 1, 2, 8, 16, 32, 64, 128


def conv_tc2(srcs, out, out_ld, B, H, W, Cout, flags=0, bias=0, residual=0, res_ld=0, out2=0, out2_ld=0, stat=0, stat_ld=0,
             aux=0, aux_ld=0, aux_scale=0, aux_shift=0, dual_fwd=0, stream=None):
    """tsr_conv2d_tc2 with srcs = [(in_ptr, in_ld, Cin, KS, w_packed_ptr), ...] (one or two sources)."""
    a = ConvTc2()
    for i, (ip, ild, cin, ks, wp) in enumerate(srcs):
        a.src[i].inp, a.src[i].in_ld, a.src[i].Cin, a.src[i].KS, a.src[i].w_packed = ip, ild, cin, ks, wp
    a.nsrc, a.dual_fwd = len(srcs), dual_fwd
    a.bias, a.residual, a.res_ld = bias or None, residual or None, res_ld
    a.out, a.out_ld, a.out2_bf16, a.out2_ld = out, out_ld, out2 or None, out2_ld
    a.stat, a.stat_ld = stat or None, stat_ld
    a.aux, a.aux_ld, a.aux_scale, a.aux_shift = aux or None, aux_ld, aux_scale or None, aux_shift or None
    a.B, a.H, a.W, a.Cout, a.flags = B, H, W, Cout, flags
    call("tsr_conv2d_tc2", ctypes.addressof(a), stream_ptr() if stream is None else stream)


_lib = None


class TsrError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the library (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TsrError(
                f"{LIB_PATH} not found: build it with `python -m tactilesr_b200.csrc.build` "
                "(or __graft_entry__.build()); tactilesr_b200 has no CPU / PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().tsr_last_error().decode("utf-8", "replace")
        raise TsrError(f"{what or 'tactilesr_b200'} failed (code {rc}): {msg}")


def call(name: str, *args):
    """Call an int-returning entry point and raise on a non-zero code."""
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        check(rc, name)


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().tsr_launch_count())
