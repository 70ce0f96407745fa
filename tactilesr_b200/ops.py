"""``torch.ops.tactilesr.*``: the C-ABI kernels registered as torch custom ops with autograd.

Each op is a thin wrapper: allocate the outputs with PyTorch's caching allocator, call the C entry point on the current
stream, nothing else.  Autograd formulas are registered with ``torch.library.register_autograd`` and fake (meta)
implementations with ``register_fake``; there is no CPU / CompositeImplicit fallback -- the ops exist for CUDA only.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor
from torch.library import custom_op, register_autograd, register_fake

from . import _lib


def _cuda_only(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _lib.TsrError(f"{what} runs on CUDA (sm_100a) tensors only; there is no CPU fallback")


# ---------------------------------------------------------------------------------------------------------------
# fused HR-label preparation + MSE (reference train/tactileSR_train.py:44-45,49)
# ---------------------------------------------------------------------------------------------------------------
@custom_op("tactilesr::mse_hr_loss", mutates_args=(), device_types="cuda")
def mse_hr_loss(out: Tensor, hr_raw: Tensor, scale_num: float) -> Tuple[Tensor, Tensor]:
    """(loss, d loss / d out): loss = mean((out - bilinear_resize(hr_raw / scale_num))^2)."""
    B, _, H, W = out.shape
    o = out.detach().contiguous().float()
    hr = hr_raw.detach().reshape(B, hr_raw.shape[-2], hr_raw.shape[-1]).contiguous().float()
    loss = torch.empty((), dtype=torch.float32, device=out.device)
    dout = torch.empty_like(o)
    ws = torch.empty(int(_lib.lib().tsr_mse_hr_workspace()), dtype=torch.uint8, device=out.device)
    _lib.call("tsr_mse_hr_loss", o.data_ptr(), hr.data_ptr(), float(scale_num), B, H, W, hr.shape[-2], hr.shape[-1],
              loss.data_ptr(), dout.data_ptr(), 1.0, ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    return loss, dout


@register_fake("tactilesr::mse_hr_loss")
def _(out, hr_raw, scale_num):
    return out.new_empty((), dtype=torch.float32), torch.empty_like(out, dtype=torch.float32)


def _mse_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])


def _mse_backward(ctx, g_loss, g_dout):
    (dout,) = ctx.saved_tensors
    return dout * g_loss, None, None


register_autograd("tactilesr::mse_hr_loss", _mse_backward, setup_context=_mse_setup)


# ---------------------------------------------------------------------------------------------------------------
# evaluation metrics (reference train/tactileSR_train.py:76-94, utility/tools.py:49-81)
# ---------------------------------------------------------------------------------------------------------------
@custom_op("tactilesr::eval_metrics", mutates_args=(), device_types="cuda")
def eval_metrics(out: Tensor, hr_raw: Tensor, scale_num: float, max_value: float, c1: float, c2: float) -> Tensor:
    """(3, B): per-sample sum of squared errors, PSNR, SSIM of out vs the prepared HR label."""
    B, C, H, W = out.shape
    o = out.detach().contiguous().float()
    hr = hr_raw.detach().reshape(B, hr_raw.shape[-2], hr_raw.shape[-1]).contiguous().float()
    res = torch.empty((3, B), dtype=torch.float32, device=out.device)
    _lib.call("tsr_eval_metrics", o.data_ptr(), hr.data_ptr(), float(scale_num), B, H, W, hr.shape[-2], hr.shape[-1],
              float(max_value), float(C * H), float(c1), float(c2), res[0].data_ptr(), res[1].data_ptr(),
              res[2].data_ptr(), _lib.stream_ptr())
    return res


@register_fake("tactilesr::eval_metrics")
def _(out, hr_raw, scale_num, max_value, c1, c2):
    return out.new_empty((3, out.shape[0]), dtype=torch.float32)


# ---------------------------------------------------------------------------------------------------------------
# the per-sample PSF model of tPSFNet.forward (reference model/tPSFNet.py:118-125) and its backward
# ---------------------------------------------------------------------------------------------------------------
@custom_op("tactilesr::psf_model", mutates_args=(), device_types="cuda")
def psf_model(alpha_beta: Tensor, depth: Tensor, want_aux: bool, f16: bool = False) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """alphaBeta (B,3), depth (B,100,100) -> HR (B,1,100,100), LR_degrade (B,1,4,4), psf (B,1,99,99), aux.
    aux: the forward -> backward hand-over of the tcgen05 path (per-row statistics of HR), empty when not requested or
    when the FFMA forward is selected (tsr_set_psf_mode(1)).  f16: the single-pass fp16 tensor-core kernels (16-bit
    precision modes, ~3e-4) instead of the fp32-accurate split-operand ones."""
    L = _lib.lib()
    B = alpha_beta.shape[0]
    dev = alpha_beta.device
    ab = alpha_beta.detach().contiguous().float()
    d = depth.detach().reshape(B, 100, 100).contiguous().float()
    HR = torch.empty((B, 1, 100, 100), dtype=torch.float32, device=dev)
    LRd = torch.empty((B, 1, 4, 4), dtype=torch.float32, device=dev)
    psf = torch.empty((B, 1, 99, 99), dtype=torch.float32, device=dev)
    st = _lib.stream_ptr()
    if L.tsr_get_psf_mode() != 1:
        aux = torch.empty((B, int(L.tsr_psf_aux_floats()) if want_aux else 0), dtype=torch.float32, device=dev)
        _lib.call("tsr_psf_forward_tc_f16" if (f16 or L.tsr_get_psf_mode() == 2) else "tsr_psf_forward_tc", ab.data_ptr(), d.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(),
                  aux.data_ptr() if want_aux else 0, B, st)
    else:
        aux = torch.empty((B, 0), dtype=torch.float32, device=dev)
        _lib.call("tsr_psf_forward_ffma", ab.data_ptr(), d.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(), B, st)
    return HR, LRd, psf, aux


@register_fake("tactilesr::psf_model")
def _(alpha_beta, depth, want_aux, f16=False):
    B = alpha_beta.shape[0]
    f = dict(dtype=torch.float32, device=alpha_beta.device)
    return (torch.empty((B, 1, 100, 100), **f), torch.empty((B, 1, 4, 4), **f), torch.empty((B, 1, 99, 99), **f),
            torch.empty((B, 1212 if want_aux else 0), **f))


@custom_op("tactilesr::psf_model_backward", mutates_args=(), device_types="cuda")
def psf_model_backward(alpha_beta: Tensor, depth: Tensor, HR: Tensor, aux: Tensor, dLRd: Tensor, dHR: Tensor,
                       dpsf: Tensor, f16: bool = False) -> Tensor:
    """d alphaBeta (B,3).  Absent upstream gradients are passed as empty tensors.  The training case (gradient through
    LR_degrade only, train/tPSFNet_train.py:186-189) with a forward hand-over runs the tcgen05 backward; everything else
    the general FFMA backward."""
    B = alpha_beta.shape[0]
    ab = alpha_beta.detach().contiguous().float()
    d = depth.detach().reshape(B, 100, 100).contiguous().float()
    dab = torch.empty((B, 3), dtype=torch.float32, device=ab.device)
    st = _lib.stream_ptr()
    keep = [t.detach().contiguous().float() for t in (dLRd, dHR, dpsf)]       # keep converted copies alive
    ptrs = [0 if t.numel() == 0 else t.data_ptr() for t in keep]
    if aux.numel() > 0 and ptrs[0] and not ptrs[1] and not ptrs[2]:
        _lib.call("tsr_psf_backward_tc_f16" if f16 else "tsr_psf_backward_tc", ab.data_ptr(), d.data_ptr(), aux.data_ptr(), ptrs[0],
                  dab.data_ptr(), B, st)
    else:
        _lib.call("tsr_psf_backward", ab.data_ptr(), d.data_ptr(), HR.detach().contiguous().data_ptr(), ptrs[0], ptrs[1],
                  ptrs[2], dab.data_ptr(), B, st)
    return dab


@register_fake("tactilesr::psf_model_backward")
def _(alpha_beta, depth, HR, aux, dLRd, dHR, dpsf, f16=False):
    return alpha_beta.new_empty((alpha_beta.shape[0], 3), dtype=torch.float32)


def _psf_setup(ctx, inputs, output):
    alpha_beta, depth, _, f16 = inputs
    HR, _, _, aux = output
    ctx.f16 = bool(f16) or _lib.lib().tsr_get_psf_mode() == 2
    # outputs that take no part in the loss must reach backward as None, not as materialised zeros: the training case
    # (gradient through LR_degrade only) is what selects the tcgen05 backward
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(alpha_beta, depth, HR, aux)


def _psf_backward(ctx, dHR, dLRd, dpsf, daux):
    alpha_beta, depth, HR, aux = ctx.saved_tensors
    if dHR is None and dLRd is None and dpsf is None:
        return None, None, None, None
    e = alpha_beta.new_empty((0,))
    dab = torch.ops.tactilesr.psf_model_backward(alpha_beta, depth, HR, aux, e if dLRd is None else dLRd,
                                                 e if dHR is None else dHR, e if dpsf is None else dpsf, ctx.f16)
    return dab, None, None, None


register_autograd("tactilesr::psf_model", _psf_backward, setup_context=_psf_setup)
