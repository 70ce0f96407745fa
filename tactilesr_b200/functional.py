"""Fused loss for the SR trainer: HR / HR_scale_num, bilinear resize to the SR resolution and
``nn.MSELoss`` (reference train/tactileSR_train.py:39,44-45,49) in one kernel that also emits d(out); and the batched
evaluation metrics of ``eval_func`` (reference train/tactileSR_train.py:76-94, utility/tools.py:49-81)."""
from __future__ import annotations

import torch

from . import _lib


class _MseHrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, hr_raw, scale_num):
        if not out.is_cuda:
            raise _lib.TsrError("mse_hr_loss runs on CUDA tensors only (no CPU fallback)")
        B, _, H, W = out.shape
        o = out.detach().contiguous().float()
        hr = hr_raw.detach().reshape(B, hr_raw.shape[-2], hr_raw.shape[-1]).contiguous().float()
        loss = torch.empty((), dtype=torch.float32, device=out.device)
        dout = torch.empty_like(o)
        ws = torch.empty(int(_lib.lib().tsr_mse_hr_workspace()), dtype=torch.uint8, device=out.device)
        _lib.call("tsr_mse_hr_loss", o.data_ptr(), hr.data_ptr(), float(scale_num), B, H, W, hr.shape[-2], hr.shape[-1],
                  loss.data_ptr(), dout.data_ptr(), 1.0, ws.data_ptr(), ws.numel(), _lib.stream_ptr())
        ctx.save_for_backward(dout)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dout,) = ctx.saved_tensors
        return dout * g, None, None


def mse_hr_loss(out: torch.Tensor, hr_raw: torch.Tensor, hr_scale_num: float = 10.0) -> torch.Tensor:
    """mean((out - resize(hr_raw / hr_scale_num, out.shape[-2:]))**2)."""
    return _MseHrFn.apply(out, hr_raw, hr_scale_num)


def eval_metrics(out: torch.Tensor, hr_raw: torch.Tensor, hr_scale_num: float = 10.0, max_value: float = 250.0,
                 C1: float = 0.01 ** 2, C2: float = 0.03 ** 2):
    """Per-sample metrics of one evaluation batch, computed by one kernel (one block per sample) instead of the
    reference's python loop over samples: returns device tensors ``(mse, psnr, ssim)`` -- ``mse`` the scalar
    ``nn.MSELoss(out, HR)`` of the batch, ``psnr`` / ``ssim`` (B,) exactly as ``calculationPSNR(out[i], HR[i], max_value)``
    / ``calculationSSIM(out[i], HR[i])`` on the reference's (1, H, W) slices (so the PSNR divisor is H, not H*W)."""
    if not out.is_cuda:
        raise _lib.TsrError("eval_metrics runs on CUDA tensors only (no CPU fallback)")
    B, _, H, W = out.shape
    o = out.detach().contiguous().float()
    hr = hr_raw.detach().reshape(B, hr_raw.shape[-2], hr_raw.shape[-1]).contiguous().float()
    res = torch.empty((3, B), dtype=torch.float32, device=out.device)
    _lib.call("tsr_eval_metrics", o.data_ptr(), hr.data_ptr(), float(hr_scale_num), B, H, W, hr.shape[-2], hr.shape[-1],
              float(max_value), float(out.shape[1] * H), float(C1), float(C2), res[0].data_ptr(), res[1].data_ptr(),
              res[2].data_ptr(), _lib.stream_ptr())
    return res[0].sum() / (B * out.shape[1] * H * W), res[1], res[2]
