"""Fused loss for the SR trainer: HR / HR_scale_num, bilinear resize to the SR resolution and
``nn.MSELoss`` (reference train/tactileSR_train.py:39,44-45,49) in one kernel that also emits d(out)."""
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
