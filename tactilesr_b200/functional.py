"""Fused loss for the SR trainer: HR / HR_scale_num, bilinear resize to the SR resolution and
``nn.MSELoss`` (reference train/tactileSR_train.py:39,44-45,49) in one kernel that also emits d(out); and the batched
evaluation metrics of ``eval_func`` (reference train/tactileSR_train.py:76-94, utility/tools.py:49-81)."""
from __future__ import annotations

import torch

from . import _lib
from . import ops  # noqa: F401  (registers torch.ops.tactilesr.*)


def mse_hr_loss(out: torch.Tensor, hr_raw: torch.Tensor, hr_scale_num: float = 10.0) -> torch.Tensor:
    """mean((out - resize(hr_raw / hr_scale_num, out.shape[-2:]))**2)   (torch.ops.tactilesr.mse_hr_loss)."""
    if not out.is_cuda:
        raise _lib.TsrError("mse_hr_loss runs on CUDA tensors only (no CPU fallback)")
    return torch.ops.tactilesr.mse_hr_loss(out, hr_raw, float(hr_scale_num))[0]


def eval_metrics(out: torch.Tensor, hr_raw: torch.Tensor, hr_scale_num: float = 10.0, max_value: float = 250.0,
                 C1: float = 0.01 ** 2, C2: float = 0.03 ** 2):
    """Per-sample metrics of one evaluation batch, computed by one kernel (one block per sample) instead of the
    reference's python loop over samples: returns device tensors ``(mse, psnr, ssim)`` -- ``mse`` the scalar
    ``nn.MSELoss(out, HR)`` of the batch, ``psnr`` / ``ssim`` (B,) exactly as ``calculationPSNR(out[i], HR[i], max_value)``
    / ``calculationSSIM(out[i], HR[i])`` on the reference's (1, H, W) slices (so the PSNR divisor is H, not H*W)."""
    if not out.is_cuda:
        raise _lib.TsrError("eval_metrics runs on CUDA tensors only (no CPU fallback)")
    B, C, H, W = out.shape
    res = torch.ops.tactilesr.eval_metrics(out, hr_raw, float(hr_scale_num), float(max_value), float(C1), float(C2))
    return res[0].sum() / (B * C * H * W), res[1], res[2]
