"""CPU restatement of the tPSFNet hot path -- TEST INFRASTRUCTURE ONLY.

Functional restatement of ``/root/reference/model/tPSFNet.py`` (constants :38-55, `_sdf`
:67-76, `tactilePSF` :78-83, `depth2tactile` :85-100, `forward` :102-127,
`degradation_process` :129-141) and of ``Trainer_tPSF.train_cal_loss``
(``train/tPSFNet_train.py:180-190``).  The per-sample python loop of the reference is
restated batched (one grouped dense correlation); the arithmetic per sample is the same
dense 99x99 correlation.  fp32 or fp64, CPU tensors, autograd for gradients.

Pinned against the unmodified reference by ``oracle/make_golden.py`` ->
``tests/golden/tpsf_*.npz``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def tpsf_layout():
    return [
        ("MLP_layer.1.weight", (256, 48)), ("MLP_layer.1.bias", (256,)),
        ("MLP_layer.3.weight", (1024, 256)), ("MLP_layer.3.bias", (1024,)),
        ("MLP_layer.5.weight", (256, 1024)), ("MLP_layer.5.bias", (256,)),
        ("MLP_layer.7.weight", (3, 256)), ("MLP_layer.7.bias", (3,)),
    ]


def make_state(seed: int, dtype=torch.float32) -> "OrderedDict[str, Tensor]":
    """Seeded MLP state: weights N(0, 0.03) (tPSFNet.py:65), biases U(+-1/sqrt(fan_in))."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape in tpsf_layout():
        if key.endswith("weight"):
            sd[key] = (torch.randn(shape, generator=g, dtype=torch.float32) * 0.03).to(dtype)
            fan_in = shape[1]
        else:
            sd[key] = ((torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) / math.sqrt(fan_in)).to(dtype)
    return sd


def psf_sdf(dtype=torch.float32) -> Tensor:
    """PSF_sdf (tPSFNet.py:42-46): distance to pixel (49,49), rescaled to [0, 10]."""
    r = torch.arange(99, dtype=torch.float64) - 49.0
    d = torch.sqrt(r[:, None] ** 2 + r[None, :] ** 2).to(torch.float32)  # reference fills an fp32 tensor
    d = 10 * (d - d.min()) / (d.max() - d.min())
    return d.to(dtype)[None, None]


def masking_sdf(dtype=torch.float32) -> Tensor:
    """LR_masking_sdf (tPSFNet.py:49-55): (4,4,100,100) distances to (12+25i, 12+25j),
    first index = row, jointly rescaled to [0, 10]."""
    x = torch.arange(100, dtype=torch.float64)
    out = torch.zeros(4, 4, 100, 100, dtype=torch.float32)
    for i in range(4):
        for j in range(4):
            out[i, j] = torch.sqrt((x[:, None] - (12 + 25 * i)) ** 2 + (x[None, :] - (12 + 25 * j)) ** 2).to(torch.float32)
    out = 10 * (out - out.min()) / (out.max() - out.min())
    return out.to(dtype)


def mlp(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    """MLP_layer (tPSFNet.py:26-36): Flatten, 48-256-1024-256-3 with ReLU, Softplus."""
    h = x.flatten(1)
    h = torch.relu(F.linear(h, sd["MLP_layer.1.weight"], sd["MLP_layer.1.bias"]))
    h = torch.relu(F.linear(h, sd["MLP_layer.3.weight"], sd["MLP_layer.3.bias"]))
    h = torch.relu(F.linear(h, sd["MLP_layer.5.weight"], sd["MLP_layer.5.bias"]))
    return F.softplus(F.linear(h, sd["MLP_layer.7.weight"], sd["MLP_layer.7.bias"]))


def tpsf_forward(sd: Dict[str, Tensor], x: Tensor, depth: Tensor):
    """tPSFNet.forward (tPSFNet.py:102-127) -> (HR, LR_degrade, psf, alphaBeta)."""
    assert x.shape[0] == depth.shape[0], "Batch size of LR tactile and depth should be the same!"
    B = x.shape[0]
    dt = x.dtype
    ab = mlp(sd, x)                                                     # (B,3)
    alpha, beta, gamma = ab[:, 0], ab[:, 1], ab[:, 2]
    sdf = psf_sdf(dt).to(x.device)
    psf = alpha[:, None, None, None] * torch.exp(-sdf ** 2 / (beta ** 2)[:, None, None, None])   # :83
    # depth2tactile (:85-100)
    dmax = depth.flatten(1).max(dim=1).values
    mask = depth > (dmax - 1e-3)[:, None, None, None]
    padded = F.pad(depth, (48, 48, 48, 48))
    conv = F.conv2d(padded.transpose(0, 1), psf, padding=1, groups=B).transpose(0, 1)  # (B,1,100,100)
    second_max = conv.detach().masked_fill(mask, 0).flatten(1).max(dim=1).values          # :95-97
    HR = torch.where(mask, second_max[:, None, None, None], conv)
    # degradation_process (:129-141)
    msdf = masking_sdf(dt).to(x.device)
    masking = torch.exp(-msdf[None] ** 2 / gamma[:, None, None, None, None])             # (B,4,4,100,100)
    mn = masking.flatten(1).min(dim=1).values[:, None, None, None, None]
    mx = masking.flatten(1).max(dim=1).values[:, None, None, None, None]
    masking = (masking - mn) / (mx - mn)
    LRd = (HR[:, 0, None, None] * masking).sum(dim=(-1, -2)) * 1e-4                       # (B,4,4)
    return HR, LRd[:, None], psf, ab[:, None, :]


def loss_and_grads(sd, LR_raw: Tensor, depth: Tensor, scale_num: float = 100.0):
    """Trainer_tPSF.train_cal_loss (train/tPSFNet_train.py:180-190) + backward.
    ``depth`` is (B,100,100) as the dataset yields it."""
    leaf = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in sd.items())
    LR = LR_raw / scale_num
    HR, LRd, psf, ab = tpsf_forward(leaf, LR, depth.unsqueeze(1))
    loss = torch.mean((LR[:, 2:3] - LRd) ** 2)
    grads = torch.autograd.grad(loss, list(leaf.values()))
    g = OrderedDict(zip(leaf.keys(), grads))
    return loss.detach(), (HR.detach(), LRd.detach(), psf.detach(), ab.detach()), g


def synthetic_depth(B: int, seed: int, dtype=torch.float32) -> Tensor:
    """Synthetic contact maps: binary discs / rectangles covering
    5-40 % of a frame drawn at 50x50 then bilinearly resized to 100x100 so edges are
    fractional (mirrors utility/raw_data_process.py:105-107); max is exactly 1.0 and every
    sample has contact and non-contact pixels."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(50, dtype=torch.float32), torch.arange(50, dtype=torch.float32), indexing="ij")
    out = torch.zeros(B, 50, 50, dtype=torch.float32)
    for b in range(B):
        kind = int(torch.randint(0, 2, (1,), generator=g))
        cy, cx = (torch.rand(2, generator=g, dtype=torch.float32) * 24 + 13).tolist()
        if kind == 0:
            r = float(torch.rand(1, generator=g, dtype=torch.float32) * 9 + 7)
            out[b] = (((yy - cy) ** 2 + (xx - cx) ** 2) <= r * r).float()
        else:
            hh, ww = (torch.rand(2, generator=g, dtype=torch.float32) * 10 + 6).tolist()
            out[b] = ((yy - cy).abs() <= hh).float() * ((xx - cx).abs() <= ww).float()
    d = F.interpolate(out[:, None], size=(100, 100), mode="bilinear", align_corners=False)[:, 0]
    return d.to(dtype)
