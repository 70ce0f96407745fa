"""CPU restatement of the TactileSR hot path -- TEST INFRASTRUCTURE ONLY.

Functional (state-dict driven) restatement of
``/root/reference/model/tactileSR_model.py`` (TactileSR :18-98, TactileSRCNN
:101-153, MSRB :157-214, ResBlock :216-225), of ``Trainer_tactileSR.train_cal_loss``
(``train/tactileSR_train.py:41-51``) and of ``torch.optim.Adam`` as constructed at
``train/tactileSR_train.py:212``.  Works in fp32 or fp64 on CPU tensors; gradients
come from torch autograd over this restatement.

Pinned against the unmodified reference by ``oracle/make_golden.py`` ->
``tests/golden/tactilesr_*.npz`` (see tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# parameter construction (key layout = SURVEY.md Appendix A; reference
# tactileSR_model.py:22-65 registration order)
# --------------------------------------------------------------------------------------
def _bn_entries(prefix: str, c: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    return [
        (prefix + ".weight", (c,), "bn_w"),
        (prefix + ".bias", (c,), "bn_b"),
        (prefix + ".running_mean", (c,), "bn_rm"),
        (prefix + ".running_var", (c,), "bn_rv"),
        (prefix + ".num_batches_tracked", (), "bn_nbt"),
    ]


def _msrb_entries(prefix: str, n: int = 64):
    e = []
    for name, c, k in (("conv_3_1", n, 3), ("conv_5_1", n, 5), ("conv_3_2", 2 * n, 3), ("conv_5_2", 2 * n, 5)):
        e.append((f"{prefix}.{name}.0.weight", (c, c, k, k), "conv_w"))
        e.append((f"{prefix}.{name}.0.bias", (c,), "conv_b"))
        e += _bn_entries(f"{prefix}.{name}.1", c)
    e.append((f"{prefix}.confusion.weight", (n, 4 * n, 1, 1), "conv_w"))
    e.append((f"{prefix}.confusion.bias", (n,), "conv_b"))
    return e


def tactilesr_layout(seqsCnt: int = 1, axisCnt: int = 3, n_msrb: int = 6, n_res: int = 1):
    """(key, shape, kind) for every state_dict entry of TactileSR, in registration order
    (reference tactileSR_model.py:29-63)."""
    e = []
    for i in range(n_msrb):
        e += _msrb_entries(f"patternFeatureExtra_layer.{i}")
    for i in range(n_res):
        for cv in ("conv1", "conv2"):
            e.append((f"forceFeatureExtra_layer.{i}.{cv}.weight", (64, 64, 3, 3), "conv_w"))
            e.append((f"forceFeatureExtra_layer.{i}.{cv}.bias", (64,), "conv_b"))
    for s in range(seqsCnt):
        p = f"inputLayer_pattern_list.{s}"
        e.append((p + ".1.weight", (64, axisCnt, 3, 3), "conv_w"))
        e += _bn_entries(p + ".2", 64)
        e.append((p + ".4.weight", (64, 64, 3, 3), "conv_w"))
        e += _bn_entries(p + ".5", 64)
    e.append(("inputContact_layer.0.weight", (64, 64 * seqsCnt, 3, 3), "conv_w"))
    e += _bn_entries("inputContact_layer.1", 64)
    e.append(("output_layer.0.weight", (128, 128, 3, 3), "conv_w"))
    e.append(("output_layer.2.weight", (1, 128, 3, 3), "conv_w"))
    e.append(("input_layer_force.1.weight", (64, axisCnt, 3, 3), "conv_w"))
    return e


def tactilesrcnn_layout():
    """state_dict layout of TactileSRCNN (reference tactileSR_model.py:105-128)."""
    e = []
    for i in range(6):
        e += _msrb_entries(f"msrb_layer.{i}")
    for j, cin in ((0, 3), (3, 64), (6, 64)):
        e.append((f"input_zyx.{j}.weight", (64, cin, 3, 3), "conv_w"))
        e += _bn_entries(f"input_zyx.{j + 1}", 64)
    e.append(("output.0.weight", (1, 64, 3, 3), "conv_w"))
    return e


def make_state(layout, seed: int, dtype=torch.float32, nondegenerate: bool = True) -> "OrderedDict[str, Tensor]":
    """Seeded synthetic state_dict with the reference's key layout.

    ``nondegenerate=True`` draws BN affine params / running stats and biases from ranges
    that keep ReLUs alive, so the SR output is not the ~97 %-zeros map the default init
    gives (parity pitfall 1).  Conv weights follow the reference's
    Kaiming-normal(fan_out) scale (tactileSR_model.py:92-98).
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in layout:
        if kind == "conv_w":
            fan_out = shape[0] * shape[2] * shape[3]
            std = math.sqrt(2.0 / fan_out)
            if nondegenerate and shape[0] == 1:
                std *= 4.0
            t = torch.randn(shape, generator=g, dtype=torch.float32) * std
        elif kind == "conv_b":
            t = (torch.rand(shape, generator=g, dtype=torch.float32) - 0.5) * 0.2
        elif kind == "bn_w":
            t = 0.5 + torch.rand(shape, generator=g, dtype=torch.float32) if nondegenerate else torch.full(shape, 0.1)
        elif kind == "bn_b":
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 0.6 - 0.1) if nondegenerate else torch.full(shape, 0.1)
        elif kind == "bn_rm":
            t = (torch.rand(shape, generator=g, dtype=torch.float32) - 0.5) * 0.2 if nondegenerate else torch.zeros(shape)
        elif kind == "bn_rv":
            t = 0.5 + torch.rand(shape, generator=g, dtype=torch.float32) if nondegenerate else torch.ones(shape)
        elif kind == "bn_nbt":
            sd[key] = torch.tensor(0, dtype=torch.int64)
            continue
        else:
            raise ValueError(kind)
        sd[key] = t.to(dtype)
    return sd


# --------------------------------------------------------------------------------------
# fixed interpolation tables
# --------------------------------------------------------------------------------------
def bilinear_matrix(n_in: int, n_out: int, dtype=torch.float64) -> Tensor:
    """(n_out, n_in) matrix of ``F.interpolate(mode='bilinear', align_corners=False)`` along
    one axis: src = max((d + 0.5) * n_in / n_out - 0.5, 0); taps (floor, floor+1 clamped)."""
    m = torch.zeros(n_out, n_in, dtype=dtype)
    scale = n_in / n_out
    for d in range(n_out):
        s = max((d + 0.5) * scale - 0.5, 0.0)
        i0 = min(int(math.floor(s)), n_in - 1)
        i1 = min(i0 + 1, n_in - 1)
        l1 = s - i0
        m[d, i0] += 1.0 - l1
        m[d, i1] += l1
    return m


def upsample_bilinear(x: Tensor, n_out: int) -> Tensor:
    """nn.Upsample(scale_factor, 'bilinear', align_corners=False) (tactileSR_model.py:35,60):
    separable, rows then columns."""
    mh = bilinear_matrix(x.shape[-2], n_out, x.dtype).to(x.device)
    mw = bilinear_matrix(x.shape[-1], n_out, x.dtype).to(x.device)
    return torch.einsum("yh,bchw,xw->bcyx", mh, x, mw)


# --------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------
def batchnorm(x: Tensor, sd: Dict[str, Tensor], prefix: str, training: bool,
              new_stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """nn.BatchNorm2d (tactileSR_model.py:38 etc.): train = batch mean / biased var, running
    stats updated with momentum 0.1 and the unbiased var; eval = running stats."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if new_stats is not None:
            n = x.shape[0] * x.shape[2] * x.shape[3]
            with torch.no_grad():
                new_stats[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * sd[prefix + ".running_mean"] + BN_MOMENTUM * mean
                new_stats[prefix + ".running_var"] = (1 - BN_MOMENTUM) * sd[prefix + ".running_var"] + BN_MOMENTUM * var * (n / max(n - 1, 1))
                new_stats[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    inv = torch.rsqrt(var + BN_EPS)
    return (x - mean[None, :, None, None]) * (inv * w)[None, :, None, None] + b[None, :, None, None]


def conv(x: Tensor, sd, prefix: str, pad: int) -> Tensor:
    return F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), padding=pad)


def conv_bn_relu(x, sd, conv_prefix, bn_prefix, pad, training, new_stats):
    return torch.relu(batchnorm(conv(x, sd, conv_prefix, pad), sd, bn_prefix, training, new_stats))


def msrb(x: Tensor, sd, p: str, training: bool, new_stats=None) -> Tensor:
    """MSRB.forward (tactileSR_model.py:196-206)."""
    o31 = conv_bn_relu(x, sd, p + ".conv_3_1.0", p + ".conv_3_1.1", 1, training, new_stats)
    o51 = conv_bn_relu(x, sd, p + ".conv_5_1.0", p + ".conv_5_1.1", 2, training, new_stats)
    i2 = torch.cat([o31, o51], 1)
    o32 = conv_bn_relu(i2, sd, p + ".conv_3_2.0", p + ".conv_3_2.1", 1, training, new_stats)
    o52 = conv_bn_relu(i2, sd, p + ".conv_5_2.0", p + ".conv_5_2.1", 2, training, new_stats)
    i3 = torch.cat([o32, o52], 1)
    return torch.relu(conv(i3, sd, p + ".confusion", 0) + x)


def resblock(x: Tensor, sd, p: str) -> Tensor:
    """ResBlock.forward (tactileSR_model.py:222-225)."""
    y = torch.relu(conv(x, sd, p + ".conv1", 1))
    y = conv(y, sd, p + ".conv2", 1)
    return torch.relu(x + y)


def _count(sd, prefix: str) -> int:
    idx = set()
    for k in sd:
        if k.startswith(prefix + "."):
            idx.add(int(k[len(prefix) + 1:].split(".")[0]))
    return len(idx)


def tactilesr_forward(sd: Dict[str, Tensor], x: Tensor, training: bool, scale_factor: int = 10,
                      axisCnt: int = 3, new_stats: Optional[dict] = None,
                      taps: Optional[dict] = None) -> Tensor:
    """TactileSR.forward (tactileSR_model.py:67-84).  ``taps`` (if a dict) receives the
    intermediate activations named in SURVEY.md section 7 step 0."""
    S = _count(sd, "inputLayer_pattern_list")
    assert x.shape[1] == S * axisCnt, "input channel should be same with seqsCnt x axisCnt!"
    hw = x.shape[-1] * scale_factor
    feats = []
    for s in range(S):
        p = f"inputLayer_pattern_list.{s}"
        u = upsample_bilinear(x[:, axisCnt * s:axisCnt * (s + 1)], hw)
        h = conv_bn_relu(u, sd, p + ".1", p + ".2", 1, training, new_stats)
        h = conv_bn_relu(h, sd, p + ".4", p + ".5", 1, training, new_stats)
        feats.append(h)
    h = torch.cat(feats, 1)
    h = conv_bn_relu(h, sd, "inputContact_layer.0", "inputContact_layer.1", 1, training, new_stats)
    if taps is not None:
        taps["inputContact"] = h
    for i in range(_count(sd, "patternFeatureExtra_layer")):
        h = msrb(h, sd, f"patternFeatureExtra_layer.{i}", training, new_stats)
        if taps is not None:
            taps[f"msrb{i}"] = h
    f = torch.relu(conv(upsample_bilinear(x[:, :axisCnt], hw), sd, "input_layer_force.1", 1))
    for i in range(_count(sd, "forceFeatureExtra_layer")):
        f = resblock(f, sd, f"forceFeatureExtra_layer.{i}")
    if taps is not None:
        taps["force"] = f
    o = torch.cat([f, h], 1)
    o = torch.relu(conv(o, sd, "output_layer.0", 1))
    if taps is not None:
        taps["output0"] = o
    pre = conv(o, sd, "output_layer.2", 1)
    if taps is not None:
        taps["pre_relu"] = pre
    o = torch.relu(pre)
    # final F.interpolate(size=(4*sf, 4*sf)) (:83) maps 4*sf -> 4*sf: the identity.
    return o


def tactilesrcnn_forward(sd, x, training: bool, new_stats=None) -> Tensor:
    """TactileSRCNN.forward (tactileSR_model.py:148-153)."""
    h = upsample_bilinear(x, x.shape[-1] * 10)
    for j in (0, 3, 6):
        h = conv_bn_relu(h, sd, f"input_zyx.{j}", f"input_zyx.{j + 1}", 1, training, new_stats)
    for i in range(6):
        h = msrb(h, sd, f"msrb_layer.{i}", training, new_stats)
    return torch.relu(conv(h, sd, "output.0", 1))


# --------------------------------------------------------------------------------------
# training step (train/tactileSR_train.py:41-51 + cpu/trainer.py:346-362 + optim.Adam)
# --------------------------------------------------------------------------------------
def prep_hr(HR_raw: Tensor, HR_scale_num: float = 10.0, hw: int = 40) -> Tensor:
    """HR / HR_scale_num then bilinear resize to (hw, hw) (train/tactileSR_train.py:44-45)."""
    HR = HR_raw / HR_scale_num
    mh = bilinear_matrix(HR.shape[-2], hw, HR.dtype).to(HR.device)
    mw = bilinear_matrix(HR.shape[-1], hw, HR.dtype).to(HR.device)
    return torch.einsum("yh,bchw,xw->bcyx", mh, HR, mw)


def param_keys(sd) -> List[str]:
    return [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]


def loss_and_grads(sd, LR, HR_raw, training: bool = True, HR_scale_num: float = 10.0,
                   scale_factor: int = 10, taps: Optional[dict] = None):
    """Returns (loss, out, grads{key}, new_stats) for one ``train_cal_loss`` + backward."""
    leaf = OrderedDict()
    for k, v in sd.items():
        leaf[k] = v.detach().clone().requires_grad_(True) if k in set(param_keys(sd)) else v
    new_stats = {}
    out = tactilesr_forward(leaf, LR, training, scale_factor, new_stats=new_stats, taps=taps)
    HR = prep_hr(HR_raw, HR_scale_num, out.shape[-1])
    loss = torch.mean((out - HR) ** 2)
    keys = param_keys(sd)
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys], allow_unused=True)
    g = OrderedDict((k, (gi if gi is not None else torch.zeros_like(sd[k]))) for k, gi in zip(keys, grads))
    return loss.detach(), out.detach(), g, new_stats


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, wd: float,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam single-tensor update, coupled L2 decay (torch/optim/adam.py
    ``_single_tensor_adam``): g += wd*p; m, v EMA; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)."""
    g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def train_steps(sd, batches, lr: float = 1e-3, wd: float = 1e-2, HR_scale_num: float = 10.0):
    """Run len(batches) steps of fwd + MSE + bwd + Adam; returns (losses, final state)."""
    sd = OrderedDict((k, v.clone()) for k, v in sd.items())
    keys = param_keys(sd)
    m = {k: torch.zeros_like(sd[k]) for k in keys}
    v = {k: torch.zeros_like(sd[k]) for k in keys}
    losses = []
    for t, (LR, HR_raw) in enumerate(batches, start=1):
        loss, _, g, new_stats = loss_and_grads(sd, LR, HR_raw, True, HR_scale_num)
        losses.append(float(loss))
        for k in keys:
            sd[k], m[k], v[k] = adam_step(sd[k], g[k], m[k], v[k], t, lr, wd)
        sd.update(new_stats)
    return losses, sd


# --------------------------------------------------------------------------------------
# evaluation metrics (reference utility/tools.py:49-81, train/tactileSR_train.py:76-94)
# --------------------------------------------------------------------------------------
def psnr(p1: Tensor, p2: Tensor, max_value: float) -> Tensor:
    """calculationPSNR (utility/tools.py:49-62): note the divisor is shape[0] * shape[1] of whatever is passed in --
    eval_func passes (1, H, W) slices, so it is H."""
    mse = ((p1 - p2) ** 2).sum() / (p1.shape[0] * p1.shape[1])
    return 10 * torch.log10(max_value ** 2 / mse)


def ssim(p1: Tensor, p2: Tensor, C1: float = 0.01 ** 2, C2: float = 0.03 ** 2) -> Tensor:
    """calculationSSIM (utility/tools.py:65-81): whole-image means / variances, no window."""
    mu1, mu2 = p1.mean(), p2.mean()
    s1, s2, s12 = (p1 * p1).mean() - mu1 * mu1, (p2 * p2).mean() - mu2 * mu2, (p1 * p2).mean() - mu1 * mu2
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))


def eval_batch(out: Tensor, HR_raw: Tensor, HR_scale_num: float = 10.0, max_value: float = 250.0):
    """One iteration of eval_func's loop body (train/tactileSR_train.py:76-94): (mse, per-sample psnr, per-sample ssim)."""
    HR = prep_hr(HR_raw, HR_scale_num, out.shape[-1])
    mse = ((out - HR) ** 2).mean()
    ps = torch.stack([psnr(out[i], HR[i], max_value) for i in range(out.shape[0])])
    ss = torch.stack([ssim(out[i], HR[i]) for i in range(out.shape[0])])
    return mse, ps, ss
