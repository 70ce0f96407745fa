"""Drop-in tPSFNet: same class, constructor, attributes and ``state_dict`` (``MLP_layer.{1,3,5,7}``) as
reference ``model/tPSFNet.py:13-141``; ``forward`` runs the MLP and the fused per-sample PSF kernels
(csrc/mlp.cu, csrc/conv_f32.cu / conv_tc.cu for the wide layers, csrc/psf_tc.cu, csrc/psf.cu) instead of the python
``for i in range(B)`` loop (:118-125).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from .. import ops  # noqa: F401  (registers torch.ops.tactilesr.*)
from ..engine import get_precision

_ACT = {"none": 0, "relu": 1, "softplus": 2}


def _distance_table(h: int, w: int, cy: float, cx: float) -> torch.Tensor:
    # reference `_sdf` (:67-76): Euclidean distance of (row, col) to the centre, evaluated in double
    r = torch.arange(h, dtype=torch.float64)[:, None] - cy
    c = torch.arange(w, dtype=torch.float64)[None, :] - cx
    return torch.sqrt(r * r + c * c).to(torch.float32)


def _linear_fwd(x, w, b, act, st):
    """y = act(x W^T + b).  Layers whose shape fits the fp32 implicit-GEMM convolution kernels (K % 16 == 0,
    N % 64 == 0: 48->256, 256->1024, 1024->256) run as 1x1 convolutions over M = batch "pixels" (register-prefetching
    128x64 FFMA tiles, csrc/conv_f32.cu); the 256->3 softplus head uses the small strided SGEMM (csrc/mlp.cu)."""
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    if act != "softplus" and K % 16 == 0 and N % 64 == 0:
        wT = torch.empty((K * N,), dtype=torch.float32, device=x.device)        # [ci][co] = W^T
        _lib.call("tsr_pack_conv_weight_f32", w.data_ptr(), wT.data_ptr(), 0, N, K, 1, st)
        _lib.call("tsr_conv2d_f32", x.data_ptr(), K, wT.data_ptr(), b.data_ptr(), 0, 0, y.data_ptr(), N, M, 1, 1, K, N, 1,
                  1 if act == "relu" else 0, st)
    else:
        _lib.call("tsr_linear_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, _ACT[act], st)
    return y


def _linear_bwd(dy, out, x, w, act, need_dx, st):
    """(dW, db, dx) of y = act(x W^T + b) given dy and the stored output; same kernel choice as _linear_fwd."""
    M, K = x.shape
    N = w.shape[0]
    dev = x.device
    L = _lib.lib()
    dw = torch.empty_like(w)
    db = torch.empty((N,), dtype=torch.float32, device=dev)
    dx = torch.empty((M, K), dtype=torch.float32, device=dev) if need_dx else None
    if act == "relu" and K % 64 == 0 and N % 64 == 0:
        dpre = torch.empty((M, N), dtype=torch.float32, device=dev)
        _lib.call("tsr_relu_backward", dy.data_ptr(), N, out.data_ptr(), N, dpre.data_ptr(), N, 0, M, N, st)
        nws = max(int(L.tsr_conv2d_wgrad_f32_workspace(M, 1, 1, K, N, 1)), int(L.tsr_colsum_workspace(M, N)), 256)
        ws = torch.empty((nws,), dtype=torch.uint8, device=dev)
        _lib.call("tsr_conv2d_wgrad_f32", x.data_ptr(), K, dpre.data_ptr(), N, dw.data_ptr(), ws.data_ptr(), nws, M, 1, 1, K, N,
                  1, 0, st)
        _lib.call("tsr_colsum", dpre.data_ptr(), N, 0, M, N, db.data_ptr(), ws.data_ptr(), nws, 0, st)
        if need_dx:     # dx = dpre W: W [N][K] is already the [ci = N][co = K] image the kernel wants
            _lib.call("tsr_conv2d_f32", dpre.data_ptr(), N, w.data_ptr(), 0, 0, 0, dx.data_ptr(), K, M, 1, 1, N, K, 1, 0, st)
    else:
        dpre = torch.empty((M, N), dtype=torch.float32, device=dev)
        nws = int(L.tsr_linear_bwd_workspace(M, N, K))
        wsp = torch.empty((max(nws, 16),), dtype=torch.uint8, device=dev)
        _lib.call("tsr_linear_bwd", dy.data_ptr(), out.data_ptr(), x.data_ptr(), w.data_ptr(), dpre.data_ptr(),
                  dw.data_ptr(), db.data_ptr(), 0 if dx is None else dx.data_ptr(), M, N, K, _ACT[act], 0,
                  wsp.data_ptr(), wsp.numel(), st)
    return dw, db, dx


# ---------------------------------------------------------------------------------------------------------------------
# 16-bit precision modes ("fp16" / "bf16", tactilesr_b200.set_precision): the two wide layers (256 -> 1024 -> 256, 98 % of
# the MLP's FLOPs) run on the tcgen05 kernels of the SR path as 1x1 convolutions over an (M/64, 8, 8) "image" of the batch:
# forward in the mode's activation type, data and weight gradients on bf16 tensors, fp32 accumulation, fp32 master weights
# and gradients.  The 48 -> 256 input layer and the 256 -> 3 softplus head stay on the fp32 kernels.
# ---------------------------------------------------------------------------------------------------------------------
_TC_MIN_BATCH = 1024


def _tc_codes(mode):
    act = 2 if mode == "fp16" else 1
    return act, (torch.float16 if act == 2 else torch.bfloat16)


def _convert(x, C, src_code, dst_dtype, dst_code, st):
    y = torch.empty((x.shape[0], C), dtype=dst_dtype, device=x.device)
    _lib.call("tsr_copy_channels", x.data_ptr(), C, src_code, y.data_ptr(), C, dst_code, x.shape[0], C, st)
    return y


def _tc_pack(w, act, need_dgrad, st):
    """(forward image in the activation type, data-gradient image in bf16) of a Linear weight [N][K] = OIHW with 1x1 taps."""
    N, K = w.shape
    wf = torch.empty((N * K,), dtype=torch.float16 if act == 2 else torch.bfloat16, device=w.device)
    wd = torch.empty((N * K,), dtype=torch.bfloat16, device=w.device) if need_dgrad else None
    if act == 2:
        _lib.call("tsr_pack_conv_weight_f16", w.data_ptr(), wf.data_ptr(), 0, N, K, 1, st)
        if need_dgrad:
            _lib.call("tsr_pack_conv_weight_bf16", w.data_ptr(), 0, wd.data_ptr(), N, K, 1, st)
    else:
        _lib.call("tsr_pack_conv_weight_bf16", w.data_ptr(), wf.data_ptr(), 0 if wd is None else wd.data_ptr(), N, K, 1, st)
    return wf, wd


def _tc_conv(x, K, wpk, bias, N, out_dtype, flags, st, out2=None):
    """y[M][N] = x[M][K] W^T (+ bias, ReLU per flags) on tsr_conv2d_tc; out2: optional bf16 copy of the result."""
    M = x.shape[0]
    y = torch.empty((M, N), dtype=out_dtype, device=x.device)
    _lib.call("tsr_conv2d_tc", x.data_ptr(), K, wpk.data_ptr(), 0 if bias is None else bias.data_ptr(), 0, 0, y.data_ptr(), N,
              1, M // 8, 8, K, N, 1, flags, 0, 0, 0, 0 if out2 is None else out2.data_ptr(), N, st)
    return y


def _tc_linear_bwd(dy_bf16, out, act, x_bf16, w, wd, need_dx, st):
    """Backward of y = relu(x W^T + b) on the tensor cores: dy (bf16, overwritten with the pre-activation gradient),
    stored output `out` (activation type), bf16 input -> (dW fp32, db fp32, dx bf16 or None)."""
    M, N = dy_bf16.shape
    K = x_bf16.shape[1]
    dev = dy_bf16.device
    L = _lib.lib()
    _lib.call("tsr_relu_backward", dy_bf16.data_ptr(), N, out.data_ptr(), N, dy_bf16.data_ptr(), N, act, M, N, st)
    dw = torch.empty_like(w)
    db = torch.empty((N,), dtype=torch.float32, device=dev)
    nws = max(int(L.tsr_conv2d_wgrad_tc_workspace(M // 64, 8, 8, K, N, 1)), int(L.tsr_colsum_workspace(M, N)), 256)
    ws = torch.empty((nws,), dtype=torch.uint8, device=dev)
    _lib.call("tsr_conv2d_wgrad_tc", x_bf16.data_ptr(), K, dy_bf16.data_ptr(), N, dw.data_ptr(), ws.data_ptr(), nws, M // 64, 8, 8,
              K, N, 1, 0, st)
    _lib.call("tsr_colsum", dy_bf16.data_ptr(), N, 1, M, N, db.data_ptr(), ws.data_ptr(), nws, 0, st)
    dx = _tc_conv(dy_bf16, N, wd, None, K, torch.bfloat16, 0, st) if need_dx else None
    return dw, db, dx


class _MLPTcFn(torch.autograd.Function):
    """MLP_layer in the 16-bit precision modes (see above); same interface as _MLPFn."""

    @staticmethod
    def forward(ctx, mode, x, w1, b1, w2, b2, w3, b3, w4, b4):
        st = _lib.stream_ptr()
        act, adt = _tc_codes(mode)
        B = x.shape[0]
        need_grad = any(ctx.needs_input_grad[2:])
        x2 = x.detach().reshape(B, -1).contiguous().float()
        ws = [w.detach().contiguous() for w in (w1, w2, w3, w4)]
        bs = [b.detach().contiguous() for b in (b1, b2, b3, b4)]
        f16 = 2 if act == 2 else 0
        x1 = _linear_fwd(x2, ws[0], bs[0], "relu", st)                         # fp32 [B][256]
        x1a = _convert(x1, 256, 0, adt, act, st)
        x1s = _convert(x1, 256, 0, torch.bfloat16, 1, st) if (act == 2 and need_grad) else x1a
        w2f, w2d = _tc_pack(ws[1], act, need_grad, st)
        w3f, w3d = _tc_pack(ws[2], act, need_grad, st)
        h2s = torch.empty((B, 1024), dtype=torch.bfloat16, device=x.device) if (act == 2 and need_grad) else None
        h2 = _tc_conv(x1a, 256, w2f, bs[1], 1024, adt, 1 | f16, st, out2=h2s)
        h3 = _tc_conv(h2, 1024, w3f, bs[2], 256, adt, 1 | f16, st)
        h3f = _convert(h3, 256, act, torch.float32, 0, st)
        ab = _linear_fwd(h3f, ws[3], bs[3], "softplus", st)
        if need_grad:
            ctx.mode = mode
            ctx.save_for_backward(x2, x1, x1s, h2, h2 if h2s is None else h2s, h3, h3f, ab, ws[0], w2d, w3d, ws[1], ws[2], ws[3])
        return ab.clone()

    @staticmethod
    def backward(ctx, dab):
        st = _lib.stream_ptr()
        act, _ = _tc_codes(ctx.mode)
        x2, x1, x1s, h2, h2s, h3, h3f, ab, w1, w2d, w3d, w2, w3, w4 = ctx.saved_tensors
        dw4, db4, dx3 = _linear_bwd(dab.detach().contiguous().float(), ab, h3f, w4, "softplus", True, st)
        g3 = _convert(dx3, 256, 0, torch.bfloat16, 1, st)
        dw3, db3, g2 = _tc_linear_bwd(g3, h3, act, h2s, w3, w3d, True, st)
        dw2, db2, g1 = _tc_linear_bwd(g2, h2, act, x1s, w2, w2d, True, st)
        dx1 = _convert(g1, 256, 1, torch.float32, 0, st)
        dw1, db1, _ = _linear_bwd(dx1, x1, x2, w1, "relu", False, st)
        return (None, None, dw1, db1, dw2, db2, dw3, db3, dw4, db4)


class _MLPFn(torch.autograd.Function):
    """MLP_layer (reference model/tPSFNet.py:26-36): Flatten -> 48-256-1024-256-3 with ReLU / Softplus, forward and
    hand-written backward on our GEMM kernels."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, w4, b4):
        if not x.is_cuda:
            raise _lib.TsrError("tactilesr_b200.tPSFNet runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        st = _lib.stream_ptr()
        B = x.shape[0]
        x2 = x.detach().reshape(B, -1).contiguous().float()
        ws = [w.detach().contiguous() for w in (w1, w2, w3, w4)]
        bs = [b.detach().contiguous() for b in (b1, b2, b3, b4)]
        acts = [x2]
        for i, (w, b) in enumerate(zip(ws, bs)):
            acts.append(_linear_fwd(acts[-1], w, b, "softplus" if i == 3 else "relu", st))
        ctx.save_for_backward(*acts, *ws)
        return acts[-1].clone()

    @staticmethod
    def backward(ctx, dab):
        st = _lib.stream_ptr()
        saved = ctx.saved_tensors
        acts, ws = list(saved[0:5]), list(saved[5:9])
        grads = []
        dy = dab.detach().contiguous().float()
        for i in (3, 2, 1, 0):
            dw, db, dx = _linear_bwd(dy, acts[i + 1], acts[i], ws[i], "softplus" if i == 3 else "relu", i > 0, st)
            grads = [dw, db] + grads
            dy = dx
        return (None, *grads)


class tPSFNet(nn.Module):
    precision = None   # None -> global tactilesr_b200.get_precision(); "fp32" pins the fp32 MLP

    def __init__(self, gama, perception_scale, size=(100, 100), device=None):
        super().__init__()
        self.gama = gama
        self.perception_scale = perception_scale
        # reference :21-23 stores the argument as given (None stays None)
        self.device = device

        self.MLP_layer = nn.Sequential(
            nn.Flatten(),
            nn.Linear(16 * 3, 256), nn.ReLU(),
            nn.Linear(256, 1024), nn.ReLU(),
            nn.Linear(1024, 256), nn.ReLU(),
            nn.Linear(256, 3), nn.Softplus())
        self._init_weights(self.MLP_layer)
        self.zeroPad_func = nn.ZeroPad2d(padding=(48, 48, 48, 48))

        # constant distance fields (reference :41-55): plain attributes, not buffers, not in state_dict.
        sdf = _distance_table(99, 99, 49, 49)
        sdf = 10 * (sdf - sdf.min()) / (sdf.max() - sdf.min())
        self.PSF_sdf = sdf[None, None].to(device) if device is not None else sdf[None, None]
        m = torch.stack([torch.stack([_distance_table(100, 100, 12 + 25 * i, 12 + 25 * j) for j in range(4)])
                         for i in range(4)])
        m = 10 * (m - m.min()) / (m.max() - m.min())
        self.LR_masking_sdf = m.to(device) if device is not None else m

    def _init_weights(self, modules):
        for m in modules:
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0, std=0.03)

    def forward(self, x, depth):
        assert x.shape[0] == depth.shape[0], "Batch size of LR tactile and depth should be the same!"
        L = self.MLP_layer
        B = x.shape[0]
        mode = self.precision or get_precision()
        params = (L[1].weight, L[1].bias, L[3].weight, L[3].bias, L[5].weight, L[5].bias, L[7].weight, L[7].bias)
        if mode in ("fp16", "bf16") and B % 64 == 0 and B >= _TC_MIN_BATCH and x.is_cuda:
            ab = _MLPTcFn.apply(mode, x, *params)
        else:            # fp32 mode; batches the (M/64, 8, 8) tiling of the tensor-core kernels does not divide; and
            ab = _MLPFn.apply(x, *params)      # small batches, which are launch-bound (the fp32 path has half the launches)
        # the python ``for i in range(B)`` loop of the reference (:118-125) is one custom op with autograd
        # (torch.ops.tactilesr.psf_model); the forward -> backward hand-over is only produced when a backward can follow
        # 16-bit precision modes: the PSF products run as one fp16 tensor-core pass (~3e-4 on HR / LR_degrade, inside the
        # modes' 1e-2 tolerance); "fp32": the split-operand kernels (~1e-6)
        HR, LRd, psf, _ = torch.ops.tactilesr.psf_model(ab, depth, torch.is_grad_enabled() and ab.requires_grad,
                                                        mode in ("fp16", "bf16"))
        return HR, LRd, psf, ab.view(B, 1, 3)
