"""Drop-in tPSFNet: same class, constructor, attributes and ``state_dict`` (``MLP_layer.{1,3,5,7}``) as
reference ``model/tPSFNet.py:13-141``; ``forward`` runs the MLP and the fused per-sample PSF kernels
(csrc/mlp.cu, csrc/psf.cu) instead of the python ``for i in range(B)`` loop (:118-125).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from .. import ops  # noqa: F401  (registers torch.ops.tactilesr.*)

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
        ab = _MLPFn.apply(x, L[1].weight, L[1].bias, L[3].weight, L[3].bias, L[5].weight, L[5].bias, L[7].weight,
                          L[7].bias)
        # the python ``for i in range(B)`` loop of the reference (:118-125) is one custom op with autograd
        # (torch.ops.tactilesr.psf_model); the forward -> backward hand-over is only produced when a backward can follow
        HR, LRd, psf, _ = torch.ops.tactilesr.psf_model(ab, depth, torch.is_grad_enabled() and ab.requires_grad)
        return HR, LRd, psf, ab.view(B, 1, 3)
