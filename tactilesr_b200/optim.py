"""Fused Adam with the stock ``torch.optim.Adam`` interface and ``state_dict`` layout
(``state[i] = {step, exp_avg, exp_avg_sq}``, same ``param_groups`` keys), replacing the optimizer built at
reference train/tactileSR_train.py:212, train/tactileSRSeqs_train.py:74 and train/tPSFNet_train.py:201.

Parameters, both moments and (when the engine produced them) the gradients of a param group live in
flat fp32 buffers, so one kernel launch (csrc/elementwise.cu: adam_kernel) updates the whole group and
the data-parallel all-reduce runs over the same flat gradient buffer.
"""
from __future__ import annotations

import math
from typing import List

import numpy as np
import torch

from . import _lib
from .engine import bump_weight_epoch


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *,
                 foreach=None, maximize=False, capturable=False, differentiable=False, fused=None,
                 decoupled_weight_decay=False):
        if amsgrad or maximize or decoupled_weight_decay:
            raise NotImplementedError("FusedAdam implements the reference's configuration: plain Adam with coupled L2")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, maximize=maximize,
                        foreach=foreach, capturable=capturable, differentiable=differentiable, fused=fused,
                        decoupled_weight_decay=decoupled_weight_decay)
        super().__init__(params, defaults)
        self._flat = {}
        self._graph_hyper = None       # CUDA-graph mode: per-group device {lr / bc1, 1 / sqrt(bc2)} (see enable_graph_mode)

    # -- flat storage -------------------------------------------------------------------------
    def _flatten_group(self, gi: int, group) -> dict:
        params: List[torch.nn.Parameter] = [p for p in group["params"]]
        if not params:
            return {}
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.TsrError("FusedAdam runs on CUDA parameters only (no CPU fallback)")
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(params, offs):
            k = p.numel()
            flat_p[o:o + k].copy_(p.data.reshape(-1))
            p.data = flat_p[o:o + k].view(p.shape)
            p._tsr_flat_grad = flat_g[o:o + k].view(p.shape)
            st = self.state[p]
            if "exp_avg" in st:   # resumed from a checkpoint
                flat_m[o:o + k].copy_(st["exp_avg"].reshape(-1))
                flat_v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
            st["exp_avg"] = flat_m[o:o + k].view(p.shape)
            st["exp_avg_sq"] = flat_v[o:o + k].view(p.shape)
            if "step" not in st:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
        f = dict(p=flat_p, g=flat_g, m=flat_m, v=flat_v, offs=offs, n=n, params=params)
        self._flat[gi] = f
        return f

    def flat_grad(self, gi: int = 0) -> torch.Tensor:
        """Flat gradient buffer of param group ``gi`` (all-reduce target of the data-parallel trainer)."""
        if gi not in self._flat:
            self._flatten_group(gi, self.param_groups[gi])
        return self._flat[gi]["g"]

    # -- CUDA-graph support -----------------------------------------------------------------------
    def enable_graph_mode(self) -> None:
        """From now on ``step`` reads the step-dependent scalars (bias corrections, learning rate) from device memory, so
        the launch can be captured in a CUDA graph; the host refreshes them with ``update_graph_hyper`` before every
        replay and advances its step counters with ``graph_advance`` afterwards.  Requires an ordinary step to have run
        (flat buffers exist) and every parameter of a group to share one step count."""
        self._graph_hyper = []
        for gi, group in enumerate(self.param_groups):
            if gi not in self._flat:
                self._flatten_group(gi, group)
            dev = self._flat[gi]["p"].device if self._flat.get(gi) else torch.device("cuda")
            self._graph_hyper.append(torch.zeros(2, dtype=torch.float32, device=dev))

    def disable_graph_mode(self) -> None:
        self._graph_hyper = None

    def update_graph_hyper(self) -> None:
        """Scalars of the *next* step -> device (stream-ordered before the replay)."""
        for gi, group in enumerate(self.param_groups):
            f = self._flat.get(gi)
            if not f:
                continue
            # bit-for-bit the scalars tsr_adam_step derives on the host: betas and lr as floats, bias corrections in
            # double, their reciprocals rounded to float, the lr product in float
            b1, b2 = (float(np.float32(b)) for b in group["betas"])
            t = int(self.state[f["params"][0]]["step"]) + 1
            inv_bc1 = np.float32(1.0 / (1.0 - math.pow(b1, float(t))))
            vals = [float(np.float32(group["lr"]) * inv_bc1), float(np.float32(1.0 / math.sqrt(1.0 - math.pow(b2, float(t)))))]
            # a fresh pageable source per call: the copy is staged before it returns, so the host (which runs many
            # iterations ahead of the GPU) can never overwrite scalars an earlier replay has not consumed yet
            self._graph_hyper[gi].copy_(torch.tensor(vals, dtype=torch.float32))

    def graph_advance(self) -> None:
        """Host bookkeeping of one replayed step: step counters and the packed-weight epoch."""
        for group in self.param_groups:
            for p in group["params"]:
                self.state[p]["step"] += 1
        bump_weight_epoch()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._flat = {}   # re-flatten around the loaded moments on the next step

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        st_ptr = _lib.stream_ptr()
        for gi, group in enumerate(self.param_groups):
            f = self._flat.get(gi)
            if f is None or len(f["params"]) != len(group["params"]) or any(
                    a is not b for a, b in zip(f["params"], group["params"])):
                f = self._flatten_group(gi, group)
            if not f:
                continue
            b1, b2 = group["betas"]
            lr, eps, wd = float(group["lr"]), group["eps"], group["weight_decay"]
            params = f["params"]
            with_grad = [p for p in params if p.grad is not None]
            if not with_grad:
                continue
            steps = {int(self.state[p]["step"].item()) if self.state[p]["step"].is_cuda else int(self.state[p]["step"])
                     for p in with_grad}
            all_flat = len(with_grad) == len(params) and len(steps) == 1
            if all_flat:
                # gradients the engine did not write in place (foreign autograd graph) are gathered first
                for p, o in zip(params, f["offs"]):
                    if p.grad.data_ptr() != f["g"].data_ptr() + o * 4:
                        f["g"][o:o + p.numel()].copy_(p.grad.reshape(-1))
                step = steps.pop() + 1
                if self._graph_hyper is not None:
                    _lib.call("tsr_adam_step_dev", f["p"].data_ptr(), f["g"].data_ptr(), f["m"].data_ptr(),
                              f["v"].data_ptr(), f["n"], self._graph_hyper[gi].data_ptr(), b1, b2, eps, wd, 1.0, st_ptr)
                else:
                    _lib.call("tsr_adam_step", f["p"].data_ptr(), f["g"].data_ptr(), f["m"].data_ptr(), f["v"].data_ptr(),
                              f["n"], lr, b1, b2, eps, wd, step, 1.0, st_ptr)
                for p in params:
                    self.state[p]["step"] += 1
            elif self._graph_hyper is not None:
                raise _lib.TsrError("FusedAdam graph mode needs every parameter of a group to receive a gradient")
            else:
                # stock Adam skips parameters without a gradient (e.g. the transplanted stacks of
                # tactileSRSeqs_train.py:74-77 are not even in the optimizer): per-parameter launches.
                for p, o in zip(params, f["offs"]):
                    if p.grad is None:
                        continue
                    k = p.numel()
                    g = p.grad
                    if g.data_ptr() != f["g"].data_ptr() + o * 4:
                        f["g"][o:o + k].copy_(g.reshape(-1))
                    s = self.state[p]
                    s["step"] += 1
                    _lib.call("tsr_adam_step", f["p"].data_ptr() + o * 4, f["g"].data_ptr() + o * 4,
                              f["m"].data_ptr() + o * 4, f["v"].data_ptr() + o * 4, k, lr, b1, b2, eps, wd,
                              int(s["step"]), 1.0, st_ptr)
        bump_weight_epoch()
        return loss


Adam = FusedAdam
