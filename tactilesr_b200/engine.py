"""Layer-program executor for the SR conv stacks.

A module's forward is lowered to a short list of ops over NHWC activation buffers; each op's
``fwd`` / ``bwd`` launches the hand-written CUDA kernels through the C ABI (``_lib``).  The whole
program is one ``torch.autograd.Function`` node (``_ProgramFn``), so the reference's
``loss.backward()`` / ``optimizer.step()`` protocol (cpu/trainer.py:346-362) keeps working while
the forward+backward of reference model/tactileSR_model.py runs entirely in our kernels.

Three numeric modes (``tactilesr_b200.set_precision``):
  * "fp32": activations fp32, FFMA implicit-GEMM kernels (conv_f32.cu) -- <= 1e-5 parity mode.
  * "bf16": activations and gradients bf16, tcgen05 implicit-GEMM kernels (conv_tc.cu), fp32 accumulate/statistics.
  * "fp16": same tcgen05 kernels at the same speed, but activations and forward weights are stored as fp16
    (10 mantissa bits, the TF32 mantissa: ~8x less rounding error than bf16 -- the <= 1e-2 tensor-core mode);
    gradients stay bf16 (range matters there); tcgen05 kind::f16 rejects fp16 x bf16 operands, so when gradients
    are needed every conv keeps a bf16 copy of its input for the weight-gradient kernel.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib

_PRECISION = "fp32"
_WEIGHT_EPOCH = 0          # bumped by FusedAdam.step (in-place kernel updates do not bump ._version)
_STATS_EPOCH = 0           # bumped whenever a training forward updates BatchNorm running statistics in place


def set_precision(mode: str) -> None:
    global _PRECISION
    if mode not in ("fp32", "bf16", "fp16"):
        raise ValueError("precision must be 'fp32', 'bf16' or 'fp16'")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def bump_weight_epoch() -> None:
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------
# symbolic buffers / views
# ---------------------------------------------------------------------------------------------
@dataclass(eq=False)
class Buf:
    name: str
    C: int
    kind: str = "act"       # "act": NHWC activation in the mode dtype; "plane": (B,H,W) fp32 single channel


@dataclass(eq=False)
class View:
    buf: Buf
    c0: int
    C: int

    @staticmethod
    def of(buf: Buf) -> "View":
        return View(buf, 0, buf.C)


class RunCtx:
    """Per-call state: concrete tensors of the symbolic buffers, saved statistics, gradients."""

    def __init__(self, mode: str, B: int, H: int, W: int, device, training: bool, need_grad: bool):
        self.mode, self.B, self.H, self.W = mode, B, H, W
        self.device, self.training, self.need_grad = device, training, need_grad
        # storage-type codes of the C ABI (0 float, 1 bf16, 2 fp16): activations, gradients, and the combined code of
        # the kernels that read both (tsr_bn_backward / tsr_relu_backward)
        self.act = {"fp32": 0, "bf16": 1, "fp16": 2}[mode]
        self.grd = 0 if mode == "fp32" else 1
        self.mix = self.act
        self.tc = mode != "fp32"
        self.act_dtype = (torch.float32, torch.bfloat16, torch.float16)[self.act]
        self.grad_dtype = (torch.float32, torch.bfloat16)[self.grd]
        self.esize = 2 if self.tc else 4
        self.npix = B * H * W
        self.bufs: Dict[Buf, torch.Tensor] = {}
        self.grads: Dict[Buf, torch.Tensor] = {}
        self.grad_written: Dict[Buf, bool] = {}
        self.saved: Dict[object, tuple] = {}
        self.x: Optional[torch.Tensor] = None
        self.param_grads: Dict[torch.nn.Parameter, torch.Tensor] = {}
        # "fp16" mode with gradients: bf16 shadow of every buffer some conv reads (the weight-gradient operand; tcgen05
        # kind::f16 cannot mix an fp16 x with a bf16 dy).  Producers write their channel slice of the shadow.
        self.bn_partials: Dict[object, torch.Tensor] = {}      # BNReLUOp -> statistics partials from its conv's epilogue
        self.folded: set = set()                               # BNReLUOps already applied by their (inference) conv
        self.keep_taps = False
        self.shadow: Dict[Buf, torch.Tensor] = {}
        self.conv_inputs: set = set()
        self._ws: Optional[torch.Tensor] = None

    # -- buffers ------------------------------------------------------------------------------
    def alloc(self, buf: Buf) -> torch.Tensor:
        t = self.bufs.get(buf)
        if t is None:
            if buf.kind == "plane":
                t = torch.empty((self.npix,), dtype=torch.float32, device=self.device)
            else:
                t = torch.empty((self.npix, buf.C), dtype=self.act_dtype, device=self.device)
            self.bufs[buf] = t
        return t

    def vptr(self, v: View, grad: bool = False) -> Tuple[int, int]:
        t = self.galloc(v.buf) if grad else self.alloc(v.buf)
        es = 4 if v.buf.kind == "plane" else self.esize
        return t.data_ptr() + v.c0 * es, v.buf.C

    def galloc(self, buf: Buf) -> torch.Tensor:
        t = self.grads.get(buf)
        if t is None:
            if buf.kind == "plane":
                t = torch.empty((self.npix,), dtype=torch.float32, device=self.device)
            else:
                t = torch.empty((self.npix, buf.C), dtype=self.grad_dtype, device=self.device)
            self.grads[buf] = t
            self.grad_written[buf] = False
        return t

    def wants_shadow(self, buf: Buf) -> bool:
        return self.act == 2 and self.need_grad and buf in self.conv_inputs

    def sptr(self, v: View) -> Tuple[int, int]:
        """(pointer, ld) of view v inside the bf16 shadow of its buffer."""
        t = self.shadow.get(v.buf)
        if t is None:
            t = torch.empty((self.npix, v.buf.C), dtype=torch.bfloat16, device=self.device)
            self.shadow[v.buf] = t
        return t.data_ptr() + v.c0 * 2, v.buf.C

    def shadow_fill(self, v: View) -> None:
        """Copy the freshly written fp16 view into the shadow (producers without a fused second output)."""
        if self.wants_shadow(v.buf):
            ip, ild = self.vptr(v)
            sp, sld = self.sptr(v)
            _lib.call("tsr_copy_channels", ip, ild, 2, sp, sld, 1, self.npix, v.C, _lib.stream_ptr())

    def workspace(self, nbytes: int) -> Tuple[int, int]:
        nbytes = max(int(nbytes), 256)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        return self._ws.data_ptr(), self._ws.numel()

    def pgrad(self, p: torch.nn.Parameter) -> Tuple[torch.Tensor, int]:
        """(gradient tensor for p, accumulate flag).  A parameter used by several ops accumulates."""
        g = self.param_grads.get(p)
        if g is None:
            # FusedAdam exposes a view of its flat gradient buffer; write there directly when autograd will
            # adopt the tensor as p.grad (p.grad is None), otherwise hand autograd a fresh tensor to accumulate.
            g = getattr(p, "_tsr_flat_grad", None) if p.grad is None else None
            if g is None or g.device != p.device:
                g = torch.empty_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
            self.param_grads[p] = g
            return g, 0
        return g, 1


# ---------------------------------------------------------------------------------------------
# packed-weight cache
# ---------------------------------------------------------------------------------------------
class _PackCache:
    """Packed weights live on the Parameter object itself (``_tsr_pack``), so they die with it -- a process-wide dict
    keyed by id()/data_ptr would alias a freed parameter with a new one that reuses both."""

    def __init__(self):
        self._groups = {}      # (mode, ids of a program's conv weights) -> persistent buffers + device descriptor table

    def get(self, w: torch.Tensor, mode: str, need_dgrad: bool):
        tag = (w.data_ptr(), w._version, _WEIGHT_EPOCH, tuple(w.shape), w.device)
        store = w.__dict__.setdefault("_tsr_pack", {})
        hit = store.get(mode)
        if hit is not None and hit[0] == tag and (hit[2] is not None or not need_dgrad):
            return hit[1], hit[2]
        Cout, Cin, K, _ = w.shape
        dt = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[mode]
        wf = torch.empty((K * K * Cin * Cout,), dtype=dt, device=w.device)
        wc = w.detach().contiguous()
        st = _lib.stream_ptr()
        if mode == "fp16":       # forward weights fp16, data-gradient weights bf16 (gradients are bf16 tensors)
            wd = torch.empty_like(wf, dtype=torch.bfloat16) if need_dgrad else None
            _lib.call("tsr_pack_conv_weight_f16", wc.data_ptr(), wf.data_ptr(), 0, Cout, Cin, K, st)
            if need_dgrad:
                _lib.call("tsr_pack_conv_weight_bf16", wc.data_ptr(), 0, wd.data_ptr(), Cout, Cin, K, st)
        else:
            wd = torch.empty_like(wf) if need_dgrad else None
            fn = "tsr_pack_conv_weight_bf16" if mode == "bf16" else "tsr_pack_conv_weight_f32"
            _lib.call(fn, wc.data_ptr(), wf.data_ptr(), _ptr(wd), Cout, Cin, K, st)
        store[mode] = (tag, wf, wd)
        return wf, wd


    # -- all conv weights of a program in one launch ---------------------------------------------------------------
    def prepack(self, weights, mode: str, need_dgrad: bool) -> None:
        """Tensor-core modes: bring the packed images of every weight in ``weights`` up to date with ONE kernel launch
        (the optimizer changes all of them at every step) and leave the per-parameter cache entries current, so the
        ConvOps' ``get`` calls all hit.  Buffers and the device descriptor table are persistent per weight set."""
        import numpy as np
        key = (mode,) + tuple(id(w) for w in weights)
        grp = self._groups.get(key)
        if grp is not None and any(r() is None for r in grp["refs"]):
            grp = None
        if grp is None:
            import weakref
            grp = {"refs": [weakref.ref(w) for w in weights], "sig": None, "ptrs": None, "table": None, "bufs": [],
                   "dgrad": False}
            for k in [k for k, g in self._groups.items() if any(r() is None for r in g["refs"])]:
                del self._groups[k]
            self._groups[key] = grp
        sig = (_WEIGHT_EPOCH, tuple(w._version for w in weights))
        ptrs = tuple(w.data_ptr() for w in weights)
        want_d = grp["dgrad"] or need_dgrad
        if grp["sig"] == sig and grp["ptrs"] == ptrs and grp["dgrad"] == want_d:
            return
        dt_f = torch.float16 if mode == "fp16" else torch.bfloat16
        if grp["table"] is None or grp["ptrs"] != ptrs or grp["dgrad"] != want_d:
            desc = np.zeros(len(weights), dtype=np.dtype([("w", "<u8"), ("wf", "<u8"), ("wd", "<u8"), ("Cout", "<i4"),
                                                          ("Cin", "<i4"), ("KS", "<i4"), ("dt_f", "<i4"), ("dt_d", "<i4"),
                                                          ("pad", "<i4")]))
            bufs = []
            for i, w in enumerate(weights):
                Cout, Cin, K, _ = w.shape
                old = grp["bufs"][i] if i < len(grp["bufs"]) else (None, None)
                wf = old[0] if old[0] is not None else torch.empty((K * K * Cin * Cout,), dtype=dt_f, device=w.device)
                wd = old[1]
                if want_d and wd is None:
                    wd = torch.empty((K * K * Cin * Cout,), dtype=torch.bfloat16, device=w.device)
                bufs.append((wf, wd))
                assert w.is_contiguous()
                desc[i] = (w.data_ptr(), wf.data_ptr(), 0 if wd is None else wd.data_ptr(), Cout, Cin, K,
                           2 if mode == "fp16" else 1, 1, 0)
            grp["bufs"], grp["dgrad"], grp["ptrs"] = bufs, want_d, ptrs
            grp["table"] = torch.from_numpy(desc.view(np.uint8).copy()).to(weights[0].device)
            grp["max"] = max(int(w.numel()) for w in weights)
        _lib.call("tsr_pack_conv_weights_multi", grp["table"].data_ptr(), len(weights), grp["max"], _lib.stream_ptr())
        grp["sig"] = sig
        for w, (wf, wd) in zip(weights, grp["bufs"]):
            tag = (w.data_ptr(), w._version, _WEIGHT_EPOCH, tuple(w.shape), w.device)
            w.__dict__.setdefault("_tsr_pack", {})[mode] = (tag, wf, wd)

    def get_folded(self, conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d, mode: str):
        """Inference: forward pack of ``conv`` with the eval-mode BatchNorm ``bn`` folded in -> (packed weights, bias)."""
        w = conv.weight
        tag = (w.data_ptr(), w._version, _WEIGHT_EPOCH, _STATS_EPOCH, tuple(w.shape), w.device, bn.weight.data_ptr(),
               bn.weight._version, bn.bias._version, bn.running_mean.data_ptr(), bn.running_mean._version,
               bn.running_var._version, bn.eps, None if conv.bias is None else conv.bias._version)
        store = w.__dict__.setdefault("_tsr_pack", {})
        hit = store.get(mode + ":folded")
        if hit is not None and hit[0] == tag:
            return hit[1], hit[2]
        Cout, Cin, K, _ = w.shape
        st = _lib.stream_ptr()
        coef = torch.empty((4, Cout), dtype=torch.float32, device=w.device)
        _lib.call("tsr_bn_eval_coeffs", Cout, bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                  bn.running_var.data_ptr(), bn.eps, coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(),
                  coef[3].data_ptr(), st)
        wf = torch.empty((K * K * Cin * Cout,), dtype=torch.float16 if mode == "fp16" else torch.bfloat16, device=w.device)
        bias = torch.empty((Cout,), dtype=torch.float32, device=w.device)
        _lib.call("tsr_pack_conv_weight_folded", w.detach().contiguous().data_ptr(), _ptr(conv.bias), coef[0].data_ptr(),
                  coef[1].data_ptr(), wf.data_ptr(), bias.data_ptr(), Cout, Cin, K, 2 if mode == "fp16" else 1, st)
        store[mode + ":folded"] = (tag, wf, bias)
        return wf, bias


_PACK = _PackCache()


# ---------------------------------------------------------------------------------------------
# ops
# ---------------------------------------------------------------------------------------------
class Op:
    def fwd(self, c: RunCtx) -> None:
        raise NotImplementedError

    def bwd(self, c: RunCtx) -> None:
        raise NotImplementedError

    def params(self) -> Sequence[torch.nn.Parameter]:
        return ()

    def reads(self) -> Sequence[Buf]:
        return ()

    def writes(self) -> Sequence[Buf]:
        return ()


def _first_write(c: RunCtx, buf: Buf) -> bool:
    """True if this is the first gradient contribution to ``buf`` in the current backward."""
    c.galloc(buf)
    first = not c.grad_written[buf]
    c.grad_written[buf] = True
    return first


class HeadOp(Op):
    """Upsample(x sf, bilinear) + Conv2d(3 -> 64, 3x3, no bias) [+ ReLU]
    (reference tactileSR_model.py:35-37, 60-62, 107 + 122)."""

    def __init__(self, ch0: int, weight, out: View, relu: bool, sf: int):
        self.ch0, self.weight, self.out, self.relu, self.sf = ch0, weight, out, relu, sf

    def params(self):
        return (self.weight,)

    def writes(self):
        return (self.out.buf,)

    def _x(self, c: RunCtx):
        x = c.x
        return x.data_ptr() + self.ch0 * 16 * 4, x.shape[1] * 16

    def fwd(self, c):
        xp, xbs = self._x(c)
        op, old = c.vptr(self.out)
        _lib.call("tsr_head_fwd", xp, xbs, self.weight.data_ptr(), op, old, c.act, c.B, self.sf,
                  1 if self.relu else 0, _lib.stream_ptr())
        c.shadow_fill(self.out)

    def bwd(self, c):
        st = _lib.stream_ptr()
        gp, gld = c.vptr(self.out, grad=True)
        if self.relu:
            ap, ald = c.vptr(self.out)
            _lib.call("tsr_relu_backward", gp, gld, ap, ald, gp, gld, c.mix, c.npix, self.out.C, st)
        xp, xbs = self._x(c)
        g, acc = c.pgrad(self.weight)
        ws, wsb = c.workspace(_lib.lib().tsr_head_wgrad_workspace(c.B))
        _lib.call("tsr_head_wgrad", xp, xbs, gp, gld, c.grd, g.data_ptr(), ws, wsb, c.B, self.sf, acc, st)


class ConvOp(Op):
    """Conv2d(Cin -> Cout, k in {1,3,5}, same padding) with a fused epilogue: + bias, + residual, ReLU
    (reference tactileSR_model.py:41,47,53,168,174,180,186,191+205-206,219-225)."""

    def __init__(self, src: View, conv: torch.nn.Conv2d, out: View, relu: bool = False,
                 residual: Optional[View] = None, src_needs_grad: bool = True):
        self.src, self.conv, self.out, self.relu, self.residual = src, conv, out, relu, residual
        self.src_needs_grad = src_needs_grad
        self.bn_consumer = None      # set by the program builder when the output feeds a BNReLUOp directly
        self.K = conv.kernel_size[0]
        self.Cin, self.Cout = conv.in_channels, conv.out_channels
        assert src.C == self.Cin and out.C == self.Cout
        assert conv.stride == (1, 1) and conv.padding == (self.K // 2, self.K // 2) and conv.groups == 1

    def params(self):
        return (self.conv.weight,) if self.conv.bias is None else (self.conv.weight, self.conv.bias)

    def reads(self):
        return (self.src.buf,) if self.residual is None else (self.src.buf, self.residual.buf)

    def writes(self):
        return (self.out.buf,)

    def _conv(self, c: RunCtx, inp, inld, wpk, bias, res, resld, outp, outld, Cin, Cout, flags, on_grads=False,
              bn_partial=0, out2=0, out2_ld=0):
        st = _lib.stream_ptr()
        if c.tc:
            need = _lib.lib().tsr_conv2d_tc_workspace(c.B, c.H, c.W, Cin, Cout, self.K)
            ws, wsb = c.workspace(need)
            if c.act == 2 and not on_grads:
                flags |= 2      # fp16 activations / weights (forward); data gradients run on bf16 tensors
            _lib.call("tsr_conv2d_tc", inp, inld, wpk, bias, res, resld, outp, outld, c.B, c.H, c.W, Cin, Cout,
                      self.K, flags, ws, wsb, bn_partial, out2, out2_ld, st)
        else:
            _lib.call("tsr_conv2d_f32", inp, inld, wpk, bias, res, resld, outp, outld, c.B, c.H, c.W, Cin, Cout,
                      self.K, flags, st)

    def _wgrad_input(self, c: RunCtx):
        """(pointer, ld) of the conv input as the weight-gradient kernel wants it (bf16 in both tensor-core modes)."""
        return c.sptr(self.src) if c.act == 2 else c.vptr(self.src)

    def fwd(self, c):
        ip, ild = c.vptr(self.src)
        bn = self.bn_consumer
        if (c.tc and bn is not None and not c.training and not c.need_grad and not c.keep_taps and not self.relu
                and self.residual is None and bn.bn.running_mean is not None and bn.bn.weight is not None
                and bn.src.buf is self.out.buf
                and bn.src.c0 == self.out.c0 and bn.src.C == self.Cout):
            # inference: the eval-mode BatchNorm (+ReLU) that follows is folded into the weights / bias, and the result
            # goes straight into the BatchNorm's output slice -- no BN kernels, no intermediate tensor
            wf, bias = _PACK.get_folded(self.conv, bn.bn, c.mode)
            op, old = c.vptr(bn.out)
            self._conv(c, ip, ild, wf.data_ptr(), bias.data_ptr(), 0, 0, op, old, self.Cin, self.Cout,
                       1 if bn.relu else 0)
            c.folded.add(bn)
            return
        wf, _ = _PACK.get(self.conv.weight, c.mode, False)
        op, old = c.vptr(self.out)
        rp, rld = c.vptr(self.residual) if self.residual is not None else (0, 0)
        # batch statistics of the BatchNorm that consumes this output come out of the conv epilogue (tensor-core modes)
        part = 0
        if (c.tc and bn is not None and not (_lib.lib().tsr_get_tc_desc_mode() & 128)      # bit 7: separate statistics pass
                and not self.relu and self.residual is None and bn.src.buf is self.out.buf
                and bn.src.c0 == self.out.c0 and (c.training or bn.bn.running_mean is None)):
            t = torch.empty((_lib.lib().tsr_conv2d_tc_stat_rows(), 2, self.Cout), dtype=torch.float32, device=c.device)
            c.bn_partials[bn] = t
            part = t.data_ptr()
        if c.tc and c.wants_shadow(self.out.buf):     # bf16 shadow of an fp16 output straight from the epilogue
            o2, o2ld = c.sptr(self.out)
            self._conv(c, ip, ild, wf.data_ptr(), _ptr(self.conv.bias), rp, rld, op, old, self.Cin, self.Cout,
                       1 if self.relu else 0, bn_partial=part, out2=o2, out2_ld=o2ld)
        else:
            self._conv(c, ip, ild, wf.data_ptr(), _ptr(self.conv.bias), rp, rld, op, old, self.Cin, self.Cout,
                       1 if self.relu else 0, bn_partial=part)

    def bwd(self, c):
        st = _lib.stream_ptr()
        gp, gld = c.vptr(self.out, grad=True)
        if self.relu:
            ap, ald = c.vptr(self.out)
            _lib.call("tsr_relu_backward", gp, gld, ap, ald, gp, gld, c.mix, c.npix, self.Cout, st)
        # residual branch: d(residual) += dz
        if self.residual is not None:
            first = _first_write(c, self.residual.buf)
            rp, rld = c.vptr(self.residual, grad=True)
            if first:
                _lib.call("tsr_copy_channels", gp, gld, c.grd, rp, rld, c.grd, c.npix, self.Cout, st)
            else:
                raise NotImplementedError("residual gradient accumulation after another writer")
        # weight / bias gradients
        ip, ild = c.vptr(self.src)
        g, acc = c.pgrad(self.conv.weight)
        if c.tc:
            need = _lib.lib().tsr_conv2d_wgrad_tc_workspace(c.B, c.H, c.W, self.Cin, self.Cout, self.K)
            ws, wsb = c.workspace(need)
            xp, xld = self._wgrad_input(c)
            _lib.call("tsr_conv2d_wgrad_tc", xp, xld, gp, gld, g.data_ptr(), ws, wsb, c.B, c.H, c.W, self.Cin,
                      self.Cout, self.K, acc, st)
        else:
            need = _lib.lib().tsr_conv2d_wgrad_f32_workspace(c.B, c.H, c.W, self.Cin, self.Cout, self.K)
            ws, wsb = c.workspace(need)
            _lib.call("tsr_conv2d_wgrad_f32", ip, ild, gp, gld, g.data_ptr(), ws, wsb, c.B, c.H, c.W, self.Cin,
                      self.Cout, self.K, acc, st)
        if self.conv.bias is not None:
            gb, accb = c.pgrad(self.conv.bias)
            bn = self.bn_consumer
            if bn is not None and c.saved[bn][1]:
                # The gradient reaching a bias that feeds a batch-statistics BatchNorm is sum_pix dy with dy the BN
                # backward output, which is identically 0 (the reference computes fp32 rounding noise ~1e-8 of the layer's
                # gradient scale here, SURVEY section 0 pitfall 2): write the exact value instead of reducing 2 GB of zeros.
                if not accb:
                    gb.zero_()
            else:
                ws, wsb = c.workspace(_lib.lib().tsr_colsum_workspace(c.npix, self.Cout))
                _lib.call("tsr_colsum", gp, gld, c.grd, c.npix, self.Cout, gb.data_ptr(), ws, wsb, accb, st)
        # data gradient
        if self.src_needs_grad:
            _, wd = _PACK.get(self.conv.weight, c.mode, True)
            first = _first_write(c, self.src.buf)
            dp, dld = c.vptr(self.src, grad=True)
            self._conv(c, gp, gld, wd.data_ptr(), 0, 0 if first else dp, 0 if first else dld, dp, dld, self.Cout,
                       self.Cin, 0, on_grads=True)


class BNReLUOp(Op):
    """BatchNorm2d (train: batch statistics + running-stat update; eval: running stats) [+ ReLU]
    (reference tactileSR_model.py:38-39, 42-43, 48-49, 169-170, 175-176, 181-182, 187-188)."""

    def __init__(self, src: View, bn: torch.nn.BatchNorm2d, out: View, relu: bool = True):
        self.src, self.bn, self.out, self.relu = src, bn, out, relu
        assert src.C == bn.num_features == out.C

    def params(self):
        return (self.bn.weight, self.bn.bias)

    def reads(self):
        return (self.src.buf,)

    def writes(self):
        return (self.out.buf,)

    def fwd(self, c):
        if self in c.folded:
            return
        st = _lib.stream_ptr()
        bn, C = self.bn, self.src.C
        coef = torch.empty((4, C), dtype=torch.float32, device=c.device)
        sc, sh, mu, iv = (coef[i].data_ptr() for i in range(4))
        yp, yld = c.vptr(self.src)
        use_batch = c.training or bn.running_mean is None
        part = c.bn_partials.pop(self, None)
        if use_batch and c.training and bn.track_running_stats and bn.running_mean is not None:
            global _STATS_EPOCH
            _STATS_EPOCH += 1
        if use_batch and part is not None:
            track = c.training and bn.track_running_stats and bn.running_mean is not None
            mom = 0.1 if bn.momentum is None else bn.momentum
            _lib.call("tsr_bn_finalize_partials", part.data_ptr(), part.shape[0], c.npix, C, bn.weight.data_ptr(),
                      bn.bias.data_ptr(), _ptr(bn.running_mean) if track else 0, _ptr(bn.running_var) if track else 0,
                      _ptr(bn.num_batches_tracked) if track else 0, mom, bn.eps, sc, sh, mu, iv, st)
        elif use_batch:
            ws, wsb = c.workspace(_lib.lib().tsr_bn_workspace(c.npix, C))
            track = c.training and bn.track_running_stats and bn.running_mean is not None
            mom = 0.1 if bn.momentum is None else bn.momentum
            _lib.call("tsr_bn_train_stats", yp, yld, c.act, c.npix, C, bn.weight.data_ptr(), bn.bias.data_ptr(),
                      _ptr(bn.running_mean) if track else 0, _ptr(bn.running_var) if track else 0,
                      _ptr(bn.num_batches_tracked) if track else 0, mom, bn.eps, sc, sh, mu, iv, ws, wsb, st)
        else:
            _lib.call("tsr_bn_eval_coeffs", C, bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                      bn.running_var.data_ptr(), bn.eps, sc, sh, mu, iv, st)
        op, old = c.vptr(self.out)
        o2, o2ld = c.sptr(self.out) if c.wants_shadow(self.out.buf) else (0, 0)
        _lib.call("tsr_bn_apply", yp, yld, c.act, sc, sh, op, old, c.act, c.npix, C, 1 if self.relu else 0, o2, o2ld, st)
        c.saved[self] = (coef, use_batch)

    def bwd(self, c):
        st = _lib.stream_ptr()
        C = self.src.C
        coef, use_batch = c.saved[self]
        sc, sh, mu, iv = (coef[i].data_ptr() for i in range(4))
        gp, gld = c.vptr(self.out, grad=True)
        yp, yld = c.vptr(self.src)
        first = _first_write(c, self.src.buf)
        assert first, "BN input has a single consumer"
        dp, dld = c.vptr(self.src, grad=True)
        gw, accw = c.pgrad(self.bn.weight)
        gb, accb = c.pgrad(self.bn.bias)
        assert accw == accb
        ws, wsb = c.workspace(_lib.lib().tsr_bn_backward_workspace(c.npix, C))
        _lib.call("tsr_bn_backward", gp, gld, yp, yld, dp, dld, c.mix, sc, sh, mu, iv, gw.data_ptr(), gb.data_ptr(),
                  accw, c.npix, C, 1 if self.relu else 0, 1 if use_batch else 0, ws, wsb, st)


class TailOp(Op):
    """Conv2d(Cin -> 1, 3x3, no bias) + ReLU writing the (B,1,H,W) fp32 result
    (reference tactileSR_model.py:55-56, 125-126)."""

    def __init__(self, src: View, conv: torch.nn.Conv2d, out: Buf, relu: bool = True):
        self.src, self.conv, self.out, self.relu = src, conv, out, relu
        assert conv.out_channels == 1 and conv.kernel_size == (3, 3) and conv.bias is None

    def params(self):
        return (self.conv.weight,)

    def reads(self):
        return (self.src.buf,)

    def writes(self):
        return (self.out,)

    def fwd(self, c):
        ip, ild = c.vptr(self.src)
        _lib.call("tsr_tail_fwd", ip, ild, c.act, self.conv.weight.data_ptr(), c.alloc(self.out).data_ptr(), c.B, c.H,
                  c.W, self.src.C, 1 if self.relu else 0, _lib.stream_ptr())

    def bwd(self, c):
        st = _lib.stream_ptr()
        dout = c.grads[self.out]
        out = c.bufs[self.out]
        ip, ild = c.vptr(self.src)
        g, acc = c.pgrad(self.conv.weight)
        ws, wsb = c.workspace(_lib.lib().tsr_tail_wgrad_workspace(c.B, c.H, c.W, self.src.C))
        _lib.call("tsr_tail_wgrad", ip, ild, c.act, dout.data_ptr(), out.data_ptr(), g.data_ptr(), ws, wsb, c.B, c.H,
                  c.W, self.src.C, 1 if self.relu else 0, acc, st)
        first = _first_write(c, self.src.buf)
        assert first
        dp, dld = c.vptr(self.src, grad=True)
        _lib.call("tsr_tail_dgrad", dout.data_ptr(), out.data_ptr(), self.conv.weight.data_ptr(), dp, dld, c.grd, c.B,
                  c.H, c.W, self.src.C, 1 if self.relu else 0, st)


class InputOp(Op):
    """NCHW fp32 module input -> NHWC activation buffer (standalone MSRB / ResBlock use)."""

    def __init__(self, out: Buf):
        self.out = out

    def writes(self):
        return (self.out,)

    def fwd(self, c):
        x = c.x
        t = c.alloc(self.out)
        _lib.call("tsr_nchw_to_nhwc", x.data_ptr(), t.data_ptr(), self.out.C, c.act, c.B, self.out.C, c.H * c.W,
                  _lib.stream_ptr())
        c.shadow_fill(View.of(self.out))

    def bwd(self, c):
        pass


# ---------------------------------------------------------------------------------------------
# program
# ---------------------------------------------------------------------------------------------
@dataclass
class Program:
    ops: List[Op] = field(default_factory=list)
    out: Optional[Buf] = None          # result buffer: "plane" (B,1,H,W) or "act" (returned as NCHW fp32)
    sf: int = 10                       # spatial size = 4*sf for head programs; None => taken from the input
    input_is_taxel: bool = True
    wants_input_grad: bool = False
    in_buf: Optional[Buf] = None
    taps: Dict[str, View] = field(default_factory=dict)

    def add(self, op: Op) -> Op:
        self.ops.append(op)
        return op

    def parameters(self) -> List[torch.nn.Parameter]:
        seen, out = set(), []
        for op in self.ops:
            for p in op.params():
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
        return out


def _check_device(x: torch.Tensor) -> None:
    if not x.is_cuda:
        raise _lib.TsrError("tactilesr_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback")


def run_forward(prog: Program, x: torch.Tensor, training: bool, need_grad: bool, mode: Optional[str] = None,
                keep_taps: bool = False):
    _check_device(x)
    mode = mode or _PRECISION
    x = x.detach().contiguous().float()
    B = x.shape[0]
    if prog.input_is_taxel:
        H = W = x.shape[-1] * prog.sf
    else:
        H, W = x.shape[-2], x.shape[-1]
    c = RunCtx(mode, B, H, W, x.device, training, need_grad)
    c.x = x
    c.conv_inputs = {op.src.buf for op in prog.ops if isinstance(op, ConvOp)}
    c.keep_taps = keep_taps
    if c.tc and (training or need_grad):
        seen, ws = set(), []
        for op in prog.ops:
            if isinstance(op, ConvOp) and id(op.conv.weight) not in seen and op.conv.weight.is_contiguous():
                seen.add(id(op.conv.weight))
                ws.append(op.conv.weight)
        if ws:
            _PACK.prepack(ws, mode, need_grad)
    keep = need_grad or keep_taps
    last_use: Dict[Buf, int] = {}
    if not keep:
        for i, op in enumerate(prog.ops):
            for b in op.reads():
                last_use[b] = i
    for i, op in enumerate(prog.ops):
        op.fwd(c)
        if not keep:
            for b in op.reads():
                if last_use.get(b) == i and b is not prog.out:
                    c.bufs.pop(b, None)
    if prog.out.kind == "plane":
        out = c.bufs[prog.out].view(B, 1, H, W)
    else:
        out = torch.empty((B, prog.out.C, H, W), dtype=torch.float32, device=x.device)
        t = c.bufs[prog.out]
        _lib.call("tsr_nhwc_to_nchw", t.data_ptr(), prog.out.C, c.act, out.data_ptr(), B, prog.out.C, H * W,
                  _lib.stream_ptr())
    return out, c


def run_backward(prog: Program, c: RunCtx, dout: torch.Tensor, hooks=None) -> Dict[torch.nn.Parameter, torch.Tensor]:
    dout = dout.detach().contiguous().float()
    if prog.out.kind == "plane":
        c.grads[prog.out] = dout.view(-1)
        c.grad_written[prog.out] = True
    else:
        g = c.galloc(prog.out)
        c.grad_written[prog.out] = True
        _lib.call("tsr_nchw_to_nhwc", dout.data_ptr(), g.data_ptr(), prog.out.C, c.grd, c.B, prog.out.C, c.H * c.W,
                  _lib.stream_ptr())
    # reverse sweep; gradient buffers are dropped as soon as their producer has consumed them
    for i in range(len(prog.ops) - 1, -1, -1):
        op = prog.ops[i]
        needed = [b for b in op.writes()]
        if any(b not in c.grads for b in needed):
            continue   # dead branch (no gradient reaches this op)
        op.bwd(c)
        if hooks is not None:
            hooks(i, op, c)
        for b in needed:
            if not (prog.wants_input_grad and b is prog.in_buf) and not _written_earlier(prog, i, b):
                c.grads.pop(b, None)
    return c.param_grads


def _written_earlier(prog: Program, i: int, b: Buf) -> bool:
    """True if an op before index i also writes (a slice of) buffer b, i.e. its gradient is still needed."""
    for j in range(i):
        if b in prog.ops[j].writes():
            return True
    return False


class _ProgramFn(torch.autograd.Function):
    """One autograd node for a whole layer program: forward + hand-written backward."""

    @staticmethod
    def forward(ctx, holder, x, *params):
        prog: Program = holder["prog"]
        need_grad = holder["need_grad"]
        out, c = run_forward(prog, x, holder["training"], need_grad, holder.get("mode"))
        if need_grad:
            ctx.prog, ctx.c, ctx.params, ctx.holder = prog, c, params, holder
        holder["ctx"] = c
        return out

    @staticmethod
    def backward(ctx, dout):
        prog, c = ctx.prog, ctx.c
        if c is None:
            raise RuntimeError("tactilesr_b200: backward through the same forward twice is not supported")
        hooks = ctx.holder.get("grad_hook")
        pg = run_backward(prog, c, dout, hooks)
        # Parameter gradients are delivered straight into ``.grad`` (first contribution: the tensor the kernels wrote --
        # for FusedAdam parameters a view of its flat gradient buffer, so no copy and the data-parallel all-reduce works
        # in place; later contributions: accumulated).  Returning them through autograd instead would make
        # AccumulateGrad clone every tensor we still hold a reference to.
        grads = [None] * len(ctx.params)
        for p in ctx.params:
            g = pg.get(p)
            if g is None or not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = g
            else:
                p.grad.add_(g)
        gx = None
        if prog.wants_input_grad and ctx.needs_input_grad[1]:
            gb = c.grads.get(prog.in_buf)
            if gb is not None:
                gx = torch.empty((c.B, prog.in_buf.C, c.H, c.W), dtype=torch.float32, device=c.device)
                _lib.call("tsr_nhwc_to_nchw", gb.data_ptr(), prog.in_buf.C, c.grd, gx.data_ptr(), c.B, prog.in_buf.C,
                          c.H * c.W, _lib.stream_ptr())
        fin = ctx.holder.get("grad_done")
        if fin is not None:
            fin(pg)
        ctx.c = None   # release activations
        return (None, gx, *grads)


def apply_program(prog: Program, x: torch.Tensor, training: bool, mode: Optional[str] = None,
                  extra: Optional[dict] = None) -> torch.Tensor:
    params = prog.parameters()
    need_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or
                                             (prog.wants_input_grad and x.requires_grad))
    holder = {"prog": prog, "training": training, "need_grad": need_grad, "mode": mode}
    if extra:
        holder.update(extra)
    return _ProgramFn.apply(holder, x, *params)
