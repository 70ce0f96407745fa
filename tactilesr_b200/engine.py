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
    are needed every conv keeps a bf16 copy ("shadow") of its input for the weight-gradient kernel.  (TSR_TC_MODE bit 10
    trades them for an in-kernel conversion: 9 MB per sample less memory, but the 128-channel weight gradients run ~1.8x
    slower -- their MMAs already take the whole shared-memory bandwidth, measured in round 2.)
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


_OVERFLOW_FLAG: Dict[int, torch.Tensor] = {}       # device index -> int32 flag registered with the library


def _arm_overflow_guard(device) -> None:
    """"fp16" mode: hand the library a sticky device flag that its fp16-storing kernels raise on a non-finite value."""
    idx = torch.device(device).index or 0
    t = _OVERFLOW_FLAG.get(idx)
    if t is None:
        t = torch.zeros((), dtype=torch.int32, device=device)
        _OVERFLOW_FLAG[idx] = t
    _lib.lib().tsr_set_f16_overflow_flag(t.data_ptr())


def check_fp16_overflow(device=None) -> None:
    """Raise if a stored fp16 activation overflowed since the last check (one device -> host read: call it at logging
    cadence, as the trainer does, not per step).  fp16 holds |x| <= 65504; un-normalised activations beyond that need the
    bf16 or fp32 mode."""
    for idx, t in _OVERFLOW_FLAG.items():
        if device is not None and (torch.device(device).index or 0) != idx:
            continue
        if int(t.item()) != 0:
            t.zero_()
            raise _lib.TsrError("tactilesr_b200: an fp16 activation overflowed (|x| > 65504 or NaN) in the 'fp16' precision "
                                "mode -- use set_precision('bf16') or set_precision('fp32') for this model / data")


def bump_weight_epoch() -> None:
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1


def invalidate_packed_weights() -> None:
    """Drop every cached packed / folded copy of the convolution weights at the next forward.

    The caches are keyed on ``(data_ptr, Tensor._version, optimizer epoch)``.  ``FusedAdam.step`` bumps the epoch itself and
    ``load_state_dict`` / ``copy_`` / ``nn.init`` bump ``_version``; writes that bypass both -- ``p.data.copy_(...)``,
    ``p.data.mul_(...)`` (EMA, manual re-initialisation), a foreign optimizer implemented with raw kernels -- must call this
    once afterwards."""
    bump_weight_epoch()


# ---------------------------------------------------------------------------------------------
# optional per-kernel-class timing (bench.py: step_breakdown_ms); off by default, no cost when off
# ---------------------------------------------------------------------------------------------
_PROF: Optional[Dict[str, list]] = None


def profile_begin() -> None:
    global _PROF
    _PROF = {}


def profile_end() -> Dict[str, float]:
    """Milliseconds per kernel class since profile_begin() (CUDA events on the launching stream)."""
    global _PROF
    rec, _PROF = _PROF, None
    torch.cuda.synchronize()
    return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in (rec or {}).items()}


class _prof:
    def __init__(self, key: str):
        self.key = key

    def __enter__(self):
        if _PROF is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if _PROF is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _PROF.setdefault(self.key, []).append((self.e0, e1))


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
        self.shadow: Dict[Buf, torch.Tensor] = {}
        self.conv_inputs: set = set()
        self.no_shadow = bool(_lib.lib().tsr_get_tc_desc_mode() & 1024) if self.tc else False
        self.bn_partials: Dict[object, torch.Tensor] = {}      # BNReLUOp -> statistics partials from its convs' epilogues
        self.bn_partials_parts: Dict[object, set] = {}         # ... and which of its parts they cover
        self.folded: Dict[object, set] = {}                    # BNReLUOp -> parts already applied by their (inference) conv
        self.grad_masked: set = set()                          # buffers whose gradient already carries the ReLU mask
        self.grad_alias: Dict[Buf, tuple] = {}                 # residual gradients handed over by reference (ptr, ld, keepalive)
        self.bnb: Dict[object, torch.Tensor] = {}              # BNReLUOp -> (sum g, sum g*y) partials from a dgrad epilogue
        self.keep: list = []                                   # small temporaries that must outlive the launches reading them
        self.keep_taps = False
        self._ws: Optional[torch.Tensor] = None
        # zero-initialised fp32 scratch (statistics tables the conv epilogues add to): carved out of one arena per sweep, so a
        # sweep costs ONE fill launch instead of one per table (24 per training step of TactileSR)
        self.zero_cap = 0
        self._zarena: Optional[torch.Tensor] = None
        self._zoff = 0
        self.zero_later: list = []                             # gradients that are identically zero: cleared by one launch
        self.defer_zero = True

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
        return self.act == 2 and self.need_grad and not self.no_shadow and buf in self.conv_inputs

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

    def zeros_f32(self, n: int) -> torch.Tensor:
        """n fp32 zeros from the sweep's arena (256-byte aligned); a too-small estimate just opens another zeroed block."""
        n4 = (n + 63) // 64 * 64
        if self._zarena is None or self._zoff + n4 > self._zarena.numel():
            self._zarena = torch.zeros(max(self.zero_cap, n4), dtype=torch.float32, device=self.device)
            self._zoff = 0
        t = self._zarena[self._zoff:self._zoff + n]
        self._zoff += n4
        return t

    def new_sweep(self) -> None:
        """Backward starts with a fresh arena (tables of the forward may still be referenced)."""
        self._zarena, self._zoff = None, 0

    def stat_table(self, bnop, part: int) -> torch.Tensor:
        """Zeroed [rows][2][channels of bnop] table that the conv epilogue(s) feeding ``bnop`` add their statistics to."""
        t = self.bn_partials.get(bnop)
        if t is None:
            rows = _lib.lib().tsr_conv2d_tc2_stat_rows()
            t = self.zeros_f32(rows * 2 * bnop.src.C).view(rows, 2, bnop.src.C)
            self.bn_partials[bnop] = t
            self.bn_partials_parts[bnop] = set()
        self.bn_partials_parts[bnop].add(part)
        return t

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

    @staticmethod
    def _tag(w: torch.Tensor):
        return (w.data_ptr(), w._version, _WEIGHT_EPOCH, tuple(w.shape), w.device)

    def get(self, w: torch.Tensor, mode: str, need_dgrad: bool, need_fwd: bool = True):
        """(forward image, data-gradient image) of conv weight ``w``; an image that is not asked for may be None."""
        tag = self._tag(w)
        store = w.__dict__.setdefault("_tsr_pack", {})
        hit = store.get(mode)
        if hit is not None and hit[0] != tag:
            hit = None
        wf = hit[1] if hit is not None else None
        wd = hit[2] if hit is not None else None
        if (wf is not None or not need_fwd) and (wd is not None or not need_dgrad):
            return wf, wd
        Cout, Cin, K, _ = w.shape
        n = K * K * Cin * Cout
        wc = w.detach().contiguous()
        st = _lib.stream_ptr()
        mk_f, mk_d = need_fwd and wf is None, need_dgrad and wd is None
        if mode == "fp16":       # forward weights fp16, data-gradient weights bf16 (gradients are bf16 tensors)
            if mk_f:
                wf = torch.empty((n,), dtype=torch.float16, device=w.device)
                _lib.call("tsr_pack_conv_weight_f16", wc.data_ptr(), wf.data_ptr(), 0, Cout, Cin, K, st)
            if mk_d:
                wd = torch.empty((n,), dtype=torch.bfloat16, device=w.device)
                _lib.call("tsr_pack_conv_weight_bf16", wc.data_ptr(), 0, wd.data_ptr(), Cout, Cin, K, st)
        else:
            dt = torch.bfloat16 if mode == "bf16" else torch.float32
            if mk_f:
                wf = torch.empty((n,), dtype=dt, device=w.device)
            if mk_d:
                wd = torch.empty((n,), dtype=dt, device=w.device)
            fn = "tsr_pack_conv_weight_bf16" if mode == "bf16" else "tsr_pack_conv_weight_f32"
            _lib.call(fn, wc.data_ptr(), wf.data_ptr() if mk_f else 0, wd.data_ptr() if mk_d else 0, Cout, Cin, K, st)
        store[mode] = (tag, wf, wd)
        return wf, wd

    def get_dual(self, w3: torch.Tensor, w5: torch.Tensor, mode: str) -> torch.Tensor:
        """Dual-branch forward image of a (3x3, 5x5) weight pair (tsr_pack_conv_weight_dual); kept on ``w3``."""
        tag = (self._tag(w3), self._tag(w5))
        store = w3.__dict__.setdefault("_tsr_pack", {})
        hit = store.get(mode + ":dual")
        if hit is not None and hit[0] == tag:
            return hit[1]
        Cin = w3.shape[1]
        img = hit[1] if hit is not None else torch.empty((_lib.lib().tsr_pack_conv_weight_dual_elems(Cin),), device=w3.device,
                                                         dtype=torch.float16 if mode == "fp16" else torch.bfloat16)
        _lib.call("tsr_pack_conv_weight_dual", w3.detach().contiguous().data_ptr(), w5.detach().contiguous().data_ptr(),
                  img.data_ptr(), Cin, 2 if mode == "fp16" else 1, _lib.stream_ptr())
        store[mode + ":dual"] = (tag, img)
        return img


    # -- all conv weights of a program in one launch ---------------------------------------------------------------
    def prepack(self, items, mode: str, need_dgrad: bool) -> None:
        """Tensor-core modes: bring the packed images of every weight in ``items`` = [(weight, role)] up to date with ONE
        kernel launch (the optimizer changes all of them at every step) and leave the per-parameter cache entries current,
        so the ops' ``get`` / ``get_dual`` calls all hit.  role 0: standard forward image; roles 1 / 2: the 3x3 / 5x5 branch
        of a dual-branch image shared by two consecutive items.  Buffers and the device descriptor table are persistent
        per weight set."""
        import numpy as np
        weights = [w for w, _ in items]
        key = (mode,) + tuple((id(w), r) for w, r in items)
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
                                                          ("mode", "<i4")]))
            bufs = []
            for i, (w, role) in enumerate(items):
                Cout, Cin, K, _ = w.shape
                old = grp["bufs"][i] if i < len(grp["bufs"]) else (None, None)
                if role == 2:
                    wf = bufs[i - 1][0]                       # the image of the preceding role-1 item
                elif old[0] is not None:
                    wf = old[0]
                elif role == 1:
                    wf = torch.empty((_lib.lib().tsr_pack_conv_weight_dual_elems(Cin),), dtype=dt_f, device=w.device)
                else:
                    wf = torch.empty((K * K * Cin * Cout,), dtype=dt_f, device=w.device)
                wd = old[1]
                if want_d and wd is None:
                    wd = torch.empty((K * K * Cin * Cout,), dtype=torch.bfloat16, device=w.device)
                bufs.append((wf, wd))
                assert w.is_contiguous()
                desc[i] = (w.data_ptr(), wf.data_ptr(), 0 if wd is None else wd.data_ptr(), Cout, Cin, K,
                           2 if mode == "fp16" else 1, 1, role)
            grp["bufs"], grp["dgrad"], grp["ptrs"] = bufs, want_d, ptrs
            grp["table"] = torch.from_numpy(desc.view(np.uint8).copy()).to(weights[0].device)
            grp["max"] = max(int(w.numel()) for w in weights)
        _lib.call("tsr_pack_conv_weights_multi", grp["table"].data_ptr(), len(weights), grp["max"], _lib.stream_ptr())
        grp["sig"] = sig
        for i, ((w, role), (wf, wd)) in enumerate(zip(items, grp["bufs"])):
            store = w.__dict__.setdefault("_tsr_pack", {})
            if role == 0:
                store[mode] = (self._tag(w), wf, wd)
            else:
                store[mode] = (self._tag(w), None, wd)
                if role == 1:
                    store[mode + ":dual"] = ((self._tag(w), self._tag(items[i + 1][0])), wf)

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

    # -- static structure used by Program.plan() --------------------------------------------------
    def out_views(self) -> Sequence[View]:
        """Activation views this op produces."""
        return ()

    def grad_targets(self) -> Sequence[Tuple[View, str]]:
        """(view, kind) pairs whose gradient this op's backward writes: kind 'conv' (data gradient of a convolution),
        'res' (residual hand-over), 'other'."""
        return ()


def _first_write(c: RunCtx, buf: Buf) -> bool:
    """True if this is the first gradient contribution to ``buf`` in the current backward."""
    c.galloc(buf)
    first = not c.grad_written[buf]
    c.grad_written[buf] = True
    return first


@dataclass
class Sink:
    """Post-processing the LAST data-gradient writer of a buffer applies in its epilogue (Program.plan()):
    kind 'relu': the buffer is the ReLU output of convolutions -> zero the gradient where the activation is <= 0
    (replaces the tsr_relu_backward pass of the producing ops); kind 'bn': the buffer is the output of one BatchNorm(+ReLU)
    op -> mask through scale*y + shift > 0 and reduce (sum g, sum g*y) for its backward (replaces the bn_bwd_partial pass)."""
    kind: str
    bn: Optional["BNReLUOp"] = None


def _stat_rows() -> int:
    return _lib.lib().tsr_conv2d_tc2_stat_rows()


class HeadOp(Op):
    """Upsample(x sf, bilinear) + Conv2d(3 -> 64, 3x3, no bias) [+ ReLU]
    (reference tactileSR_model.py:35-37, 60-62, 107 + 122)."""

    def __init__(self, ch0: int, conv: torch.nn.Conv2d, out: View, relu: bool, sf: int):
        self.ch0, self.conv, self.out, self.relu, self.sf = ch0, conv, out, relu, sf

    @property
    def weight(self):              # read through the module at run time (programs are cached across parameter updates)
        return self.conv.weight

    def params(self):
        return (self.weight,)

    def writes(self):
        return (self.out.buf,)

    def out_views(self):
        return (self.out,)

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
        if self.relu and self.out.buf not in c.grad_masked:
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
        self.bn_consumer: Optional["BNReLUOp"] = None   # set by the program builder when the output feeds a BNReLUOp directly
        self.bn_part = 0                                # ... as its part number `bn_part`
        self.sink: Optional[Sink] = None                # Program.plan(): fused epilogue of the data gradient into `src`
        self.alias_residual_grad = False                # Program.plan(): d(residual) may be handed over by reference
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

    def out_views(self):
        return (self.out,)

    def grad_targets(self):
        t = [(self.src, "conv")] if self.src_needs_grad else []
        if self.residual is not None:
            t.append((self.residual, "res"))
        return t

    # -- forward ------------------------------------------------------------------------------------
    def _bn_slice(self):
        """(BatchNorm module, output view) of the consumer part this convolution feeds."""
        bnop = self.bn_consumer
        bn, off, C = bnop.parts[self.bn_part]
        return bn, View(bnop.out.buf, bnop.out.c0 + off, C)

    def _feeds_bn(self) -> bool:
        bnop = self.bn_consumer
        if bnop is None or self.relu or self.residual is not None or bnop.src.buf is not self.out.buf:
            return False
        _, off, C = bnop.parts[self.bn_part]
        return bnop.src.c0 + off == self.out.c0 and C == self.Cout

    def fwd(self, c):
        ip, ild = c.vptr(self.src)
        st = _lib.stream_ptr()
        if not c.tc:
            wf, _ = _PACK.get(self.conv.weight, c.mode, False)
            op, old = c.vptr(self.out)
            rp, rld = c.vptr(self.residual) if self.residual is not None else (0, 0)
            _lib.call("tsr_conv2d_f32", ip, ild, wf.data_ptr(), _ptr(self.conv.bias), rp, rld, op, old, c.B, c.H, c.W, self.Cin,
                      self.Cout, self.K, 1 if self.relu else 0, st)
            return
        f16 = _lib.TC2_F16 if c.act == 2 else 0
        feeds = self._feeds_bn()
        if feeds and not c.training and not c.need_grad and not c.keep_taps:
            bn, bn_out = self._bn_slice()
            if bn.running_mean is not None and bn.weight is not None:
                # inference: the eval-mode BatchNorm (+ReLU) that follows is folded into the weights / bias, and the result
                # goes straight into the BatchNorm's output slice -- no BN kernels, no intermediate tensor
                wf, bias = _PACK.get_folded(self.conv, bn, c.mode)
                op, old = c.vptr(bn_out)
                _lib.conv_tc2([(ip, ild, self.Cin, self.K, wf.data_ptr())], op, old, c.B, c.H, c.W, self.Cout,
                              flags=f16 | (_lib.TC2_RELU if self.bn_consumer.relu else 0), bias=bias.data_ptr(), stream=st)
                c.folded.setdefault(self.bn_consumer, set()).add(self.bn_part)
                return
        wf, _ = _PACK.get(self.conv.weight, c.mode, False)
        op, old = c.vptr(self.out)
        rp, rld = c.vptr(self.residual) if self.residual is not None else (0, 0)
        kw = {}
        flags = f16 | (_lib.TC2_RELU if self.relu else 0)
        # batch statistics of the BatchNorm that consumes this output come out of the conv epilogue
        if feeds and not (_lib.lib().tsr_get_tc_desc_mode() & 128):      # bit 7: separate statistics pass
            bn, _ = self._bn_slice()
            if c.training or bn.running_mean is None:
                t = c.stat_table(self.bn_consumer, self.bn_part)
                off = self.bn_consumer.parts[self.bn_part][1]
                kw.update(stat=t.data_ptr() + off * 4, stat_ld=t.shape[2])
                flags |= _lib.TC2_STAT_PRECLEARED
        if c.wants_shadow(self.out.buf):     # bf16 shadow of an fp16 output straight from the epilogue
            o2, o2ld = c.sptr(self.out)
            kw.update(out2=o2, out2_ld=o2ld)
        _lib.conv_tc2([(ip, ild, self.Cin, self.K, wf.data_ptr())], op, old, c.B, c.H, c.W, self.Cout, flags=flags,
                      bias=_ptr(self.conv.bias), residual=rp, res_ld=rld, stream=st, **kw)

    # -- backward -----------------------------------------------------------------------------------
    def bwd(self, c):
        self.bwd_pre(c)
        self.bwd_weights(c)
        if self.src_needs_grad:
            _dgrad(c, [self], self.sink)

    def bwd_pre(self, c):
        """ReLU backward of the fused epilogue (unless the producer of the gradient already masked it) and the residual
        branch: d(residual) += dz."""
        st = _lib.stream_ptr()
        gp, gld = c.vptr(self.out, grad=True)
        if self.relu and self.out.buf not in c.grad_masked:
            ap, ald = c.vptr(self.out)
            _lib.call("tsr_relu_backward", gp, gld, ap, ald, gp, gld, c.mix, c.npix, self.Cout, st)
        if self.residual is not None:
            rbuf = self.residual.buf
            if self.alias_residual_grad and c.tc and rbuf not in c.grads and rbuf not in c.grad_alias:
                # the next (and last) writer of this gradient is a tensor-core data gradient: it reads dz through its
                # residual input instead of a copy being made here
                c.grad_alias[rbuf] = (gp, gld, c.grads[self.out.buf])
                return
            first = _first_write(c, rbuf)
            rp, rld = c.vptr(self.residual, grad=True)
            if first:
                _lib.call("tsr_copy_channels", gp, gld, c.grd, rp, rld, c.grd, c.npix, self.Cout, st)
            else:
                raise NotImplementedError("residual gradient accumulation after another writer")

    def bwd_weights(self, c):
        with _prof("bwd:wgrad"):
            self._bwd_weights(c)

    def _bwd_weights(self, c):
        st = _lib.stream_ptr()
        gp, gld = c.vptr(self.out, grad=True)
        g, acc = c.pgrad(self.conv.weight)
        if c.tc:
            need = _lib.lib().tsr_conv2d_wgrad_tc_workspace(c.B, c.H, c.W, self.Cin, self.Cout, self.K)
            ws, wsb = c.workspace(need)
            # x as bf16: the activation itself (bf16 mode), its bf16 shadow (fp16 mode), or -- without shadows -- the fp16
            # activation, converted in shared memory by the kernel
            if c.act == 2 and not c.no_shadow:
                (xp, xld), xdt = c.sptr(self.src), 1
            else:
                (xp, xld), xdt = c.vptr(self.src), c.act
            _lib.call("tsr_conv2d_wgrad_tc_x", xp, xld, xdt, gp, gld, g.data_ptr(), ws, wsb, c.B, c.H, c.W, self.Cin,
                      self.Cout, self.K, acc, st)
        else:
            ip, ild = c.vptr(self.src)
            need = _lib.lib().tsr_conv2d_wgrad_f32_workspace(c.B, c.H, c.W, self.Cin, self.Cout, self.K)
            ws, wsb = c.workspace(need)
            _lib.call("tsr_conv2d_wgrad_f32", ip, ild, gp, gld, g.data_ptr(), ws, wsb, c.B, c.H, c.W, self.Cin,
                      self.Cout, self.K, acc, st)
        if self.conv.bias is not None:
            gb, accb = c.pgrad(self.conv.bias)
            bnop = self.bn_consumer
            if bnop is not None and self._feeds_bn() and c.saved[bnop][1][self.bn_part]:
                # The gradient reaching a bias that feeds a batch-statistics BatchNorm is sum_pix dy with dy the BN
                # backward output, which is identically 0 (the reference computes fp32 rounding noise ~1e-8 of the layer's
                # gradient scale here, SURVEY section 0 pitfall 2): write the exact value instead of reducing 2 GB of zeros.
                if not accb:
                    if c.defer_zero:
                        c.zero_later.append(gb)
                    else:
                        gb.zero_()
            else:
                ws, wsb = c.workspace(_lib.lib().tsr_colsum_workspace(c.npix, self.Cout))
                _lib.call("tsr_colsum", gp, gld, c.grd, c.npix, self.Cout, gb.data_ptr(), ws, wsb, accb, st)


def _dgrad(c: RunCtx, ops: List[ConvOp], sink: Optional[Sink]) -> None:
    with _prof("bwd:dgrad"):
        _dgrad_impl(c, ops, sink)


def _dgrad_impl(c: RunCtx, ops: List[ConvOp], sink: Optional[Sink]) -> None:
    """Data gradient of one convolution, or of two convolutions that read the same view (their contributions are
    K-concatenated into ONE tensor-core launch), into the gradient of ``ops[0].src`` -- adding to what earlier writers
    left there (through the epilogue's residual input) and applying the buffer's Sink when this is its last writer."""
    st = _lib.stream_ptr()
    src = ops[0].src
    buf = src.buf
    if not c.tc:
        for op in ops:
            _, wd = _PACK.get(op.conv.weight, c.mode, True)
            first = _first_write(c, buf)
            dp, dld = c.vptr(src, grad=True)
            gp, gld = c.vptr(op.out, grad=True)
            _lib.call("tsr_conv2d_f32", gp, gld, wd.data_ptr(), 0, 0 if first else dp, 0 if first else dld, dp, dld, c.B, c.H,
                      c.W, op.Cout, op.Cin, op.K, 0, st)
        return
    alias = c.grad_alias.pop(buf, None)
    first = _first_write(c, buf)
    dp, dld = c.vptr(src, grad=True)
    if alias is not None:
        rp, rld = alias[0], alias[1]
    elif not first:
        rp, rld = dp, dld
    else:
        rp, rld = 0, 0
    srcs = []
    for op in ops:
        _, wd = _PACK.get(op.conv.weight, c.mode, True, need_fwd=False)
        gp, gld = c.vptr(op.out, grad=True)
        srcs.append((gp, gld, op.Cout, op.K, wd.data_ptr()))
    if len(srcs) == 2 and (srcs[0][3] == 1 or srcs[1][3] == 1):        # 1x1 sources cannot be K-concatenated: two launches
        _lib.conv_tc2(srcs[:1], dp, dld, c.B, c.H, c.W, src.C, residual=rp, res_ld=rld, stream=st)
        srcs, rp, rld = srcs[1:], dp, dld
    flags, kw = 0, {}
    whole = src.c0 == 0 and src.C == buf.C
    if sink is not None and whole and not (_lib.lib().tsr_get_tc_desc_mode() & 256):     # bit 8: no fused gradient sinks
        auxf = _lib.TC2_AUX_F16 if c.act == 2 else 0
        if sink.kind == "relu":
            ap, ald = c.vptr(View.of(buf))
            flags = _lib.TC2_MASK | auxf
            kw.update(aux=ap, aux_ld=ald)
            c.grad_masked.add(buf)
        elif sink.kind == "bn" and sink.bn in c.saved:
            bnop = sink.bn
            coef = c.saved[bnop][0]
            yp, yld = c.vptr(bnop.src)
            part = c.zeros_f32(_stat_rows() * 2 * buf.C).view(_stat_rows(), 2, buf.C)
            flags = _lib.TC2_BNB | (_lib.TC2_BNB_RELU if bnop.relu else 0) | auxf | _lib.TC2_STAT_PRECLEARED
            kw.update(aux=yp, aux_ld=yld, aux_scale=coef[0].data_ptr(), aux_shift=coef[1].data_ptr(), stat=part.data_ptr(),
                      stat_ld=buf.C)
            c.bnb[bnop] = part
    _lib.conv_tc2(srcs, dp, dld, c.B, c.H, c.W, src.C, flags=flags, residual=rp, res_ld=rld, stream=st, **kw)


class DualConvOp(Op):
    """The 3x3 and the 5x5 convolution of one input, outputs side by side in one buffer (MSRB.forward, reference
    tactileSR_model.py:198-199 and :201-202; the torch.cat of :200 / :203 is the shared buffer).  Forward: ONE dual-branch
    tensor-core launch when both produce 64 channels (the 9 shared taps run at N = 128), else two launches; backward: two
    weight gradients and ONE K-concatenated data gradient."""

    def __init__(self, src: View, conv_a: torch.nn.Conv2d, conv_b: torch.nn.Conv2d, out: View):
        Ca, Cb = conv_a.out_channels, conv_b.out_channels
        assert out.C == Ca + Cb
        self.a = ConvOp(src, conv_a, View(out.buf, out.c0, Ca))
        self.b = ConvOp(src, conv_b, View(out.buf, out.c0 + Ca, Cb))
        self.src, self.out = src, out
        self.sink: Optional[Sink] = None
        self.dual_ok = (self.a.K == 3 and self.b.K == 5 and Ca == 64 and Cb == 64 and src.C % 64 == 0)

    def feed(self, bnop: "BNReLUOp") -> "BNReLUOp":
        self.a.bn_consumer, self.a.bn_part = bnop, 0
        self.b.bn_consumer, self.b.bn_part = bnop, 1
        return bnop

    def params(self):
        return tuple(self.a.params()) + tuple(self.b.params())

    def reads(self):
        return (self.src.buf,)

    def writes(self):
        return (self.out.buf,)

    def out_views(self):
        return (self.a.out, self.b.out)

    def grad_targets(self):
        return [(self.src, "conv")]

    def fwd(self, c):
        bnop = self.a.bn_consumer
        use_dual = (c.tc and self.dual_ok and (c.training or c.need_grad) and bnop is not None and bnop is self.b.bn_consumer
                    and self.a._feeds_bn() and self.b._feeds_bn() and not (_lib.lib().tsr_get_tc_desc_mode() & 512))   # bit 9
        if not use_dual:
            self.a.fwd(c)
            self.b.fwd(c)
            return
        img = _PACK.get_dual(self.a.conv.weight, self.b.conv.weight, c.mode)
        ip, ild = c.vptr(self.src)
        op, old = c.vptr(self.out)
        ba, bb = self.a.conv.bias, self.b.conv.bias
        bias = 0
        if ba is not None or bb is not None:
            z = torch.zeros(64, dtype=torch.float32, device=c.device) if (ba is None or bb is None) else None
            bias_t = torch.cat([z if ba is None else ba.detach(), z if bb is None else bb.detach()])
            c.keep.append(bias_t)
            bias = bias_t.data_ptr()
        flags = _lib.TC2_F16 if c.act == 2 else 0
        kw = {}
        bna, bnb_ = self.a._bn_slice()[0], self.b._bn_slice()[0]
        if ((c.training or (bna.running_mean is None and bnb_.running_mean is None))
                and not (_lib.lib().tsr_get_tc_desc_mode() & 128)):
            t = c.stat_table(bnop, 0)
            c.stat_table(bnop, 1)
            kw.update(stat=t.data_ptr(), stat_ld=t.shape[2])
            flags |= _lib.TC2_STAT_PRECLEARED
        _lib.conv_tc2([(ip, ild, self.src.C, 5, img.data_ptr())], op, old, c.B, c.H, c.W, 128, flags=flags, bias=bias,
                      dual_fwd=1, stream=_lib.stream_ptr(), **kw)

    def bwd(self, c):
        self.a.bwd_weights(c)
        self.b.bwd_weights(c)
        _dgrad(c, [self.a, self.b], self.sink)


class BNReLUOp(Op):
    """BatchNorm2d (train: batch statistics + running-stat update; eval: running stats) [+ ReLU]
    (reference tactileSR_model.py:38-39, 42-43, 48-49, 169-170, 175-176, 181-182, 187-188).  ``bn`` may be a list of
    BatchNorm2d modules that normalise consecutive channel slices of ``src`` (the two branches of an MSRB stage whose
    outputs are concatenated, :200 / :203): statistics are finished per module, the elementwise passes run once over all
    channels."""

    def __init__(self, src: View, bn, out: View, relu: bool = True):
        self.bns = list(bn) if isinstance(bn, (list, tuple)) else [bn]
        self.bn = self.bns[0]
        self.src, self.out, self.relu = src, out, relu
        self.parts = []
        off = 0
        for m in self.bns:
            if m.momentum is None or not m.affine:
                raise _lib.TsrError("tactilesr_b200: BatchNorm2d with momentum=None (cumulative average) or affine=False is "
                                    "not supported by the fused kernels")
            self.parts.append((m, off, m.num_features))
            off += m.num_features
        assert src.C == off == out.C

    def params(self):
        return tuple(p for m in self.bns for p in (m.weight, m.bias))

    def reads(self):
        return (self.src.buf,)

    def writes(self):
        return (self.out.buf,)

    def out_views(self):
        return (self.out,)

    def grad_targets(self):
        return [(self.src, "other")]

    def fwd(self, c):
        folded = c.folded.get(self, ())
        if len(folded) == len(self.parts):
            return
        assert not folded, "partially folded BatchNorm group"
        st = _lib.stream_ptr()
        Ct = self.src.C
        coef = torch.empty((4, Ct), dtype=torch.float32, device=c.device)
        yp, yld = c.vptr(self.src)
        table = c.bn_partials.pop(self, None)
        covered = c.bn_partials_parts.pop(self, set())
        used = []
        for i, (bn, off, C) in enumerate(self.parts):
            sc, sh, mu, iv = (coef[k].data_ptr() + off * 4 for k in range(4))
            use_batch = c.training or bn.running_mean is None
            used.append(use_batch)
            track = c.training and bn.track_running_stats and bn.running_mean is not None
            if use_batch and track:
                global _STATS_EPOCH
                _STATS_EPOCH += 1
            if use_batch and table is not None and i in covered:
                _lib.call("tsr_bn_finalize_partials", table.data_ptr() + off * 4, table.shape[2], table.shape[0], c.npix, C,
                          bn.weight.data_ptr(), bn.bias.data_ptr(), _ptr(bn.running_mean) if track else 0,
                          _ptr(bn.running_var) if track else 0, _ptr(bn.num_batches_tracked) if track else 0, bn.momentum,
                          bn.eps, sc, sh, mu, iv, st)
            elif use_batch:
                ws, wsb = c.workspace(_lib.lib().tsr_bn_workspace(c.npix, C))
                _lib.call("tsr_bn_train_stats", yp + off * c.esize, yld, c.act, c.npix, C, bn.weight.data_ptr(),
                          bn.bias.data_ptr(), _ptr(bn.running_mean) if track else 0, _ptr(bn.running_var) if track else 0,
                          _ptr(bn.num_batches_tracked) if track else 0, bn.momentum, bn.eps, sc, sh, mu, iv, ws, wsb, st)
            else:
                _lib.call("tsr_bn_eval_coeffs", C, bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                          bn.running_var.data_ptr(), bn.eps, sc, sh, mu, iv, st)
        op, old = c.vptr(self.out)
        o2, o2ld = c.sptr(self.out) if c.wants_shadow(self.out.buf) else (0, 0)
        _lib.call("tsr_bn_apply", yp, yld, c.act, coef[0].data_ptr(), coef[1].data_ptr(), op, old, c.act, c.npix, Ct,
                  1 if self.relu else 0, o2, o2ld, st)
        c.saved[self] = (coef, used)

    def bwd(self, c):
        st = _lib.stream_ptr()
        Ct = self.src.C
        coef, used = c.saved[self]
        gp, gld = c.vptr(self.out, grad=True)
        yp, yld = c.vptr(self.src)
        first = _first_write(c, self.src.buf)
        assert first, "BN input has a single consumer"
        dp, dld = c.vptr(self.src, grad=True)
        table = c.bnb.pop(self, None)
        if table is not None:
            # level 1 (masked g, sum g, sum g*y) came out of the epilogue of the data gradient that produced `gp`
            c12 = torch.empty((2, Ct), dtype=torch.float32, device=c.device)
            for i, (bn, off, C) in enumerate(self.parts):
                gw, accw = c.pgrad(bn.weight)
                gb, accb = c.pgrad(bn.bias)
                assert accw == accb
                _lib.call("tsr_bn_bwd_finalize_partials", table.data_ptr() + off * 4, table.shape[2], table.shape[0], c.npix, C,
                          coef[2].data_ptr() + off * 4, coef[3].data_ptr() + off * 4, gw.data_ptr(), gb.data_ptr(), accw,
                          c12[0].data_ptr() + off * 4, c12[1].data_ptr() + off * 4, 1 if used[i] else 0, st)
            _lib.call("tsr_bn_backward_apply", gp, gld, yp, yld, dp, dld, c.mix, coef[0].data_ptr(), coef[1].data_ptr(),
                      coef[2].data_ptr(), coef[3].data_ptr(), c12[0].data_ptr(), c12[1].data_ptr(), c.npix, Ct, 0, st)
            return
        for i, (bn, off, C) in enumerate(self.parts):
            sc, sh, mu, iv = (coef[k].data_ptr() + off * 4 for k in range(4))
            gw, accw = c.pgrad(bn.weight)
            gb, accb = c.pgrad(bn.bias)
            assert accw == accb
            ws, wsb = c.workspace(_lib.lib().tsr_bn_backward_workspace(c.npix, C))
            gsz = 4 if c.grd == 0 else 2
            _lib.call("tsr_bn_backward", gp + off * gsz, gld, yp + off * c.esize, yld, dp + off * gsz, dld, c.mix, sc, sh, mu, iv,
                      gw.data_ptr(), gb.data_ptr(), accw, c.npix, C, 1 if self.relu else 0, 1 if used[i] else 0, ws, wsb, st)


class TailOp(Op):
    """Conv2d(Cin -> 1, 3x3, no bias) + ReLU writing the (B,1,H,W) fp32 result
    (reference tactileSR_model.py:55-56, 125-126)."""

    def __init__(self, src: View, conv: torch.nn.Conv2d, out: Buf, relu: bool = True):
        self.src, self.conv, self.out, self.relu = src, conv, out, relu
        self.mask_input = False            # Program.plan(): the only writer of the gradient of a ReLU-conv output
        assert conv.out_channels == 1 and conv.kernel_size == (3, 3) and conv.bias is None

    def params(self):
        return (self.conv.weight,)

    def reads(self):
        return (self.src.buf,)

    def writes(self):
        return (self.out,)

    def grad_targets(self):
        return [(self.src, "other")]

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
        if self.mask_input and c.tc:
            # the tail's input is the ReLU output of one convolution (Program.plan()): its ReLU backward is applied here
            _lib.call("tsr_tail_dgrad_masked", dout.data_ptr(), out.data_ptr(), self.conv.weight.data_ptr(), dp, dld, c.grd, c.B,
                      c.H, c.W, self.src.C, 1 if self.relu else 0, ip, ild, c.act, st)
            c.grad_masked.add(self.src.buf)
            return
        _lib.call("tsr_tail_dgrad", dout.data_ptr(), out.data_ptr(), self.conv.weight.data_ptr(), dp, dld, c.grd, c.B,
                  c.H, c.W, self.src.C, 1 if self.relu else 0, st)


class InputOp(Op):
    """NCHW fp32 module input -> NHWC activation buffer (standalone MSRB / ResBlock use)."""

    def __init__(self, out: Buf):
        self.out = out

    def writes(self):
        return (self.out,)

    def out_views(self):
        return (View.of(self.out),)

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
    _planned: bool = False

    def add(self, op: Op) -> Op:
        self.ops.append(op)
        return op

    def plan(self) -> None:
        """Static analysis of the gradient flow, once per program: for every buffer, the LAST op (in backward order) that
        writes its gradient.  If that is a tensor-core data gradient over the whole buffer it gets the buffer's Sink (ReLU
        mask or BatchNorm-backward level 1 fused into its epilogue), and a residual hand-over whose only successor is such
        a data gradient is done by reference instead of a copy."""
        if self._planned:
            return
        self._planned = True
        producers: Dict[Buf, list] = {}
        writers: Dict[Buf, list] = {}
        for i, op in enumerate(self.ops):
            for v in op.out_views():
                producers.setdefault(v.buf, []).append((op, v))
            for v, kind in op.grad_targets():
                writers.setdefault(v.buf, []).append((i, op, v, kind))
        for buf, ws in writers.items():
            ws.sort(key=lambda t: t[0])                    # backward runs from the highest index down: ws[0] is the last writer
            i, op, v, kind = ws[0]
            whole = v.c0 == 0 and v.C == buf.C
            if kind == "conv" and whole:
                prods = producers.get(buf, [])
                if len(prods) == 1 and isinstance(prods[0][0], BNReLUOp) and prods[0][1].c0 == 0 and prods[0][1].C == buf.C:
                    # (not on a 1x1 data gradient: it has no main loop to hide the longer epilogue behind -- measured 1.10 ms
                    # fused vs 0.45 + 0.16 ms as two passes for the 64 -> 256 confusion gradient at B = 1024)
                    if getattr(op, "K", 3) != 1:
                        op.sink = Sink("bn", prods[0][0])
                elif (prods and all(isinstance(p, (ConvOp, HeadOp)) and p.relu for p, _ in prods)
                      and sum(pv.C for _, pv in prods) == buf.C):
                    op.sink = Sink("relu")
            if len(ws) == 2 and ws[1][3] == "res" and kind == "conv" and ws[1][2].c0 == v.c0 and ws[1][2].C == v.C:
                ws[1][1].alias_residual_grad = True
            if isinstance(op, TailOp) and len(ws) == 1 and whole:
                prods = producers.get(buf, [])
                if (prods and all(isinstance(pr, (ConvOp, HeadOp)) and pr.relu for pr, _ in prods)
                        and sum(pv.C for _, pv in prods) == buf.C):
                    op.mask_input = True

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
        if x.dim() != 4 or tuple(x.shape[-2:]) != (4, 4):
            raise _lib.TsrError(f"tactilesr_b200: the taxel head kernels take (B, 3*seqsCnt, 4, 4) inputs, got {tuple(x.shape)}")
        H = W = x.shape[-1] * prog.sf
    else:
        H, W = x.shape[-2], x.shape[-1]
    c = RunCtx(mode, B, H, W, x.device, training, need_grad)
    c.x = x
    if c.act == 2:
        _arm_overflow_guard(x.device)
    prog.plan()
    c.conv_inputs = {op.src.buf for op in prog.ops if isinstance(op, (ConvOp, DualConvOp))}
    if c.tc:
        c.zero_cap = _stat_rows() * 2 * (sum(op.src.C for op in prog.ops if isinstance(op, BNReLUOp)) + 16)
    c.keep_taps = keep_taps
    if c.tc and (training or need_grad):
        seen, items = set(), []
        for op in prog.ops:
            if isinstance(op, DualConvOp):
                wa, wb = op.a.conv.weight, op.b.conv.weight
                if id(wa) in seen or id(wb) in seen or not (wa.is_contiguous() and wb.is_contiguous()):
                    continue
                seen.update((id(wa), id(wb)))
                items += [(wa, 1), (wb, 2)] if op.dual_ok else [(wa, 0), (wb, 0)]
            elif isinstance(op, ConvOp) and id(op.conv.weight) not in seen and op.conv.weight.is_contiguous():
                seen.add(id(op.conv.weight))
                items.append((op.conv.weight, 0))
        if items:
            _PACK.prepack(items, mode, need_grad)
    keep = need_grad or keep_taps
    last_use: Dict[Buf, int] = {}
    if not keep:
        for i, op in enumerate(prog.ops):
            for b in op.reads():
                last_use[b] = i
    for i, op in enumerate(prog.ops):
        with _prof("fwd:" + type(op).__name__):
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
    c.new_sweep()
    c.defer_zero = hooks is None       # (a per-op hook may ship gradient buckets before the sweep ends: zero those at once)
    # reverse sweep; gradient buffers are dropped as soon as their producer has consumed them
    for i in range(len(prog.ops) - 1, -1, -1):
        op = prog.ops[i]
        needed = [b for b in op.writes()]
        if any(b not in c.grads for b in needed):
            continue   # dead branch (no gradient reaches this op)
        if isinstance(op, (ConvOp, DualConvOp)):
            op.bwd(c)                      # (timed inside: bwd:wgrad / bwd:dgrad; the rest is ReLU / residual hand-over)
        else:
            with _prof("bwd:" + type(op).__name__):
                op.bwd(c)
        if hooks is not None:
            hooks(i, op, c)
        for b in needed:
            if not (prog.wants_input_grad and b is prog.in_buf) and not _written_earlier(prog, i, b):
                c.grads.pop(b, None)
    if c.zero_later:
        torch._foreach_zero_(c.zero_later)
        c.zero_later = []
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
    if need_grad:
        # gradients are written straight into ``.grad`` (see _ProgramFn.backward): tensor hooks on parameters would never fire
        for p in params:
            if p.requires_grad and (p._backward_hooks or getattr(p, "_post_accumulate_grad_hooks", None)):
                raise _lib.TsrError("tactilesr_b200: gradient hooks on parameters (Tensor.register_hook / "
                                    "register_post_accumulate_grad_hook, e.g. DDP, GradScaler unscale hooks) are not supported: "
                                    "the kernels write parameter gradients straight into .grad; use Trainer's gradient "
                                    "all-reduce (cpu/distributed.py GradAllReduce) instead")
    holder = {"prog": prog, "training": training, "need_grad": need_grad, "mode": mode}
    if extra:
        holder.update(extra)
    return _ProgramFn.apply(holder, x, *params)
