"""Per-launch timing (CUDA events, L2 flushed between launches by rotating buffers) of the generation-2 tensor-core
convolution on the shapes of one MSRB at a given batch (B = argv[1]), with the wait-cycle breakdown of CTA 0.  usage: python tools/tc2_perf.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tactilesr_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = W = 40
L = _lib.lib()
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
NROT = 3


def pack(w, f16, dgrad=False):
    Cout, Cin, KS, _ = w.shape
    dt = torch.float16 if f16 else torch.bfloat16
    o = torch.empty(KS * KS * Cin * Cout, dtype=dt, device=dev)
    _lib.call("tsr_pack_conv_weight_f16" if f16 else "tsr_pack_conv_weight_bf16", w.data_ptr(), 0 if dgrad else o.data_ptr(),
              o.data_ptr() if dgrad else 0, Cout, Cin, KS, st)
    return o


def timeit(fn, n=6):
    for i in range(2):
        fn(i % NROT)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(i % NROT)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[len(ts) // 2] * 1e3


def act(C, dt):
    return [torch.randn(B, H, W, C, device=dev).to(dt) for _ in range(NROT)]


DBG = torch.zeros(8, dtype=torch.int64, device=dev)
LAST = [None]


def timeit_dbg(fn, n=6):
    """timeit for a generation-2 launch; afterwards one more launch with the wait-cycle counters switched on."""
    t = timeit(fn, n)
    LAST[0] = fn
    return t


def report(name, flops, us_new, us_old=None):
    s = f"{name:46s} tc2 {us_new:8.1f} us {flops / us_new / 1e6:7.1f} TF/s"
    if us_old is not None:
        s += f" | {us_old:8.1f} us"
    print(s, flush=True)
    if LAST[0] is not None:
        DBG.zero_()
        L.tsr_conv2d_tc2_debug(DBG.data_ptr())
        LAST[0](0)
        torch.cuda.synchronize()
        L.tsr_conv2d_tc2_debug(0)
        d = DBG.tolist()
        tot = max(d[5], 1)
        print(f"      CTA0 cycles: loop {tot}  blocks {d[6]}  MMA waits: accum-drained {100 * d[2] / tot:.0f}%  A tile {100 * d[3] / tot:.0f}%  "
              f"B tile {100 * d[4] / tot:.0f}% | A-producer idle {100 * d[0] / tot:.0f}%  B-producer idle {100 * d[1] / tot:.0f}%  "
              f"epilogue warp busy+wait {100 * d[7] / tot:.0f}%", flush=True)
        LAST[0] = None


rows = L.tsr_conv2d_tc2_stat_rows()
f16 = torch.float16
bf = torch.bfloat16
npix = B * H * W

# forward convs with BatchNorm statistics (fp16 operands)
for Cin, Cout, KS in [(64, 64, 3), (64, 64, 5), (128, 128, 3), (128, 128, 5)]:
    x = act(Cin, f16); out = act(Cout, f16)
    w = torch.randn(Cout, Cin, KS, KS, device=dev) * 0.05
    wf = pack(w, True)
    bias = torch.randn(Cout, device=dev)
    part = torch.empty(rows, 2, Cout, device=dev)
    new = timeit_dbg(lambda i: _lib.conv_tc2([(x[i].data_ptr(), Cin, Cin, KS, wf.data_ptr())], out[i].data_ptr(), Cout, B, H, W, Cout,
                                         flags=_lib.TC2_F16, bias=bias.data_ptr(), stat=part.data_ptr(), stat_ld=Cout))
    old = None
    report(f"fwd+stats {Cin}->{Cout} {KS}x{KS}", 2.0 * npix * Cin * Cout * KS * KS, new, old)

if len(sys.argv) > 2 and sys.argv[2] == "fwd":
    sys.exit(0)
# dual-branch forward 64 -> 64 + 64
x = act(64, f16); out = act(128, f16)
w3 = torch.randn(64, 64, 3, 3, device=dev) * 0.05; w5 = torch.randn(64, 64, 5, 5, device=dev) * 0.05
img = torch.empty(L.tsr_pack_conv_weight_dual_elems(64), dtype=f16, device=dev)
_lib.call("tsr_pack_conv_weight_dual", w3.data_ptr(), w5.data_ptr(), img.data_ptr(), 64, 2, st)
bias = torch.randn(128, device=dev)
part = torch.empty(rows, 2, 128, device=dev)
new = timeit_dbg(lambda i: _lib.conv_tc2([(x[i].data_ptr(), 64, 64, 5, img.data_ptr())], out[i].data_ptr(), 128, B, H, W, 128,
                                     flags=_lib.TC2_F16, bias=bias.data_ptr(), stat=part.data_ptr(), stat_ld=128, dual_fwd=1))
report("dual fwd+stats 64 -> 64|64 (3x3 | 5x5)", 2.0 * npix * 64 * 64 * 34, new)

# confusion 1x1 256 -> 64 + bias + residual + relu + bf16 copy
x = act(256, f16); out = act(64, f16); res = act(64, f16); o2 = act(64, bf)
w = torch.randn(64, 256, 1, 1, device=dev) * 0.05
wf = pack(w, True)
bias = torch.randn(64, device=dev)
new = timeit_dbg(lambda i: _lib.conv_tc2([(x[i].data_ptr(), 256, 256, 1, wf.data_ptr())], out[i].data_ptr(), 64, B, H, W, 64,
                                     flags=_lib.TC2_F16 | _lib.TC2_RELU, bias=bias.data_ptr(), residual=res[i].data_ptr(), res_ld=64,
                                     out2=o2[i].data_ptr(), out2_ld=64))
old = None
report("fwd 1x1 256->64 +res+relu+copy", 2.0 * npix * 256 * 64, new, old)
gb = npix * (256 + 64 + 64 + 64) * 2 / 1e9
print(f"    compulsory bytes {gb:.2f} GB -> tc2 {gb / new * 1e3:.0f} GB/s", flush=True)

# 1x1 dgrad 64 -> 256 (plain, and with the BN-backward epilogue)
dy = act(64, bf); dx = act(256, bf); y = act(256, f16)
wd = pack(w, False, dgrad=True)
coef = torch.randn(2, 256, device=dev)
part = torch.empty(rows, 2, 256, device=dev)
new = timeit_dbg(lambda i: _lib.conv_tc2([(dy[i].data_ptr(), 64, 64, 1, wd.data_ptr())], dx[i].data_ptr(), 256, B, H, W, 256))
old = None
report("dgrad 1x1 64->256", 2.0 * npix * 256 * 64, new, old)
new = timeit_dbg(lambda i: _lib.conv_tc2([(dy[i].data_ptr(), 64, 64, 1, wd.data_ptr())], dx[i].data_ptr(), 256, B, H, W, 256,
                                     flags=_lib.TC2_BNB | _lib.TC2_BNB_RELU | _lib.TC2_AUX_F16, aux=y[i].data_ptr(), aux_ld=256,
                                     aux_scale=coef[0].data_ptr(), aux_shift=coef[1].data_ptr(), stat=part.data_ptr(), stat_ld=256))
report("dgrad 1x1 64->256 + BN-bwd epilogue", 2.0 * npix * 256 * 64, new)

# dual data gradients
for C in (64, 128):
    dy3 = act(C, bf); dy5 = act(C, bf); dx = act(C, bf); y = act(C, f16); res = act(C, bf)
    w3 = torch.randn(C, C, 3, 3, device=dev) * 0.05; w5 = torch.randn(C, C, 5, 5, device=dev) * 0.05
    wd3, wd5 = pack(w3, False, True), pack(w5, False, True)
    coef = torch.randn(2, C, device=dev)
    part = torch.empty(rows, 2, C, device=dev)
    fl = 2.0 * npix * C * C * 34

    old = None
    srcs = lambda i: [(dy3[i].data_ptr(), C, C, 3, wd3.data_ptr()), (dy5[i].data_ptr(), C, C, 5, wd5.data_ptr())]
    new = timeit_dbg(lambda i: _lib.conv_tc2(srcs(i), dx[i].data_ptr(), C, B, H, W, C))
    report(f"dual dgrad {C} (3x3 + 5x5), plain", fl, new, old)
    new = timeit_dbg(lambda i: _lib.conv_tc2(srcs(i), dx[i].data_ptr(), C, B, H, W, C, flags=_lib.TC2_BNB | _lib.TC2_BNB_RELU | _lib.TC2_AUX_F16,
                                         aux=y[i].data_ptr(), aux_ld=C, aux_scale=coef[0].data_ptr(), aux_shift=coef[1].data_ptr(),
                                         stat=part.data_ptr(), stat_ld=C))
    report(f"dual dgrad {C} + BN-bwd epilogue", fl, new)
    new = timeit_dbg(lambda i: _lib.conv_tc2(srcs(i), dx[i].data_ptr(), C, B, H, W, C, flags=_lib.TC2_MASK | _lib.TC2_AUX_F16,
                                         aux=y[i].data_ptr(), aux_ld=C, residual=res[i].data_ptr(), res_ld=C))
    report(f"dual dgrad {C} + residual + ReLU mask", fl, new)

# weight gradients: bf16 x (no conversion) vs fp16 x (converted to bf16 in shared memory by the idle warps)
for Cin, Cout, KS in [(64, 64, 3), (64, 64, 5), (128, 128, 3), (128, 128, 5), (256, 64, 1)]:
    xb = act(Cin, bf); xh = act(Cin, f16); dy = act(Cout, bf)
    need = L.tsr_conv2d_wgrad_tc_workspace(B, H, W, Cin, Cout, KS)
    ws = torch.empty(max(int(need), 256), dtype=torch.uint8, device=dev)
    dw = torch.zeros(Cout, Cin, KS, KS, device=dev)
    tb_ = timeit(lambda i: _lib.call("tsr_conv2d_wgrad_tc_x", xb[i].data_ptr(), Cin, 1, dy[i].data_ptr(), Cout, dw.data_ptr(), ws.data_ptr(),
                                     ws.numel(), B, H, W, Cin, Cout, KS, 0, st))
    th_ = timeit(lambda i: _lib.call("tsr_conv2d_wgrad_tc_x", xh[i].data_ptr(), Cin, 2, dy[i].data_ptr(), Cout, dw.data_ptr(), ws.data_ptr(),
                                     ws.numel(), B, H, W, Cin, Cout, KS, 0, st))
    fl = 2.0 * npix * Cin * Cout * KS * KS
    print(f"wgrad {Cin}->{Cout} {KS}x{KS}: bf16 x {tb_:8.1f} us {fl / tb_ / 1e6:7.1f} TF/s | fp16 x (in-smem conversion) {th_:8.1f} us {fl / th_ / 1e6:7.1f} TF/s",
          flush=True)
