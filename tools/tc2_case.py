"""One convolution shape launched a few times (for ncu).  usage: python tools/tc2_case.py <gen:1|2> <case> [B]
cases: dgrad1x1 | fwd3x3_128 | fwd3x3_64 | dual64 | dualdgrad128"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tactilesr_b200 import _lib

gen, case = int(sys.argv[1]), sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
H = W = 40
L = _lib.lib()
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
f16, bf = torch.float16, torch.bfloat16


def pack(w, half, dgrad=False):
    Cout, Cin, KS, _ = w.shape
    o = torch.empty(KS * KS * Cin * Cout, dtype=f16 if half else bf, device=dev)
    _lib.call("tsr_pack_conv_weight_f16" if half else "tsr_pack_conv_weight_bf16", w.data_ptr(), 0 if dgrad else o.data_ptr(),
              o.data_ptr() if dgrad else 0, Cout, Cin, KS, st)
    return o


rows = L.tsr_conv2d_tc2_stat_rows()
if case == "dgrad1x1":
    w = torch.randn(64, 256, 1, 1, device=dev) * 0.05
    wd = pack(w, False, True)
    dy = torch.randn(B, H, W, 64, device=dev).to(bf)
    dx = torch.empty(B, H, W, 256, dtype=bf, device=dev)
    if gen == 2:
        fn = lambda: _lib.conv_tc2([(dy.data_ptr(), 64, 64, 1, wd.data_ptr())], dx.data_ptr(), 256, B, H, W, 256)
    else:
        fn = lambda: _lib.call("tsr_conv2d_tc", dy.data_ptr(), 64, wd.data_ptr(), 0, 0, 0, dx.data_ptr(), 256, B, H, W, 64, 256, 1, 0, 0, 0,
                               0, 0, 0, st)
elif case in ("fwd3x3_128", "fwd3x3_64"):
    C = 128 if case.endswith("128") else 64
    w = torch.randn(C, C, 3, 3, device=dev) * 0.05
    wf = pack(w, True)
    x = torch.randn(B, H, W, C, device=dev).to(f16)
    out = torch.empty_like(x)
    bias = torch.randn(C, device=dev)
    part = torch.empty(rows, 2, C, device=dev)
    if gen == 2:
        fn = lambda: _lib.conv_tc2([(x.data_ptr(), C, C, 3, wf.data_ptr())], out.data_ptr(), C, B, H, W, C, flags=_lib.TC2_F16,
                                   bias=bias.data_ptr(), stat=part.data_ptr(), stat_ld=C)
    else:
        fn = lambda: _lib.call("tsr_conv2d_tc", x.data_ptr(), C, wf.data_ptr(), bias.data_ptr(), 0, 0, out.data_ptr(), C, B, H, W, C, C, 3, 2,
                               0, 0, part.data_ptr(), 0, 0, st)
elif case == "fwd5x5_128":        # the roofline kernel of bench.py: MSRB conv_5_2 forward shape, no epilogue extras
    w = torch.randn(128, 128, 5, 5, device=dev) * 0.02
    wf = pack(w, True)
    x = torch.randn(B, H, W, 128, device=dev).to(f16)
    out = torch.empty_like(x)
    fn = lambda: _lib.conv_tc2([(x.data_ptr(), 128, 128, 5, wf.data_ptr())], out.data_ptr(), 128, B, H, W, 128, flags=_lib.TC2_F16)
elif case.startswith("wgrad"):          # wgrad5x5_128 | wgrad3x3_128 | wgrad5x5_64 | wgrad3x3_64
    KS = int(case[5])
    C = int(case.split("_")[1])
    x = torch.randn(B, H, W, C, device=dev).to(bf)
    dy = torch.randn(B, H, W, C, device=dev).to(bf)
    need = L.tsr_conv2d_wgrad_tc_workspace(B, H, W, C, C, KS)
    ws = torch.empty(max(int(need), 256), dtype=torch.uint8, device=dev)
    dw = torch.zeros(C, C, KS, KS, device=dev)
    fn = lambda: _lib.call("tsr_conv2d_wgrad_tc", x.data_ptr(), C, dy.data_ptr(), C, dw.data_ptr(), ws.data_ptr(), ws.numel(), B, H, W,
                           C, C, KS, 0, st)
else:
    raise SystemExit("unknown case")
for _ in range(3):
    fn()
torch.cuda.synchronize()
print("done")
