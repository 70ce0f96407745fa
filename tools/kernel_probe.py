"""Launch one hot kernel a few times (for ncu --set full):  python tools/kernel_probe.py {fwd|wgrad} Cin Cout KS B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tactilesr_b200 import _lib

which, Cin, Cout, KS, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
x = torch.randn(B * 1600, Cin, device="cuda").to(torch.bfloat16)
dy = torch.randn(B * 1600, Cout, device="cuda").to(torch.bfloat16)
w = torch.randn(Cout, Cin, KS, KS, device="cuda") * 0.02
wf = torch.empty(KS * KS * Cin * Cout, dtype=torch.bfloat16, device="cuda")
out = torch.empty(B * 1600, Cout, dtype=torch.bfloat16, device="cuda")
dw = torch.empty_like(w)
_lib.call("tsr_pack_conv_weight_bf16", w.data_ptr(), wf.data_ptr(), 0, Cout, Cin, KS, st)
ws = torch.empty(max(int(L.tsr_conv2d_wgrad_tc_workspace(B, 40, 40, Cin, Cout, KS)), 256), dtype=torch.uint8, device="cuda")
if which == "fwd":
    f = lambda: _lib.call("tsr_conv2d_tc", x.data_ptr(), Cin, wf.data_ptr(), 0, 0, 0, out.data_ptr(), Cout, B, 40, 40, Cin, Cout, KS, 0, 0, 0, 0, 0, 0, st)
else:
    f = lambda: _lib.call("tsr_conv2d_wgrad_tc", x.data_ptr(), Cin, dy.data_ptr(), Cout, dw.data_ptr(), ws.data_ptr(), ws.numel(), B, 40, 40, Cin, Cout, KS, 0, st)
for _ in range(3):
    f()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5):
    f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{which} {Cin}->{Cout} k{KS} B={B}: {ms:.3f} ms {2.0*B*1600*Cin*Cout*KS*KS/ms/1e9:.1f} TFLOP/s")
