"""Time tsr_conv2d_tc on the network's conv shapes (CUDA events, warm, B from argv)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tactilesr_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
st = torch.cuda.current_stream().cuda_stream
for (Cin, Cout, KS) in [(128, 128, 5), (128, 128, 3), (64, 64, 5), (64, 64, 3), (256, 64, 1)]:
    x = torch.randn(B * 1600, Cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(Cout, Cin, KS, KS, device="cuda") * 0.02
    wf = torch.empty(KS * KS * Cin * Cout, dtype=torch.bfloat16, device="cuda")
    out = torch.empty(B * 1600, Cout, dtype=torch.bfloat16, device="cuda")
    _lib.call("tsr_pack_conv_weight_bf16", w.data_ptr(), wf.data_ptr(), 0, Cout, Cin, KS, st)
    f = lambda: _lib.call("tsr_conv2d_tc", x.data_ptr(), Cin, wf.data_ptr(), 0, 0, 0, out.data_ptr(), Cout, B, 40, 40, Cin, Cout, KS, 0, 0, 0, 0, 0, 0, st)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * B * 1600 * Cin * Cout * KS * KS
    print(f"B={B} {Cin}->{Cout} k{KS}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s  in+out {(B*1600*(Cin+Cout)*2)/ms/1e6:.0f} GB/s", flush=True)
