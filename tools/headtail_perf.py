"""Time the head / tail kernels (HBM- or FFMA-bound small-channel convs): python tools/headtail_perf.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tactilesr_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
dev = "cuda"
x = torch.rand(B, 3, 4, 4, device=dev) * 8
wh = torch.randn(64, 3, 3, 3, device=dev) * 0.1
act = torch.empty(B * 1600, 64, dtype=torch.float16, device=dev)
g64 = torch.randn(B * 1600, 64, device=dev).to(torch.bfloat16)
dwh = torch.empty_like(wh)
wsh = torch.empty(max(int(L.tsr_head_wgrad_workspace(B)), 256), dtype=torch.uint8, device=dev)
a128 = torch.randn(B * 1600, 128, device=dev).to(torch.float16)
wt = torch.randn(1, 128, 3, 3, device=dev) * 0.05
out = torch.empty(B, 1, 40, 40, device=dev)
dout = torch.randn(B, 1, 40, 40, device=dev)
din = torch.empty(B * 1600, 128, dtype=torch.bfloat16, device=dev)
dwt = torch.empty_like(wt)
wst = torch.empty(max(int(L.tsr_tail_wgrad_workspace(B, 40, 40, 128)), 256), dtype=torch.uint8, device=dev)
cases = {
    "head_fwd (out 64ch fp16)": (lambda: _lib.call("tsr_head_fwd", x.data_ptr(), 48, wh.data_ptr(), act.data_ptr(), 64, 2, B, 10, 0, st), B * 1600 * 64 * 2),
    "head_wgrad (dy 64ch bf16)": (lambda: _lib.call("tsr_head_wgrad", x.data_ptr(), 48, g64.data_ptr(), 64, 1, dwh.data_ptr(), wsh.data_ptr(), wsh.numel(), B, 10, 0, st), B * 1600 * 64 * 2),
    "tail_fwd (in 128ch fp16)": (lambda: _lib.call("tsr_tail_fwd", a128.data_ptr(), 128, 2, wt.data_ptr(), out.data_ptr(), B, 40, 40, 128, 1, st), B * 1600 * 128 * 2),
    "tail_dgrad (out 128ch bf16)": (lambda: _lib.call("tsr_tail_dgrad", dout.data_ptr(), out.data_ptr(), wt.data_ptr(), din.data_ptr(), 128, 1, B, 40, 40, 128, 1, st), B * 1600 * 128 * 2),
    "tail_wgrad (in 128ch fp16)": (lambda: _lib.call("tsr_tail_wgrad", a128.data_ptr(), 128, 2, dout.data_ptr(), out.data_ptr(), dwt.data_ptr(), wst.data_ptr(), wst.numel(), B, 40, 40, 128, 1, 0, st), B * 1600 * 128 * 2),
}
for name, (fn, nbytes) in cases.items():
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"B={B} {name}: {ms*1e3:.0f} us  {nbytes/ms/1e6:.0f} GB/s of the big tensor", flush=True)
