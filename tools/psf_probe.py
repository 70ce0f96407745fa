"""Launch the fused PSF kernels (for ncu / timing): python tools/psf_probe.py B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tactilesr_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
st = torch.cuda.current_stream().cuda_stream
ab = torch.rand(B, 3, device="cuda") * 0.5 + 0.5
yy, xx = torch.meshgrid(torch.arange(100.0, device="cuda"), torch.arange(100.0, device="cuda"), indexing="ij")
depth = torch.clamp((20 - ((yy - 50) ** 2 + (xx - 45) ** 2).sqrt()) / 2 + 0.5, 0, 1).expand(B, 100, 100).contiguous()
HR = torch.empty(B, 100, 100, device="cuda"); LRd = torch.empty(B, 16, device="cuda"); psf = torch.empty(B, 99, 99, device="cuda")
dL = torch.rand(B, 16, device="cuda"); dab = torch.empty(B, 3, device="cuda")
aux = torch.empty(B, int(_lib.lib().tsr_psf_aux_floats()), device="cuda")
f_tc = lambda: _lib.call("tsr_psf_forward_tc", ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(), 0, B, st)
f_tca = lambda: _lib.call("tsr_psf_forward_tc", ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(), aux.data_ptr(), B, st)
b_tc = lambda: _lib.call("tsr_psf_backward_tc", ab.data_ptr(), depth.data_ptr(), aux.data_ptr(), dL.data_ptr(), dab.data_ptr(), B, st)
f_16 = lambda: _lib.call("tsr_psf_forward_tc_f16", ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(), 0, B, st)
f_16a = lambda: _lib.call("tsr_psf_forward_tc_f16", ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(), aux.data_ptr(), B, st)
b_16 = lambda: _lib.call("tsr_psf_backward_tc_f16", ab.data_ptr(), depth.data_ptr(), aux.data_ptr(), dL.data_ptr(), dab.data_ptr(), B, st)
f_ff = lambda: _lib.call("tsr_psf_forward_ffma", ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(), B, st)
b = lambda: _lib.call("tsr_psf_backward", ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), dL.data_ptr(), 0, 0, dab.data_ptr(), B, st)
for name, fn in (("fwd tcgen05", f_tc), ("fwd tcgen05 + aux", f_tca), ("fwd tcgen05 f16", f_16), ("fwd tcgen05 f16 + aux", f_16a),
                 ("fwd ffma", f_ff), ("bwd tcgen05", b_tc), ("bwd tcgen05 f16", b_16), ("bwd ffma", b)):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"psf {name} B={B}: {ms:.3f} ms  {B/ms*1e3:.0f} samples/s  {B/ms*1e3*119472/1e9:.0f} GB/s compulsory(fwd)")
