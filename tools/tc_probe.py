"""GPU probe for the tcgen05 conv kernel: compares tsr_conv2d_tc with an fp32 convolution of the same
bf16-rounded operands, per descriptor mode, each case in its own subprocess (a trap must not kill the sweep).
usage: python tools/tc_probe.py            (driver)     |     python tools/tc_probe.py case KS Cin Cout B mode
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def case(KS, Cin, Cout, B, mode, H=40, W=40, flags=0):
    import torch
    import torch.nn.functional as F
    from tactilesr_b200 import _lib
    L = _lib.lib()
    L.tsr_set_tc_desc_mode(mode)
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, KS, KS, device=dev) / (Cin * KS * KS) ** 0.5)
    wb = w.to(torch.bfloat16).float()
    bias = torch.randn(Cout, device=dev)
    res = torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16)
    wf = torch.empty(KS * KS * Cin * Cout, dtype=torch.bfloat16, device=dev)
    wd = torch.empty_like(wf)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("tsr_pack_conv_weight_bf16", w.data_ptr(), wf.data_ptr(), wd.data_ptr(), Cout, Cin, KS, st)
    out = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    _lib.call("tsr_conv2d_tc", x.data_ptr(), Cin, wf.data_ptr(), bias.data_ptr(), res.data_ptr(), Cout, out.data_ptr(), Cout,
              B, H, W, Cin, Cout, KS, 1, 0, 0, 0, 0, 0, st)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wb, bias, padding=KS // 2).permute(0, 2, 3, 1) + res.float()
    ref = torch.relu(ref)
    err = (out.float() - ref).norm() / ref.norm()
    # dgrad: dx = conv(dy, flipped/transposed weights)
    dy = torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16)
    dx = torch.zeros(B, H, W, Cin, dtype=torch.bfloat16, device=dev)
    err2 = -1.0
    if Cin in (64, 128):
        _lib.call("tsr_conv2d_tc", dy.data_ptr(), Cout, wd.data_ptr(), 0, 0, 0, dx.data_ptr(), Cin, B, H, W, Cout, Cin, KS, 0, 0, 0, 0, 0, 0, st)
        torch.cuda.synchronize()
        refd = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wb, padding=KS // 2).permute(0, 2, 3, 1)
        err2 = ((dx.float() - refd).norm() / refd.norm()).item()
    # wgrad: dW = sum_pix x (x) dy
    err3 = -1.0
    if Cout in (64, 128):
        L.tsr_conv2d_wgrad_tc_workspace.restype = __import__("ctypes").c_size_t
        need = L.tsr_conv2d_wgrad_tc_workspace(B, H, W, Cin, Cout, KS)
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        dw = torch.zeros(Cout, Cin, KS, KS, device=dev)
        _lib.call("tsr_conv2d_wgrad_tc", x.data_ptr(), Cin, dy.data_ptr(), Cout, dw.data_ptr(), ws.data_ptr(), ws.numel(), B, H, W, Cin, Cout, KS, 0, st)
        torch.cuda.synchronize()
        refw = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, KS, KS), dy.float().permute(0, 3, 1, 2), padding=KS // 2)
        err3 = ((dw - refw).norm() / refw.norm()).item()
    print(f"RESULT KS={KS} Cin={Cin} Cout={Cout} B={B} mode={mode} fwd_rel_err={err.item():.3e} dgrad_rel_err={err2:.3e} wgrad_rel_err={err3:.3e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "case":
        case(*[int(a) for a in sys.argv[2:7]])
        sys.exit(0)
    cases = [(1, 64, 64, 2, 0), (1, 256, 64, 2, 0), (3, 64, 64, 2, 0), (5, 128, 128, 3, 0),
             (3, 448, 64, 2, 0), (3, 128, 128, 5, 0), (5, 64, 64, 1, 0), (3, 64, 128, 7, 0)]
    for c in cases:
        try:
            r = subprocess.run([sys.executable, __file__, "case", *map(str, c)], capture_output=True, text=True, timeout=120)
            tail = [l for l in (r.stdout + r.stderr).splitlines() if l.strip()][-3:]
            print(c, "rc", r.returncode, "|", " / ".join(tail), flush=True)
        except subprocess.TimeoutExpired:
            print(c, "TIMEOUT", flush=True)
