"""Kernel-level parity of the small-channel head / tail kernels (csrc/conv_f32.cu) through the C ABI against stock PyTorch in
fp64 on the same (rounded) operands: the register-blocked tail forward (4 x 5 output blocks per warp, incl. a ragged last
block and a short last strip), the tail data gradient with the fused ReLU mask of its input, the head forward and the head
weight gradient (gradient rows staged by asynchronous copies, several chunks and samples per CTA).
Reference lines: model/tactileSR_model.py:35-37, 55-56, 60-62, 107, 122-126."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_l2

pytestmark = pytest.mark.gpu
DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("Cin,H,W,B", [(128, 40, 40, 3), (64, 40, 40, 2), (128, 10, 24, 5), (128, 7, 8, 1)])
def test_tail_forward_matches_conv2d(dtype, Cin, H, W, B):
    from tactilesr_b200 import _lib
    torch.manual_seed(Cin + H + W)
    dev = "cuda"
    xb = torch.randn(B, H, W, Cin + 64, device=dev).to(dtype)          # the input is a channel slice of a wider buffer
    x = xb[..., 64:]
    w = torch.randn(1, Cin, 3, 3, device=dev) * 0.1
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), padding=1)
    for relu in (0, 1):
        out = torch.full((B, 1, H, W), 7.0, device=dev)
        _lib.call("tsr_tail_fwd", x.data_ptr(), Cin + 64, DT[dtype], w.data_ptr(), out.data_ptr(), B, H, W, Cin, relu, _st())
        assert rel_l2(out, torch.relu(ref) if relu else ref) < 2e-6


@pytest.mark.parametrize("adt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("Cin,B", [(128, 3), (64, 2)])
def test_tail_data_gradient_with_fused_input_relu_mask(adt, Cin, B):
    """din = conv_transpose(dout * (out > 0), w) masked by (in_act > 0), against autograd of relu(conv(relu(z)))."""
    from tactilesr_b200 import _lib
    torch.manual_seed(Cin)
    dev, H, W = "cuda", 40, 40
    act = torch.relu(torch.randn(B, H, W, Cin, device=dev)).to(adt)        # saved post-ReLU activation (zeros included)
    act[0, 0, 0, :8] = 0
    w = torch.randn(1, Cin, 3, 3, device=dev) * 0.1
    a64 = act.double().permute(0, 3, 1, 2).requires_grad_(True)
    out64 = torch.relu(F.conv2d(a64, w.double(), padding=1))
    dout = torch.randn(B, 1, H, W, device=dev)
    (ga,) = torch.autograd.grad(out64, a64, dout.double())
    ref = (ga * (a64 > 0)).permute(0, 2, 3, 1)
    out = out64.float().contiguous()
    for gdt in (torch.bfloat16, torch.float32):
        din = torch.empty(B, H, W, Cin, device=dev, dtype=gdt)
        _lib.call("tsr_tail_dgrad_masked", dout.data_ptr(), out.data_ptr(), w.data_ptr(), din.data_ptr(), Cin, DT[gdt], B, H, W, Cin, 1,
                  act.data_ptr(), Cin, DT[adt], _st())
        assert rel_l2(din, ref) < (4e-3 if gdt == torch.bfloat16 else 2e-6)
        assert bool((din.float()[act == 0] == 0).all())              # masked exactly, not approximately


@pytest.mark.parametrize("gdt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("sf,B", [(10, 5), (10, 300), (3, 2)])
def test_head_forward_and_weight_gradient(gdt, sf, B):
    from tactilesr_b200 import _lib
    torch.manual_seed(sf + B)
    dev, H = "cuda", 4 * sf
    x = torch.rand(B, 3, 4, 4, device=dev) * 8
    w = torch.randn(64, 3, 3, 3, device=dev) * 0.2
    w64 = w.double().requires_grad_(True)
    up = F.interpolate(x.double(), scale_factor=sf, mode="bilinear", align_corners=False)
    y64 = F.conv2d(up, w64, padding=1)
    # forward (fp32 output, with and without ReLU) into a channel slice of a 128-wide buffer
    if gdt == torch.float32:
        for relu in (0, 1):
            ob = torch.zeros(B, H, H, 128, device=dev)
            _lib.call("tsr_head_fwd", x.data_ptr(), 48, w.data_ptr(), ob[..., 64:].data_ptr(), 128, 0, B, sf, relu, _st())
            ref = (torch.relu(y64) if relu else y64).permute(0, 2, 3, 1)
            assert rel_l2(ob[..., 64:], ref.detach()) < 2e-6
            assert float(ob[..., :64].abs().max()) == 0.0
    # weight gradient from a gradient stored in `gdt`, inside a 128-wide buffer
    gb = (torch.randn(B, H, H, 128, device=dev) * 0.5).to(gdt)
    g = gb[..., 64:]
    (ref_dw,) = torch.autograd.grad(y64, w64, g.double().permute(0, 3, 1, 2))
    ws = torch.empty(max(int(_lib.lib().tsr_head_wgrad_workspace(B)), 256), dtype=torch.uint8, device=dev)
    dw0 = torch.randn_like(w)
    for acc in (0, 1):
        dw = dw0.clone()
        _lib.call("tsr_head_wgrad", x.data_ptr(), 48, g.data_ptr(), 128, DT[gdt], dw.data_ptr(), ws.data_ptr(), ws.numel(), B, sf, acc,
                  _st())
        assert rel_l2(dw, ref_dw + (dw0.double() if acc else 0)) < 2e-5
    dw2 = dw0.clone()
    _lib.call("tsr_head_wgrad", x.data_ptr(), 48, g.data_ptr(), 128, DT[gdt], dw2.data_ptr(), ws.data_ptr(), ws.numel(), B, sf, 1, _st())
    assert torch.equal(dw, dw2)
