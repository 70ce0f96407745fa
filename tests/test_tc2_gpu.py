"""Kernel-level parity of the generation-2 tensor-core convolution (csrc/conv_tc2.cu) and of the tcgen05 weight gradient
at batch sizes that make the persistent kernels LOOP (B = 64 ... 1024: tens of blocks per CTA pair, both TMEM accumulator
buffers, ring wrap across blocks, statistics accumulated across blocks, multi-group outputs) -- against
torch.nn.functional.conv2d / conv_transpose2d / torch.nn.grad.conv2d_weight in fp32 (TF32 off) on the same rounded
operands.  Reference semantics: nn.Conv2d forward / autograd of model/tactileSR_model.py:67-84, 196-206.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    """On the device (the B = 1024 tensors have 2e8 elements)."""
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / b.norm()).item()

H = W = 40


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _st():
    return torch.cuda.current_stream().cuda_stream


def _pack(w, f16, fwd=True, dgrad=False):
    from tactilesr_b200 import _lib
    Cout, Cin, KS, _ = w.shape
    dt = torch.float16 if f16 else torch.bfloat16
    wf = torch.empty(KS * KS * Cin * Cout, dtype=dt, device=w.device) if fwd else None
    wd = torch.empty(KS * KS * Cin * Cout, dtype=dt, device=w.device) if dgrad else None
    _lib.call("tsr_pack_conv_weight_f16" if f16 else "tsr_pack_conv_weight_bf16", w.data_ptr(), wf.data_ptr() if fwd else 0,
              wd.data_ptr() if dgrad else 0, Cout, Cin, KS, _st())
    return wf, wd


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1)


def _rand_w(Cout, Cin, KS, dt):
    w = torch.randn(Cout, Cin, KS, KS, device="cuda") / (Cin * KS * KS) ** 0.5
    return w, w.to(dt).float()


@pytest.mark.parametrize("Cin,Cout,KS,B,f16", [
    (64, 64, 3, 1024, 1), (128, 128, 5, 1024, 0), (64, 64, 5, 256, 0), (128, 128, 3, 256, 1), (256, 64, 1, 256, 1),
    (448, 64, 3, 64, 0), (64, 256, 1, 256, 0), (64, 64, 3, 3, 1), (128, 128, 5, 1, 0), (64, 128, 3, 37, 1),
    (64, 448, 3, 9, 0), (64, 192, 1, 70, 1)])
def test_tc2_forward_bias_residual_relu(Cin, Cout, KS, B, f16):
    """Forward with the whole epilogue (+ bias, + residual, ReLU, bf16 copy) incl. ragged / tiny batches."""
    from tactilesr_b200 import _lib
    torch.manual_seed(Cin + Cout + KS + B)
    dt = torch.float16 if f16 else torch.bfloat16
    x = torch.randn(B, H, W, Cin, device="cuda").to(dt)
    w, wr = _rand_w(Cout, Cin, KS, dt)
    bias = torch.randn(Cout, device="cuda")
    res = torch.randn(B, H, W, Cout, device="cuda").to(dt)
    wf, _ = _pack(w, f16)
    out = torch.full((B, H, W, Cout), 7.0, dtype=dt, device="cuda")
    out2 = torch.full((B, H, W, Cout), 7.0, dtype=torch.bfloat16, device="cuda") if f16 else None
    _lib.conv_tc2([(x.data_ptr(), Cin, Cin, KS, wf.data_ptr())], out.data_ptr(), Cout, B, H, W, Cout,
                  flags=_lib.TC2_RELU | (_lib.TC2_F16 if f16 else 0), bias=bias.data_ptr(), residual=res.data_ptr(), res_ld=Cout,
                  out2=out2.data_ptr() if f16 else 0, out2_ld=Cout)
    ref = torch.relu(_nhwc(F.conv2d(_nchw(x), wr, bias, padding=KS // 2)) + res.float())
    e = rel_l2(out.float(), ref)
    assert e < (5e-4 if f16 else 4e-3), e          # rounding of the stored output only
    if f16:
        assert torch.equal(out2, ref.to(torch.bfloat16)) or rel_l2(out2.float(), ref) < 4e-3


@pytest.mark.parametrize("Cin,Cout,KS,B,f16", [(64, 64, 3, 1024, 1), (128, 128, 5, 256, 1), (128, 128, 3, 1024, 0),
                                               (64, 64, 5, 64, 0), (64, 384, 3, 64, 1), (64, 64, 3, 2, 0)])
def test_tc2_forward_bn_statistics(Cin, Cout, KS, B, f16):
    """Batch statistics out of the epilogue == statistics of the stored tensor; bit-deterministic across launches; the
    finalised scale / shift / mean / invstd match nn.BatchNorm2d's training-mode arithmetic (:169,175,181,187)."""
    from tactilesr_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(Cin + Cout + KS + B)
    dt = torch.float16 if f16 else torch.bfloat16
    x = torch.randn(B, H, W, Cin, device="cuda").to(dt)
    w, wr = _rand_w(Cout, Cin, KS, dt)
    bias = torch.randn(Cout, device="cuda")
    wf, _ = _pack(w, f16)
    rows = L.tsr_conv2d_tc2_stat_rows()
    runs = []
    for _ in range(2):
        out = torch.zeros(B, H, W, Cout, dtype=dt, device="cuda")
        part = torch.full((rows, 2, Cout), 7.0, device="cuda")      # the call must clear it
        _lib.conv_tc2([(x.data_ptr(), Cin, Cin, KS, wf.data_ptr())], out.data_ptr(), Cout, B, H, W, Cout,
                      flags=_lib.TC2_F16 if f16 else 0, bias=bias.data_ptr(), stat=part.data_ptr(), stat_ld=Cout)
        runs.append((out, part))
    out, part = runs[0]
    assert torch.equal(out, runs[1][0]) and torch.equal(part, runs[1][1])
    ref = _nhwc(F.conv2d(_nchw(x), wr, bias, padding=KS // 2))
    assert rel_l2(out.float(), ref) < (5e-4 if f16 else 4e-3)
    y = out.double().reshape(-1, Cout)
    s, ss = part[:, 0].double().sum(0), part[:, 1].double().sum(0)
    assert ((s - y.sum(0)).abs().max() / y.abs().sum(0).max()).item() < 1e-6
    assert ((ss - (y * y).sum(0)).abs().max() / (y * y).sum(0).max()).item() < 1e-6
    n = y.shape[0]
    gamma = torch.rand(Cout, device="cuda") + 0.5; beta = torch.randn(Cout, device="cuda")
    rm = torch.zeros(Cout, device="cuda"); rv = torch.ones(Cout, device="cuda")
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    coef = torch.empty(4, Cout, device="cuda")
    _lib.call("tsr_bn_finalize_partials", part.data_ptr(), Cout, rows, n, Cout, gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(),
              rv.data_ptr(), nbt.data_ptr(), 0.1, 1e-5, coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(),
              coef[3].data_ptr(), _st())
    mean, var = y.mean(0), y.var(0, unbiased=False)
    inv = 1 / (var + 1e-5).sqrt()
    assert (coef[2].double() - mean).abs().max() < 1e-5 * max(1.0, mean.abs().max().item())
    assert ((coef[3].double() - inv).abs() / inv).max() < 1e-5
    assert ((coef[0].double() - gamma.double() * inv).abs() / (gamma.double() * inv).abs()).max() < 1e-5
    assert int(nbt.item()) == 1 and (rm.double() - 0.1 * mean).abs().max() < 1e-5
    assert ((rv.double() - (0.9 + 0.1 * var * n / (n - 1))).abs()).max() < 1e-5


@pytest.mark.parametrize("C,KS,B", [(64, 3, 1024), (128, 3, 256), (128, 5, 256), (64, 5, 64), (64, 3, 5)])
def test_tc2_dgrad_with_accumulate(C, KS, B):
    """Data gradient (conv_transpose2d) that adds onto an existing gradient tensor through the residual input."""
    from tactilesr_b200 import _lib
    torch.manual_seed(C + KS + B)
    bf = torch.bfloat16
    dy = torch.randn(B, H, W, C, device="cuda").to(bf)
    w, wr = _rand_w(C, C, KS, bf)
    _, wd = _pack(w, 0, fwd=False, dgrad=True)
    dx = torch.randn(B, H, W, C, device="cuda").to(bf)
    prev = dx.clone()
    _lib.conv_tc2([(dy.data_ptr(), C, C, KS, wd.data_ptr())], dx.data_ptr(), C, B, H, W, C, residual=dx.data_ptr(), res_ld=C)
    ref = _nhwc(F.conv_transpose2d(_nchw(dy), wr, padding=KS // 2)) + prev.float()
    assert rel_l2(dx.float(), ref) < 4e-3


@pytest.mark.parametrize("C,B,epi", [(64, 1024, "mask"), (128, 256, "bnb"), (128, 1024, "bnb"), (64, 256, "bnb"), (128, 64, "mask"),
                                     (64, 3, "bnb"), (128, 2, "none")])
def test_tc2_dual_dgrad_kconcat(C, B, epi):
    """ONE launch for d x = conv3^T(dy3) + conv5^T(dy5) (+ residual) with the fused post-ops:
    "mask": ReLU backward by a saved fp16 activation; "bnb": BatchNorm(+ReLU) backward level 1 -- masked g stored, (sum g,
    sum g*y) partials -> dgamma, dbeta, c1, c2 -> dy through tsr_bn_backward_apply, all against torch autograd of
    relu(batch_norm(y)) in fp64."""
    from tactilesr_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(C + B)
    bf = torch.bfloat16
    dy3 = torch.randn(B, H, W, C, device="cuda").to(bf)
    dy5 = torch.randn(B, H, W, C, device="cuda").to(bf)
    w3, w3r = _rand_w(C, C, 3, bf)
    w5, w5r = _rand_w(C, C, 5, bf)
    _, wd3 = _pack(w3, 0, fwd=False, dgrad=True)
    _, wd5 = _pack(w5, 0, fwd=False, dgrad=True)
    res = torch.randn(B, H, W, C, device="cuda").to(bf)
    da_ref = _nhwc(F.conv_transpose2d(_nchw(dy3), w3r, padding=1) + F.conv_transpose2d(_nchw(dy5), w5r, padding=2)) + res.float()
    g = torch.full((B, H, W, C), 7.0, dtype=bf, device="cuda")
    srcs = [(dy3.data_ptr(), C, C, 3, wd3.data_ptr()), (dy5.data_ptr(), C, C, 5, wd5.data_ptr())]
    if epi == "none":
        _lib.conv_tc2(srcs, g.data_ptr(), C, B, H, W, C, residual=res.data_ptr(), res_ld=C)
        assert rel_l2(g.float(), da_ref) < 4e-3
        return
    if epi == "mask":
        a = torch.randn(B, H, W, C, device="cuda").clamp_min(0).to(torch.float16)
        _lib.conv_tc2(srcs, g.data_ptr(), C, B, H, W, C, flags=_lib.TC2_MASK | _lib.TC2_AUX_F16, residual=res.data_ptr(), res_ld=C,
                      aux=a.data_ptr(), aux_ld=C)
        ref = da_ref * (a > 0)
        assert rel_l2(g.float(), ref) < 4e-3
        assert (g[a <= 0] == 0).all()
        return
    # BatchNorm(+ReLU) backward through the fused epilogue
    y = (torch.randn(B, H, W, C, device="cuda") * 1.5 + 0.3).to(torch.float16)
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda") * 0.3
    yd = y.double().reshape(-1, C)
    mean, var = yd.mean(0), yd.var(0, unbiased=False)
    inv = 1 / (var + 1e-5).sqrt()
    coef = torch.stack([(gamma.double() * inv), beta.double() - mean * gamma.double() * inv, mean, inv]).float().contiguous()
    rows = L.tsr_conv2d_tc2_stat_rows()
    runs = []
    for _ in range(2):
        part = torch.full((rows, 2, C), 3.0, device="cuda")
        _lib.conv_tc2(srcs, g.data_ptr(), C, B, H, W, C, flags=_lib.TC2_BNB | _lib.TC2_BNB_RELU | _lib.TC2_AUX_F16,
                      residual=res.data_ptr(), res_ld=C, aux=y.data_ptr(), aux_ld=C, aux_scale=coef[0].data_ptr(),
                      aux_shift=coef[1].data_ptr(), stat=part.data_ptr(), stat_ld=C)
        runs.append((g.clone(), part))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1]), "must be bit-deterministic"
    part = runs[0][1]
    z = torch.addcmul(coef[1], y.float(), coef[0])          # fp32 fma as the kernels evaluate the mask
    mask = z > 0
    assert rel_l2(g.float(), da_ref * mask) < 4e-3
    gd = g.double().reshape(-1, C)
    assert ((part[:, 0].double().sum(0) - gd.sum(0)).abs().max() / gd.abs().sum(0).max()).item() < 1e-6
    assert ((part[:, 1].double().sum(0) - (gd * yd).sum(0)).abs().max() / (gd * yd).abs().sum(0).max()).item() < 1e-6
    n = B * H * W
    dgamma = torch.zeros(C, device="cuda"); dbeta = torch.zeros(C, device="cuda")
    c12 = torch.empty(2, C, device="cuda")
    _lib.call("tsr_bn_bwd_finalize_partials", part.data_ptr(), C, rows, n, C, coef[2].data_ptr(), coef[3].data_ptr(),
              dgamma.data_ptr(), dbeta.data_ptr(), 0, c12[0].data_ptr(), c12[1].data_ptr(), 1, _st())
    dyo = torch.empty(B, H, W, C, dtype=bf, device="cuda")
    _lib.call("tsr_bn_backward_apply", g.data_ptr(), C, y.data_ptr(), C, dyo.data_ptr(), C, 2, coef[0].data_ptr(),
              coef[1].data_ptr(), coef[2].data_ptr(), coef[3].data_ptr(), c12[0].data_ptr(), c12[1].data_ptr(), n, C, 0, _st())
    # autograd reference in fp64 on the stored g (so that only the BN backward arithmetic is compared)
    yr = y.double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    br = beta.double().requires_grad_(True)
    o = F.batch_norm(yr.permute(0, 3, 1, 2), None, None, gr, br, True, 0.0, 1e-5)
    o.backward(g.double().permute(0, 3, 1, 2))              # g is already masked: the ReLU's backward is done
    assert rel_l2(dgamma, gr.grad) < 1e-4 and rel_l2(dbeta, br.grad) < 1e-4
    assert rel_l2(dyo.float(), yr.grad) < 5e-3


@pytest.mark.parametrize("Cin,B,f16", [(64, 1024, 1), (64, 256, 0), (64, 2, 1), (128, 64, 1)])
def test_tc2_dual_branch_forward(Cin, B, f16):
    """conv3x3 and conv5x5 of the same input (MSRB.forward :198-199) from one halo tile: channels [0, 64) = conv3,
    [64, 128) = conv5, with both BatchNorm statistics tables."""
    from tactilesr_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(Cin + B)
    dt = torch.float16 if f16 else torch.bfloat16
    x = torch.randn(B, H, W, Cin, device="cuda").to(dt)
    w3, w3r = _rand_w(64, Cin, 3, dt)
    w5, w5r = _rand_w(64, Cin, 5, dt)
    bias = torch.randn(128, device="cuda")
    img = torch.empty(L.tsr_pack_conv_weight_dual_elems(Cin), dtype=dt, device="cuda")
    _lib.call("tsr_pack_conv_weight_dual", w3.data_ptr(), w5.data_ptr(), img.data_ptr(), Cin, 2 if f16 else 1, _st())
    rows = L.tsr_conv2d_tc2_stat_rows()
    out = torch.full((B, H, W, 128), 7.0, dtype=dt, device="cuda")
    part = torch.empty(rows, 2, 128, device="cuda")
    _lib.conv_tc2([(x.data_ptr(), Cin, Cin, 5, img.data_ptr())], out.data_ptr(), 128, B, H, W, 128, flags=_lib.TC2_F16 if f16 else 0,
                  bias=bias.data_ptr(), stat=part.data_ptr(), stat_ld=128, dual_fwd=1)
    ref = torch.cat([_nhwc(F.conv2d(_nchw(x), w3r, bias[:64], padding=1)), _nhwc(F.conv2d(_nchw(x), w5r, bias[64:], padding=2))], -1)
    tol = 5e-4 if f16 else 4e-3
    assert rel_l2(out[..., :64].float(), ref[..., :64]) < tol and rel_l2(out[..., 64:].float(), ref[..., 64:]) < tol
    y = out.double().reshape(-1, 128)
    assert ((part[:, 0].double().sum(0) - y.sum(0)).abs().max() / y.abs().sum(0).max()).item() < 1e-6
    assert ((part[:, 1].double().sum(0) - (y * y).sum(0)).abs().max() / (y * y).sum(0).max()).item() < 1e-6
    # the multi-tensor packer writes the same image (one launch per optimizer step in the engine)
    import numpy as np
    img2 = torch.zeros_like(img)
    desc = np.zeros(2, dtype=np.dtype([("w", "<u8"), ("wf", "<u8"), ("wd", "<u8"), ("Cout", "<i4"), ("Cin", "<i4"), ("KS", "<i4"),
                                       ("dt_f", "<i4"), ("dt_d", "<i4"), ("mode", "<i4")]))
    desc[0] = (w3.data_ptr(), img2.data_ptr(), 0, 64, Cin, 3, 2 if f16 else 1, 1, 1)
    desc[1] = (w5.data_ptr(), img2.data_ptr(), 0, 64, Cin, 5, 2 if f16 else 1, 1, 2)
    table = torch.from_numpy(desc.view(np.uint8).copy()).cuda()
    _lib.call("tsr_pack_conv_weights_multi", table.data_ptr(), 2, 64 * Cin * 25, _st())
    assert torch.equal(img, img2)


@pytest.mark.parametrize("Cin,Cout,KS,B", [(64, 64, 3, 1024), (64, 64, 5, 256), (128, 128, 3, 256), (128, 128, 5, 1024),
                                           (256, 64, 1, 256), (448, 64, 3, 64)])
def test_wgrad_tc_large_batch(Cin, Cout, KS, B):
    """tcgen05 weight gradient at loop-exercising batch sizes (many pixel tiles per CTA, the full split grid), incl.
    accumulation onto an existing gradient; bit-deterministic."""
    from tactilesr_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(Cin + Cout + KS + B)
    bf = torch.bfloat16
    x = torch.randn(B, H, W, Cin, device="cuda").to(bf)
    dy = torch.randn(B, H, W, Cout, device="cuda").to(bf)
    need = L.tsr_conv2d_wgrad_tc_workspace(B, H, W, Cin, Cout, KS)
    ws = torch.empty(max(int(need), 256), dtype=torch.uint8, device="cuda")
    dw = torch.zeros(Cout, Cin, KS, KS, device="cuda")
    _lib.call("tsr_conv2d_wgrad_tc", x.data_ptr(), Cin, dy.data_ptr(), Cout, dw.data_ptr(), ws.data_ptr(), ws.numel(), B, H, W,
              Cin, Cout, KS, 0, _st())
    # exact bf16 products, fp32 accumulation over up to 1.6e6 pixels: against the fp64 result the error is the summation
    # order's (~sqrt(K) * 2^-24); cuDNN's own fp32 weight gradient sits at the same distance from fp64
    ref = torch.nn.grad.conv2d_weight(_nchw(x).double(), (Cout, Cin, KS, KS), _nchw(dy).double(), padding=KS // 2)
    assert rel_l2(dw, ref) < max(5e-5, 1.5e-7 * (B * H * W) ** 0.5), rel_l2(dw, ref)      # 1.9e-4 at B = 1024 (K = 1.6e6)
    dw2 = dw.clone()
    _lib.call("tsr_conv2d_wgrad_tc", x.data_ptr(), Cin, dy.data_ptr(), Cout, dw2.data_ptr(), ws.data_ptr(), ws.numel(), B, H, W,
              Cin, Cout, KS, 1, _st())
    assert torch.equal(dw2, dw + dw)


@pytest.mark.parametrize("Cin,Cout,KS,B", [(64, 64, 3, 256), (64, 64, 5, 64), (128, 128, 3, 64), (128, 128, 5, 256), (256, 64, 1, 64),
                                           (64, 64, 3, 1)])
def test_wgrad_tc_fp16_activations(Cin, Cout, KS, B):
    """The "fp16" mode hands the weight gradient its fp16 forward activations: the kernel converts every x halo tile to bf16
    in shared memory (tcgen05 kind::f16 cannot mix fp16 x with bf16 dy).  Must equal -- bit for bit -- the gradient computed
    from an explicit bf16 copy of x, and the fp64 gradient of that copy to summation-order accuracy."""
    from tactilesr_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(Cin + Cout + KS + B)
    x16 = torch.randn(B, H, W, Cin, device="cuda").to(torch.float16)
    xb = x16.to(torch.bfloat16)
    dy = torch.randn(B, H, W, Cout, device="cuda").to(torch.bfloat16)
    need = L.tsr_conv2d_wgrad_tc_workspace(B, H, W, Cin, Cout, KS)
    ws = torch.empty(max(int(need), 256), dtype=torch.uint8, device="cuda")
    dw16 = torch.zeros(Cout, Cin, KS, KS, device="cuda")
    dwb = torch.zeros_like(dw16)
    _lib.call("tsr_conv2d_wgrad_tc_x", x16.data_ptr(), Cin, 2, dy.data_ptr(), Cout, dw16.data_ptr(), ws.data_ptr(), ws.numel(), B, H, W,
              Cin, Cout, KS, 0, _st())
    _lib.call("tsr_conv2d_wgrad_tc_x", xb.data_ptr(), Cin, 1, dy.data_ptr(), Cout, dwb.data_ptr(), ws.data_ptr(), ws.numel(), B, H, W,
              Cin, Cout, KS, 0, _st())
    assert torch.equal(dw16, dwb)
    ref = torch.nn.grad.conv2d_weight(_nchw(xb).double(), (Cout, Cin, KS, KS), _nchw(dy).double(), padding=KS // 2)
    assert rel_l2(dw16, ref) < max(5e-5, 1.5e-7 * (B * H * W) ** 0.5)
