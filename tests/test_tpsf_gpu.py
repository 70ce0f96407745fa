"""GPU parity of the tPSFNet path against the golden vectors from the unmodified reference and the CPU oracle.
Tolerances: outputs 2e-5 rel-L2 vs fp64 (reference fp32 noise 4e-7); gradients 1e-3 on the l2 norm and rel-L2."""
import numpy as np
import pytest
import torch

from tests.util import load_golden, rel_l2, summarize

pytestmark = pytest.mark.gpu


def _setup(g):
    from oracle import tpsf_oracle as po
    from tactilesr_b200.model import tPSFNet
    m = tPSFNet(gama=1.4, perception_scale=None, device="cuda")
    m.load_state_dict(po.make_state(int(g["seed_w"])), strict=True)
    m = m.cuda()
    LR_raw = torch.from_numpy(g["LR_raw"])
    depth = po.synthetic_depth(int(g["B"]), int(g["seed_x"]) + 1)
    return m, LR_raw, depth


def test_tpsf_forward_backward_matches_reference():
    g = load_golden("tpsf_fwdbwd.npz")
    m, LR_raw, depth = _setup(g)
    LR = LR_raw.cuda() / 100
    HR, LRd, psf, ab = m(LR, depth.cuda().unsqueeze(1))
    assert HR.shape == (4, 1, 100, 100) and LRd.shape == (4, 1, 4, 4) and psf.shape == (4, 1, 99, 99) and ab.shape == (4, 1, 3)
    assert rel_l2(ab, g["f64/alphaBeta"]) < 2e-6
    assert rel_l2(HR, g["f64/HR"]) < 2e-5
    assert rel_l2(LRd, g["f64/LRd"]) < 2e-5
    assert rel_l2(psf[:, 0, 49], g["f64/psf_center_row"]) < 2e-5
    for b in range(4):
        got = summarize(psf[b])
        assert abs(got[0] - g["f64/psf_summary"][b][0]) / g["f64/psf_summary"][b][0] < 2e-5
    loss = torch.nn.functional.mse_loss(LR[:, 2:3], LRd)
    assert abs(loss.item() - float(g["f64/loss"])) / float(g["f64/loss"]) < 2e-5
    loss.backward()
    for (n, p), want in zip(m.named_parameters(), g["f64/grad_summary"]):
        got = summarize(p.grad)
        assert abs(got[0] - want[0]) / want[0] < 1e-3, (n, got[0], want[0])
        k = min(len(got), len(want))
        scale = max(np.abs(want[3:k]).max(), 1e-30)
        assert np.abs(got[3:k] - want[3:k]).max() / scale < 1e-3, n


def test_tpsf_all_output_gradients_match_cpu_oracle():
    """Gradients through every output (HR, LR_degrade, psf, alphaBeta), with gamma large enough that the
    mask minimum m = exp(-100/gamma) does not underflow."""
    from oracle import tpsf_oracle as po
    from tactilesr_b200.model import tPSFNet
    B = 3
    sd = po.make_state(9)
    sd["MLP_layer.7.bias"] = sd["MLP_layer.7.bias"] + torch.tensor([0.5, 1.0, 30.0])   # alpha, beta, gamma up
    gen = torch.Generator().manual_seed(10)
    LR = torch.rand(B, 3, 4, 4, generator=gen) * 13
    depth = po.synthetic_depth(B, 11)
    wHR = torch.rand(B, 1, 100, 100, generator=gen)
    wL = torch.rand(B, 1, 4, 4, generator=gen) * 100
    wP = torch.rand(B, 1, 99, 99, generator=gen)
    wA = torch.rand(B, 1, 3, generator=gen)
    leaf = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    HRo, LRdo, psfo, abo = po.tpsf_forward(leaf, LR.double(), depth.double().unsqueeze(1))
    tot = (HRo * wHR).sum() + (LRdo * wL).sum() + (psfo * wP).sum() + (abo * wA).sum()
    go = torch.autograd.grad(tot, list(leaf.values()))
    m = tPSFNet(1.4, None, device="cuda")
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    HR, LRd, psf, ab = m(LR.cuda(), depth.cuda().unsqueeze(1))
    assert rel_l2(HR, HRo) < 2e-5 and rel_l2(LRd, LRdo) < 2e-5 and rel_l2(psf, psfo) < 2e-5 and rel_l2(ab, abo) < 2e-6
    ((HR * wHR.cuda()).sum() + (LRd * wL.cuda()).sum() + (psf * wP.cuda()).sum() + (ab * wA.cuda()).sum()).backward()
    for (n, p), gref in zip(m.named_parameters(), go):
        assert rel_l2(p.grad, gref) < 1e-3, (n, rel_l2(p.grad, gref))


def test_tpsf_edge_cases():
    """single-sample batch; all-contact and single-pixel-contact depth maps (mask edge cases)."""
    from oracle import tpsf_oracle as po
    from tactilesr_b200.model import tPSFNet
    sd = po.make_state(3)
    m = tPSFNet(1.4, None, device="cuda")
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    depth = torch.zeros(3, 1, 100, 100)
    depth[0] = 1.0                       # every pixel is contact: HR is the constant second max (= 0)
    depth[1, 0, 50, 50] = 1.0            # one contact pixel
    depth[2, 0, :, :50] = 0.5            # no pixel reaches 1: contact = the 0.5 plateau
    LR = torch.rand(3, 3, 4, 4, generator=torch.Generator().manual_seed(4)) * 13
    HRo, LRdo, psfo, abo = po.tpsf_forward({k: v.double() for k, v in sd.items()}, LR.double(), depth.double())
    for sl in (slice(0, 3), slice(1, 2)):
        HR, LRd, psf, ab = m(LR[sl].cuda(), depth[sl].cuda())
        assert torch.isfinite(HR).all() and torch.isfinite(LRd).all()
        assert (HR - HRo[sl].float().cuda()).abs().max() <= 2e-5 * max(HRo[sl].abs().max().item(), 1e-6) + 1e-7
        assert rel_l2(LRd, LRdo[sl]) < 5e-5 or LRdo[sl].abs().max() < 1e-12
    with pytest.raises(AssertionError):
        m(LR.cuda(), depth[:2].cuda())


@pytest.mark.parametrize("B", [1, 5, 700])
def test_psf_tensor_core_forward_matches_ffma_forward(B):
    """The tcgen05 forward (fp16 hi/lo split operands, fp32 TMEM accumulation) against the FFMA forward on the same
    inputs, through the C ABI: wide beta / gamma ranges, ragged batch (700 > 2 CTAs x 148 SMs => persistent loop),
    a depth map scaled by 3 (power-of-two pre-scaling path) and an all-zero one."""
    from oracle import tpsf_oracle as po
    from tactilesr_b200 import _lib
    g = torch.Generator().manual_seed(B)
    ab = torch.stack([torch.rand(B, generator=g) * 2 + 0.2, torch.rand(B, generator=g) * 3 + 0.25,
                      torch.rand(B, generator=g) * 40 + 0.4], 1).cuda().contiguous()
    depth = po.synthetic_depth(min(B, 16), 5).repeat((B + 15) // 16, 1, 1)[:B].clone()
    if B > 2:
        depth[1] *= 3.0
        depth[2] = 0.0
    depth = depth.cuda().contiguous()
    st = torch.cuda.current_stream().cuda_stream
    outs = {}
    for name in ("tsr_psf_forward_ffma", "tsr_psf_forward_tc"):
        HR = torch.empty(B, 100, 100, device="cuda"); LRd = torch.empty(B, 16, device="cuda"); psf = torch.empty(B, 99, 99, device="cuda")
        if name.endswith("_tc"):
            _lib.call(name, ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(), 0, B, st)
        else:
            _lib.call(name, ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(), B, st)
        outs[name] = (HR, LRd, psf)
    (h0, l0, p0), (h1, l1, p1) = outs["tsr_psf_forward_ffma"], outs["tsr_psf_forward_tc"]
    assert torch.isfinite(h1).all() and torch.isfinite(l1).all()
    assert torch.equal(p0, p1)
    scale = h0.flatten(1).abs().amax(1).clamp_min(1e-20)[:, None, None]
    err = ((h1 - h0).abs() / scale).max().item()
    print(f"psf tc vs ffma B={B}: HR max err / sample max {err:.2e}, rel-L2 {rel_l2(h1, h0):.2e}, LRd rel-L2 {rel_l2(l1, l0):.2e}")
    assert err < 1e-5, err
    assert rel_l2(h1, h0) < 3e-6 and rel_l2(l1, l0) < 1e-5


@pytest.mark.parametrize("B", [3, 700])
def test_psf_tensor_core_backward_matches_ffma_backward(B):
    """tcgen05 backward (training case: gradient through LR_degrade only; P3 = E2 D E + E D E2 as four split-operand
    contractions, d gamma / d alpha from the forward's per-row statistics) against the FFMA backward, through the C ABI."""
    from oracle import tpsf_oracle as po
    from tactilesr_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(100 + B)
    ab = torch.stack([torch.rand(B, generator=g) * 2 + 0.2, torch.rand(B, generator=g) * 3 + 0.25,
                      torch.rand(B, generator=g) * 40 + 0.4], 1).cuda().contiguous()
    depth = po.synthetic_depth(min(B, 16), 6).repeat((B + 15) // 16, 1, 1)[:B].clone()
    depth[1] *= 3.0
    depth = depth.cuda().contiguous()
    dL = (torch.randn(B, 16, generator=g) * 0.3).cuda().contiguous()
    st = torch.cuda.current_stream().cuda_stream
    HR = torch.empty(B, 100, 100, device="cuda"); LRd = torch.empty(B, 16, device="cuda")
    aux = torch.empty(B, int(L.tsr_psf_aux_floats()), device="cuda")
    _lib.call("tsr_psf_forward_tc", ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), LRd.data_ptr(), 0, aux.data_ptr(), B, st)
    d0 = torch.empty(B, 3, device="cuda"); d1 = torch.empty(B, 3, device="cuda")
    _lib.call("tsr_psf_backward", ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), dL.data_ptr(), 0, 0, d0.data_ptr(), B, st)
    _lib.call("tsr_psf_backward_tc", ab.data_ptr(), depth.data_ptr(), aux.data_ptr(), dL.data_ptr(), d1.data_ptr(), B, st)
    assert torch.isfinite(d1).all()
    for c, name in enumerate(("alpha", "beta", "gamma")):
        e = rel_l2(d1[:, c], d0[:, c])
        worst = ((d1[:, c] - d0[:, c]).abs() / d0[:, c].abs().clamp_min(1e-3 * d0[:, c].abs().max())).max().item()
        print(f"psf bwd tc vs ffma B={B} d{name}: rel-L2 {e:.2e}, worst per-sample {worst:.2e}")
        assert e < 2e-4 and worst < 2e-3, (name, e, worst)


def test_trainer_tpsf_graph_and_eager_agree():
    """Trainer_tPSF (train/tPSFNet_train.py:173-190 mirror): 6 iterations eager vs cuda_graph=True are bit-identical, and
    the loss goes down; eval_func returns the reference's first-sample metrics."""
    from tactilesr_b200.train.tPSFNet_train import Trainer_tPSF, build_model_and_optimizer, eval_func
    from oracle import tpsf_oracle as po
    cfg = dict(gama=1.4, perception_scale=None, lr=1e-3, weight_decay=1e-5, scale_num=100)
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(21)
    data = [(torch.rand(8, 3, 4, 4, generator=g) * 1300, po.synthetic_depth(8, 30 + i)) for i in range(6)]
    res = []
    for use_graph in (False, True):
        torch.manual_seed(3)
        model, opt = build_model_and_optimizer(cfg, dev)
        tr = Trainer_tPSF(100, model=model, optimizer=opt, lr_scheduler=torch.optim.lr_scheduler.StepLR(opt, 1, 0.9),
                          data_loader=data, max_iters=100, log_period=10 ** 9, device=dev, cuda_graph=use_graph)
        losses = []
        for it in range(6):
            tr.cur_iter = it
            tr.train_one_iter()
            losses.append(tr._loss_acc.clone()); tr._loss_acc, tr._loss_cnt = None, 0
        res.append((torch.stack(losses).cpu(), [p.detach().clone() for p in model.parameters()]))
        assert (len(tr._graphs) == 1) == use_graph
    assert torch.equal(res[0][0], res[1][0]), (res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)
    mse, ssim = eval_func(model, data[:2], cfg)
    assert mse > 0 and -1.0 <= ssim <= 1.0


def test_psf_kernels_size_independent_properties():
    """Properties at a large ragged batch (B = 5000: 17 samples per CTA in the persistent loops): (a) permuting the batch
    permutes the results bit for bit (forward and backward); (b) doubling alpha doubles HR, LR_degrade and psf exactly
    (every scaling on the path is a power of two); (c) results do not depend on the launch's batch size."""
    from oracle import tpsf_oracle as po
    from tactilesr_b200 import _lib
    L = _lib.lib()
    B = 5000
    g = torch.Generator().manual_seed(77)
    ab = torch.stack([torch.rand(B, generator=g) + 0.3, torch.rand(B, generator=g) * 2 + 0.3,
                      torch.rand(B, generator=g) * 20 + 0.5], 1).cuda().contiguous()
    depth = po.synthetic_depth(32, 9).repeat((B + 31) // 32, 1, 1)[:B].cuda().contiguous()
    dL = torch.randn(B, 16, generator=g).cuda().contiguous()
    st = torch.cuda.current_stream().cuda_stream
    naux = int(L.tsr_psf_aux_floats())

    def run(ab_, depth_, dL_):
        n = ab_.shape[0]
        HR = torch.empty(n, 100, 100, device="cuda"); LRd = torch.empty(n, 16, device="cuda"); psf = torch.empty(n, 99, 99, device="cuda")
        aux = torch.empty(n, naux, device="cuda"); dab = torch.empty(n, 3, device="cuda")
        _lib.call("tsr_psf_forward_tc", ab_.data_ptr(), depth_.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(), aux.data_ptr(), n, st)
        _lib.call("tsr_psf_backward_tc", ab_.data_ptr(), depth_.data_ptr(), aux.data_ptr(), dL_.data_ptr(), dab.data_ptr(), n, st)
        return HR, LRd, psf, dab
    HR, LRd, psf, dab = run(ab, depth, dL)
    assert all(torch.isfinite(t).all() for t in (HR, LRd, psf, dab))
    perm = torch.randperm(B, generator=g).cuda()
    HRp, LRdp, psfp, dabp = run(ab[perm].contiguous(), depth[perm].contiguous(), dL[perm].contiguous())
    assert torch.equal(HRp, HR[perm]) and torch.equal(LRdp, LRd[perm]) and torch.equal(psfp, psf[perm]) and torch.equal(dabp, dab[perm])
    ab2 = ab.clone(); ab2[:, 0] *= 2
    HR2, LRd2, psf2, _ = run(ab2, depth, dL)
    assert torch.equal(HR2, 2 * HR) and torch.equal(LRd2, 2 * LRd)
    big = psf > 1e-30                       # (below, alpha * e(u) e(v) is rounded on the subnormal grid: doubling is not exact)
    assert torch.equal(psf2[big], 2 * psf[big]) and torch.allclose(psf2, 2 * psf, rtol=0, atol=1e-37)
    HRs, LRds, _, dabs = run(ab[:37].contiguous(), depth[:37].contiguous(), dL[:37].contiguous())
    assert torch.equal(HRs, HR[:37]) and torch.equal(LRds, LRd[:37]) and torch.equal(dabs, dab[:37])


@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_tpsf_16bit_modes_train_step_within_tensor_core_bound(mode, monkeypatch):
    """16-bit precision modes: the two wide MLP layers run on the tcgen05 kernels (forward in fp16 / bf16, gradients on
    bf16 tensors).  north_star's tensor-core bound is 1e-2 relative on the outputs; parameter gradients are held to 5e-2
    rel-L2 (bf16 operands) against the CPU oracle's autograd; the fp32 mode of the same module is the yardstick."""
    import tactilesr_b200 as tb
    from oracle import tpsf_oracle as po
    from tactilesr_b200 import _lib
    from tactilesr_b200.model import tPSFNet
    import sys
    import tactilesr_b200.model  # noqa: F401
    tmod = sys.modules["tactilesr_b200.model.tPSFNet"]       # (the package attribute of that name is the class)
    monkeypatch.setattr(tmod, "_TC_MIN_BATCH", 64)      # (production gate: 1024 -- smaller batches are launch-bound)
    B = 64
    sd = po.make_state(31)
    gen = torch.Generator().manual_seed(32)
    LR = torch.rand(B, 3, 4, 4, generator=gen) * 13
    depth = po.synthetic_depth(B, 33)
    leaf = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    HRo, LRdo, _, abo = po.tpsf_forward(leaf, LR.double(), depth.double().unsqueeze(1))
    torch.nn.functional.mse_loss(LR.double()[:, 2:3], LRdo).backward()
    m = tPSFNet(1.4, None, device="cuda")
    m.load_state_dict(sd)
    m = m.cuda()
    tol_out = {"fp16": 5e-3, "bf16": 1e-2}[mode]
    try:
        tb.set_precision(mode)
        n0 = _lib.launch_count()
        HR, LRd, _, ab = m(LR.cuda(), depth.cuda().unsqueeze(1))
        torch.nn.functional.mse_loss(LR.cuda()[:, 2:3], LRd).backward()
        assert _lib.launch_count() - n0 > 20
        e = {k: rel_l2(v, r) for k, v, r in (("ab", ab, abo), ("HR", HR, HRo), ("LRd", LRd, LRdo))}
        eg = {n: rel_l2(p.grad, leaf[n].grad) for n, p in m.named_parameters()}
        print(mode, e, {k: f"{v:.1e}" for k, v in eg.items()})
        assert max(e.values()) < tol_out, e
        assert max(eg.values()) < 5e-2, eg
        # ragged batch (not a multiple of 64): the fp32 MLP kernels take over; the PSF products stay on the mode's single
        # fp16 pass (measured 8e-5 .. 3e-4)
        HR5, _, _, _ = m(LR[:5].cuda(), depth[:5].cuda().unsqueeze(1))
        assert rel_l2(HR5, HRo[:5]) < 2e-3
        # inference under no_grad takes the same tensor-core path
        with torch.no_grad():
            HRn, _, _, _ = m(LR.cuda(), depth.cuda().unsqueeze(1))
        assert torch.equal(HRn, HR.detach())
    finally:
        tb.set_precision("fp32")


def test_tpsf_fp16_mode_loss_curve_tracks_fp32_mode(monkeypatch):
    """200 Adam steps of train/tPSFNet_train.py's iteration in fp32 and fp16 modes from the same weights and batches: the
    loss curves stay within 2 % of each other (mean over the last 50 steps) and both decrease."""
    import tactilesr_b200 as tb
    from oracle import tpsf_oracle as po
    from tactilesr_b200.model import tPSFNet
    from tactilesr_b200.optim import FusedAdam
    import sys
    import tactilesr_b200.model  # noqa: F401
    tmod = sys.modules["tactilesr_b200.model.tPSFNet"]       # (the package attribute of that name is the class)
    monkeypatch.setattr(tmod, "_TC_MIN_BATCH", 64)
    B, steps = 256, 200
    depth = po.synthetic_depth(B, 41).cuda().unsqueeze(1)
    gen = torch.Generator().manual_seed(42)
    LRs = [(torch.rand(B, 3, 4, 4, generator=gen) * 13).cuda() for _ in range(8)]
    curves = {}
    try:
        for mode in ("fp32", "fp16"):
            tb.set_precision(mode)
            m = tPSFNet(1.4, None, device="cuda")
            m.load_state_dict(po.make_state(43))
            m = m.cuda()
            opt = FusedAdam(m.parameters(), lr=1e-4, weight_decay=1e-5)
            losses = []
            for i in range(steps):
                LR = LRs[i % len(LRs)]
                _, LRd, _, _ = m(LR, depth)
                loss = torch.nn.functional.mse_loss(LR[:, 2:3], LRd)
                opt.zero_grad(); loss.backward(); opt.step()
                losses.append(loss)
            curves[mode] = torch.stack(losses).float().cpu()
    finally:
        tb.set_precision("fp32")
    a, b = curves["fp32"], curves["fp16"]
    print("tPSF loss fp32 first/last", a[0].item(), a[-50:].mean().item(), "fp16", b[0].item(), b[-50:].mean().item())
    assert a[-50:].mean() < a[:10].mean() and b[-50:].mean() < b[:10].mean()
    assert abs(a[-50:].mean() - b[-50:].mean()) / a[-50:].mean() < 2e-2
    assert ((a - b).abs() / a).max() < 0.1


@pytest.mark.parametrize("B", [5, 700])
def test_psf_single_pass_fp16_kernels_within_tensor_core_mode_tolerance(B):
    """The one-pass fp16 tcgen05 kernels (16-bit precision modes, north_star tolerance 1e-2) against the FFMA forward /
    backward through the C ABI: HR and LR_degrade within 2e-3 (measured ~3e-4), d(alpha, beta, gamma) within 1e-2."""
    from oracle import tpsf_oracle as po
    from tactilesr_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(300 + B)
    ab = torch.stack([torch.rand(B, generator=g) * 2 + 0.2, torch.rand(B, generator=g) * 3 + 0.25,
                      torch.rand(B, generator=g) * 40 + 0.4], 1).cuda().contiguous()
    depth = po.synthetic_depth(min(B, 16), 7).repeat((B + 15) // 16, 1, 1)[:B].clone()
    depth[1] *= 3.0
    depth[2] = 0.0
    depth = depth.cuda().contiguous()
    dL = (torch.randn(B, 16, generator=g) * 0.3).cuda().contiguous()
    st = torch.cuda.current_stream().cuda_stream
    mk = lambda: (torch.empty(B, 100, 100, device="cuda"), torch.empty(B, 16, device="cuda"), torch.empty(B, 99, 99, device="cuda"))
    h0, l0, p0 = mk(); h1, l1, p1 = mk()
    aux = torch.empty(B, int(L.tsr_psf_aux_floats()), device="cuda")
    _lib.call("tsr_psf_forward_ffma", ab.data_ptr(), depth.data_ptr(), h0.data_ptr(), l0.data_ptr(), p0.data_ptr(), B, st)
    _lib.call("tsr_psf_forward_tc_f16", ab.data_ptr(), depth.data_ptr(), h1.data_ptr(), l1.data_ptr(), p1.data_ptr(), aux.data_ptr(), B, st)
    assert torch.isfinite(h1).all() and torch.isfinite(l1).all() and torch.equal(p0, p1)
    scale = h0.flatten(1).abs().amax(1).clamp_min(1e-20)[:, None, None]
    err = ((h1 - h0).abs() / scale).max().item()
    print(f"psf f16 vs ffma B={B}: HR max err / sample max {err:.2e}, rel-L2 {rel_l2(h1, h0):.2e}, LRd rel-L2 {rel_l2(l1, l0):.2e}")
    assert err < 4e-3 and rel_l2(h1, h0) < 2e-3 and rel_l2(l1, l0) < 2e-3
    d0 = torch.empty(B, 3, device="cuda"); d1 = torch.empty(B, 3, device="cuda")
    _lib.call("tsr_psf_backward", ab.data_ptr(), depth.data_ptr(), h0.data_ptr(), dL.data_ptr(), 0, 0, d0.data_ptr(), B, st)
    _lib.call("tsr_psf_backward_tc_f16", ab.data_ptr(), depth.data_ptr(), aux.data_ptr(), dL.data_ptr(), d1.data_ptr(), B, st)
    assert torch.isfinite(d1).all()
    for c, name in enumerate(("alpha", "beta", "gamma")):
        e = rel_l2(d1[:, c], d0[:, c])
        print(f"psf bwd f16 vs ffma B={B} d{name}: rel-L2 {e:.2e}")
        assert e < 1e-2, (name, e)


def test_tpsf_module_fp16_mode_within_1e2_of_reference():
    """tPSFNet under set_precision("fp16"): single-pass PSF kernels (and, at B >= 1024, the tensor-core MLP) against the
    golden vectors of the unmodified reference: outputs and loss within north_star's 1e-2 tensor-core-mode bound."""
    import tactilesr_b200 as tb
    g = load_golden("tpsf_fwdbwd.npz")
    m, LR_raw, depth = _setup(g)
    tb.set_precision("fp16")
    try:
        LR = LR_raw.cuda() / 100
        HR, LRd, psf, ab = m(LR, depth.cuda().unsqueeze(1))
        assert rel_l2(HR, g["f64/HR"]) < 1e-2 and rel_l2(LRd, g["f64/LRd"]) < 1e-2
        loss = torch.nn.functional.mse_loss(LR[:, 2:3], LRd)
        assert abs(loss.item() - float(g["f64/loss"])) / float(g["f64/loss"]) < 1e-2
        loss.backward()
        for (n, p), want in zip(m.named_parameters(), g["f64/grad_summary"]):
            got = summarize(p.grad)
            assert abs(got[0] - want[0]) / want[0] < 2e-2, (n, got[0], want[0])
    finally:
        tb.set_precision("fp32")
