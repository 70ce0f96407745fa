"""north_star's tolerance clause "matching loss curves over 1k steps in the tensor-core mode", as a test: 1 000 training
steps of the "fp16" tensor-core mode next to the REFERENCE ARITHMETIC -- the oracle's functional restatement of
model/tactileSR_model.py / model/tPSFNet.py (pinned to the unmodified reference by tests/test_oracle_golden.py) executed
on the same GPU by stock ATen / cuDNN in fp32 with TF32 off, stock torch.optim.Adam -- from identical initial weights and
an identical data stream at the reference's batch size 32, with the reference's schedule: Adam(1e-3, wd 1e-2), 2 000-iteration
'auto' warm-up from lr * 1e-4 towards StepLR(2, 0.8) (train/tactileSR_train.py:212-227, cpu/lr_scheduler.py:106-166);
tPSFNet: Adam(1e-4, wd 1e-5), no warm-up (train/tPSFNet_train.py:201-217).

The data are a learnable synthetic task (the HR label follows the taxel frame), so the loss falls by orders of magnitude
over the run and the curves are compared where it matters: 20-step moving averages, mean and max relative difference."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

STEPS, B, EPOCH_LEN, WIN = 1000, 32, 250, 20


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _smooth(v):
    return np.convolve(np.asarray(v, dtype=np.float64), np.ones(WIN) / WIN, mode="valid")


def _compare(name, got, ref, mean_tol, max_tol):
    got, ref = np.asarray(got), np.asarray(ref)
    first = np.abs(got[:10] - ref[:10]).max() / np.abs(ref[:10]).max()
    rel = np.abs(_smooth(got) - _smooth(ref)) / _smooth(ref)
    print(f"{name}: first-10-step max rel diff {first:.2e}; smoothed({WIN}) rel diff mean {rel.mean():.4f} max {rel.max():.4f}; "
          f"loss {ref[0]:.4g} -> {_smooth(ref)[-1]:.4g} (reference arithmetic) / {_smooth(got)[-1]:.4g} (fp16 mode)")
    assert first < 2e-2, (name, first)
    assert rel.mean() < mean_tol and rel.max() < max_tol, (name, rel.mean(), rel.max())
    assert _smooth(got)[-1] < 0.5 * got[0], "the run must actually learn"


def _sr_batch(i):
    g = torch.Generator().manual_seed(10_000 + i)
    LR = torch.rand(B, 3, 4, 4, generator=g) * 8
    up = torch.nn.functional.interpolate(LR[:, 2:3], size=(100, 100), mode="bilinear", align_corners=False)
    HR = up * 25 + torch.rand(B, 1, 100, 100, generator=g) * 5
    return LR.cuda(), HR.cuda()


def _schedule(opt):
    from tactilesr_b200.cpu.trainer import LRWarmupScheduler
    return LRWarmupScheduler(torch.optim.lr_scheduler.StepLR(opt, 2, 0.8), True, EPOCH_LEN, 2000, False, "auto", 1e-5, 1e-4)


def _tick(sch, i):
    sch.iter_update()
    if (i + 1) % EPOCH_LEN == 0:
        sch.epoch_update()


def _reference_sr_curve(sd0, batches):
    """The reference's training loop in its own arithmetic: oracle forward (stock conv2d / batch_norm / interpolate on CUDA,
    fp32), autograd, stock Adam."""
    from oracle import tactilesr_oracle as so
    keys = set(so.param_keys(sd0))
    leaf = OrderedDict((k, v.detach().clone().cuda().requires_grad_(k in keys)) for k, v in sd0.items())
    opt = torch.optim.Adam([leaf[k] for k in sd0 if k in keys], lr=1e-3, weight_decay=1e-2)
    sch = _schedule(opt)
    losses = []
    for i, (LR, HR) in enumerate(batches):
        out = so.tactilesr_forward(leaf, LR, True, 10, new_stats={})
        loss = torch.mean((out - so.prep_hr(HR, 10.0, out.shape[-1])) ** 2)
        opt.zero_grad()
        loss.backward()
        opt.step()
        _tick(sch, i)
        losses.append(loss.detach())
    return torch.stack(losses).cpu().numpy()


def test_sr_1k_step_loss_curve_matches_reference_arithmetic():
    import tactilesr_b200 as tb
    from tactilesr_b200.functional import mse_hr_loss
    from tactilesr_b200.model import TactileSR
    from tactilesr_b200.optim import FusedAdam
    torch.manual_seed(42)
    m = TactileSR().cuda().train()
    sd0 = OrderedDict((k, v.detach().clone()) for k, v in m.state_dict().items())
    batches = (_sr_batch(i) for i in range(STEPS))
    ref = _reference_sr_curve(sd0, batches)
    tb.set_precision("fp16")
    try:
        opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-2)
        sch = _schedule(opt)
        got = []
        for i in range(STEPS):
            LR, HR = _sr_batch(i)
            loss = mse_hr_loss(m(LR), HR, 10.0)
            opt.zero_grad()
            loss.backward()
            opt.step()
            _tick(sch, i)
            got.append(loss.detach())
        got = torch.stack(got).cpu().numpy()
        tb.check_fp16_overflow()
    finally:
        tb.set_precision("fp32")
    _compare("TactileSR", got, ref, mean_tol=0.02, max_tol=0.10)        # measured on B200: 0.45 % / 2.8 %


yy, xx = torch.meshgrid(torch.arange(100.0), torch.arange(100.0), indexing="ij")


def _c5_batch(i):
    """contact discs with fractional edges (max exactly 1) and a taxel frame that follows them (tools/joint_c5.py)"""
    g = torch.Generator().manual_seed(20_000 + i)
    cx, cy = torch.rand(B, generator=g) * 60 + 20, torch.rand(B, generator=g) * 60 + 20
    r = torch.rand(B, generator=g) * 18 + 8
    depth = torch.clamp((r[:, None, None] - ((yy - cy[:, None, None]) ** 2 + (xx - cx[:, None, None]) ** 2).sqrt()) / 2 + 0.5, 0, 1)
    pooled = torch.nn.functional.avg_pool2d(depth[:, None], 25)
    LR = torch.cat([pooled * 2 + torch.rand(B, 1, 4, 4, generator=g) * 0.2 for _ in range(2)] +
                   [pooled * 12 + torch.rand(B, 1, 4, 4, generator=g) * 0.5], 1)
    return LR.cuda(), depth.unsqueeze(1).cuda()


def test_joint_c5_loss_curves_match_reference_arithmetic():
    """BASELINE.json configs[4] ("C5") data flow per step: tPSFNet fwd + MSE(LR_z, LR_degrade) + bwd + Adam, then
    HR = HR_tactile.detach() feeds one TactileSR step (depth2tactile.py:107-119 -> tactileSR_train.py:41-51).  Both loss
    curves of the fp16 mode (PSF kernels fp32-accurate, MLP and SR stacks on the tensor cores) against the reference
    arithmetic.  300 steps keep the reference side (a grouped 99x99 convolution per sample) within a minute."""
    import tactilesr_b200 as tb
    from oracle import tactilesr_oracle as so
    from oracle import tpsf_oracle as po
    from tactilesr_b200.functional import mse_hr_loss
    from tactilesr_b200.model import TactileSR, tPSFNet
    from tactilesr_b200.optim import FusedAdam
    steps = 300
    torch.manual_seed(42)
    sr = TactileSR().cuda().train()
    psf = tPSFNet(gama=1.4, perception_scale=None, device="cuda").cuda().train()
    sr0 = OrderedDict((k, v.detach().clone()) for k, v in sr.state_dict().items())
    psf0 = OrderedDict((k, v.detach().clone()) for k, v in psf.state_dict().items())
    # ---- reference arithmetic
    keys = set(so.param_keys(sr0))
    leaf = OrderedDict((k, v.clone().cuda().requires_grad_(k in keys)) for k, v in sr0.items())
    pleaf = OrderedDict((k, v.clone().cuda().requires_grad_(True)) for k, v in psf0.items())
    o_sr = torch.optim.Adam([leaf[k] for k in sr0 if k in keys], lr=1e-3, weight_decay=1e-2)
    o_psf = torch.optim.Adam(list(pleaf.values()), lr=1e-4, weight_decay=1e-5)
    sch = _schedule(o_sr)
    ref_p, ref_s = [], []
    for i in range(steps):
        LR, depth = _c5_batch(i)
        HR, LRd, _, _ = po.tpsf_forward(pleaf, LR, depth)
        lp = torch.mean((LR[:, 2:3] - LRd) ** 2)
        o_psf.zero_grad(); lp.backward(); o_psf.step()
        out = so.tactilesr_forward(leaf, LR, True, 10, new_stats={})
        ls = torch.mean((out - so.prep_hr(HR.detach(), 10.0, out.shape[-1])) ** 2)
        o_sr.zero_grad(); ls.backward(); o_sr.step()
        _tick(sch, i)
        ref_p.append(lp.detach()); ref_s.append(ls.detach())
    ref_p, ref_s = torch.stack(ref_p).cpu().numpy(), torch.stack(ref_s).cpu().numpy()
    # ---- ours
    tb.set_precision("fp16")
    try:
        f_sr = FusedAdam(sr.parameters(), lr=1e-3, weight_decay=1e-2)
        f_psf = FusedAdam(psf.parameters(), lr=1e-4, weight_decay=1e-5)
        sch = _schedule(f_sr)
        got_p, got_s = [], []
        for i in range(steps):
            LR, depth = _c5_batch(i)
            HR, LRd, _, _ = psf(LR, depth)
            lp = torch.nn.functional.mse_loss(LR[:, 2:3], LRd)
            f_psf.zero_grad(); lp.backward(); f_psf.step()
            ls = mse_hr_loss(sr(LR), HR.detach(), 10.0)
            f_sr.zero_grad(); ls.backward(); f_sr.step()
            _tick(sch, i)
            got_p.append(lp.detach()); got_s.append(ls.detach())
        got_p, got_s = torch.stack(got_p).cpu().numpy(), torch.stack(got_s).cpu().numpy()
    finally:
        tb.set_precision("fp32")
    _compare("C5 tPSFNet", got_p, ref_p, mean_tol=0.005, max_tol=0.02)      # measured: 0.01 % / 0.05 %
    _compare("C5 TactileSR", got_s, ref_s, mean_tol=0.02, max_tol=0.08)     # measured: 0.13 % / 0.54 %
