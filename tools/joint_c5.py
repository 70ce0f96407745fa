"""C5 ("joint" tPSFNet + TactileSR training): the reference's offline data flow executed in one step --
(i) tPSFNet fwd + MSE(LR[:,2:3], LR_degrade) + bwd + Adam(1e-4, wd 1e-5) on (LR, depth)   (train/tPSFNet_train.py:180-201)
(ii) HR = HR_tactile.detach()                                                              (depth2tactile.py:107-119)
(iii) TactileSR fwd + HR/10 + resize + MSE + bwd + Adam(1e-3, wd 1e-2) on (LR, HR)          (train/tactileSR_train.py:41-51,212)
with the reference's TactileSR schedule (2000-iteration 'auto' warm-up from lr*1e-4, StepLR(2, 0.8) per epoch) and no
warm-up for tPSFNet.  Runs the same seeded stream in the fp32 mode (parity-proven against the reference) and in the
tensor-core modes and writes both loss curves.

    python tools/joint_c5.py curves [steps] [B]      -> gpurun_out/joint_c5_curves.csv + summary
    python tools/joint_c5.py speed [B]               -> joint step throughput in the fp16 mode
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tactilesr_b200 as tb
from tactilesr_b200.cpu.trainer import LRWarmupScheduler
from tactilesr_b200.functional import mse_hr_loss
from tactilesr_b200.model import TactileSR, tPSFNet
from tactilesr_b200.optim import FusedAdam

dev = torch.device("cuda", 0)
yy, xx = torch.meshgrid(torch.arange(100.0), torch.arange(100.0), indexing="ij")


def batch(i, B):
    """contact blobs (discs with fractional edges, max exactly 1) and a taxel frame that follows them"""
    g = torch.Generator().manual_seed(20_000 + i)
    cx, cy = torch.rand(B, generator=g) * 60 + 20, torch.rand(B, generator=g) * 60 + 20
    r = torch.rand(B, generator=g) * 18 + 8
    depth = torch.clamp((r[:, None, None] - ((yy - cy[:, None, None]) ** 2 + (xx - cx[:, None, None]) ** 2).sqrt()) / 2 + 0.5, 0, 1)
    pooled = torch.nn.functional.avg_pool2d(depth[:, None], 25)                       # (B,1,4,4)
    LR = torch.cat([pooled * 2 + torch.rand(B, 1, 4, 4, generator=g) * 0.2 for _ in range(2)] +
                   [pooled * 12 + torch.rand(B, 1, 4, 4, generator=g) * 0.5], 1)     # z axis carries the force, 0..13
    return LR.to(dev), depth.unsqueeze(1).to(dev)


def build(seed=42):
    torch.manual_seed(seed)
    sr = TactileSR().to(dev).train()
    psf = tPSFNet(gama=1.4, perception_scale=None, device=dev).to(dev).train()
    o_sr = FusedAdam(sr.parameters(), lr=1e-3, weight_decay=1e-2)
    o_psf = FusedAdam(psf.parameters(), lr=1e-4, weight_decay=1e-5)
    return sr, psf, o_sr, o_psf


def joint_step(sr, psf, o_sr, o_psf, LR, depth):
    HR, LRd, _, _ = psf(LR, depth)
    l_psf = torch.nn.functional.mse_loss(LR[:, 2:3], LRd)
    o_psf.zero_grad(); l_psf.backward(); o_psf.step()
    l_sr = mse_hr_loss(sr(LR), HR.detach(), 10.0)
    o_sr.zero_grad(); l_sr.backward(); o_sr.step()
    return l_psf.detach(), l_sr.detach()


what = sys.argv[1] if len(sys.argv) > 1 else "curves"
if what == "curves":
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    epoch_len = 250
    curves = {}
    for mode in ("fp32", "fp16", "bf16"):
        tb.set_precision(mode)
        sr, psf, o_sr, o_psf = build()
        sch = LRWarmupScheduler(torch.optim.lr_scheduler.StepLR(o_sr, 2, 0.8), True, epoch_len, 2000, False, "auto", 1e-5, 1e-4)
        a, b = [], []
        for i in range(steps):
            lp, ls = joint_step(sr, psf, o_sr, o_psf, *batch(i, B))
            sch.iter_update()
            if (i + 1) % epoch_len == 0:
                sch.epoch_update()
            a.append(lp); b.append(ls)
        curves[mode] = (torch.stack(a).cpu().numpy(), torch.stack(b).cpu().numpy())
        print(mode, "psf loss first/last:", curves[mode][0][:2], curves[mode][0][-2:], " sr loss first/last:", curves[mode][1][:2], curves[mode][1][-2:], flush=True)
    win = 20
    sm = lambda v: np.convolve(v, np.ones(win) / win, mode="valid")
    for mode in ("fp16", "bf16"):
        for k, name in ((0, "tPSFNet"), (1, "TactileSR")):
            ref, got = curves["fp32"][k], curves[mode][k]
            rel = np.abs(sm(got) - sm(ref)) / sm(ref)
            print(f"{mode} vs fp32 mode, {name} loss: first-10-step max rel diff {np.abs(got[:10] - ref[:10]).max() / ref[:10].max():.2e}; "
                  f"smoothed({win}) rel diff mean {rel.mean():.4f} max {rel.max():.4f}; final smoothed {sm(ref)[-1]:.5g} (fp32) {sm(got)[-1]:.5g} ({mode})")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "joint_c5_curves.csv")
    cols = [np.arange(steps)] + [curves[m][k] for m in ("fp32", "fp16", "bf16") for k in (0, 1)]
    np.savetxt(out, np.stack(cols, 1), delimiter=",", comments="",
               header="step,psf_fp32,sr_fp32,psf_fp16,sr_fp16,psf_bf16,sr_bf16")
else:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    tb.set_precision("fp16")
    sr, psf, o_sr, o_psf = build()
    data = [batch(i, B) for i in range(3)]
    for i in range(3):
        joint_step(sr, psf, o_sr, o_psf, *data[i % 3])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    n = 6
    for i in range(n):
        joint_step(sr, psf, o_sr, o_psf, *data[i % 3])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"joint C5 step (fp16 mode) B={B}: {ms:.2f} ms/step, {B / ms * 1e3:.0f} samples/s per GPU")
