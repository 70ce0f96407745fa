"""1k-step loss curves from identical init / data: fp32 mode (parity-proven against the reference) vs bf16 tensor-core
mode, with the reference's schedule for TactileSR (Adam lr 1e-3, wd 1e-2; iteration warm-up over 2000 iters in 'auto'
mode from lr*1e-4, StepLR(2, 0.8) per epoch).  Writes profiles/r01_loss_curve_1k.csv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tactilesr_b200 as tb
from tactilesr_b200.cpu.trainer import LRWarmupScheduler
from tactilesr_b200.functional import mse_hr_loss
from tactilesr_b200.model import TactileSR
from tactilesr_b200.optim import FusedAdam

steps, B, epoch_len = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 32, 250
gen = torch.Generator().manual_seed(0)
# a learnable synthetic task: HR = smooth blob whose amplitude / position follow the taxel frame
def batch(i):
    g = torch.Generator().manual_seed(10_000 + i)
    LR = torch.rand(B, 3, 4, 4, generator=g) * 8
    up = torch.nn.functional.interpolate(LR[:, 2:3], size=(100, 100), mode="bilinear", align_corners=False)
    HR = (up * 25 + torch.rand(B, 1, 100, 100, generator=g) * 5)
    return LR.cuda(), HR.cuda()
curves = {}
for mode in ("fp32", "bf16"):
    tb.set_precision(mode)
    torch.manual_seed(42)
    m = TactileSR().cuda().train()
    opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-2)
    sch = LRWarmupScheduler(torch.optim.lr_scheduler.StepLR(opt, 2, 0.8), True, epoch_len, 2000, False, "auto", 1e-5, 1e-4)
    acc = []
    for i in range(steps):
        LR, HR = batch(i)
        loss = mse_hr_loss(m(LR), HR, 10.0)
        opt.zero_grad(); loss.backward(); opt.step()
        sch.iter_update()
        if (i + 1) % epoch_len == 0:
            sch.epoch_update()
        acc.append(loss.detach())
    curves[mode] = torch.stack(acc).cpu().numpy()
    print(mode, "first/last 5:", curves[mode][:5], curves[mode][-5:])
a, b = curves["fp32"], curves["bf16"]
win = 20
sm = lambda v: np.convolve(v, np.ones(win) / win, mode="valid")
rel = np.abs(sm(b) - sm(a)) / sm(a)
print(f"smoothed({win}) relative difference: mean {rel.mean():.4f} max {rel.max():.4f}; final smoothed loss fp32 {sm(a)[-1]:.4f} bf16 {sm(b)[-1]:.4f}")
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "loss_curve_1k.csv")
np.savetxt(out, np.stack([np.arange(steps), a, b], 1), delimiter=",", header="step,loss_fp32_mode,loss_bf16_mode", comments="")
