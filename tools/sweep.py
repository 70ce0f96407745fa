"""Configuration sweep of BASELINE.json (C2-C4 shapes) on one GPU; prints a markdown table (commit it under profiles/).
  C3: TactileSR eval forward, batch 1k..256k, fp16 / bf16 / fp32 modes
  C4-shape: TactileSR(seqsCnt=7) train step (single GPU part)
  C2: tPSFNet train step, batch 256..8192
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tactilesr_b200 as tb
from tactilesr_b200.functional import mse_hr_loss
from tactilesr_b200.model import TactileSR, tPSFNet
from tactilesr_b200.optim import FusedAdam

dev = torch.device("cuda", 0)
FWD = {1: 14.642e9, 7: 16.091e9}
TRAIN = {1: 43.916e9, 7: 48.229e9}


def timeit(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("| config | mode | batch | ms | samples/s | TFLOP/s (algorithmic) |")
print("|---|---|---|---|---|---|")
torch.manual_seed(0)
m = TactileSR().to(dev)
m.train()
with torch.no_grad():
    tb.set_precision("fp32")
    m(torch.rand(64, 3, 4, 4, device=dev) * 8)       # non-trivial running stats
m.eval()
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
for mode, batches in (("fp16", [1024, 4096, 16384, 65536, 262144]), ("bf16", [1024, 16384, 262144]), ("fp32", [1024, 4096, 16384])):
    tb.set_precision(mode)
    for B in batches:
        if quick and B > 16384:
            continue
        LR = torch.rand(B, 3, 4, 4, device=dev) * 8
        with torch.no_grad():
            ms = timeit(lambda: m(LR), 1 if B >= 65536 else 3, warm=1)
        print(f"| C3 TactileSR S=1 eval forward | {mode} | {B} | {ms:.1f} | {B / ms * 1e3:.0f} | {B / ms * 1e3 * FWD[1] / 1e12:.0f} |", flush=True)
for S in (1, 7):
    for mode, B in (("fp16", 1024), ("bf16", 1024), ("fp32", 256)):
        tb.set_precision(mode)
        torch.manual_seed(1)
        ms_ = TactileSR(seqsCnt=S).to(dev).train()
        opt = FusedAdam(ms_.parameters(), lr=1e-4 if S == 7 else 1e-3, weight_decay=1e-2)
        LR = torch.rand(B, 3 * S, 4, 4, device=dev) * 8
        HR = torch.rand(B, 1, 100, 100, device=dev) * 250

        def step():
            loss = mse_hr_loss(ms_(LR), HR, 10.0)
            opt.zero_grad(); loss.backward(); opt.step()
        ms = timeit(step, 4)
        print(f"| C{'1' if S == 1 else '4'}-shape TactileSR S={S} train step | {mode} | {B} | {ms:.1f} | {B / ms * 1e3:.0f} | {B / ms * 1e3 * TRAIN[S] / 1e12:.0f} |", flush=True)
        del ms_, opt
tb.set_precision("fp32")
for B in (256, 2048, 8192, 32768):
    pm = tPSFNet(1.4, None, device=dev).to(dev)
    popt = FusedAdam(pm.parameters(), lr=1e-4, weight_decay=1e-5)
    x = torch.rand(B, 3, 4, 4, device=dev) * 13
    yy, xx = torch.meshgrid(torch.arange(100.0, device=dev), torch.arange(100.0, device=dev), indexing="ij")
    cx = torch.rand(B, device=dev) * 50 + 25
    r = torch.rand(B, device=dev) * 20 + 10
    depth = torch.clamp((r[:, None, None] - ((yy - 50) ** 2 + (xx - cx[:, None, None]) ** 2).sqrt()) / 2 + 0.5, 0, 1).unsqueeze(1)

    def pstep():
        HR, LRd, _, _ = pm(x, depth)
        loss = torch.nn.functional.mse_loss(x[:, 2:3], LRd)
        popt.zero_grad(); loss.backward(); popt.step()
    ms = timeit(pstep, 4)
    with torch.no_grad():
        msf = timeit(lambda: pm(x, depth), 4)
    print(f"| C2 tPSFNet train step | fp32 | {B} | {ms:.2f} | {B / ms * 1e3:.0f} | fwd only: {B / msf * 1e3:.0f} samples/s = {B / msf * 1e3 * 119472 / 1e9:.0f} GB/s compulsory |", flush=True)
