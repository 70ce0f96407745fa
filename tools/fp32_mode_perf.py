"""fp32 (FFMA) mode throughput: python tools/fp32_mode_perf.py  -> eval forward at B=1024 and train step at B=64"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tactilesr_b200 as tb
from tactilesr_b200.functional import mse_hr_loss
from tactilesr_b200.model import TactileSR
from tactilesr_b200.optim import FusedAdam

tb.set_precision("fp32")
dev = "cuda"


def timeit(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


m = TactileSR().to(dev)
opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-2)
for B in (64, 256):
    LR = torch.rand(B, 3, 4, 4, device=dev) * 8
    HR = torch.rand(B, 1, 100, 100, device=dev) * 250

    def step():
        loss = mse_hr_loss(m(LR), HR, 10.0)
        opt.zero_grad(); loss.backward(); opt.step()
    m.train()
    ms = timeit(step, 5)
    print(f"fp32 train B={B}: {ms:.1f} ms  {B / ms * 1e3:.0f} samples/s  {B / ms * 1e3 * 43.916e9 / 1e12:.1f} TFLOP/s")
m.eval()
for B in (1024,):
    LR = torch.rand(B, 3, 4, 4, device=dev) * 8
    with torch.no_grad():
        ms = timeit(lambda: m(LR), 3)
    print(f"fp32 eval  B={B}: {ms:.1f} ms  {B / ms * 1e3:.0f} samples/s  {B / ms * 1e3 * 14.642e9 / 1e12:.1f} TFLOP/s")
