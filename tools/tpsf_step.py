"""tPSFNet training-step timing (C2): [TSR_PRECISION=fp16] python tools/tpsf_step.py [B ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tactilesr_b200.model import tPSFNet
from tactilesr_b200.optim import FusedAdam
dev = "cuda"
import tactilesr_b200 as tb
tb.set_precision(os.environ.get("TSR_PRECISION", "fp32"))
for Bp in [int(a) for a in sys.argv[1:]] or [256, 2048, 8192]:
    pm = tPSFNet(gama=1.4, perception_scale=None, device=dev).to(dev)
    popt = FusedAdam(pm.parameters(), lr=1e-4, weight_decay=1e-5)
    x = torch.rand(Bp, 3, 4, 4, device=dev) * 13
    yy, xx = torch.meshgrid(torch.arange(100.0, device=dev), torch.arange(100.0, device=dev), indexing="ij")
    depth = torch.clamp((20 - ((yy - 50) ** 2 + (xx - 45) ** 2).sqrt()) / 2 + 0.5, 0, 1).expand(Bp, 1, 100, 100).contiguous()

    def step():
        HR, LRd, _, _ = pm(x, depth)
        loss = torch.nn.functional.mse_loss(x[:, 2:3], LRd)
        popt.zero_grad(); loss.backward(); popt.step()
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"tPSF train B={Bp}: {ms:.3f} ms/step {Bp/ms*1e3:.0f} samples/s")
