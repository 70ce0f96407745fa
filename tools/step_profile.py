"""One profiled training step (after warm-up) for ncu: ncu --profile-from-start off ... python tools/step_profile.py B mode"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tactilesr_b200 as tb
from tactilesr_b200.functional import mse_hr_loss
from tactilesr_b200.model import TactileSR
from tactilesr_b200.optim import FusedAdam

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
tb.set_precision(mode)
torch.manual_seed(0)
m = TactileSR().cuda().train()
opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-2)
LR = torch.rand(B, 3, 4, 4, device="cuda") * 8
HR = torch.rand(B, 1, 100, 100, device="cuda") * 250


def step():
    loss = mse_hr_loss(m(LR), HR, 10.0)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
l = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", l.item())
