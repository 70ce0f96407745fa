import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import tactilesr_b200 as tb
from tactilesr_b200 import _lib
from tactilesr_b200.functional import mse_hr_loss
from tactilesr_b200.model import TactileSR
from tactilesr_b200.optim import FusedAdam
from tests.util import sr_inputs
def run(mode, dm, seed0=900):
    _lib.lib().tsr_set_tc_desc_mode(dm)
    tb.set_precision(mode)
    torch.manual_seed(42)
    m = TactileSR().cuda().train()
    opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-2)
    out = []
    for t in range(40):
        LR, HR_raw = sr_inputs(16, 1, seed0 + t)
        loss = mse_hr_loss(m(LR.cuda()), HR_raw.cuda(), 10.0)
        opt.zero_grad(); loss.backward(); opt.step()
        out.append(loss.item())
    return np.array(out)
for seed0 in (900, 2000, 3000):
    ref = run("fp32", 0, seed0)
    for mode in ("bf16", "fp16"):
        for dm in (0, 128):
            c = run(mode, dm, seed0)
            rel = np.abs(c - ref) / ref
            print(f"seed {seed0} {mode} fused={dm==0}: mean {rel.mean():.4f} max {rel.max():.4f} argmax {rel.argmax()}")
