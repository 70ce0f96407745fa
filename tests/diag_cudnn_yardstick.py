"""Stock PyTorch eager (cuDNN) running the reference's TactileSR training step on the same B200 -- "the thing to beat"
(SURVEY section 8d).  The reference files cannot travel to the GPU box, so the step is the oracle's functional restatement
(F.conv2d / F.batch_norm / F.interpolate: the same ATen / cuDNN calls nn.Module issues) with torch.optim.Adam.
    python tests/diag_cudnn_yardstick.py [B ...]
Not collected by pytest (a timing diagnostic that uses the oracle, hence under tests/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import tactilesr_oracle as so

dev = "cuda"


def run(B, mode, steps=6):
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    torch.backends.cudnn.benchmark = True
    sd = so.make_state(so.tactilesr_layout(1), 42, nondegenerate=False)
    keys = set(so.param_keys(sd))
    leaf = {k: (v.to(dev).requires_grad_(True) if k in keys else v.to(dev)) for k, v in sd.items()}
    opt = torch.optim.Adam([leaf[k] for k in sd if k in keys], lr=1e-3, weight_decay=1e-2)
    g = torch.Generator().manual_seed(1)
    LR = (torch.rand(B, 3, 4, 4, generator=g) * 8).to(dev)
    HR = (torch.rand(B, 1, 100, 100, generator=g) * 250).to(dev)

    def step():
        new_stats = {}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "autocast-bf16"):
            out = so.tactilesr_forward(leaf, LR, True, new_stats=new_stats)
        loss = torch.mean((out.float() - so.prep_hr(HR, 10.0, 40)) ** 2)
        opt.zero_grad()
        loss.backward()
        opt.step()
        leaf.update(new_stats)
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return ms, B / ms * 1e3


print("| stock PyTorch eager (cuDNN) | batch | ms/step | samples/s |")
print("|---|---|---|---|")
for B in [int(a) for a in sys.argv[1:]] or [32, 256, 512]:
    for mode in ("fp32", "tf32", "autocast-bf16"):
        try:
            ms, sps = run(B, mode)
            print(f"| {mode} | {B} | {ms:.1f} | {sps:.0f} |", flush=True)
        except torch.OutOfMemoryError:
            print(f"| {mode} | {B} | OOM | - |", flush=True)
        torch.cuda.empty_cache()
