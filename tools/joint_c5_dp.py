"""BASELINE.json configs[4] ("C5") on the machine it is defined on: joint tPSFNet + TactileSR training, data parallel, at
8 192 samples per GPU (global batch 65 536 on 8 GPUs).  Per step and rank (the reference's offline data flow):
  (i)   tPSFNet fwd + MSE(LR_z, LR_degrade) + bwd on the rank's 8 192 (LR, depth) samples, gradient average over ranks, Adam
  (ii)  HR = HR_tactile.detach()
  (iii) TactileSR fwd + HR/10 + resize + MSE + bwd on (LR, HR) in MICRO micro-batches (saved activations of 8 192 samples do
        not fit 180 GB: 26 MB/sample), gradients accumulated in FusedAdam's flat buffer, ONE NCCL average, Adam.
        BatchNorm batch = one micro-batch on one rank (DDP semantics).
    torchrun --nproc-per-node N tools/joint_c5_dp.py [per_gpu_batch] [micro_batches] [steps]
Prints one JSON line (rank 0): global samples/s (max over ranks of the CUDA-event time), both losses per step."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import tactilesr_b200 as tb
from tactilesr_b200.cpu import distributed as D
from tactilesr_b200.functional import mse_hr_loss
from tactilesr_b200.model import TactileSR, tPSFNet
from tactilesr_b200.optim import FusedAdam

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
MICRO = int(sys.argv[2]) if len(sys.argv) > 2 else 4
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 4
rank, local, world = D.init_distributed()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
tb.set_precision("fp16")
torch.manual_seed(42)
sr = TactileSR().to(dev).train()
psf = tPSFNet(gama=1.4, perception_scale=None, device=dev).to(dev).train()
o_sr = FusedAdam(sr.parameters(), lr=1e-3, weight_decay=1e-2)
o_psf = FusedAdam(psf.parameters(), lr=1e-4, weight_decay=1e-5)
yy, xx = torch.meshgrid(torch.arange(100.0, device=dev), torch.arange(100.0, device=dev), indexing="ij")


def batch(i):
    g = torch.Generator(device=dev).manual_seed(20_000 + 97 * i + rank)
    cx, cy = torch.rand(B, generator=g, device=dev) * 60 + 20, torch.rand(B, generator=g, device=dev) * 60 + 20
    r = torch.rand(B, generator=g, device=dev) * 18 + 8
    depth = torch.clamp((r[:, None, None] - ((yy - cy[:, None, None]) ** 2 + (xx - cx[:, None, None]) ** 2).sqrt()) / 2 + 0.5, 0, 1)
    pooled = torch.nn.functional.avg_pool2d(depth[:, None], 25)
    LR = torch.cat([pooled * 2 + torch.rand(B, 1, 4, 4, generator=g, device=dev) * 0.2 for _ in range(2)] +
                   [pooled * 12 + torch.rand(B, 1, 4, 4, generator=g, device=dev) * 0.5], 1)
    return LR, depth.unsqueeze(1)


def avg(t):
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.AVG)


def step(LR, depth):
    HR, LRd, _, _ = psf(LR, depth)
    l_psf = torch.nn.functional.mse_loss(LR[:, 2:3], LRd)
    o_psf.zero_grad()
    l_psf.backward()
    for p in psf.parameters():
        avg(p.grad)
    o_psf.step()
    HR = HR.detach()
    o_sr.zero_grad()
    mb = B // MICRO
    l_sr = torch.zeros((), device=dev)
    for k in range(MICRO):
        l = mse_hr_loss(sr(LR[k * mb:(k + 1) * mb]), HR[k * mb:(k + 1) * mb], 10.0) / MICRO
        l.backward()
        l_sr += l.detach()
    avg(o_sr.flat_grad(0))
    o_sr.step()
    return l_psf.detach(), l_sr


data = [batch(i) for i in range(2)]
for i in range(2):
    step(*data[i % 2])
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
losses = []
for i in range(STEPS):
    losses.append(step(*data[i % 2]))
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
lp = torch.stack([a for a, _ in losses]); ls = torch.stack([b for _, b in losses])
avg(lp); avg(ls)
if rank == 0:
    ms = float(ms) / STEPS
    print(json.dumps({"config": "C5 joint tPSFNet + TactileSR, fp16 mode", "n_gpus": world, "per_gpu_batch": B, "global_batch": B * world,
                      "micro_batches": MICRO, "bn_batch": B // MICRO, "ms_per_step": ms, "samples_per_s": B * world / ms * 1e3,
                      "psf_loss": [round(float(v), 5) for v in lp.tolist()], "sr_loss": [round(float(v), 5) for v in ls.tolist()],
                      "max_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}), flush=True)
if world > 1:
    dist.destroy_process_group()
