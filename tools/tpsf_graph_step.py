"""tPSFNet training iteration through Trainer_tPSF, eager and CUDA-graph: python tools/tpsf_graph_step.py MODE B [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tactilesr_b200 as tb
from tactilesr_b200.train.tPSFNet_train import Trainer_tPSF, build_model_and_optimizer

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
tb.set_precision(mode)
dev = torch.device("cuda", 0)
cfg = dict(gama=1.4, perception_scale=None, lr=1e-4, weight_decay=1e-5, scale_num=100)
yy, xx = torch.meshgrid(torch.arange(100.0), torch.arange(100.0), indexing="ij")
plane = torch.clamp((20 - ((yy - 50) ** 2 + (xx - 45) ** 2).sqrt()) / 2 + 0.5, 0, 1)
for B in [int(a) for a in sys.argv[2:]] or [8192]:
    data = [((torch.rand(B, 3, 4, 4) * 1300).to(dev), plane.expand(B, 100, 100).contiguous().to(dev))]
    for graph in (False, True):
        model, opt = build_model_and_optimizer(cfg, dev)
        tr = Trainer_tPSF(100, model=model, optimizer=opt, lr_scheduler=torch.optim.lr_scheduler.StepLR(opt, 10 ** 6, 0.9),
                          data_loader=data, max_iters=10 ** 6, log_period=10 ** 9, device=dev, cuda_graph=graph)
        it = 0
        for _ in range(6):
            tr.cur_iter = it; it += 1
            tr.train_one_iter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            tr.cur_iter = it; it += 1
            tr.train_one_iter()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"tPSF {mode} B={B} graph={graph}: {ms:.3f} ms/iter {B / ms * 1e3:.0f} samples/s")
