"""Headline step (B per GPU) eager vs Trainer(cuda_graph=True): python tools/graph_ab.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tactilesr_b200 as tb
from tactilesr_b200.train.tactileSR_train import Trainer_tactileSR, build_model_and_optimizer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = dict(seqsCnt=1, axisCnt=3, HR_scale_num=10, scale_factor=10, patternFeatureExtraLayerCnt=6, forceFeatureExtraLayerCnt=1, lr=1e-3, weight_decay=1e-2)
dev = torch.device("cuda", 0)
tb.set_precision("fp16")
data = [(torch.rand(B, 3, 4, 4, device=dev) * 8, torch.rand(B, 1, 100, 100, device=dev) * 250) for _ in range(4)]
class L:
    def __len__(self): return 4
    def __iter__(self):
        while True:
            yield from data
for g in (False, True):
    torch.manual_seed(0)
    m, o = build_model_and_optimizer(cfg, dev)
    t = Trainer_tactileSR(cfg, model=m, optimizer=o, lr_scheduler=torch.optim.lr_scheduler.StepLR(o, 2, 0.8), data_loader=L(),
                          max_iters=10 ** 9, log_period=10 ** 9, device=dev, cuda_graph=g)
    for _ in range(5):
        t.train_one_iter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        t.train_one_iter()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"B={B} cuda_graph={g}: {ms:.2f} ms/step {B / ms * 1e3:.0f} samples/s", flush=True)
    del t, m, o
