"""Eager vs CUDA-graph training iteration at small batches: python tools/graph_speed.py [B ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tactilesr_b200 as tb
from tactilesr_b200.train.tactileSR_train import Trainer_tactileSR, build_model_and_optimizer
CFG = dict(seqsCnt=1, axisCnt=3, HR_scale_num=10, scale_factor=10, patternFeatureExtraLayerCnt=6,
           forceFeatureExtraLayerCnt=1, lr=1e-3, weight_decay=1e-2)
dev = torch.device("cuda", 0)
tb.set_precision("fp16")
for B in [int(a) for a in sys.argv[1:]] or [32, 128, 256]:
    for use_graph in (False, True):
        torch.manual_seed(0)
        model, opt = build_model_and_optimizer(CFG, dev)
        data = [(torch.rand(B, 3, 4, 4, device=dev) * 8, torch.rand(B, 1, 100, 100, device=dev) * 250) for _ in range(4)]

        class Loader:
            def __len__(self): return 4
            def __iter__(self):
                while True:
                    yield from data
        tr = Trainer_tactileSR(CFG, model=model, optimizer=opt, lr_scheduler=torch.optim.lr_scheduler.StepLR(opt, 2, 0.8),
                               data_loader=Loader(), max_iters=10 ** 9, log_period=10 ** 9, device=dev, cuda_graph=use_graph)
        for _ in range(5):
            tr.train_one_iter()
        n = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            tr.train_one_iter()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"B={B} {'graph' if use_graph else 'eager'}: {ms:.2f} ms/iter {B / ms * 1e3:.0f} samples/s", flush=True)
