"""What the three ``train/*.py`` entry points share: process-group / device setup under ``torchrun``, seeding
(reference cpu/misc.py ``set_random_seed``), the data loaders with the batch sharded across ranks
(``DistributedSampler``; the reference's ``DataLoader(shuffle=True)`` at train/tactileSR_train.py:28-38 on one rank),
synthetic datasets of the reference's shapes for machines without its data files, the evaluation hook and the command
line."""
from __future__ import annotations

import argparse
import logging
import os
import random
from typing import Optional

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, DistributedSampler

from ..cpu import distributed as D
from ..cpu.trainer import HookBase

logger = logging.getLogger(__name__)


def set_random_seed(seed: Optional[int], rank: int = 0) -> None:
    """Same generators as the reference seeds (python, numpy, torch CPU + CUDA); ranks get distinct streams."""
    if seed is None:
        return
    seed = int(seed) + rank
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def setup_device():
    """(rank, world size, device).  Under torchrun: NCCL process group, one rank per GPU (LOCAL_RANK)."""
    rank, local, world = D.init_distributed(auto=True)
    if not torch.cuda.is_available():
        raise RuntimeError("tactilesr_b200 entry points need a CUDA device (sm_100a); there is no CPU path")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    logging.basicConfig(level=logging.INFO if rank == 0 else logging.WARNING, format="[%(asctime)s %(name)s] %(message)s")
    return rank, world, dev


def shutdown() -> None:
    """Leave the process group cleanly at the end of an entry point."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


class SyntheticSRDataset(Dataset):
    """Random (LR, HR) records with the shapes and value ranges of the reference's SR dataset files
    (LR (3*seqsCnt, 4, 4) taxel frames in 0..8, HR (1, 100, 100) in 0..250)."""

    def __init__(self, n: int, seqsCnt: int = 1, seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        self.LR = (torch.rand(n, 3 * seqsCnt, 4, 4, generator=g) * 8).numpy()
        self.HR = (torch.rand(n, 1, 100, 100, generator=g) * 250).numpy()

    def __len__(self):
        return len(self.LR)

    def __getitem__(self, i):
        return self.LR[i], self.HR[i]


class SyntheticPSFDataset(Dataset):
    """Random (LR, depth) records as ``tPSFNetDataSet`` yields them: LR (3, 4, 4) raw taxel units (the trainer divides by
    scale_num), depth (100, 100) contact masks with fractional edges and maximum exactly 1."""

    def __init__(self, n: int, seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        self.LR = (torch.rand(n, 3, 4, 4, generator=g) * 1300).numpy()
        yy, xx = torch.meshgrid(torch.arange(100.0), torch.arange(100.0), indexing="ij")
        cx = 20 + 60 * torch.rand(n, 1, 1, generator=g)
        cy = 20 + 60 * torch.rand(n, 1, 1, generator=g)
        r = 8 + 22 * torch.rand(n, 1, 1, generator=g)
        d = ((xx - cx) ** 2 + (yy - cy) ** 2).sqrt()
        depth = (r + 1.5 - d).clamp(0, 3) / 3          # discs with a 3-pixel soft edge
        self.depth = (depth / depth.amax(dim=(1, 2), keepdim=True)).numpy()

    def __len__(self):
        return len(self.LR)

    def __getitem__(self, i):
        return self.LR[i], self.depth[i]


def make_loader(dataset, batch_size: int, train: bool, world: int, rank: int, seed: int = 0, workers: int = 0):
    """Training: the global batch is sharded -- every rank draws ``batch_size`` samples of its own 1/world slice of a
    per-epoch permutation (``DistributedSampler``; ``DistributedHook`` calls ``set_epoch``).  Evaluation: every rank sees
    the whole set in order, as the reference's single-process loader does."""
    sampler = None
    if train and world > 1:
        sampler = DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=True, seed=seed, drop_last=True)
    return DataLoader(dataset, batch_size=batch_size, shuffle=train and sampler is None, sampler=sampler, num_workers=workers,
                      pin_memory=True, drop_last=train)


class EvalHook(HookBase):
    """reference cpu/hooks/eval_hook.py: run ``eval_func`` every ``period`` epochs (and at the end) and log what it returns."""
    priority = 8

    def __init__(self, period: int, eval_func, names=("test_loss", "test_ssim", "test_psnr")):
        self._period, self._eval_func, self._names = period, eval_func, names

    def after_epoch(self):
        if self.every_n_epochs(self._period) or self.is_last_epoch():
            was_training = self.trainer.model.training
            res = self._eval_func()
            self.trainer.model.train(was_training)
            if res is not None and D.is_main_process():
                res = res if isinstance(res, (tuple, list)) else (res,)
                logger.info("==> [test] " + ", ".join(f"{n}: {v:.4f}" for n, v in zip(self._names, res)))
                self.trainer.log(self.trainer.cur_iter, smooth=False, **{n: float(v) for n, v in zip(self._names, res)})


def parse_cli(description: str, config: dict, argv=None) -> dict:
    """Command line of an entry point: any ``--key value`` overrides ``config[key]`` (typed like the default), plus
    ``--synthetic N`` (N random training records instead of the dataset files), ``--precision`` and ``--cuda-graph``."""
    ap = argparse.ArgumentParser(description=description)
    ap.add_argument("--synthetic", type=int, default=0, help="train on N synthetic records of the reference's shapes")
    ap.add_argument("--precision", default="fp16", choices=("fp32", "fp16", "bf16"))
    ap.add_argument("--cuda-graph", action="store_true", help="replay the iteration as a CUDA graph")
    ap.add_argument("--max-iters", type=int, default=0, help="stop after this many iterations (0: run all epochs)")
    for k, v in config.items():
        if isinstance(v, bool):
            ap.add_argument(f"--{k}", type=lambda s: s.lower() in ("1", "true", "yes"), default=None)
        elif isinstance(v, (int, float, str)):
            ap.add_argument(f"--{k}", type=type(v), default=None)
    ns = vars(ap.parse_args(argv))
    cfg = dict(config)
    for k in config:
        if ns.get(k) is not None:
            cfg[k] = ns[k]
    cfg["_synthetic"], cfg["_precision"], cfg["_cuda_graph"], cfg["_max_iters"] = (ns["synthetic"], ns["precision"],
                                                                                   ns["cuda_graph"], ns["max_iters"])
    return cfg
