"""Reference train/tactileSR_train.py on the tactilesr_b200 kernels: ``Trainer_tactileSR.train_cal_loss`` (:41-51),
``build_dataloader`` (:28-38), ``eval_func`` (:64-101) and the entry point ``main(config)`` (:199-242) -- runnable as
``python -m tactilesr_b200.train.tactileSR_train`` or, data parallel,
``torchrun --nproc-per-node N -m tactilesr_b200.train.tactileSR_train`` (the batch is sharded over the ranks with a
``DistributedSampler``).  The matplotlib PNG inference hook of that file is outside the hot path."""
from __future__ import annotations

import torch

from ..cpu.trainer import Trainer
from ..functional import eval_metrics, mse_hr_loss
from ..model.tactileSR_model import TactileSR
from ..optim import FusedAdam


class Trainer_tactileSR(Trainer):
    def __init__(self, config, *args, device=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.config = config
        self.seqsCnt = config["seqsCnt"]
        self.axisCnt = config["axisCnt"]
        self.HR_scale_num = config["HR_scale_num"]
        self.scale_factor = config["scale_factor"]
        self.device = device if device is not None else next(self.model_or_module.parameters()).device

    def train_cal_loss(self, batch):
        LR, HR = batch
        # H2D (non_blocking from pinned memory); HR / HR_scale_num, the bilinear resize to 4*sf and the MSE are one
        # fused kernel (functional.mse_hr_loss) instead of 4 ATen launches (reference :43-45, :49).
        LR = LR.to(self.device, non_blocking=True).float()
        HR = HR.to(self.device, non_blocking=True).float()
        LR = LR[:, :self.seqsCnt * self.axisCnt]
        out = self.model(LR)
        loss = mse_hr_loss(out, HR, self.HR_scale_num)
        return loss, {"total_loss": loss}


def build_model_and_optimizer(config, device):
    """reference main() :204-213 (the TactileSR branch): model on the rank-local device + Adam(lr, weight_decay)."""
    model = TactileSR(scale_factor=config["scale_factor"], seqsCnt=config["seqsCnt"], axisCnt=config["axisCnt"],
                      patternFeatureExtraLayerCnt=config["patternFeatureExtraLayerCnt"],
                      forceFeatureExtraLayerCnt=config["forceFeatureExtraLayerCnt"]).to(device)
    optimizer = FusedAdam(model.parameters(), lr=config["lr"], weight_decay=config["weight_decay"])
    return model, optimizer


def eval_func(model, test_loader, config, device=None):
    """reference eval_func (train/tactileSR_train.py:64-101) without its per-sample python loop and per-batch host
    syncs: label preparation, MSE, PSNR and SSIM of a batch are one kernel; the three running sums stay on the device
    and are read back once.  Returns (loss, ssim, psnr) averaged over the batches exactly as the reference logs them.
    Like the reference it runs the model in eval mode (and, unlike it, under ``torch.no_grad()``)."""
    device = device if device is not None else next(model.parameters()).device
    seqsCnt, axisCnt = config["seqsCnt"], config["axisCnt"]
    acc = torch.zeros(3, dtype=torch.float64, device=device)
    n = 0
    model.eval()
    with torch.no_grad():
        for LR, HR in test_loader:
            LR = LR.to(device, non_blocking=True).float()[:, :seqsCnt * axisCnt]
            HR = HR.to(device, non_blocking=True).float()
            out = model(LR)
            mse, psnr, ssim = eval_metrics(out, HR, config["HR_scale_num"], config["sensorMaxVaule_factor"])
            acc += torch.stack([mse.double(), ssim.double().mean(), psnr.double().mean()])
            n += 1
    loss, ssim, psnr = (acc / max(n, 1)).tolist()
    return loss, ssim, psnr


def build_dataloader(config, rank: int = 0, world: int = 1):
    """reference :28-38: (train_loader, test_loader) over the pickled-dict ``.npy`` files -- or, with ``_synthetic`` = N,
    over N random records of the same shapes.  ``train_batch_size`` is the per-rank batch, as in the reference under DDP."""
    from ..data.srdataset import TactileSRDataset
    from .common import SyntheticSRDataset, make_loader
    n = int(config.get("_synthetic", 0))
    if n > 0:
        train_set = SyntheticSRDataset(n, config["seqsCnt"], seed=config["random_seed"])
        test_set = SyntheticSRDataset(max(config["test_batch_size"] * 2, 16), config["seqsCnt"], seed=config["random_seed"] + 1)
    else:
        train_set, test_set = TactileSRDataset(config["train_dataset_dir"]), TactileSRDataset(config["test_dataset_dir"])
    train_loader = make_loader(train_set, config["train_batch_size"], True, world, rank, seed=config["random_seed"])
    test_loader = make_loader(test_set, config["test_batch_size"], False, world, rank)
    return train_loader, test_loader


def make_trainer(config, model, optimizer, train_loader, device):
    """reference main() :213-228: StepLR + the warm-up arguments exactly as that call passes them (``warmup_by_epoch`` is
    not forwarded, so the 2000-step warm-up runs per iteration)."""
    lr_scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=config["lr_scheduler_step_size"],
                                                   gamma=config["lr_scheduler_gamma"])
    kw = {}
    if config.get("_max_iters", 0) > 0:
        kw["max_iters"] = config["_max_iters"]
    else:
        kw["max_epochs"] = config["epochs"]
    warm = {k: config[k] for k in ("warmup_t", "warmup_mode", "warmup_init_lr", "warmup_factor") if k in config}
    if "max_iters" in kw:
        warm["by_epoch"] = False
    return Trainer_tactileSR(config, model, optimizer, lr_scheduler, train_loader, work_dir=config["save_dir"],
                             checkpoint_period=config["checkpoint_period"], device=device,
                             cuda_graph=bool(config.get("_cuda_graph", False)), **kw, **warm)


def main(config):
    """reference main() :199-242."""
    from .. import set_precision
    from .common import EvalHook, set_random_seed, setup_device, shutdown
    rank, world, device = setup_device()
    set_precision(config.get("_precision", "fp16"))
    set_random_seed(config["random_seed"])           # identical initial weights on every rank
    train_loader, test_loader = build_dataloader(config, rank, world)
    model, optimizer = build_model_and_optimizer(config, device)
    trainer = make_trainer(config, model, optimizer, train_loader, device)
    if trainer.train_by_epoch:
        trainer.register_hooks([EvalHook(1, lambda: eval_func(model, test_loader, config, device))])
    trainer.train(auto_resume=False)
    shutdown()
    return trainer


if __name__ == "__main__":
    from ..config import tactileSR_config
    from .common import parse_cli
    main(parse_cli("TactileSR (single frame) training on the tactilesr_b200 kernels", tactileSR_config))
