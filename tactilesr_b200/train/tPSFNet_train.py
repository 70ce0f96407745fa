"""Hot-path part of reference train/tPSFNet_train.py: ``Trainer_tPSF.train_cal_loss`` (:173-190), ``eval_func``'s loss
(:59-75) and the model / optimizer construction of ``main`` (:193-201), running on the tactilesr_b200 kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..cpu.trainer import Trainer
from ..model.tPSFNet import tPSFNet
from ..optim import FusedAdam


class Trainer_tPSF(Trainer):
    def __init__(self, scale_num, *args, device=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.criterion = nn.MSELoss()
        self.scale_num = scale_num
        self.device = device if device is not None else next(self.model_or_module.parameters()).device

    def train_cal_loss(self, batch):
        LR, depth = batch
        LR = LR.to(self.device, non_blocking=True).float() / self.scale_num
        depth = depth.to(self.device, non_blocking=True).float().unsqueeze(1)
        HR_tactile, LR_tactile_degrade, ret_psf, ret_alphaBeta = self.model(LR, depth)
        # only the z-axis frame is supervised and only LR_degrade carries gradient (reference :186-189): this is the
        # case the tcgen05 PSF backward handles
        loss = self.criterion(LR[:, 2:3], LR_tactile_degrade)
        return loss, {"total_loss": loss}


def build_model_and_optimizer(config, device):
    """reference main() :193-201: tPSFNet(gama, perception_scale) on the device + Adam(lr, weight_decay)."""
    model = tPSFNet(gama=config["gama"], perception_scale=config["perception_scale"], device=device).to(device)
    optimizer = FusedAdam(model.parameters(), lr=config["lr"], weight_decay=config["weight_decay"])
    return model, optimizer


@torch.no_grad()
def eval_func(model, test_loader, config, device=None):
    """reference eval_func (:51-69): per batch, MSE and whole-image SSIM (utility/tools.py:65-81) between the FIRST sample's
    z-axis taxel frame and its LR_degrade (4x4 each), averaged over the batches.  The batches run through the model on
    the GPU; the first-sample frames are collected on the device and read back once.  Returns (mse_ave, ssim_ave)."""
    device = device if device is not None else next(model.parameters()).device
    zs, ds = [], []
    model.eval()
    for LR, depth in test_loader:
        LR = LR.to(device, non_blocking=True).float() / config["scale_num"]
        depth = depth.to(device, non_blocking=True).float().unsqueeze(1)
        _, LRd, _, _ = model(LR, depth)
        zs.append(LR[0, 2])
        ds.append(LRd[0, 0])
    if not zs:
        return 0.0, 0.0
    z = torch.stack(zs).double().cpu().numpy()          # (n, 4, 4)
    d = torch.stack(ds).double().cpu().numpy()
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mse = ((d - z) ** 2).mean(axis=(1, 2))
    mu1, mu2 = d.mean(axis=(1, 2)), z.mean(axis=(1, 2))
    s1 = (d * d).mean(axis=(1, 2)) - mu1 * mu1
    s2 = (z * z).mean(axis=(1, 2)) - mu2 * mu2
    s12 = (d * z).mean(axis=(1, 2)) - mu1 * mu2
    ssim = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return float(mse.mean()), float(ssim.mean())


def build_dataloader(config, rank: int = 0, world: int = 1):
    """reference :30-48 (train / test loaders; the two single-tap inference loaders belong to the PNG hook).  The raw-data
    reader (utility/raw_data_process.py) is outside the hot path, so real data comes as a dataset object in
    ``config['train_dataset']`` / ``config['test_dataset']`` yielding (LR raw (3,4,4), depth (100,100)); ``_synthetic`` = N
    builds N random records of those shapes."""
    from .common import SyntheticPSFDataset, make_loader
    n = int(config.get("_synthetic", 0))
    if n > 0:
        train_set = SyntheticPSFDataset(n, seed=config["random_seed"])
        test_set = SyntheticPSFDataset(max(config["test_batch_size"] * 2, 16), seed=config["random_seed"] + 1)
    elif "train_dataset" in config:
        train_set, test_set = config["train_dataset"], config["test_dataset"]
    else:
        raise FileNotFoundError("tPSFNet training data: pass dataset objects in config['train_dataset'] / ['test_dataset'] "
                                "or use --synthetic N (the reference's raw-data reader is not part of this package)")
    return (make_loader(train_set, config["train_batch_size"], True, world, rank, seed=config["random_seed"]),
            make_loader(test_set, config["test_batch_size"], False, world, rank))


def main(config):
    """reference main() :193-229."""
    from .. import set_precision
    from .common import EvalHook, set_random_seed, setup_device, shutdown
    rank, world, device = setup_device()
    set_precision(config.get("_precision", "fp32"))
    set_random_seed(config["random_seed"])
    train_loader, test_loader = build_dataloader(config, rank, world)
    model, optimizer = build_model_and_optimizer(config, device)
    lr_scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=config["lr_scheduler_step_size"],
                                                   gamma=config["lr_scheduler_gamma"])
    kw = {"max_iters": config["_max_iters"], "by_epoch": False} if config.get("_max_iters", 0) > 0 else {"max_epochs": config["epochs"]}
    trainer = Trainer_tPSF(config["scale_num"], model, optimizer, lr_scheduler, train_loader, work_dir=config["save_dir"],
                           checkpoint_period=config["checkpoint_period"], device=device, **kw)
    if trainer.train_by_epoch:
        trainer.register_hooks([EvalHook(1, lambda: eval_func(model, test_loader, config, device), names=("test_mse", "test_ssim"))])
    trainer.train(auto_resume=False)
    shutdown()
    return trainer


if __name__ == "__main__":
    from ..config import tPSFNet_config
    from .common import parse_cli
    main(parse_cli("tPSFNet training on the tactilesr_b200 kernels", tPSFNet_config))
