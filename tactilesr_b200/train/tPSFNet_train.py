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
