"""Hot-path part of reference train/tactileSR_train.py: ``Trainer_tactileSR.train_cal_loss`` (:41-51) and the
model / optimizer construction of ``main`` (:204-213), running on the tactilesr_b200 kernels.  Plotting, PNG
inference hooks and dataset readers (the rest of that file) are outside the hot path."""
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
