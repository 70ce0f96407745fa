"""Hot-path part of reference train/tactileSRSeqs_train.py: ``model_param_init`` (:43-59, the transplant of a trained
single-frame model's two feature stacks into the 7-frame model) and the model / optimizer construction of ``main``
(:62-77).  The trainer is the same ``Trainer_tactileSR`` (the reference imports it from train/tactileSR_train.py).

Faithful to the reference's order of operations: the optimizer is built over the sequence model's *own* parameters
BEFORE the transplant (:74 then :77), so the transplanted ``patternFeatureExtra_layer`` / ``forceFeatureExtra_layer``
receive gradients but are never updated (frozen pre-trained feature extractors), while the per-frame heads,
``inputContact_layer`` and ``output_layer`` train.  ``FusedAdam`` keeps that behaviour: parameters that never see a
gradient are skipped exactly like ``torch.optim.Adam`` skips them.
"""
from __future__ import annotations

import torch

from ..model.tactileSR_model import TactileSR
from ..optim import FusedAdam
from .tactileSR_train import Trainer_tactileSR, eval_func  # noqa: F401  (same trainer / evaluation as the reference)


def model_param_init(singleSR_config, seqsSR_config, seqsSR_model, device=None):
    """reference :43-59.  ``seqsSR_config['load_checkpoint_dir']`` is a checkpoint written by the single-frame trainer
    (ours or the reference's: same ``{'model': state_dict}`` layout)."""
    device = device if device is not None else next(seqsSR_model.parameters()).device
    checkpoint = torch.load(seqsSR_config["load_checkpoint_dir"], map_location=device, weights_only=False)
    singleSR_model = TactileSR(scale_factor=singleSR_config["scale_factor"], seqsCnt=singleSR_config["seqsCnt"],
                               axisCnt=singleSR_config["axisCnt"],
                               patternFeatureExtraLayerCnt=singleSR_config["patternFeatureExtraLayerCnt"],
                               forceFeatureExtraLayerCnt=singleSR_config["forceFeatureExtraLayerCnt"]).to(device)
    singleSR_model.load_state_dict(checkpoint["model"], strict=False)
    seqsSR_model.patternFeatureExtra_layer = singleSR_model.patternFeatureExtra_layer
    seqsSR_model.forceFeatureExtra_layer = singleSR_model.forceFeatureExtra_layer
    return seqsSR_model


def build_model_and_optimizer(config, singleSR_config, device):
    """reference main() :66-77: sequence model, Adam over ITS parameters, then the transplant."""
    model = TactileSR(scale_factor=config["scale_factor"], seqsCnt=config["seqsCnt"], axisCnt=config["axisCnt"],
                      patternFeatureExtraLayerCnt=config["patternFeatureExtraLayerCnt"],
                      forceFeatureExtraLayerCnt=config["forceFeatureExtraLayerCnt"]).to(device)
    optimizer = FusedAdam(model.parameters(), lr=config["lr"], weight_decay=config["weight_decay"])
    model = model_param_init(singleSR_config, config, model, device)
    return model, optimizer


def main(config, single_config=None):
    """reference main() :62-98: the optimizer is built BEFORE ``model_param_init`` transplants the pretrained stacks (so they
    stay frozen), no warm-up.  Without a ``load_checkpoint_dir`` file the transplant is skipped (synthetic smoke runs)."""
    import os

    from .. import set_precision
    from ..config import tactileSR_config
    from .common import EvalHook, set_random_seed, setup_device, shutdown
    from .tactileSR_train import build_dataloader, make_trainer
    from .tactileSR_train import build_model_and_optimizer as build_plain
    rank, world, device = setup_device()
    set_precision(config.get("_precision", "fp16"))
    set_random_seed(config["random_seed"])
    train_loader, test_loader = build_dataloader(config, rank, world)
    if os.path.exists(config.get("load_checkpoint_dir", "")):
        model, optimizer = build_model_and_optimizer(config, single_config or tactileSR_config, device)
    else:
        model, optimizer = build_plain(config, device)
    cfg = {k: v for k, v in config.items() if not k.startswith("warmup_")}          # :79-87 passes no warm-up arguments
    trainer = make_trainer(cfg, model, optimizer, train_loader, device)
    if trainer.train_by_epoch:
        trainer.register_hooks([EvalHook(1, lambda: eval_func(model, test_loader, config, device))])
    trainer.train(auto_resume=False)
    shutdown()
    return trainer


if __name__ == "__main__":
    from ..config import tactileSeqs_config
    from .common import parse_cli
    main(parse_cli("TactileSR (7-frame sequence) training on the tactilesr_b200 kernels", tactileSeqs_config))
