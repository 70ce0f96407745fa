"""Checkpoint interoperability with the unmodified reference.  CPU part (here, needs
/root/reference): state_dicts of every model class load into the reference's classes and back with strict=True, the
Seqs transplant of tactileSRSeqs_train.py:43-59 works on our classes, and the optimizer state layout is stock Adam's.
GPU part: a checkpoint written by our Trainer after real steps is consumed by stock torch.optim.Adam."""
import os
import sys

import pytest
import torch

REF = os.environ.get("TACTILESR_REFERENCE", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model", "tactileSR_model.py")),
                               reason="reference sources not present on this machine")


def _ref_models():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    rm = importlib.import_module("model.tactileSR_model")
    rp = importlib.import_module("model.tPSFNet")
    return rm, rp


@needs_ref
@pytest.mark.parametrize("kind", ["sr1", "sr7", "srcnn", "tpsf"])
def test_state_dicts_load_both_ways(kind):
    rm, rp = _ref_models()
    from tactilesr_b200.model import TactileSR, TactileSRCNN, tPSFNet
    torch.manual_seed(3)
    if kind == "sr1":
        ours, ref = TactileSR(), rm.TactileSR()
    elif kind == "sr7":
        ours, ref = TactileSR(seqsCnt=7), rm.TactileSR(seqsCnt=7)
    elif kind == "srcnn":
        ours, ref = TactileSRCNN(), rm.TactileSRCNN()
    else:
        ours, ref = tPSFNet(1.4, None, device="cpu"), rp.tPSFNet(1.4, None, device="cpu")
    so, sr = ours.state_dict(), ref.state_dict()
    assert list(so.keys()) == list(sr.keys())
    assert [tuple(v.shape) for v in so.values()] == [tuple(v.shape) for v in sr.values()]
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]   # optimizer index <-> param
    ref.load_state_dict(so, strict=True)          # ours -> reference
    for k, v in ref.state_dict().items():
        assert torch.equal(v, so[k]), k
    torch.manual_seed(4)
    ref2 = type(ref)() if kind in ("sr1", "srcnn") else (rm.TactileSR(seqsCnt=7) if kind == "sr7" else rp.tPSFNet(1.4, None, device="cpu"))
    ours.load_state_dict(ref2.state_dict(), strict=True)    # reference -> ours
    for k, v in ours.state_dict().items():
        assert torch.equal(v, ref2.state_dict()[k]), k


@needs_ref
def test_seqs_transplant_matches_reference():
    """tactileSRSeqs_train.py:43-59: load a single-frame checkpoint, then assign its two stacks into the 7-frame model."""
    rm, _ = _ref_models()
    from tactilesr_b200.model import TactileSR
    torch.manual_seed(11)
    single_ref = rm.TactileSR()
    ck = {"model": single_ref.state_dict()}
    torch.manual_seed(12)
    seq_ref = rm.TactileSR(seqsCnt=7)
    torch.manual_seed(12)
    seq_ours = TactileSR(seqsCnt=7)
    for seq, cls in ((seq_ref, rm.TactileSR), (seq_ours, TactileSR)):
        single = cls()
        single.load_state_dict(ck["model"], strict=False)
        seq.patternFeatureExtra_layer = single.patternFeatureExtra_layer
        seq.forceFeatureExtra_layer = single.forceFeatureExtra_layer
    a, b = seq_ours.state_dict(), seq_ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


@needs_ref
def test_optimizer_state_layout_is_stock_adam():
    """A state_dict written by torch.optim.Adam over the reference model loads into FusedAdam over ours (same param
    order) and comes back unchanged: checkpoints move between the two code bases in both directions."""
    rm, _ = _ref_models()
    from tactilesr_b200.model import TactileSR
    from tactilesr_b200.optim import FusedAdam
    torch.manual_seed(1)
    ref = rm.TactileSR()
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-2)
    for p in ref.parameters():
        p.grad = torch.randn_like(p) * 1e-3
    opt_ref.step()
    sd = opt_ref.state_dict()
    ours = TactileSR()
    opt = FusedAdam(ours.parameters(), lr=5e-4, weight_decay=0.0)
    opt.load_state_dict(sd)
    back = opt.state_dict()
    assert back["param_groups"][0]["lr"] == 1e-3 and back["param_groups"][0]["weight_decay"] == 1e-2
    assert set(sd["param_groups"][0].keys()) <= set(back["param_groups"][0].keys())
    assert back["param_groups"][0]["params"] == sd["param_groups"][0]["params"]
    assert sorted(back["state"].keys()) == sorted(sd["state"].keys())
    for i, st in sd["state"].items():
        assert set(st.keys()) == set(back["state"][i].keys()) == {"step", "exp_avg", "exp_avg_sq"}
        assert torch.equal(st["exp_avg"], back["state"][i]["exp_avg"]) and float(st["step"]) == float(back["state"][i]["step"])


@pytest.mark.gpu
def test_trainer_checkpoint_is_consumed_by_stock_adam(tmp_path):
    """Train 3 steps with our trainer, save_checkpoint, then: the file holds the reference's keys (cpu/trainer.py:
    394-431), stock torch.optim.Adam loads the optimizer state over a fresh copy of the parameters and its next step
    equals FusedAdam's next step on the same gradient."""
    import tactilesr_b200 as tb
    from tactilesr_b200.train.tactileSR_train import Trainer_tactileSR, build_model_and_optimizer
    from tests.util import sr_inputs
    cfg = dict(seqsCnt=1, axisCnt=3, HR_scale_num=10, scale_factor=10, patternFeatureExtraLayerCnt=6,
               forceFeatureExtraLayerCnt=1, lr=1e-3, weight_decay=1e-2)
    tb.set_precision("fp32")
    torch.manual_seed(42)
    dev = torch.device("cuda", 0)
    model, opt = build_model_and_optimizer(cfg, dev)
    loader = [sr_inputs(4, 1, 70 + i) for i in range(3)]
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.8)
    tr = Trainer_tactileSR(cfg, model=model, optimizer=opt, lr_scheduler=sched, data_loader=loader, max_iters=3,
                           log_period=10 ** 9, device=dev, work_dir=str(tmp_path))
    for _ in range(3):
        tr.train_one_iter()
    tr.save_checkpoint("iter_2.pth")
    ck = torch.load(os.path.join(tr.ckpt_dir, "iter_2.pth"), map_location="cpu", weights_only=False)
    assert {"num_gpus", "model", "optimizer", "lr_scheduler", "metric_storage"} <= set(ck.keys())
    assert list(ck["model"].keys()) == list(model.state_dict().keys()) and len(ck["model"]) == 205
    # stock Adam over a copy of the parameters, resumed from our optimizer state
    params = [torch.nn.Parameter(p.detach().clone()) for p in model.parameters()]
    stock = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-2)
    stock.load_state_dict(ck["optimizer"])
    g = [torch.randn_like(p) * 1e-3 for p in params]
    for p, q, gi in zip(params, model.parameters(), g):
        p.grad = gi.clone()
        q.grad = gi.clone()
    stock.step()
    opt.step()
    worst = max(((p - q).abs().max() / q.abs().max().clamp_min(1e-12)).item() for p, q in zip(params, model.parameters()))
    assert worst < 5e-6, worst


@pytest.mark.gpu
def test_seqs_training_with_transplanted_stacks(tmp_path):
    """train/tactileSRSeqs_train.py flow on the GPU: single-frame checkpoint -> 7-frame model with transplanted stacks ->
    training steps.  As in the reference the optimizer was built before the transplant, so the transplanted stacks stay
    bit-identical to the checkpoint while the heads / inputContact / output layers move, and the loss goes down."""
    import tactilesr_b200 as tb
    from tactilesr_b200.model import TactileSR
    from tactilesr_b200.train.tactileSRSeqs_train import Trainer_tactileSR, build_model_and_optimizer
    from tests.util import sr_inputs
    tb.set_precision("fp16")
    try:
        dev = torch.device("cuda", 0)
        single_cfg = dict(seqsCnt=1, axisCnt=3, HR_scale_num=10, scale_factor=10, patternFeatureExtraLayerCnt=6,
                          forceFeatureExtraLayerCnt=1, lr=1e-3, weight_decay=1e-2)
        torch.manual_seed(7)
        single = TactileSR().to(dev)
        ck = str(tmp_path / "epoch_50.pth")
        torch.save({"model": single.state_dict()}, ck)
        cfg = dict(single_cfg, seqsCnt=7, lr=1e-4, load_checkpoint_dir=ck)
        torch.manual_seed(8)
        model, opt = build_model_and_optimizer(cfg, single_cfg, dev)
        ref_stack = {k: v.detach().clone() for k, v in model.patternFeatureExtra_layer.state_dict().items()}
        head0 = model.inputContact_layer[0].weight.detach().clone()
        g = torch.Generator().manual_seed(9)
        loader = [(torch.rand(8, 21, 4, 4, generator=g) * 8, sr_inputs(8, 1, 50)[1]) for _ in range(2)] * 3
        tr = Trainer_tactileSR(cfg, model=model, optimizer=opt, lr_scheduler=torch.optim.lr_scheduler.StepLR(opt, 2, 0.8),
                               data_loader=loader, max_iters=100, log_period=10 ** 9, device=dev)
        losses = []
        for it in range(6):
            tr.cur_iter = it
            tr.train_one_iter()
            losses.append(tr._loss_acc.item()); tr._loss_acc, tr._loss_cnt = None, 0
        assert all(torch.isfinite(torch.tensor(losses)))
        for k, v in model.patternFeatureExtra_layer.state_dict().items():
            if "running" in k or "num_batches" in k:
                continue            # BatchNorm statistics of the transplanted stacks do follow the new data
            assert torch.equal(v, ref_stack[k]), k
        assert not torch.equal(model.inputContact_layer[0].weight, head0)
        assert model.patternFeatureExtra_layer[0].conv_3_1[0].weight.grad is not None
    finally:
        tb.set_precision("fp32")
