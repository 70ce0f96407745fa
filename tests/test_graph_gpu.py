"""CUDA-graph replay of the training iteration (Trainer(cuda_graph=True)): bit-identical to the eager iteration --
same kernels, same order, the step-dependent Adam scalars and the learning rate arrive through device memory."""
import pytest
import torch

from tests.util import sr_inputs

pytestmark = pytest.mark.gpu

CFG = dict(seqsCnt=1, axisCnt=3, HR_scale_num=10, scale_factor=10, patternFeatureExtraLayerCnt=6,
           forceFeatureExtraLayerCnt=1, lr=1e-3, weight_decay=1e-2)


def _run(use_graph, mode, iters=7, B=8):
    import tactilesr_b200 as tb
    from tactilesr_b200.train.tactileSR_train import Trainer_tactileSR, build_model_and_optimizer
    tb.set_precision(mode)
    torch.manual_seed(42)
    dev = torch.device("cuda", 0)
    model, opt = build_model_and_optimizer(CFG, dev)
    loader = [sr_inputs(B, 1, 300 + i) for i in range(iters)]
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)       # lr changes while the graph is live
    tr = Trainer_tactileSR(CFG, model=model, optimizer=opt, lr_scheduler=sched, data_loader=loader, max_epochs=2,
                           log_period=10 ** 9, device=dev, cuda_graph=use_graph, warmup_t=4, warmup_mode="auto",
                           warmup_init_lr=1e-5, warmup_factor=1e-4)
    losses = []
    for it in range(iters):
        tr.cur_iter = it
        tr.train_one_iter()
        tr.lr_scheduler.iter_update()
        if it == 4:
            tr.lr_scheduler.epoch_update()
        losses.append(tr._loss_acc.clone())
        tr._loss_acc, tr._loss_cnt = None, 0
    torch.cuda.synchronize()
    assert (len(tr._graphs) == 1) == use_graph
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    steps = {int(opt.state[p]["step"]) for p in model.parameters()}
    return torch.stack(losses).cpu(), sd, steps, opt.state_dict()


@pytest.mark.parametrize("mode", ["fp16", "fp32"])
def test_graphed_training_is_bit_identical_to_eager(mode):
    import tactilesr_b200 as tb
    try:
        l0, sd0, st0, _ = _run(False, mode)
        l1, sd1, st1, osd = _run(True, mode)
    finally:
        tb.set_precision("fp32")
    assert st0 == st1 == {7}
    assert torch.equal(l0, l1), (l0, l1)
    for k in sd0:
        assert torch.equal(sd0[k], sd1[k]), k
    assert float(osd["state"][0]["step"]) == 7.0


def test_graph_training_interleaved_with_eager_evaluation():
    """Cache-invalidation check: graph-replayed training changes weights and BatchNorm statistics without running any
    Python, while evaluation in between runs eagerly with BatchNorm folded into cached packed weights.  The interleaved
    sequence must produce exactly the evaluation outputs of an all-eager run."""
    import tactilesr_b200 as tb
    from tactilesr_b200.train.tactileSR_train import Trainer_tactileSR, build_model_and_optimizer
    outs = []
    try:
        for use_graph in (False, True):
            tb.set_precision("fp16")
            torch.manual_seed(42)
            dev = torch.device("cuda", 0)
            from oracle import tactilesr_oracle as so
            cfg = dict(CFG, lr=2e-5)
            model, opt = build_model_and_optimizer(cfg, dev)
            model.load_state_dict(so.make_state(so.tactilesr_layout(1), 5), strict=True)   # every ReLU path alive
            loader = [sr_inputs(8, 1, 400 + i) for i in range(8)]
            tr = Trainer_tactileSR(cfg, model=model, optimizer=opt, lr_scheduler=torch.optim.lr_scheduler.StepLR(opt, 1, 0.9),
                                   data_loader=loader, max_iters=100, log_period=10 ** 9, device=dev, cuda_graph=use_graph)
            x = sr_inputs(5, 1, 999)[0].cuda()
            evals = []
            for it in range(8):
                tr.cur_iter = it
                model.train()
                tr.train_one_iter()
                if it in (3, 4, 7):
                    model.eval()
                    with torch.no_grad():
                        evals.append(model(x).clone())
            outs.append(evals)
    finally:
        tb.set_precision("fp32")
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    assert not torch.equal(outs[0][0], outs[0][1])      # the weights did move between evaluations
