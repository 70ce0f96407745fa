"""GPU: data-parallel training step with two ranks (both on cuda:0, gloo transport so that it runs on a 1-GPU box).
Parity is per shard (BatchNorm statistics are rank-local, DDP semantics -- SURVEY section 8e): after the overlapped
bucketed all-reduce every rank must hold the mean of the per-rank gradients, bit-identical across ranks, and equal to
running the two shards one after the other in a single process."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CFG = dict(seqsCnt=1, axisCnt=3, HR_scale_num=10, scale_factor=10, patternFeatureExtraLayerCnt=2,
           forceFeatureExtraLayerCnt=1, lr=1e-3, weight_decay=1e-2)


def _batch(rank, step, B=4):
    g = torch.Generator().manual_seed(1000 + 10 * step + rank)
    return torch.rand(B, 3, 4, 4, generator=g) * 8, torch.rand(B, 1, 100, 100, generator=g) * 250


def _build(dev):
    import tactilesr_b200 as tb
    from tactilesr_b200.train.tactileSR_train import build_model_and_optimizer
    tb.set_precision("fp32")
    torch.manual_seed(7)
    return build_model_and_optimizer(CFG, dev)


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from tactilesr_b200.cpu import distributed as D
    from tactilesr_b200.train.tactileSR_train import Trainer_tactileSR
    torch.cuda.set_device(0)
    D.init_distributed(backend="gloo")
    dev = torch.device("cuda", 0)
    model, opt = _build(dev)
    loader = [_batch(rank, s) for s in range(2)]
    tr = Trainer_tactileSR(CFG, model=model, optimizer=opt, lr_scheduler=torch.optim.lr_scheduler.StepLR(opt, 2, 0.8),
                           data_loader=loader, max_iters=2, log_period=1, device=dev, grad_bucket_bytes=64 << 10)
    tr.train(auto_resume=False)
    assert tr._dp is not None and tr._dp.world == 2
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu()
    q.put((rank, flat, tr.metric_storage["total_loss"].latest if rank == 0 else None))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_training_matches_sequential_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, flat, loss = q.get(timeout=300)
        got[r] = (flat, loss)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(got[0][0], got[1][0]), "ranks diverged"
    # single-process emulation: per-rank gradients (rank-local BN statistics), averaged, one Adam step per iteration
    from tactilesr_b200.functional import mse_hr_loss
    dev = torch.device("cuda", 0)
    model, opt = _build(dev)
    import copy
    for step in range(2):
        grads = []
        bn_state = copy.deepcopy({k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k})
        for rank in range(2):
            model.load_state_dict(bn_state, strict=False)      # each rank starts the step from the same BN buffers
            LR, HR = _batch(rank, step)
            opt.zero_grad()
            mse_hr_loss(model(LR.to(dev)), HR.to(dev), 10.0).backward()
            grads.append([p.grad.detach().clone() for p in model.parameters()])
        opt.zero_grad()
        for p, g0, g1 in zip(model.parameters(), *grads):
            p.grad = (g0 + g1) / 2
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu()
    err = (got[0][0] - ref).abs().max().item()
    assert err < 2e-6, err
    assert got[0][1] is not None and got[0][1] > 0
