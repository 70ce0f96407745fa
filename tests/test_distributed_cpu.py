"""CPU, gloo, world_size 2: the data-parallel plumbing (init_distributed, reduce_dict, gather, bucketed overlapped
GradAllReduce) and the FusedAdam-independent host logic of the trainer (LR warm-up schedule, MetricStorage)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from tactilesr_b200.cpu import distributed as D
    r, lr, w = D.init_distributed(backend="gloo")
    assert (r, w) == (rank, world) and D.get_rank() == rank and D.get_world_size() == world
    assert D.is_main_process() == (rank == 0)
    # reduce_dict: one stacked all-reduce, averaged
    red = D.reduce_dict({"a": torch.tensor(float(rank + 1)), "b": torch.tensor(10.0 * rank)})
    assert abs(float(red["a"]) - 1.5) < 1e-6 and abs(float(red["b"]) - 5.0) < 1e-6
    # gather / all_gather of python objects
    got = D.gather({"rank": rank})
    assert (len(got) == world and [g["rank"] for g in got] == [0, 1]) if rank == 0 else got == []
    assert [g for g in D.all_gather(rank * 7)] == [0, 7]
    # bucketed gradient averaging: ranges become ready back-to-front like a backward pass
    torch.manual_seed(100 + rank)
    flat = torch.randn(10_000)
    mine = flat.clone()
    other = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(other, mine)
    want = sum(other) / world
    # both forms of the gradient average: bucketed while "backward" reports ranges, and one call after backward (default)
    for overlap in (True, False):
        flat.copy_(mine)
        ar = D.GradAllReduce(flat, bucket_bytes=4 * 3000, overlap=overlap)
        edges = [10_000, 9_000, 6_500, 6_400, 3_000, 1_234, 0]
        for hi, lo in zip(edges[:-1], edges[1:]):
            ar.ready(lo, hi)
        ar.finish()
        assert torch.allclose(flat, want, atol=1e-6), (overlap, (flat - want).abs().max())
    out.put((rank, float(flat.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_distributed_helpers():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert abs(res[0] - res[1]) < 1e-3      # both ranks hold the same averaged gradient


def test_not_distributed_defaults():
    from tactilesr_b200.cpu import distributed as D
    for k in ("RANK", "WORLD_SIZE", "SLURM_PROCID"):
        os.environ.pop(k, None)
    assert D.init_distributed() == (0, 0, 1)
    assert D.get_world_size() == 1 and D.get_rank() == 0 and D.is_main_process()
    d = {"x": torch.tensor(3.0)}
    assert D.reduce_dict(d) is d and D.all_gather(5) == [5] and D.gather(5) == [5]


def test_lr_warmup_matches_reference_schedule():
    """TactileSR's schedule (train/tactileSR_train.py:224-227): iteration warm-up over 2000 iters, mode 'auto',
    factor 1e-4, on top of StepLR(2, 0.8) stepped per epoch."""
    from tactilesr_b200.cpu.trainer import LRWarmupScheduler
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.8)
    epoch_len = 500
    w = LRWarmupScheduler(sched, by_epoch=True, epoch_len=epoch_len, warmup_t=2000, warmup_by_epoch=False,
                          warmup_mode="auto", warmup_init_lr=1e-5, warmup_factor=1e-4)
    assert abs(opt.param_groups[0]["lr"] - 1e-7) < 1e-15          # base_lr * factor
    end = 1e-3 * 0.8 ** 2                                          # StepLR value after 4 epochs
    lrs = []
    for it in range(1, 2501):
        w.iter_update()
        if it % epoch_len == 0:
            w.epoch_update()
        lrs.append(opt.param_groups[0]["lr"])
    a = 1000 / 2000
    assert abs(lrs[999] - (1e-7 * (1 - a) + end * a)) < 1e-12      # linear from base*factor to the post-warm-up value
    assert abs(lrs[1999] - end) < 1e-12
    assert all(x <= y + 1e-15 for x, y in zip(lrs[:1999], lrs[1:2000]))


def test_metric_storage_window_and_average():
    from tactilesr_b200.cpu.trainer import MetricStorage
    m = MetricStorage(window_size=3)
    for i, v in enumerate([1.0, 2.0, 3.0, 4.0]):
        m.update(i, total_loss=v)
    assert m.values_maybe_smooth["total_loss"] == (3, 3.0)      # (latest iteration, value): the reference order
    # records are HistoryBuffer-like (reference cpu/history_buffer.py): what LoggerHook / EvalHook read
    assert m["total_loss"].global_avg == 2.5 and m["total_loss"].latest == 4.0 and m["total_loss"].avg == 3.0
    assert m["total_loss"].global_sum == 10.0
    m.update(7, lr=0.5, smooth=False)
    assert m.values_maybe_smooth["lr"] == (7, 0.5)
    with pytest.raises(AssertionError):
        m.update(7, lr=0.4, smooth=False)                        # iterations must increase, as in the reference


@pytest.mark.skipif(not os.path.exists("/root/reference/cpu/history_buffer.py"), reason="reference tree not mounted")
def test_metric_storage_matches_reference_history_buffer():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_history_buffer", "/root/reference/cpu/history_buffer.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from tactilesr_b200.cpu.trainer import HistoryBuffer
    a, b = HistoryBuffer(5), ref.HistoryBuffer(5)
    for v in [0.1, 0.5, 2.0, 3.0, -1.0, 4.0, 0.25]:
        a.update(v); b.update(v)
        assert (a.latest, a.avg, a.global_avg, a.global_sum) == (b.latest, float(b.avg), b.global_avg, b.global_sum)


class _TinyModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(3))

    def forward(self, batch):
        return ((self.w * batch[0]).sum() - 1.0) ** 2


def _tiny_trainer(tmp_path, max_epochs=3, **kw):
    from tactilesr_b200.cpu.trainer import Trainer
    torch.manual_seed(0)
    data = [(torch.rand(3),) for _ in range(4)]
    model = _TinyModel()
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    sched = torch.optim.lr_scheduler.StepLR(opt, 1, 0.9)
    return Trainer(model, opt, sched, data, max_epochs=max_epochs, work_dir=str(tmp_path), log_period=2, **kw)


def test_train_writes_checkpoints_logs_and_resumes(tmp_path):
    """Default hooks as in the reference (cpu/trainer.py:194-200): checkpoints every ``checkpoint_period`` epochs with
    pruning, latest.pth resumes (iteration counter, optimizer, scheduler, metric storage)."""
    from tactilesr_b200.cpu.trainer import CheckpointHook, DistributedHook, LoggerHook
    tr = _tiny_trainer(tmp_path, max_epochs=3, checkpoint_period=1, max_num_checkpoints=2)
    names = [type(h).__name__ for h in tr._hooks]
    assert {"_LRUpdateHook", "DistributedHook", "CheckpointHook", "LoggerHook"} <= set(names) and names[-1] == "LoggerHook"
    tr.train(auto_resume=False)
    ck = sorted(os.listdir(tr.ckpt_dir))
    assert ck == ["epoch_1.pth", "epoch_2.pth", "latest.pth"], ck          # epoch_0 pruned (max_num_checkpoints = 2)
    assert tr.metric_storage["total_loss"].latest >= 0.0
    w_end = tr.model.w.detach().clone()
    tr2 = _tiny_trainer(tmp_path, max_epochs=5, checkpoint_period=1, max_num_checkpoints=2)
    tr2.load_checkpoint(auto_resume=True)
    assert tr2.start_iter == 3 * 4 and torch.equal(tr2.model.w.detach(), w_end)
    assert "total_loss" in tr2.metric_storage and tr2.metric_storage["total_loss"].latest == tr.metric_storage["total_loss"].latest
    tr2.metric_storage.update(tr2.start_iter, total_loss=0.0)              # a resumed storage keeps working
    tr2.train()
    assert tr2.cur_iter == 5 * 4 - 1


def test_foreign_hooks_are_accepted(tmp_path):
    """A hook that does not derive from our HookBase (the reference's EvalHook does not) only needs the six stage methods."""
    calls = []

    class Foreign:
        priority = 1

        def __getattr__(self, name):
            if name in ("before_train", "after_train", "before_epoch", "after_epoch", "before_iter", "after_iter"):
                return lambda: calls.append(name)
            raise AttributeError(name)

    tr = _tiny_trainer(tmp_path, max_epochs=1)
    tr.register_hooks([Foreign()])
    assert type(tr._hooks[0]).__name__ == "Foreign"                         # priority 1 sorts first
    tr.train(auto_resume=False)
    assert calls.count("after_iter") == 4 and calls[0] == "before_train" and calls[-1] == "after_train"


def test_distributed_hook_sets_sampler_epoch(tmp_path):
    from torch.utils.data import DataLoader, DistributedSampler, TensorDataset
    from tactilesr_b200.cpu.trainer import Trainer
    ds = TensorDataset(torch.rand(8, 3))
    sampler = DistributedSampler(ds, num_replicas=2, rank=0, shuffle=True)
    dl = DataLoader(ds, batch_size=2, sampler=sampler)
    model = _TinyModel()
    opt = torch.optim.SGD(model.parameters(), lr=0.01)
    tr = Trainer(model, opt, torch.optim.lr_scheduler.StepLR(opt, 1, 0.9), dl, max_epochs=3, work_dir=str(tmp_path))
    seen = []
    orig = sampler.set_epoch
    sampler.set_epoch = lambda e: (seen.append(e), orig(e))
    tr.train(auto_resume=False)
    assert seen == [0, 1, 2]


@pytest.mark.skipif(not os.path.exists("/root/reference/cpu/lr_scheduler.py"), reason="reference tree not mounted")
@pytest.mark.parametrize("mode,by_epoch,wbe", [("auto", True, False), ("factor", True, False), ("fix", False, False), ("fix", True, True)])
def test_lr_warmup_equals_reference_implementation(mode, by_epoch, wbe):
    """Step our LRWarmupScheduler and the reference's (loaded straight from its file) side by side."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_lr_scheduler", "/root/reference/cpu/lr_scheduler.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from tactilesr_b200.cpu.trainer import LRWarmupScheduler

    def make(cls):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([p], lr=1e-3)
        sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.8)
        return opt, cls(sched, by_epoch, 50 if by_epoch else None, 3 if wbe else 120, wbe, mode, 1e-5, 1e-4)
    (oa, a), (ob, b) = make(LRWarmupScheduler), make(ref.LRWarmupScheduler)
    for it in range(1, 401):
        a.iter_update(); b.iter_update()
        if by_epoch and it % 50 == 0:
            a.epoch_update(); b.epoch_update()
        assert abs(oa.param_groups[0]["lr"] - ob.param_groups[0]["lr"]) < 1e-15, (it, oa.param_groups[0]["lr"], ob.param_groups[0]["lr"])
