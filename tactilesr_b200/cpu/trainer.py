"""Training loop of the hot path: the in-scope part of reference cpu/trainer.py (``Trainer.__init__`` :82-142,
``train_one_iter`` :319-364, ``_log_iter_metrics`` :251-288, ``train`` :366-392, checkpoint save/load :394-498,
``MetricStorage`` :501-567), rebuilt so that the per-iteration path has no device->host synchronisation:

  * the loss of every iteration is accumulated on the device and read back (one ``.item()``, plus one 1-float
    all-reduce under data parallelism) only every ``log_period`` iterations -- the reference does
    ``loss.detach().cpu().item()`` and a gloo ``gather_object`` every iteration (:259, :262);
  * under data parallelism the flat gradient buffer of ``FusedAdam`` is averaged over NCCL in buckets while backward is
    still running (``cpu.distributed.GradAllReduce``), instead of wrapping the model in DistributedDataParallel;
  * AMP (``enable_amp``) is not offered: reduced precision is a property of the kernels (``set_precision('bf16')``).

The checkpoint dictionary keeps the reference's keys (``model``, ``optimizer``, ``lr_scheduler``, ``metric_storage``,
``epoch`` | ``iter``, ``hooks``, ``num_gpus``) so checkpoints interoperate.
"""
from __future__ import annotations

import logging
import os
import time
from collections import deque
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch.nn.utils import clip_grad_norm_

from . import distributed as D

logger = logging.getLogger(__name__)


class HookBase:
    """Callback protocol of the reference (cpu/hooks/hookbase.py): six stages, priority 1 (first) .. 10 (last)."""
    priority = 5
    trainer: "Trainer" = None

    def before_train(self): pass
    def after_train(self): pass
    def before_epoch(self): pass
    def after_epoch(self): pass
    def before_iter(self): pass
    def after_iter(self): pass

    @property
    def checkpointable(self) -> bool:
        return callable(getattr(self, "state_dict", None))

    @property
    def class_name(self) -> str:
        return self.__class__.__name__

    def every_n_epochs(self, n: int) -> bool:
        return (self.trainer.cur_epoch + 1) % n == 0 if n > 0 else False

    def every_n_iters(self, n: int) -> bool:
        return (self.trainer.cur_iter + 1) % n == 0 if n > 0 else False

    def every_n_inner_iters(self, n: int) -> bool:
        return (self.trainer.inner_iter + 1) % n == 0 if n > 0 else False

    def is_last_epoch(self) -> bool:
        return self.trainer.cur_epoch == self.trainer.max_epochs - 1

    def is_last_iter(self) -> bool:
        return self.trainer.cur_iter == self.trainer.max_iters - 1


class LRWarmupScheduler:
    """Warm-up wrapper around a torch scheduler with the reference's semantics (cpu/lr_scheduler.py:40-182):
    modes fix / factor / auto, warm-up counted in iterations or epochs, the wrapped scheduler stepped per epoch
    (``by_epoch``) or per iteration once warm-up is over.  It only mutates ``param_group['lr']`` (host scalars)."""

    def __init__(self, torch_scheduler, by_epoch=True, epoch_len=None, warmup_t=0, warmup_by_epoch=False,
                 warmup_mode="fix", warmup_init_lr=None, warmup_factor=None):
        self.torch_scheduler, self.by_epoch, self.epoch_len = torch_scheduler, by_epoch, epoch_len
        self.warmup_t, self.warmup_by_epoch, self.warmup_mode = warmup_t, warmup_by_epoch, warmup_mode
        self.warmup_init_lr, self.warmup_factor = warmup_init_lr, warmup_factor
        if warmup_by_epoch:
            assert by_epoch
        if by_epoch and warmup_t and not warmup_by_epoch:
            assert epoch_len is not None
        self.param_groups = torch_scheduler.optimizer.param_groups
        self.base_lrs = [g["lr"] for g in self.param_groups]
        self.last_iter = self.last_epoch = 0
        self.in_iter_warmup = False
        if warmup_t:
            horizon = warmup_t // epoch_len if (by_epoch and not warmup_by_epoch) else warmup_t
            table = [list(self.base_lrs)]
            for _ in range(horizon):          # the lr the wrapped scheduler would give with no warm-up
                torch_scheduler.step()
                table.append([g["lr"] for g in self.param_groups])
            self.regular_lrs_per_t = table
            if warmup_mode == "fix":
                assert isinstance(warmup_init_lr, float)
                self._set(warmup_init_lr)
            elif warmup_mode in ("factor", "auto"):
                assert isinstance(warmup_factor, float)
                if warmup_mode == "auto":
                    self.warmup_end_lrs = table[-1]
                self._set([b * warmup_factor for b in self.base_lrs])
            else:
                raise ValueError(f"Invalid warmup mode: {warmup_mode}")

    def _set(self, lrs) -> None:
        if not isinstance(lrs, (list, tuple)):
            lrs = [lrs] * len(self.param_groups)
        for g, lr in zip(self.param_groups, lrs):
            g["lr"] = lr

    def _warm(self, t: int, regular: List[float]) -> List[float]:
        a = t / self.warmup_t
        if self.warmup_mode == "fix":
            return [self.warmup_init_lr * (1 - a) + b * a for b in self.base_lrs]
        if self.warmup_mode == "factor":
            f = self.warmup_factor * (1 - a) + a
            return [lr * f for lr in regular]
        return [b * self.warmup_factor * (1 - a) + e * a for b, e in zip(self.base_lrs, self.warmup_end_lrs)]

    def epoch_update(self, metric=None) -> None:
        if not self.by_epoch:
            return
        self.last_epoch += 1
        if self.warmup_by_epoch and self.last_epoch < self.warmup_t:
            self._set(self._warm(self.last_epoch, self.regular_lrs_per_t[self.last_epoch]))
        elif self.warmup_by_epoch and self.last_epoch == self.warmup_t:
            self._set(self.regular_lrs_per_t[-1])
        elif not self.in_iter_warmup:
            self.torch_scheduler.step()

    def iter_update(self) -> None:
        if self.warmup_by_epoch:
            return
        self.last_iter += 1
        if self.last_iter < self.warmup_t:
            self.in_iter_warmup = True
            t = self.last_iter // self.epoch_len if self.by_epoch else self.last_iter
            self._set(self._warm(self.last_iter, self.regular_lrs_per_t[t]))
        elif self.last_iter == self.warmup_t:
            self._set(self.regular_lrs_per_t[-1])
        else:
            self.in_iter_warmup = False
            if not self.by_epoch:
                self.torch_scheduler.step()

    def state_dict(self):
        st = {k: v for k, v in self.__dict__.items() if k not in ("torch_scheduler", "param_groups")}
        st["torch_scheduler"] = self.torch_scheduler.state_dict()
        return st

    def load_state_dict(self, state):
        state = dict(state)
        self.torch_scheduler.load_state_dict(state.pop("torch_scheduler"))
        state.pop("param_groups", None)
        self.__dict__.update(state)


class _LRUpdateHook(HookBase):
    priority = 2

    def after_epoch(self):
        self.trainer.lr_scheduler.epoch_update()

    def after_iter(self):
        self.trainer.lr_scheduler.iter_update()


class HistoryBuffer:
    """Window + global statistics of one metric, same surface as reference cpu/history_buffer.py (``update``, ``latest``,
    ``avg``, ``global_avg``, ``global_sum``), so the reference's own LoggerHook / EvalHook read our storage unchanged."""

    def __init__(self, window_size: int = 20) -> None:
        self._history = deque(maxlen=window_size)
        self._count = 0
        self._sum = 0.0

    def update(self, value: float) -> None:
        self._history.append(value)
        self._count += 1
        self._sum += value

    @property
    def latest(self) -> float:
        return self._history[-1]

    @property
    def avg(self) -> float:
        return float(np.mean(self._history))

    @property
    def global_avg(self) -> float:
        return self._sum / self._count

    @property
    def global_sum(self) -> float:
        return self._sum


class MetricStorage(dict):
    """name -> HistoryBuffer, reference cpu/trainer.py:501-567: ``update(iter, smooth, **values)`` and
    ``values_maybe_smooth`` = {name: (latest iteration, window average | latest value)}."""

    def __init__(self, window_size: int = 20) -> None:
        super().__init__()
        self._window_size = window_size
        self._history = self
        self._smooth: Dict[str, bool] = {}
        self._latest_iter: Dict[str, int] = {}

    def update(self, iter: Optional[int] = None, smooth: bool = True, **kwargs) -> None:
        for key, value in kwargs.items():
            if key in self._smooth:
                assert self._smooth[key] == smooth, f"metric {key!r} changed its smoothing"
            else:
                self._smooth[key] = smooth
                self[key] = HistoryBuffer(window_size=self._window_size)
                self._latest_iter[key] = -1
            if iter is not None:
                assert iter > self._latest_iter[key], f"metric {key!r}: iteration {iter} not after {self._latest_iter[key]}"
                self._latest_iter[key] = iter
            else:
                self._latest_iter[key] += 1
            self[key].update(value)

    @property
    def values_maybe_smooth(self) -> Dict[str, Tuple[int, float]]:
        return {k: (self._latest_iter[k], b.avg if self._smooth[k] else b.latest) for k, b in self.items()}


class CheckpointHook(HookBase):
    """Periodic checkpoints with pruning (reference cpu/hooks/checkpoint_hook.py): every ``period`` epochs / iterations and
    at the end, keeping the ``max_to_keep`` most recent files."""

    def __init__(self, period: int, max_to_keep: Optional[int] = None) -> None:
        assert max_to_keep is None or max_to_keep > 0
        self._period, self._max_to_keep = period, max_to_keep
        self._recent_checkpoints: List[str] = []

    def _save(self, name: str) -> None:
        self.trainer.save_checkpoint(name)
        if self._max_to_keep is not None:
            self._recent_checkpoints.append(name)
            while len(self._recent_checkpoints) > self._max_to_keep:
                old = os.path.join(self.trainer.ckpt_dir, self._recent_checkpoints.pop(0))
                if os.path.exists(old):
                    os.remove(old)

    def after_iter(self) -> None:
        if not self.trainer.train_by_epoch and (self.every_n_iters(self._period) or self.is_last_iter()):
            self._save(f"iter_{self.trainer.cur_iter}.pth")

    def after_epoch(self) -> None:
        if self.trainer.train_by_epoch and (self.every_n_epochs(self._period) or self.is_last_epoch()):
            self._save(f"epoch_{self.trainer.cur_epoch}.pth")

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "trainer"}

    def load_state_dict(self, state) -> None:
        self.__dict__.update(state)


class DistributedHook(HookBase):
    """``DistributedSampler.set_epoch`` before every epoch (reference cpu/hooks/distributed_hook.py:7-13)."""

    def before_epoch(self) -> None:
        dl = self.trainer.data_loader
        for owner in (getattr(dl, "sampler", None), getattr(getattr(dl, "batch_sampler", None), "sampler", None)):
            if hasattr(owner, "set_epoch"):
                owner.set_epoch(self.trainer.cur_epoch)
                return


class LoggerHook(HookBase):
    """Console (and, when tensorboard is importable and ``tb_log_dir`` is given, TensorBoard) logging of the metric storage
    every ``period`` iterations (reference cpu/hooks/logger_hook.py); lowest priority so it sees what other hooks logged."""
    priority = 10

    def __init__(self, period: int = 50, tb_log_dir: Optional[str] = None) -> None:
        self._period = period
        self._tb = None
        self._last_write: Dict[str, int] = {}
        if tb_log_dir is not None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self._tb = SummaryWriter(tb_log_dir)
            except Exception:      # tensorboard absent: console only
                self._tb = None

    def before_train(self) -> None:
        self._t0 = time.perf_counter()

    def _write(self) -> None:
        ms = self.trainer.metric_storage
        tr = self.trainer
        where = (f"Epoch: [{tr.cur_epoch}][{tr.inner_iter}/{tr.epoch_len - 1}]" if tr.train_by_epoch
                 else f"Iter: [{tr.cur_iter}/{tr.max_iters - 1}]")
        parts = [where]
        if "lr" in ms:
            parts.append(f"lr: {ms['lr'].latest:.4g}")
        parts += [f"{k}: {b.avg:.4g}" for k, b in ms.items() if "loss" in k]
        if "iter_time" in ms:
            parts.append(f"iter_time: {ms['iter_time'].avg:.4f}")
        logger.info("  ".join(parts))
        if self._tb is not None:
            for k, (it, v) in ms.values_maybe_smooth.items():
                if it > self._last_write.get(k, -1):
                    self._tb.add_scalar(k, v, it)
                    self._last_write[k] = it

    def after_iter(self) -> None:
        if self.every_n_iters(self._period) or self.is_last_iter():
            self._write()

    def after_train(self) -> None:
        if self._tb is not None:
            self._tb.close()
        logger.info("Total training time: %.1f s", time.perf_counter() - self._t0)


class Trainer:
    def __init__(self, model: nn.Module, optimizer, lr_scheduler, data_loader, unpack_batch_dict: bool = False,
                 max_epochs: int = 0, max_iters: int = 0, work_dir: str = "work_dir",
                 max_num_checkpoints: Optional[int] = None, checkpoint_period: int = 1, log_period: int = 50,
                 clip_grad_norm: float = 0.0, enable_amp: bool = False, by_epoch: bool = True, warmup_t: int = 0,
                 warmup_by_epoch: bool = False, warmup_mode: str = "fix", warmup_init_lr: float = 0.0,
                 warmup_factor: float = 0.0, grad_bucket_bytes: int = 8 << 20, cuda_graph: bool = False,
                 grad_overlap: Optional[bool] = None):
        if enable_amp:
            raise NotImplementedError("enable_amp: use tactilesr_b200.set_precision('bf16') instead of autocast")
        model.train()
        assert (max_epochs > 0) ^ (max_iters > 0), "Please specify either max_epochs or max_iters."
        self.train_by_epoch = max_epochs > 0
        self.model, self.optimizer, self.data_loader = model, optimizer, data_loader
        epoch_len = len(data_loader) if self.train_by_epoch else None
        self.lr_scheduler = LRWarmupScheduler(lr_scheduler, by_epoch, epoch_len, warmup_t, warmup_by_epoch,
                                              warmup_mode, warmup_init_lr, warmup_factor)
        self.unpack_batch_dict, self.work_dir = unpack_batch_dict, work_dir
        self.metric_storage = MetricStorage()
        if self.train_by_epoch:
            self.epoch_len, self.max_epochs = len(data_loader), max_epochs
            self.max_iters = max_epochs * self.epoch_len
        else:
            self.max_iters = max_iters
        self.cur_iter = self.start_iter = 0
        self._hooks: List[HookBase] = []
        self._data_iter = iter(data_loader)
        self._max_num_checkpoints, self._checkpoint_period = max_num_checkpoints, checkpoint_period
        self._log_period, self._clip_grad_norm = log_period, clip_grad_norm
        self._loss_acc: Optional[torch.Tensor] = None      # device-side running sum of the loss
        self._loss_cnt = 0
        self._time_acc = {"data_time": 0.0, "iter_time": 0.0}
        self._dp: Optional[D.GradAllReduce] = None
        self._grad_bucket_bytes = grad_bucket_bytes
        self._grad_overlap = grad_overlap      # None: cpu.distributed.GradAllReduce's default (one call after backward)
        # cuda_graph (extension; the reference has no such switch): after two ordinary iterations the whole iteration
        # (train_cal_loss + backward + optimizer step, ~450 kernel launches) is captured once per batch shape and
        # replayed -- at the reference's batch size 32 the iteration is launch-bound, not GPU-bound.  Under data parallelism
        # the NCCL all-reduces are part of the captured graph.
        self._use_graph = cuda_graph
        self._graphs: Dict[tuple, tuple] = {}
        self._eager_iters = 0
        # default hooks as the reference registers them (cpu/trainer.py:194-200): LR update and sampler epoch on every rank,
        # checkpoints and logging on the main process
        hooks: List[HookBase] = [_LRUpdateHook(), DistributedHook()]
        if D.is_main_process():
            hooks += [CheckpointHook(checkpoint_period, max_num_checkpoints), LoggerHook(log_period)]
        self.register_hooks(hooks)

    # -- bookkeeping identical in meaning to the reference ------------------------------------------
    @property
    def lr(self) -> float:
        return self.optimizer.param_groups[0]["lr"]

    @property
    def model_or_module(self) -> nn.Module:
        m = self.model
        return m.module if isinstance(m, (nn.parallel.DistributedDataParallel, nn.DataParallel)) else m

    @property
    def cur_epoch(self) -> int:
        assert self.train_by_epoch
        return self.cur_iter // self.epoch_len

    @property
    def inner_iter(self) -> int:
        assert self.train_by_epoch
        return self.cur_iter % self.epoch_len

    @property
    def hook_info(self) -> List[str]:
        return [f"{type(h).__name__} (priority {getattr(h, 'priority', 5)})" for h in self._hooks]

    def log(self, *args, **kwargs) -> None:
        self.metric_storage.update(*args, **kwargs)

    def register_hooks(self, hooks: List[Optional[HookBase]]) -> None:
        for h in hooks:
            if h is None:
                continue
            # duck-typed: the reference's own hooks (cpu/hooks/*.py, e.g. EvalHook at train/tactileSR_train.py:230) derive
            # from ITS HookBase, not ours
            assert all(callable(getattr(h, st, None)) for st in ("before_train", "after_train", "before_epoch", "after_epoch",
                                                                 "before_iter", "after_iter")), f"{h!r} is not a hook"
            assert 1 <= getattr(h, "priority", 5) <= 10
            h.trainer = self
            pos = len(self._hooks)
            while pos > 0 and getattr(self._hooks[pos - 1], "priority", 5) > getattr(h, "priority", 5):
                pos -= 1
            self._hooks.insert(pos, h)

    def _call_hooks(self, stage: str) -> None:
        for h in self._hooks:
            getattr(h, stage)()

    # -- data parallel -------------------------------------------------------------------------------
    def _setup_dp(self) -> None:
        """Attach the overlapped gradient all-reduce to the model's layer program (world size > 1 only)."""
        if D.get_world_size() < 2 or self._dp is not None or not hasattr(self.optimizer, "flat_grad"):
            return
        flat = self.optimizer.flat_grad(0)
        # as DistributedDataParallel does at construction: every rank starts from rank 0's parameters and buffers
        # (BatchNorm running statistics stay rank-local afterwards: DDP semantics)
        import torch.distributed as dist
        with torch.no_grad():
            for f in self.optimizer._flat.values():
                if f:
                    dist.broadcast(f["p"], 0)
            flat_ids = {id(p) for f in self.optimizer._flat.values() if f for p in f["params"]}
            for p in self.model_or_module.parameters():
                if id(p) not in flat_ids:
                    dist.broadcast(p.data, 0)
            for b in self.model_or_module.buffers():
                dist.broadcast(b, 0)
        dp = D.GradAllReduce(flat, self._grad_bucket_bytes, overlap=self._grad_overlap)
        base = flat.data_ptr()
        end = base + flat.numel() * 4

        def hook(i, op, c):
            for p in op.params():
                g = c.param_grads.get(p)
                if g is not None and base <= g.data_ptr() < end:
                    lo = (g.data_ptr() - base) // 4
                    dp.ready(lo, lo + g.numel())

        self.model_or_module._engine_extra = {"grad_hook": hook}
        self._dp = dp

    # -- the hot loop ---------------------------------------------------------------------------------
    def train_cal_loss(self, batch):
        loss_dict = self.model(**batch) if self.unpack_batch_dict else self.model(batch)
        if isinstance(loss_dict, torch.Tensor):
            return loss_dict, {"total_loss": loss_dict}
        return sum(loss_dict.values()), loss_dict

    def train_one_iter(self) -> None:
        t0 = time.perf_counter()
        try:
            batch = next(self._data_iter)
        except StopIteration:
            self._data_iter = iter(self.data_loader)
            batch = next(self._data_iter)
        data_time = time.perf_counter() - t0

        if self._graph_eligible(batch):
            loss_dict = self._graphed_iter(batch)
            self._log_iter_metrics(loss_dict, data_time, time.perf_counter() - t0)
            return
        self._eager_iters += 1
        losses, loss_dict = self.train_cal_loss(batch)
        self.optimizer.zero_grad()
        losses.backward()
        if self._dp is not None:
            self._dp.finish()
            if any(p.grad is not None and not (self._dp.flat.data_ptr() <= p.grad.data_ptr() <
                                               self._dp.flat.data_ptr() + self._dp.flat.numel() * 4)
                   for g in self.optimizer.param_groups for p in g["params"]):
                raise RuntimeError("data-parallel step: a gradient was produced outside the flat buffer")
        if self._clip_grad_norm > 0:
            clip_grad_norm_(self.model.parameters(), self._clip_grad_norm)
        self.optimizer.step()
        self._log_iter_metrics(loss_dict, data_time, time.perf_counter() - t0)

    # -- CUDA-graph replay of the iteration ---------------------------------------------------------------
    def _graph_eligible(self, batch) -> bool:
        return (self._use_graph and self._eager_iters >= 2 and (self._dp is not None or D.get_world_size() == 1)
                and self._clip_grad_norm <= 0 and hasattr(self.optimizer, "enable_graph_mode")
                and isinstance(batch, (tuple, list)) and all(torch.is_tensor(t) for t in batch))

    def _graphed_iter(self, batch) -> Dict[str, torch.Tensor]:
        dev = next(self.model_or_module.parameters()).device
        key = tuple((tuple(t.shape), t.dtype) for t in batch)
        opt = self.optimizer
        entry = self._graphs.get(key)
        if entry is None:
            if opt._graph_hyper is None:
                opt.enable_graph_mode()
            static = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in batch]
            for s_, t in zip(static, batch):
                s_.copy_(t, non_blocking=True)
            opt.update_graph_hyper()
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            # data parallel: the bucketed NCCL all-reduces of backward (side stream, forked / joined by events) are captured
            # with the iteration, so the replay keeps the overlap; every rank captures the same sequence
            from .. import _lib
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                losses, loss_dict = self.train_cal_loss(tuple(static))
                opt.zero_grad()
                losses.backward()
                if self._dp is not None:
                    self._dp.finish()
                opt.step()                   # (host side of this call advanced the step counters once)
            nk = _lib.launch_count() - n0       # kernels of this library inside the graph (counted again at every replay)
            graph.replay()
            entry = (graph, static, loss_dict, nk)
            self._graphs[key] = entry
            return loss_dict
        graph, static, loss_dict, nk = entry
        for s_, t in zip(static, batch):
            s_.copy_(t, non_blocking=True)
        opt.update_graph_hyper()
        graph.replay()
        from .. import _lib
        _lib.lib().tsr_launch_count_add(nk)
        opt.graph_advance()
        return loss_dict

    def _log_iter_metrics(self, loss_dict: Dict[str, torch.Tensor], data_time: float, iter_time: float) -> None:
        total = sum(v.detach() for v in loss_dict.values())
        self._loss_acc = total.clone() if self._loss_acc is None else self._loss_acc + total
        self._loss_cnt += 1
        self._time_acc["data_time"] = max(self._time_acc["data_time"], data_time)
        self._time_acc["iter_time"] += iter_time
        last = self.cur_iter == self.max_iters - 1
        if self._loss_cnt < self._log_period and not last:
            return
        # one device->host read (and one tiny all-reduce) per log period
        mean = self._loss_acc / self._loss_cnt
        if D.get_world_size() > 1:
            mean = D.reduce_dict({"total_loss": mean})["total_loss"]
        value = float(mean.item())
        if torch.cuda.is_available():
            from ..engine import check_fp16_overflow      # sticky flag of the "fp16" mode, read at the same cadence
            check_fp16_overflow()
        if not np.isfinite(value):
            raise FloatingPointError(f"Loss became infinite or NaN at iteration={self.cur_iter}!")
        if D.is_main_process():
            self.log(self.cur_iter, lr=self.lr, smooth=False)
            self.log(self.cur_iter, data_time=self._time_acc["data_time"])
            self.log(self.cur_iter, iter_time=self._time_acc["iter_time"] / self._loss_cnt)
            self.log(self.cur_iter, total_loss=value)
        self._loss_acc, self._loss_cnt = None, 0
        self._time_acc = {"data_time": 0.0, "iter_time": 0.0}

    def train(self, resume_from_checkpoint: Optional[str] = None, auto_resume: bool = True) -> None:
        if resume_from_checkpoint is not None:
            self.load_checkpoint(path=resume_from_checkpoint)
        else:
            self.load_checkpoint(auto_resume=auto_resume)
        self._setup_dp()
        self._call_hooks("before_train")
        for self.cur_iter in range(self.start_iter, self.max_iters):
            if self.train_by_epoch and self.cur_iter % self.epoch_len == 0:
                self._call_hooks("before_epoch")
            self._call_hooks("before_iter")
            self.train_one_iter()
            self._call_hooks("after_iter")
            if self.train_by_epoch and (self.cur_iter + 1) % self.epoch_len == 0:
                self._call_hooks("after_epoch")
        self._call_hooks("after_train")

    # -- checkpoints (reference :394-498, same keys) ----------------------------------------------------
    @property
    def ckpt_dir(self) -> str:
        return os.path.join(self.work_dir, "checkpoints")

    def save_checkpoint(self, file_name: str) -> None:
        if not D.is_main_process():
            return
        data = {"num_gpus": D.get_world_size(), "model": self.model_or_module.state_dict(),
                "optimizer": self.optimizer.state_dict(), "lr_scheduler": self.lr_scheduler.state_dict(),
                "metric_storage": self.metric_storage}
        data["epoch" if self.train_by_epoch else "iter"] = self.cur_epoch if self.train_by_epoch else self.cur_iter
        hook_states = {type(h).__name__: h.state_dict() for h in self._hooks if callable(getattr(h, "state_dict", None))}
        if hook_states:
            data["hooks"] = hook_states
        os.makedirs(self.ckpt_dir, exist_ok=True)
        path = os.path.join(self.ckpt_dir, file_name)
        torch.save(data, path)
        link = os.path.join(self.ckpt_dir, "latest.pth")
        if os.path.lexists(link):
            os.remove(link)
        os.symlink(file_name, link)

    def load_checkpoint(self, path: Optional[str] = None, auto_resume: bool = False) -> None:
        if path is None and auto_resume:
            latest = os.path.join(self.ckpt_dir, "latest.pth")
            path = latest if os.path.exists(latest) else None
        if path is None:
            return
        ck = torch.load(path, map_location="cpu", weights_only=False)
        assert ck["num_gpus"] == D.get_world_size(), "checkpoint was written with a different number of GPUs"
        if self.train_by_epoch:
            self.start_iter = (ck["epoch"] + 1) * self.epoch_len
        else:
            self.start_iter = ck["iter"] + 1
        self.optimizer.load_state_dict(ck["optimizer"])
        self.lr_scheduler.load_state_dict(ck["lr_scheduler"])
        self.model_or_module.load_state_dict(ck["model"], strict=False)
        ms = ck.get("metric_storage")
        if isinstance(ms, MetricStorage):
            self.metric_storage = ms
        elif ms is not None and all(hasattr(v, "avg") for v in getattr(ms, "values", lambda: [None])()):
            # a checkpoint written by the reference trainer: its MetricStorage / HistoryBuffer objects have the same surface
            self.metric_storage = ms
        for h in self._hooks:
            name = type(h).__name__
            if callable(getattr(h, "load_state_dict", None)) and name in ck.get("hooks", {}):
                h.load_state_dict(ck["hooks"][name])
