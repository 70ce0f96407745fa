"""Data-parallel plumbing: same helper names as reference cpu/distributed.py (:36-151, :171-217), redesigned for one
process per B200 over NCCL / NVLink 5.

Differences from the reference:
  * the device is rank-local (``cuda:LOCAL_RANK``), never picked by parsing nvidia-smi (config/default.py:101-104);
  * no gloo side group and no pickled-object gather on the per-iteration path: scalar metrics are all-reduced as
    one small tensor every ``log_period`` iterations (``reduce_dict``);
  * the gradient all-reduce that DDP would have done implicitly is explicit: ``GradAllReduce`` averages the flat fp32
    gradient buffer of ``FusedAdam`` in reverse-layer-order buckets on a side stream while backward is still running.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["all_gather", "gather", "reduce_dict", "get_world_size", "get_rank", "is_main_process", "init_distributed",
           "GradAllReduce"]


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def is_main_process() -> bool:
    return get_rank() == 0


def init_distributed(auto: bool = False, backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world_size); (0, 0, 1) when not launched by torchrun / SLURM (reference :171-217).
    ``backend`` defaults to nccl when CUDA is present (gloo is used by the CPU tests)."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        local_rank = int(os.environ.get("LOCAL_RANK", 0))
    elif "SLURM_PROCID" in os.environ:
        rank, world = int(os.environ["SLURM_PROCID"]), int(os.environ["SLURM_NTASKS"])
        local_rank = rank % max(torch.cuda.device_count(), 1)
    else:
        return 0, 0, 1
    if "MASTER_ADDR" not in os.environ or "MASTER_PORT" not in os.environ:
        raise RuntimeError("init_method='env://' requires MASTER_ADDR and MASTER_PORT")
    use_cuda = torch.cuda.is_available()
    backend = backend or ("nccl" if use_cuda else "gloo")
    if use_cuda:
        torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        kw = {"device_id": torch.device("cuda", local_rank)} if (use_cuda and backend == "nccl") else {}
        dist.init_process_group(backend=backend, init_method="env://", rank=rank, world_size=world, **kw)
    dist.barrier()
    return rank, local_rank, world


def all_gather(data: Any, group=None) -> List[Any]:
    """Gather an arbitrary picklable object from every rank (reference :36-57).  Not used per iteration."""
    if get_world_size() == 1:
        return [data]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, data, group=group)
    return out


def gather(data: Any, dst: int = 0, group=None) -> List[Any]:
    """Gather an object on rank ``dst`` (reference :60-86); other ranks get []."""
    if get_world_size() == 1:
        return [data]
    if dist.get_rank(group) == dst:
        out = [None] * dist.get_world_size(group)
        dist.gather_object(data, out, dst=dst, group=group)
        return out
    dist.gather_object(data, None, dst=dst, group=group)
    return []


def reduce_dict(input_dict: Dict[str, torch.Tensor], average: bool = True) -> Dict[str, torch.Tensor]:
    """All-reduce a dict of scalar tensors as ONE stacked tensor (reference :89-115)."""
    world = get_world_size()
    if world < 2:
        return input_dict
    with torch.no_grad():
        names = sorted(input_dict.keys())
        vals = torch.stack([input_dict[k].detach().float().reshape(()) for k in names])
        dist.all_reduce(vals)
        if average:
            vals /= world
        return {k: v for k, v in zip(names, vals)}


class GradAllReduce:
    """Bucketed, backward-overlapped average of a flat gradient buffer.

    ``flat`` is ``FusedAdam.flat_grad()``; ``ready(lo, hi)`` is called (from the engine's backward hook) when the
    element range [lo, hi) is final.  Ranges are coalesced into buckets of ``bucket_bytes`` and all-reduced on a
    dedicated stream ordered after the producing kernels by an event; ``finish()`` makes the compute stream wait
    for all outstanding buckets (call before ``optimizer.step``)."""

    def __init__(self, flat: torch.Tensor, bucket_bytes: int = 8 << 20, group=None, overlap: Optional[bool] = None):
        self.flat, self.group = flat, group
        # overlap = False (default): ONE all-reduce over the whole buffer after backward, on the compute stream.
        # overlap = True (or TSR_DP_OVERLAP=1): buckets are reduced on a side stream while backward is still running.
        # Measured on 8 x B200 (B = 1024 per GPU, profiles/r02_scaling.md): the single call costs 0.8 ms of a 54.9 ms step
        # (18.3 MB over NVLink: ~0.25 ms exposed, the rest is the max over ranks), the overlapped form 1.3 ms -- its NCCL
        # CTAs cannot co-reside with the one-CTA-per-SM persistent convolution kernels of backward (registers), so each
        # bucket takes SMs away from the cluster-pair kernels that follow; hiding 0.25 ms is not worth that.
        if overlap is None:
            overlap = os.environ.get("TSR_DP_OVERLAP", "0") == "1"
        self.overlap = overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_elems = max(bucket_bytes // 4, 1)
        self.pending: List[Tuple[int, int]] = []
        self.pending_elems = 0
        self.cuda = flat.is_cuda
        self.stream = torch.cuda.Stream(device=flat.device) if self.cuda else None
        self.handles = []
        self.launched: List[Tuple[int, int]] = []

    def ready(self, lo: int, hi: int) -> None:
        if self.world < 2 or hi <= lo or not self.overlap:
            return
        self.pending.append((lo, hi))
        self.pending_elems += hi - lo
        if self.pending_elems >= self.bucket_elems:
            self._flush()

    def _flush(self) -> None:
        if not self.pending:
            return
        # coalesce adjacent ranges (backward finishes gradients in reverse layout order => mostly one span)
        spans = sorted(self.pending)
        merged = [list(spans[0])]
        for lo, hi in spans[1:]:
            if lo <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], hi)
            else:
                merged.append([lo, hi])
        self.pending, self.pending_elems = [], 0
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.stream.wait_event(ev)
        for lo, hi in merged:
            view = self.flat[lo:hi]
            self.launched.append((lo, hi))
            if self.cuda:
                with torch.cuda.stream(self.stream):
                    dist.all_reduce(view, op=dist.ReduceOp.AVG if dist.get_backend(self.group) == "nccl" else dist.ReduceOp.SUM,
                                    group=self.group)
                    if dist.get_backend(self.group) != "nccl":
                        view /= self.world
            else:
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
                view /= self.world

    def finish(self) -> None:
        if self.world < 2:
            return
        if not self.overlap:
            nccl = dist.get_backend(self.group) == "nccl"
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM, group=self.group)
            if not nccl:
                self.flat /= self.world
            return
        self._flush()
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.launched = []
