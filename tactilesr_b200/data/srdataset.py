"""Offline SR-dataset generation with *batched* tPSFNet inference, and the reference's on-disk format.

Replaces the batch-size-1 loops of reference ``data/SRdataset/depth2tactile.py:107-160`` (single-frame records
``{LR (3,4,4), depth (1,100,100), HR (1,100,100), LR_degrade (1,4,4), alphaBeta (3,)}``) and
``data/SeqsDataset/seqsDepth2Tactile.py:47-107`` (sequence records ``{LR (21,4,4), depth (1,100,100), HR (1,100,100)}``):
the PSF model runs once per chunk of thousands of samples (the tcgen05 PSF kernels are per-sample independent, so the
records are the same as with batch 1) and only the finished chunk is copied back to the host.

On disk: ``np.save(path, [[record], [record], ...])`` -- an object array of shape (N, 1) whose items are dicts of CPU
``torch`` tensors -- read back by ``TactileSRDataset`` exactly like reference ``utility/load_tactile_dataset.py:39-48``.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch


def _chunks(n: int, size: int):
    for i in range(0, n, size):
        yield i, min(i + size, n)


@torch.no_grad()
def generate_sr_records(tpsf_model, LR_raw: torch.Tensor, depth: torch.Tensor, scale_num: float = 100.0,
                        batch_size: int = 4096, device=None) -> List[list]:
    """LR_raw (N,3,4,4) raw taxel units, depth (N,100,100) -> the reference's list of one-element lists of records.
    Per record, exactly what depth2tactile.py:107-119 stores: ``LR`` = LR_raw / scale_num, ``depth`` with a leading
    channel axis, ``HR`` / ``LR_degrade`` = model outputs of that sample, ``alphaBeta`` = ret_alphaBeta[i][0]."""
    device = device if device is not None else next(tpsf_model.parameters()).device
    N = LR_raw.shape[0]
    assert depth.shape[0] == N
    records: List[list] = []
    tpsf_model.eval()
    for a, b in _chunks(N, batch_size):
        LR = (LR_raw[a:b].to(device, non_blocking=True).float()) / scale_num
        d = depth[a:b].to(device, non_blocking=True).float().unsqueeze(1)
        HR, LRd, _, ab = tpsf_model(LR, d)
        LR_c, d_c, HR_c, LRd_c, ab_c = LR.cpu(), d.cpu(), HR.cpu(), LRd.cpu(), ab.cpu()
        for i in range(b - a):
            records.append([{"LR": LR_c[i].clone(), "depth": d_c[i].clone(), "HR": HR_c[i].clone(),
                             "LR_degrade": LRd_c[i].clone(), "alphaBeta": ab_c[i][0].clone()}])
    return records


@torch.no_grad()
def generate_seqs_sr_records(tpsf_model, LR_frames_raw: torch.Tensor, depth_last: torch.Tensor, scale_num: float = 100.0,
                             batch_size: int = 4096, device=None) -> List[list]:
    """Sequence records of seqsDepth2Tactile.py:47-107.  LR_frames_raw (N,7,3,4,4): the seven frames in acquisition order
    (0, 5, ..., 30 degrees); depth_last (N,100,100): depth of the last (30 degree) frame.  The HR label is the PSF model's
    output for the last frame; ``LR`` is the frames concatenated last-to-first (21,4,4), as the reference stores them."""
    device = device if device is not None else next(tpsf_model.parameters()).device
    N = LR_frames_raw.shape[0]
    records: List[list] = []
    tpsf_model.eval()
    for a, b in _chunks(N, batch_size):
        frames = LR_frames_raw[a:b].float() / scale_num                      # (n,7,3,4,4)
        d = depth_last[a:b].to(device, non_blocking=True).float().unsqueeze(1)
        HR, _, _, _ = tpsf_model(frames[:, -1].to(device, non_blocking=True), d)
        LR_cat = torch.flip(frames, dims=[1]).reshape(b - a, 21, 4, 4)
        d_c, HR_c = d.cpu(), HR.cpu()
        for i in range(b - a):
            records.append([{"LR": LR_cat[i].clone(), "depth": d_c[i].clone(), "HR": HR_c[i].clone()}])
    return records


def save_sr_dataset(path: str, records: Sequence[list]) -> None:
    """The reference's ``np.save(path, SRdataset)`` (depth2tactile.py:153-160): object array (N, 1) of dicts."""
    arr = np.empty((len(records), 1), dtype=object)
    for i, r in enumerate(records):
        arr[i, 0] = r[0]
    np.save(path, arr, allow_pickle=True)


class TactileSRDataset(torch.utils.data.Dataset):
    """reference utility/load_tactile_dataset.py:39-48 (also TactileSRDataset_seq :52-60): (LR, HR) of record idx."""

    def __init__(self, dataset_dir: str):
        self.SRdataset = np.load(dataset_dir, allow_pickle=True)

    def __getitem__(self, idx):
        rec = self.SRdataset[idx].item()
        return np.ascontiguousarray(rec["LR"]), np.ascontiguousarray(rec["HR"])

    def __len__(self):
        return len(self.SRdataset)


class DevicePrefetcher:
    """Wraps a loader of (tensor, ...) batches: the next batch is copied host -> device on a side stream while the current
    iteration computes, so the H2D copy (40 KB per sample for the fp32 HR label) leaves the critical path.  The consumer
    stream waits on the copy's event and the tensors are recorded on it, so the caching allocator never recycles a batch
    that is still in use.  Pin the loader's tensors (``DataLoader(pin_memory=True)``) for the copies to be asynchronous."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)
        self.stream = torch.cuda.Stream(self.device)

    def __len__(self):
        return len(self.loader)

    def _issue(self, it):
        try:
            batch = next(it)
        except StopIteration:
            return None
        with torch.cuda.stream(self.stream):
            out = tuple(t.to(self.device, non_blocking=True) if torch.is_tensor(t) else t for t in batch)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    def __iter__(self):
        it = iter(self.loader)
        nxt = self._issue(it)
        while nxt is not None:
            batch, ev = nxt
            nxt = self._issue(it)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in batch:
                if torch.is_tensor(t):
                    t.record_stream(cur)
            yield batch
