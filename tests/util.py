"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NS = 16


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def sample_idx(n, k=NS):
    return np.unique(np.linspace(0, n - 1, min(k, n)).astype(np.int64))


def summarize(t):
    f = t.detach().double().flatten().cpu()
    s = f[torch.from_numpy(sample_idx(f.numel()))]
    return np.concatenate([[f.norm().item(), f.sum().item(), f.abs().sum().item()], s.numpy()])


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64) if not torch.is_tensor(a) else a.detach().double().cpu()
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64) if not torch.is_tensor(b) else b.detach().double().cpu()
    d = (a - b).norm().item()
    n = b.norm().item()
    return d / n if n > 0 else d


def summary_close(got, want, rtol, atol_scale=1.0):
    """Compare a summarize() vector: the norm relatively, sampled values against the norm scale."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    n = min(len(got), len(want))
    got, want = got[:n], want[:n]
    norm = max(abs(want[0]), 1e-30)
    err_norm = abs(got[0] - want[0]) / norm
    # sampled values: error relative to the rms magnitude implied by the l2 norm is meaningless
    # without numel, so compare against max |sample| with a floor of a fraction of the norm
    scale = max(np.abs(want[3:]).max() if n > 3 else 0.0, 1e-3 * norm) * atol_scale
    err_s = np.abs(got[3:] - want[3:]).max() / scale if n > 3 else 0.0
    return max(err_norm, err_s) <= rtol, (err_norm, err_s)


def sr_inputs(B, S, seed):
    g = torch.Generator().manual_seed(seed)
    LR = torch.rand(B, 3 * S, 4, 4, generator=g) * 8
    HR_raw = torch.rand(B, 1, 100, 100, generator=g) * 250
    return LR, HR_raw
