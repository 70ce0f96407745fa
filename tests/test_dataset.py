"""batched dataset generation + the reference's pickled-dict .npy format.
CPU: the file we write is what the reference's own np.save call writes and its reader class reads it.
GPU: batched generation == batch-1 generation (the reference's loop) record for record."""
import os
import sys

import numpy as np
import pytest
import torch

REF = os.environ.get("TACTILESR_REFERENCE", "/root/reference")


def _fake_records(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [[{"LR": torch.rand(3, 4, 4, generator=g), "depth": torch.rand(1, 100, 100, generator=g),
              "HR": torch.rand(1, 100, 100, generator=g), "LR_degrade": torch.rand(1, 4, 4, generator=g),
              "alphaBeta": torch.rand(3, generator=g)}] for _ in range(n)]


def test_file_format_round_trip(tmp_path):
    from tactilesr_b200.data import TactileSRDataset, save_sr_dataset
    recs = _fake_records(5)
    ours = str(tmp_path / "ours.npy")
    save_sr_dataset(ours, recs)
    # what the reference's generator writes: np.save(path, list_of_[dict])  (depth2tactile.py:158)
    ref = str(tmp_path / "ref.npy")
    np.save(ref, recs)
    a, b = np.load(ours, allow_pickle=True), np.load(ref, allow_pickle=True)
    assert a.shape == b.shape == (5, 1) and a.dtype == b.dtype == object
    for i in range(5):
        ra, rb = a[i].item(), b[i].item()
        assert list(ra.keys()) == list(rb.keys())
        for k in ra:
            assert torch.equal(ra[k], rb[k]) and ra[k].dtype == rb[k].dtype
    ds = TactileSRDataset(ours)
    assert len(ds) == 5
    LR, HR = ds[3]
    assert LR.shape == (3, 4, 4) and HR.shape == (1, 100, 100)
    assert np.array_equal(LR, recs[3][0]["LR"].numpy())


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "utility", "load_tactile_dataset.py")), reason="no reference here")
def test_reference_reader_reads_our_file(tmp_path):
    from tactilesr_b200.data import save_sr_dataset
    if REF not in sys.path:
        sys.path.insert(0, REF)
    try:
        from utility.load_tactile_dataset import TactileSRDataset as RefDS
    except Exception as e:      # the reference module drags optional dependencies (cv2, sklearn) at import time
        pytest.skip(f"reference reader not importable here: {e}")
    recs = _fake_records(4, seed=2)
    path = str(tmp_path / "d.npy")
    save_sr_dataset(path, recs)
    ds = RefDS(path)
    assert len(ds) == 4
    LR, HR = ds[1]
    assert np.array_equal(LR, recs[1][0]["LR"].numpy()) and np.array_equal(HR, recs[1][0]["HR"].numpy())


@pytest.mark.gpu
def test_batched_generation_equals_batch_one_loop(tmp_path):
    from oracle import tpsf_oracle as po
    from tactilesr_b200.data import TactileSRDataset, generate_seqs_sr_records, generate_sr_records, save_sr_dataset
    from tactilesr_b200.model import tPSFNet
    m = tPSFNet(1.4, None, device="cuda")
    m.load_state_dict(po.make_state(5), strict=True)
    m = m.cuda()
    N = 11
    g = torch.Generator().manual_seed(6)
    LR_raw = torch.rand(N, 3, 4, 4, generator=g) * 1300
    depth = po.synthetic_depth(N, 7)
    batched = generate_sr_records(m, LR_raw, depth, batch_size=4)        # ragged chunks 4, 4, 3
    single = generate_sr_records(m, LR_raw, depth, batch_size=1)         # the reference's loop shape
    assert len(batched) == len(single) == N
    for rb, rs in zip(batched, single):
        assert list(rb[0].keys()) == ["LR", "depth", "HR", "LR_degrade", "alphaBeta"]
        assert rb[0]["HR"].shape == (1, 100, 100) and rb[0]["LR_degrade"].shape == (1, 4, 4) and rb[0]["alphaBeta"].shape == (3,)
        for k in rb[0]:
            assert torch.equal(rb[0][k], rs[0][k]), k
    # against the CPU oracle of the model
    HRo, LRdo, _, abo = po.tpsf_forward({k: v.double() for k, v in po.make_state(5).items()}, (LR_raw / 100).double(),
                                        depth.double().unsqueeze(1))
    for i, r in enumerate(batched):
        assert (r[0]["HR"].double() - HRo[i]).norm() / HRo[i].norm() < 2e-5
        assert (r[0]["alphaBeta"].double() - abo[i][0]).abs().max() < 1e-5
    path = str(tmp_path / "SRdataset_train.npy")
    save_sr_dataset(path, batched)
    ds = TactileSRDataset(path)
    assert len(ds) == N and ds[2][1].shape == (1, 100, 100)
    # sequence records
    frames = torch.rand(N, 7, 3, 4, 4, generator=g) * 1300
    seqs = generate_seqs_sr_records(m, frames, depth, batch_size=5)
    assert seqs[0][0]["LR"].shape == (21, 4, 4) and list(seqs[0][0].keys()) == ["LR", "depth", "HR"]
    assert torch.equal(seqs[3][0]["LR"][:3], frames[3, 6] / 100) and torch.equal(seqs[3][0]["LR"][18:], frames[3, 0] / 100)


@pytest.mark.gpu
def test_device_prefetcher_yields_the_same_batches():
    from tactilesr_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(1)
    batches = [(torch.rand(4, 3, 4, 4, generator=g).pin_memory(), torch.rand(4, 1, 100, 100, generator=g).pin_memory())
               for _ in range(5)]
    got = list(DevicePrefetcher(batches, "cuda:0"))
    assert len(got) == 5
    for (a, b), (c, d) in zip(batches, got):
        assert c.is_cuda and torch.equal(a, c.cpu()) and torch.equal(b, d.cpu())


@pytest.mark.gpu
def test_c_abi_error_paths():
    """Bad arguments come back as error codes with a message (no crash, no silent fallback)."""
    from tactilesr_b200 import TsrError, _lib
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    x = torch.zeros(1, 40, 40, 64, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(9 * 64 * 64, dtype=torch.bfloat16, device="cuda")
    o = torch.zeros(1, 40, 40, 64, dtype=torch.bfloat16, device="cuda")
    rc = L.tsr_conv2d_tc(x.data_ptr(), 64, w.data_ptr(), 0, 0, 0, o.data_ptr(), 64, 1, 40, 40, 64, 32, 3, 0, 0, 0, 0, 0, 0, st)
    assert rc == 1 and b"Cout" in L.tsr_last_error()
    rc = L.tsr_conv2d_tc(x.data_ptr(), 64, w.data_ptr(), 0, 0, 0, o.data_ptr(), 64, 1, 40, 40, 64, 64, 7, 0, 0, 0, 0, 0, 0, st)
    assert rc == 1 and b"kernel size" in L.tsr_last_error()
    with pytest.raises(TsrError):
        _lib.call("tsr_psf_forward", 0, 0, 0, 0, 0, 4, st)
    dw = torch.zeros(64, 64, 3, 3, device="cuda")
    rc = L.tsr_conv2d_wgrad_tc(x.data_ptr(), 64, o.data_ptr(), 64, dw.data_ptr(), w.data_ptr(), 16, 1, 40, 40, 64, 64, 3, 0, st)
    assert rc == 4 and b"workspace" in L.tsr_last_error()
    torch.cuda.synchronize()
