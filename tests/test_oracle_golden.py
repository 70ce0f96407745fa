"""CPU: pin the oracle restatement against vectors produced by the unmodified reference
(oracle/make_golden.py).  Tolerances: the fp64 oracle must reproduce the fp64 reference run
to 1e-9; the fp32 oracle must sit at the reference's own fp32-vs-fp64 noise floor."""
import numpy as np
import pytest
import torch

from oracle import tactilesr_oracle as so
from oracle import tpsf_oracle as po
from tests.util import load_golden, rel_l2, summarize, summary_close, sr_inputs


@pytest.mark.parametrize("S", [1, 7])
def test_sr_oracle_fp64_matches_reference(S):
    g = load_golden(f"tactilesr_fwdbwd_s{S}.npz")
    sd = so.make_state(so.tactilesr_layout(S), int(g["seed_w"]), dtype=torch.float64)
    LR, HR_raw = sr_inputs(int(g["B"]), S, int(g["seed_x"]))
    taps = {}
    loss, out, grads, new_stats = so.loss_and_grads(sd, LR.double(), HR_raw.double(), True, taps=taps)
    assert rel_l2(out, g["f64/out"]) < 1e-10
    assert abs(float(loss) - float(g["f64/loss"])) / float(g["f64/loss"]) < 1e-10
    for k, v in taps.items():
        ok, err = summary_close(summarize(v), g[f"f64/tap/{k}"], 1e-9)
        assert ok, (k, err)
    names = [str(n) for n in g["param_names"]]
    assert names == so.param_keys(sd)
    for n, want in zip(names, g["f64/grad_summary"]):
        if ".0.bias" in n and ("conv_3_" in n or "conv_5_" in n):
            # conv bias feeding a train-mode BN: true gradient is 0 (SURVEY section 0 pitfall 2)
            assert summarize(grads[n])[0] < 1e-9
            continue
        ok, err = summary_close(summarize(grads[n]), want, 1e-7)
        assert ok, (n, err)
    for n, want in zip([str(x) for x in g["bn_names"]], g["f64/bn_summary"]):
        ok, err = summary_close(summarize(new_stats[n]), want, 1e-9)
        assert ok, (n, err)
    out_e = so.tactilesr_forward(sd, LR.double(), training=False)
    assert rel_l2(out_e, g["f64/out_eval"]) < 1e-10


def test_sr_oracle_fp32_at_reference_noise_floor():
    g = load_golden("tactilesr_fwdbwd_s1.npz")
    sd = so.make_state(so.tactilesr_layout(1), int(g["seed_w"]))
    LR, HR_raw = sr_inputs(int(g["B"]), 1, int(g["seed_x"]))
    loss, out, grads, _ = so.loss_and_grads(sd, LR, HR_raw, True)
    ref_noise = rel_l2(g["f32/out"], g["f64/out"])
    assert rel_l2(out, g["f64/out"]) < max(3 * ref_noise, 1e-5)
    assert rel_l2(out, g["f32/out"]) < 1e-5
    assert (out > 0).float().mean() > 0.3, "golden weights must keep the SR output alive"


def test_sr_oracle_adam_matches_stock_adam():
    g = load_golden("tactilesr_adam_s1.npz")
    sd = so.make_state(so.tactilesr_layout(1), int(g["seed_w"]), dtype=torch.float64)
    batches = [tuple(t.double() for t in sr_inputs(int(g["B"]), 1, int(g["seed_x0"]) + t)) for t in range(int(g["steps"]))]
    losses, fin = so.train_steps(sd, batches)
    np.testing.assert_allclose(losses, g["f64/losses"], rtol=1e-9)
    for n, want in zip([str(x) for x in g["state_names"]], g["f64/state_summary"]):
        got = summarize(fin[n])
        k = min(len(got), len(want))
        np.testing.assert_allclose(got[:k], want[:k], rtol=1e-7, atol=1e-9, err_msg=n)


def test_srcnn_oracle_matches_reference():
    g = load_golden("tactilesrcnn_fwd.npz")
    sd = so.make_state(so.tactilesrcnn_layout(), int(g["seed_w"]), dtype=torch.float64)
    LR, _ = sr_inputs(int(g["B"]), 1, int(g["seed_x"]))
    assert rel_l2(so.tactilesrcnn_forward(sd, LR.double(), True), g["f64/out_train"]) < 1e-10
    assert rel_l2(so.tactilesrcnn_forward(sd, LR.double(), False), g["f64/out_eval"]) < 1e-10


def test_state_layout_matches_reference_keys():
    g = load_golden("tactilesr_init.npz")
    for tag, layout in (("s1", so.tactilesr_layout(1)), ("s7", so.tactilesr_layout(7)), ("cnn", so.tactilesrcnn_layout())):
        assert [k for k, _, _ in layout] == [str(k) for k in g[f"{tag}/keys"]]
        assert [str(tuple(s)) for _, s, _ in layout] == [str(s) for s in g[f"{tag}/shapes"]]


def _tpsf_inputs(g):
    LR_raw = torch.from_numpy(g["LR_raw"])
    depth = po.synthetic_depth(int(g["B"]), int(g["seed_x"]) + 1)
    return LR_raw, depth


@pytest.mark.parametrize("tag,dtype,tol", [("f64", torch.float64, 1e-9), ("f32", torch.float32, 2e-5)])
def test_tpsf_oracle_matches_reference(tag, dtype, tol):
    g = load_golden("tpsf_fwdbwd.npz")
    sd = po.make_state(int(g["seed_w"]), dtype=dtype)
    LR_raw, depth = _tpsf_inputs(g)
    loss, (HR, LRd, psf, ab), grads = po.loss_and_grads(sd, LR_raw.to(dtype), depth.to(dtype))
    assert rel_l2(HR, g[f"{tag}/HR"]) < tol
    assert rel_l2(LRd, g[f"{tag}/LRd"]) < tol
    assert rel_l2(ab, g[f"{tag}/alphaBeta"]) < tol
    assert rel_l2(psf[:, 0, 49], g[f"{tag}/psf_center_row"]) < tol
    assert abs(float(loss) - float(g[f"{tag}/loss"])) / float(g[f"{tag}/loss"]) < tol
    for n, want in zip([str(x) for x in g["param_names"]], g[f"{tag}/grad_summary"]):
        got = summarize(grads[n])
        k = min(len(got), len(want))
        assert abs(got[0] - want[0]) / want[0] < max(tol, 1e-7) * 50, n
    assert list(g["smoke_shapes"][0]) == [4, 1, 100, 100] and list(g["smoke_shapes"][1]) == [4, 1, 4, 4]


def test_tpsf_tables_match_reference():
    g = load_golden("tpsf_fwdbwd.npz")
    np.testing.assert_allclose(summarize(po.psf_sdf()), g["PSF_sdf_summary"], rtol=1e-6)
    np.testing.assert_allclose(summarize(po.masking_sdf()), g["LR_masking_sdf_summary"], rtol=1e-6)


def test_eval_metrics_oracle_matches_reference():
    """oracle.psnr / ssim / eval_batch vs the reference's calculationPSNR / calculationSSIM run on eval_func's slices."""
    import torch
    from oracle import tactilesr_oracle as so
    g = load_golden("eval_metrics.npz")
    gen = torch.Generator().manual_seed(int(g["seed_x"]))
    B = int(g["B"])
    out = torch.relu(torch.randn(B, 1, 40, 40, generator=gen) * 5 + 6)
    HR_raw = torch.rand(B, 1, 100, 100, generator=gen) * 250
    mse, ps, ss = so.eval_batch(out.double(), HR_raw.double(), 10.0, 250.0)
    assert abs(mse.item() - float(g["f64/mse"])) / float(g["f64/mse"]) < 1e-12
    assert np.abs(ps.numpy() - g["f64/psnr"]).max() < 1e-10
    assert np.abs(ss.numpy() - g["f64/ssim"]).max() < 1e-12
