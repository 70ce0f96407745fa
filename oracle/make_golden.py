"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py
The reference ships no tests / golden vectors, so these vectors --
produced by importing ``/root/reference/model/*.py`` and stock ``torch.optim.Adam`` on CPU --
are what pins both the oracle restatement (tests/test_oracle_golden.py, CPU) and the CUDA
path (tests/test_*_gpu.py).  Inputs and weights are regenerated from seeds by
``oracle.*.make_state`` so that only small outputs need to be committed.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("TACTILESR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from oracle import tactilesr_oracle as so  # noqa: E402
from oracle import tpsf_oracle as po  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NS = 16  # sampled values per tensor


def sample_idx(n: int, k: int = NS) -> np.ndarray:
    return np.unique(np.linspace(0, n - 1, min(k, n)).astype(np.int64))


def summarize(t: torch.Tensor) -> np.ndarray:
    """[l2 norm, sum, abs-sum, sampled values...] in fp64."""
    f = t.detach().double().flatten()
    s = f[torch.from_numpy(sample_idx(f.numel()))]
    return np.concatenate([[f.norm().item(), f.sum().item(), f.abs().sum().item()], s.numpy()])


def sr_inputs(B: int, S: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    LR = torch.rand(B, 3 * S, 4, 4, generator=g) * 8            # train/tactileSR_train.py:170 range
    HR_raw = torch.rand(B, 1, 100, 100, generator=g) * 250       # data/SRdataset/depth2tactile.py:51 range
    return LR, HR_raw


def golden_sr_init():
    from model.tactileSR_model import TactileSR, TactileSRCNN
    rec = {}
    for name, ctor in (("s1", lambda: TactileSR()), ("s7", lambda: TactileSR(seqsCnt=7)), ("cnn", lambda: TactileSRCNN())):
        torch.manual_seed(42)
        m = ctor()
        sd = m.state_dict()
        rec[f"{name}/keys"] = np.array(list(sd.keys()))
        rec[f"{name}/shapes"] = np.array([str(tuple(v.shape)) for v in sd.values()])
        rec[f"{name}/summary"] = np.stack([summarize(v)[:3 + 4] if v.numel() >= 4 else np.pad(summarize(v), (0, 7))[:7]
                                           for v in sd.values()])
        rec[f"{name}/param_names"] = np.array([k for k, _ in m.named_parameters()])
    np.savez_compressed(os.path.join(OUT, "tactilesr_init.npz"), **rec)


def run_reference_sr(S: int, B: int, dtype, seed_w: int, seed_x: int, training: bool):
    from model.tactileSR_model import TactileSR
    sd = so.make_state(so.tactilesr_layout(S), seed_w)
    m = TactileSR(seqsCnt=S)
    m.load_state_dict(sd, strict=True)
    m = m.to(dtype)
    m.train(training)
    LR, HR_raw = sr_inputs(B, S, seed_x)
    LR, HR_raw = LR.to(dtype), HR_raw.to(dtype)
    taps = {}
    hooks = [m.inputContact_layer.register_forward_hook(lambda _m, _i, o: taps.__setitem__("inputContact", o.detach().clone()))]
    for i, blk in enumerate(m.patternFeatureExtra_layer):
        hooks.append(blk.register_forward_hook(lambda _m, _i, o, i=i: taps.__setitem__(f"msrb{i}", o.detach().clone())))
    hooks.append(m.forceFeatureExtra_layer.register_forward_hook(lambda _m, _i, o: taps.__setitem__("force", o.detach().clone())))
    hooks.append(m.output_layer[1].register_forward_hook(lambda _m, _i, o: taps.__setitem__("output0", o.detach().clone())))
    hooks.append(m.output_layer[2].register_forward_hook(lambda _m, _i, o: taps.__setitem__("pre_relu", o.detach().clone())))
    # train_cal_loss (train/tactileSR_train.py:41-51)
    HR = HR_raw / 10
    HR = F.interpolate(HR, size=(40, 40), mode="bilinear", align_corners=False)
    out = m(LR)
    loss = nn.MSELoss()(out, HR)
    loss.backward()
    for h in hooks:
        h.remove()
    return m, out.detach(), loss.detach(), taps


def golden_sr_fwdbwd():
    for S, B in ((1, 2), (7, 2)):
        rec = {"S": S, "B": B, "seed_w": 11 + S, "seed_x": 101 + S}
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            m, out, loss, taps = run_reference_sr(S, B, dtype, rec["seed_w"], rec["seed_x"], True)
            rec[f"{tag}/out"] = out.double().numpy()
            rec[f"{tag}/loss"] = loss.double().numpy()
            for k, v in taps.items():
                rec[f"{tag}/tap/{k}"] = summarize(v)
            names, summ = [], []
            for k, p in m.named_parameters():
                names.append(k)
                summ.append(summarize(p.grad))
            rec["param_names"] = np.array(names)
            rec[f"{tag}/grad_summary"] = np.stack(summ)
            bn = {k: v for k, v in m.state_dict().items() if "running" in k}
            rec["bn_names"] = np.array(list(bn.keys()))
            rec[f"{tag}/bn_summary"] = np.stack([summarize(v) for v in bn.values()])
            # eval-mode forward with the *initial* (seeded) running stats
            m2, out_e, _, taps_e = run_reference_sr(S, B, dtype, rec["seed_w"], rec["seed_x"], False)
            rec[f"{tag}/out_eval"] = out_e.double().numpy()
            rec[f"{tag}/tap_eval/msrb5"] = summarize(taps_e["msrb5"])
        np.savez_compressed(os.path.join(OUT, f"tactilesr_fwdbwd_s{S}.npz"), **rec)


def _reference_step(m, LR, HR_raw, dtype):
    """One train_cal_loss + backward of an already constructed reference module; returns out, loss, taps."""
    LR, HR_raw = LR.to(dtype), HR_raw.to(dtype)
    taps = {}
    hooks = [m.inputContact_layer.register_forward_hook(lambda _m, _i, o: taps.__setitem__("inputContact", o.detach().clone()))]
    for i, blk in enumerate(m.patternFeatureExtra_layer):
        hooks.append(blk.register_forward_hook(lambda _m, _i, o, i=i: taps.__setitem__(f"msrb{i}", o.detach().clone())))
    hooks.append(m.forceFeatureExtra_layer.register_forward_hook(lambda _m, _i, o: taps.__setitem__("force", o.detach().clone())))
    hooks.append(m.output_layer[1].register_forward_hook(lambda _m, _i, o: taps.__setitem__("output0", o.detach().clone())))
    HR = F.interpolate(HR_raw / 10, size=(40, 40), mode="bilinear", align_corners=False)
    out = m(LR)
    loss = nn.MSELoss()(out, HR)
    loss.backward()
    for h in hooks:
        h.remove()
    return out.detach(), loss.detach(), taps


def _record_step(rec, tag, m, out, loss, taps):
    rec[f"{tag}/out_summary"] = summarize(out)
    rec[f"{tag}/out_nonzero"] = float((out > 0).double().mean())
    rec[f"{tag}/loss"] = loss.double().numpy()
    for k, v in taps.items():
        rec[f"{tag}/tap/{k}"] = summarize(v)
    rec["param_names"] = np.array([k for k, _ in m.named_parameters()])
    rec[f"{tag}/grad_summary"] = np.stack([summarize(p.grad) for _, p in m.named_parameters()])
    bn = {k: v for k, v in m.state_dict().items() if "running" in k}
    rec["bn_names"] = np.array(list(bn.keys()))
    rec[f"{tag}/bn_summary"] = np.stack([summarize(v) for v in bn.values()])


def golden_sr_c1():
    """BASELINE.json configs[0] (SURVEY 8d "C1") exactly: the reference's own construction at seed 42
    (config/default.py:10), train mode, B = 32, LR = rand * 8, HR = rand * 250, one train_cal_loss + backward."""
    from model.tactileSR_model import TactileSR
    rec = {"B": 32, "seed_init": 42, "seed_x": 4242}
    LR, HR_raw = sr_inputs(32, 1, rec["seed_x"])
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        torch.manual_seed(42)
        m = TactileSR().to(dtype).train()
        out, loss, taps = _reference_step(m, LR, HR_raw, dtype)
        _record_step(rec, tag, m, out, loss, taps)
        rec[f"{tag}/out"] = out.float().numpy()
    np.savez_compressed(os.path.join(OUT, "tactilesr_c1_b32.npz"), **rec)


def golden_sr_b256():
    """A batch that makes the persistent tensor-core kernels loop (B = 256: ~9 blocks per CTA pair): non-degenerate seeded
    weights, one training step of the reference in fp64 (and fp32 as the yardstick)."""
    from model.tactileSR_model import TactileSR
    rec = {"B": 256, "S": 1, "seed_w": 13, "seed_x": 2256}
    LR, HR_raw = sr_inputs(256, 1, rec["seed_x"])
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        m = TactileSR()
        m.load_state_dict(so.make_state(so.tactilesr_layout(1), rec["seed_w"]), strict=True)
        m = m.to(dtype).train()
        out, loss, taps = _reference_step(m, LR, HR_raw, dtype)
        _record_step(rec, tag, m, out, loss, taps)
        del m, out, taps
    np.savez_compressed(os.path.join(OUT, "tactilesr_b256.npz"), **rec)


def golden_srcnn_bwd():
    """TactileSRCNN (model/tactileSR_model.py:101-153) forward + MSE + backward, train mode, fp32 and fp64."""
    from model.tactileSR_model import TactileSRCNN
    rec = {"B": 3, "seed_w": 33, "seed_x": 403}
    sd = so.make_state(so.tactilesrcnn_layout(), rec["seed_w"])
    LR, HR_raw = sr_inputs(3, 1, rec["seed_x"])
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        m = TactileSRCNN()
        m.load_state_dict(sd, strict=True)
        m = m.to(dtype).train()
        HR = F.interpolate(HR_raw.to(dtype) / 10, size=(40, 40), mode="bilinear", align_corners=False)
        out = m(LR.to(dtype))
        loss = nn.MSELoss()(out, HR)
        loss.backward()
        rec[f"{tag}/out"] = out.detach().double().numpy()
        rec[f"{tag}/loss"] = loss.detach().double().numpy()
        rec["param_names"] = np.array([k for k, _ in m.named_parameters()])
        rec[f"{tag}/grad_summary"] = np.stack([summarize(p.grad) for _, p in m.named_parameters()])
        rec[f"{tag}/grad_output_w"] = m.output[0].weight.grad.detach().double().numpy()
        rec[f"{tag}/grad_in0_w"] = m.input_zyx[0].weight.grad.detach().double().numpy()
    np.savez_compressed(os.path.join(OUT, "tactilesrcnn_bwd.npz"), **rec)


def golden_sr_adam():
    """3 steps of fwd + MSE + bwd + stock Adam(lr 1e-3, wd 1e-2) (train/tactileSR_train.py:212)."""
    from model.tactileSR_model import TactileSR
    S, B, steps = 1, 2, 3
    rec = {"S": S, "B": B, "steps": steps, "seed_w": 21, "seed_x0": 301}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        sd = so.make_state(so.tactilesr_layout(S), rec["seed_w"])
        m = TactileSR(seqsCnt=S)
        m.load_state_dict(sd, strict=True)
        m = m.to(dtype).train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-2)
        losses = []
        for t in range(steps):
            LR, HR_raw = sr_inputs(B, S, rec["seed_x0"] + t)
            LR, HR_raw = LR.to(dtype), HR_raw.to(dtype)
            HR = F.interpolate(HR_raw / 10, size=(40, 40), mode="bilinear", align_corners=False)
            loss = nn.MSELoss()(m(LR), HR)
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.item())
        rec[f"{tag}/losses"] = np.array(losses)
        fin = m.state_dict()
        rec["state_names"] = np.array(list(fin.keys()))
        rec[f"{tag}/state_summary"] = np.stack([summarize(v) if v.numel() >= NS else np.pad(summarize(v), (0, 3 + NS))[:3 + NS]
                                                for v in fin.values()])
    np.savez_compressed(os.path.join(OUT, "tactilesr_adam_s1.npz"), **rec)


def golden_srcnn():
    from model.tactileSR_model import TactileSRCNN
    rec = {"B": 2, "seed_w": 31, "seed_x": 401}
    sd = so.make_state(so.tactilesrcnn_layout(), rec["seed_w"])
    LR, _ = sr_inputs(2, 1, rec["seed_x"])
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        m = TactileSRCNN()
        m.load_state_dict(sd, strict=True)
        m = m.to(dtype).train()
        rec[f"{tag}/out_train"] = m(LR.to(dtype)).detach().double().numpy()
        m.load_state_dict(sd, strict=True)
        m = m.to(dtype).eval()
        rec[f"{tag}/out_eval"] = m(LR.to(dtype)).detach().double().numpy()
    np.savez_compressed(os.path.join(OUT, "tactilesrcnn_fwd.npz"), **rec)


def golden_tpsf():
    from model.tPSFNet import tPSFNet
    B = 4
    rec = {"B": B, "seed_w": 41, "seed_x": 501}
    g = torch.Generator().manual_seed(rec["seed_x"])
    LR_raw = torch.rand(B, 3, 4, 4, generator=g) * 1300          # raw taxel units; /100 in train_cal_loss
    depth = po.synthetic_depth(B, rec["seed_x"] + 1)
    rec["LR_raw"] = LR_raw.numpy()
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        # construct under the fp32 default so both constant tables hold the reference's fp32-rounded
        # values; only the forward's scratch tensors (torch.zeros(...), tPSFNet.py:111-114,136) are
        # allocated in fp64 for the yardstick run.
        m = tPSFNet(gama=1.4, perception_scale=None, device="cpu")
        m.load_state_dict(po.make_state(rec["seed_w"]), strict=True)
        if dtype == torch.float64:
            torch.set_default_dtype(torch.float64)
        try:
            m = m.to(dtype)
            # PSF_sdf is a plain attribute pinned to fp32 (tPSFNet.py:41); for the fp64 yardstick
            # run the table itself (same fp32-rounded values) is widened so F.conv2d type-checks.
            m.PSF_sdf = m.PSF_sdf.to(dtype)
            m.LR_masking_sdf = m.LR_masking_sdf.to(dtype)
            LR = LR_raw.to(dtype) / 100                        # train/tPSFNet_train.py:183
            HR, LRd, psf, ab = m(LR, depth.to(dtype).unsqueeze(1))
            loss = nn.MSELoss()(LR[:, 2:3], LRd)
            loss.backward()
        finally:
            torch.set_default_dtype(torch.float32)
        rec[f"{tag}/HR"] = HR.detach().double().numpy()
        rec[f"{tag}/LRd"] = LRd.detach().double().numpy()
        rec[f"{tag}/psf_summary"] = np.stack([summarize(psf[b]) for b in range(B)])
        rec[f"{tag}/psf_center_row"] = psf[:, 0, 49].detach().double().numpy()
        rec[f"{tag}/alphaBeta"] = ab.detach().double().numpy()
        rec[f"{tag}/loss"] = loss.detach().double().numpy()
        rec["param_names"] = np.array([k for k, _ in m.named_parameters()])
        rec[f"{tag}/grad_summary"] = np.stack([np.pad(summarize(p.grad), (0, 3 + NS))[:3 + NS] for p in m.parameters()])
        if tag == "f32":
            rec["PSF_sdf_summary"] = summarize(m.PSF_sdf)
            rec["LR_masking_sdf_summary"] = summarize(m.LR_masking_sdf)
    # the reference's own __main__ smoke shape spec (tPSFNet.py:144-150)
    torch.manual_seed(0)
    m = tPSFNet(gama=0.5, perception_scale=None, device="cpu")
    HR, LRd, _, _ = m(torch.rand(4, 3, 4, 4), torch.rand(4, 1, 100, 100))
    rec["smoke_shapes"] = np.array([list(HR.shape), list(LRd.shape)])
    np.savez_compressed(os.path.join(OUT, "tpsf_fwdbwd.npz"), **rec)


def golden_eval():
    """eval_func's loop body (train/tactileSR_train.py:76-94) with the reference's own calculationPSNR / calculationSSIM."""
    from utility.tools import calculationPSNR, calculationSSIM
    B = 5
    rec = {"B": B, "seed_x": 601}
    g = torch.Generator().manual_seed(rec["seed_x"])
    out = torch.relu(torch.randn(B, 1, 40, 40, generator=g) * 5 + 6)
    HR_raw = torch.rand(B, 1, 100, 100, generator=g) * 250
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        o, HR = out.to(dtype), HR_raw.to(dtype) / 10
        HR = F.interpolate(HR, size=(40, 40), mode="bilinear", align_corners=False)
        rec[f"{tag}/mse"] = nn.MSELoss()(o, HR).double().numpy()
        rec[f"{tag}/psnr"] = np.array([float(calculationPSNR(o[i], HR[i], maxValue=250)) for i in range(B)])
        rec[f"{tag}/ssim"] = np.array([float(calculationSSIM(o[i], HR[i])) for i in range(B)])
    np.savez_compressed(os.path.join(OUT, "eval_metrics.npz"), **rec)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if "--eval-only" in sys.argv:
        golden_eval()
        sys.exit(0)
    if "--round2" in sys.argv:          # the fixtures added in round 2 (the others are unchanged)
        golden_srcnn_bwd()
        print("srcnn bwd done")
        golden_sr_c1()
        print("c1 done")
        golden_sr_b256()
        print("b256 done")
        sys.exit(0)
    golden_sr_init()
    print("init done")
    golden_tpsf()
    print("tpsf done")
    golden_srcnn()
    print("srcnn done")
    golden_sr_fwdbwd()
    print("fwdbwd done")
    golden_sr_adam()
    print("adam done")
    golden_srcnn_bwd()
    golden_sr_c1()
    golden_sr_b256()
    print("round-2 fixtures done")
    golden_eval()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
