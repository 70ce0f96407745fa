#!/usr/bin/env python3
"""Headline benchmark of the tactileSR hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32] [--batch B]

A "step" is one training iteration of TactileSR(seqsCnt=1) -- forward + fused HR-prep/MSE loss + backward + fused Adam
(+ the bucketed NCCL gradient all-reduce when N > 1) -- on one synthetic batch of B samples per GPU (weak scaling).
``value`` times K steps with the batches already resident in HBM; ``e2e`` times the same steps through the public
trainer API (``Trainer_tactileSR.train_one_iter``) with pinned HOST batches, so the H2D copies of LR / HR and the D2H
read of the loss are inside the timed region.  ``--impl reference`` times the CPU oracle port of the reference
(PyTorch-CPU, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE_TRAIN = 43.916e9     # SURVEY.md section 8d, TactileSR S=1 forward+backward (2*MAC of the convs)
FLOP_PER_SAMPLE_FWD = 14.642e9
SR_CONFIG = dict(seqsCnt=1, axisCnt=3, HR_scale_num=10, scale_factor=10, patternFeatureExtraLayerCnt=6,
                 forceFeatureExtraLayerCnt=1, lr=1e-3, weight_decay=1e-2)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        if os.environ.get("TSR_BENCH_NO_CLOCKS") == "1":      # A/B switch: is the sampler itself perturbing the step?
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("TSR_BENCH_CLOCK_MS", "200"), "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batches(n, B, seed, pin, S=1):
    """LR (B,3S,4,4) in taxel units 0..8 and HR_raw (B,1,100,100) in 0..250."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        LR = torch.rand(B, 3 * S, 4, 4, generator=g) * 8
        HR = torch.rand(B, 1, 100, 100, generator=g) * 250
        if pin:
            LR, HR = LR.pin_memory(), HR.pin_memory()
        out.append((LR, HR))
    return out


REFERENCE_ROOT = os.environ.get("TACTILESR_REFERENCE", "/root/reference")


def cpu_reference_steps(B, steps, warmup, seed=42):
    """The reference's CPU training step on `steps` batches of B -> (samples/s, threads, kind).  kind "reference": the
    UNMODIFIED modules of /root/reference (model/tactileSR_model.py + stock torch.optim.Adam), used whenever that tree is
    mounted (the build container); kind "port": the oracle's restatement of the same ATen / oneDNN calls (the GPU box, where
    the reference tree does not exist)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    if os.path.exists(os.path.join(REFERENCE_ROOT, "model", "tactileSR_model.py")):
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            from model.tactileSR_model import TactileSR as RefSR      # noqa: E402
            torch.manual_seed(seed)
            m = RefSR().train()
            opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-2)
            times = []
            for LR, HR in synthetic_batches(steps + warmup, B, seed + 1, False):
                t0 = time.perf_counter()
                HRp = torch.nn.functional.interpolate(HR / 10, size=(40, 40), mode="bilinear", align_corners=False)
                loss = torch.nn.functional.mse_loss(m(LR), HRp)
                opt.zero_grad()
                loss.backward()
                opt.step()
                times.append(time.perf_counter() - t0)
            times = times[warmup:]
            return B * len(times) / sum(times), torch.get_num_threads(), "reference"
        finally:
            sys.path.remove(REFERENCE_ROOT)
    from oracle import tactilesr_oracle as so
    sd = so.make_state(so.tactilesr_layout(1), seed, nondegenerate=False)
    batches = synthetic_batches(steps + warmup, B, seed + 1, False)
    keys = so.param_keys(sd)
    m = {k: torch.zeros_like(sd[k]) for k in keys}
    v = {k: torch.zeros_like(sd[k]) for k in keys}
    times = []
    for t, (LR, HR) in enumerate(batches, start=1):
        t0 = time.perf_counter()
        loss, _, g, new_stats = so.loss_and_grads(sd, LR, HR, True)
        for k in keys:
            sd[k], m[k], v[k] = so.adam_step(sd[k], g[k], m[k], v[k], t, 1e-3, 1e-2)
        sd.update(new_stats)
        times.append(time.perf_counter() - t0)
    times = times[warmup:]
    return B * len(times) / sum(times), torch.get_num_threads(), "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 32
    # bounded sample: every step is one batch of the reference's own size (0.7 s on 16 cores), at most 20 steps
    steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 3))
    val, cores, kind = cpu_reference_steps(B, steps, warmup)
    line = {
        "impl": "reference", "metric": "SR train samples/sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * B / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "TactileSR(seqsCnt=1) train step: fwd + HR-prep/MSE + bwd + Adam(lr 1e-3, wd 1e-2), fp32, CPU",
                   "batch_per_step": B, "bounded_sample": f"{steps} steps of the reference's own batch size 32"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{steps} train steps of B=32 (reference config/default.py:46) after {warmup} warm-up"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_dominant_kernel(B, reps=10, f16=False):
    """conv_tc2_kernel<128> on the MSRB conv_5_2 shape (128->128, 5x5, 53.7 % of the forward FLOPs): algorithmic FLOPs per
    launch / CUDA-event time on the launching stream."""
    import torch
    from tactilesr_b200 import _lib
    dev = "cuda"
    dt = torch.float16 if f16 else torch.bfloat16
    x = torch.randn(B * 1600, 128, device=dev).to(dt)
    w = torch.randn(128, 128, 5, 5, device=dev) * 0.02
    wf = torch.empty(25 * 128 * 128, dtype=dt, device=dev)
    out = torch.empty(B * 1600, 128, dtype=dt, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("tsr_pack_conv_weight_f16" if f16 else "tsr_pack_conv_weight_bf16", w.data_ptr(), wf.data_ptr(), 0, 128, 128, 5, st)
    flags = _lib.TC2_F16 if f16 else 0

    def launch():
        _lib.conv_tc2([(x.data_ptr(), 128, 128, 5, wf.data_ptr())], out.data_ptr(), 128, B, 40, 40, 128, flags=flags, stream=st)
    for _ in range(3):
        launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * B * 1600 * 128 * 128 * 25
    return flops / (ms * 1e-3) / 1e12, ms, flops


def committed_traffic(B):
    """DRAM bytes per launch of the roofline kernel from the committed `ncu --set full` summary
    (profiles/roofline_traffic.json, written by tools/ncu_summary.py from the .ncu-rep of that capture) -- null when this
    batch size was never captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            rec = json.load(f)
        e = rec.get(str(B))
        return (e["dram_read_bytes"] + e["dram_write_bytes"], e["source"]) if e else (None, None)
    except Exception:
        return None, None


def stock_forward(m, x):
    """The reference's TactileSR.forward (model/tactileSR_model.py:67-84, MSRB :196-206, ResBlock :222-225) written with
    stock PyTorch ops over OUR module's parameter holders -- the yardstick "reference arithmetic through ATen / cuDNN on the
    same GPU".  Measurement only: the product path never calls it."""
    import torch
    import torch.nn.functional as F

    def seq(mods, t):
        for mod in mods:
            t = mod(t)
        return t

    def msrb(b, t):
        i2 = torch.cat([seq(b.conv_3_1, t), seq(b.conv_5_1, t)], 1)
        i3 = torch.cat([seq(b.conv_3_2, i2), seq(b.conv_5_2, i2)], 1)
        return F.relu(b.confusion(i3) + t)

    frames = [seq(m.inputLayer_pattern_list[s], x[:, 3 * s:3 * s + 3]) for s in range(m.seqsCnt)]
    p = seq(m.inputContact_layer, torch.cat(frames, 1))
    for b in m.patternFeatureExtra_layer:
        p = msrb(b, p)
    f = seq(m.input_layer_force, x[:, :3])
    for b in m.forceFeatureExtra_layer:
        f = F.relu(f + b.conv2(F.relu(b.conv1(f))))
    return seq(m.output_layer, torch.cat([f, p], 1))


def cudnn_yardstick(dev, timeit):
    """Stock PyTorch eager / cuDNN running the reference's training step on this GPU (fp32, TF32, autocast bf16) at the
    reference batch 32 and at 512: the number the hand-written path has to beat on the same hardware."""
    import torch
    from tactilesr_b200.model import TactileSR
    res = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for Bc in (32, 512):
            LR = torch.rand(Bc, 3, 4, 4, device=dev) * 8
            HR = torch.rand(Bc, 1, 100, 100, device=dev) * 250
            for tag in ("fp32", "tf32", "autocast_bf16"):
                torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tag != "fp32"
                torch.manual_seed(0)
                m = TactileSR().to(dev).train()
                opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-2)

                def step():
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=tag == "autocast_bf16"):
                        out = stock_forward(m, LR)
                    HRp = torch.nn.functional.interpolate(HR / 10, size=(40, 40), mode="bilinear", align_corners=False)
                    loss = torch.nn.functional.mse_loss(out.float(), HRp)
                    opt.zero_grad()
                    loss.backward()
                    opt.step()
                ms = timeit(step, 4 if Bc == 512 else 10)
                res[f"{tag}_b{Bc}_samples_per_s"] = Bc / (ms * 1e-3)
                del m, opt
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return res


def fp32_mode_numbers(dev, timeit):
    """The <= 1e-5 parity mode (FFMA2 implicit GEMM, csrc/conv_f32.cu) on the workloads of `cudnn_yardstick`: train step at
    the reference batch 32 and at 512, eval forward at 1024 (with stock cuDNN fp32 on the same forward beside it)."""
    import torch
    import tactilesr_b200 as tb
    from tactilesr_b200.functional import mse_hr_loss
    from tactilesr_b200.model import TactileSR
    from tactilesr_b200.optim import FusedAdam
    res = {}
    prev = tb.get_precision()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        tb.set_precision("fp32")
        torch.manual_seed(0)
        m = TactileSR().to(dev).train()
        opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-2)
        for Bc in (32, 512):
            LR = torch.rand(Bc, 3, 4, 4, device=dev) * 8
            HR = torch.rand(Bc, 1, 100, 100, device=dev) * 250

            def step():
                loss = mse_hr_loss(m(LR), HR, 10.0)
                opt.zero_grad()
                loss.backward()
                opt.step()
            ms = timeit(step, 3 if Bc == 512 else 10)
            res[f"train_b{Bc}_samples_per_s"] = Bc / (ms * 1e-3)
            res[f"train_b{Bc}_TFLOPs"] = Bc / (ms * 1e-3) * 3 * FLOP_PER_SAMPLE_FWD / 1e12
        m.eval()
        LR = torch.rand(1024, 3, 4, 4, device=dev) * 8
        with torch.no_grad():
            ms = timeit(lambda: m(LR), 3)
            res["eval_b1024_samples_per_s"] = 1024 / (ms * 1e-3)
            res["eval_b1024_TFLOPs"] = 1024 / (ms * 1e-3) * FLOP_PER_SAMPLE_FWD / 1e12
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
            ms = timeit(lambda: stock_forward(m, LR), 3)
            res["cudnn_fp32_eval_b1024_samples_per_s"] = 1024 / (ms * 1e-3)
        res["fp32_fma_peak_TFLOPs_measured"] = 73.2      # tools/microbench/ffma2_rate.cu on this pool's B200 (FFMA2, 1.9 GHz)
        del m, opt
    finally:
        tb.set_precision(prev)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return res


def measure_extras(dev, model, B):
    """Other configurations of BASELINE.json (reported, not the headline): SR inference (C3) and tPSFNet train (C2)."""
    import torch
    from tactilesr_b200.model import tPSFNet
    from tactilesr_b200.optim import FusedAdam
    out = {}

    def timeit(fn, reps):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    # C3: eval-mode forward, no_grad
    model.eval()
    Bi = max(B, 1024)
    LR = torch.rand(Bi, 3, 4, 4, device=dev) * 8
    with torch.no_grad():
        ms = timeit(lambda: model(LR), 5)
    model.train()
    out["sr_infer_samples_per_s"] = Bi / (ms * 1e-3)
    out["sr_infer_batch"] = Bi
    out["sr_infer_tensor_frac_of_sustained_peak"] = Bi / (ms * 1e-3) * FLOP_PER_SAMPLE_FWD / 1e12 / peaks()["tf_sust"]
    # C1's own batch size (config/default.py:46: 32): the iteration is launch-bound there; eager vs Trainer(cuda_graph=True)
    from tactilesr_b200.train.tactileSR_train import Trainer_tactileSR, build_model_and_optimizer
    for use_graph in (False, True):
        torch.manual_seed(0)
        m32, o32 = build_model_and_optimizer(SR_CONFIG, dev)
        d32 = [(torch.rand(32, 3, 4, 4, device=dev) * 8, torch.rand(32, 1, 100, 100, device=dev) * 250) for _ in range(4)]

        class L32:
            def __len__(self): return 4
            def __iter__(self):
                while True:
                    yield from d32
        t32 = Trainer_tactileSR(SR_CONFIG, model=m32, optimizer=o32, lr_scheduler=torch.optim.lr_scheduler.StepLR(o32, 2, 0.8),
                                data_loader=L32(), max_iters=10 ** 9, log_period=10 ** 9, device=dev, cuda_graph=use_graph)
        for _ in range(4):
            t32.train_one_iter()
        ms = timeit(t32.train_one_iter, 20)
        out["sr_train_b32_cuda_graph_samples_per_s" if use_graph else "sr_train_b32_eager_samples_per_s"] = 32 / (ms * 1e-3)
        del t32, m32, o32
    out["cudnn_yardstick"] = cudnn_yardstick(dev, timeit)
    out["fp32_mode"] = fp32_mode_numbers(dev, timeit)
    # TactileSRCNN (reference model/tactileSR_model.py:101-153; same kernels, different wiring): train step at the same batch
    from tactilesr_b200.functional import mse_hr_loss
    from tactilesr_b200.model import TactileSRCNN
    cnn = TactileSRCNN().to(dev).train()
    copt = FusedAdam(cnn.parameters(), lr=1e-3, weight_decay=1e-2)
    LRc = torch.rand(B, 3, 4, 4, device=dev) * 8
    HRc = torch.rand(B, 1, 100, 100, device=dev) * 250

    def cstep():
        loss = mse_hr_loss(cnn(LRc), HRc, 10.0)
        copt.zero_grad()
        loss.backward()
        copt.step()
    ms = timeit(cstep, 5)
    out["tactilesrcnn_train_samples_per_s"] = B / (ms * 1e-3)
    del cnn, copt
    # C2: tPSFNet train step: fwd + MSE(LR[:,2:3], LR_degrade) + bwd + Adam(1e-4, wd 1e-5); B = 256 is the reference
    # batch (config/default.py:18) -- 30 MB of traffic, launch-latency bound -- and B = 8192 shows the kernels
    g = torch.Generator().manual_seed(3)
    yy, xx = torch.meshgrid(torch.arange(100.0), torch.arange(100.0), indexing="ij")
    for Bp in (256, 8192):
        pm = tPSFNet(gama=1.4, perception_scale=None, device=dev).to(dev)
        popt = FusedAdam(pm.parameters(), lr=1e-4, weight_decay=1e-5)
        x = (torch.rand(Bp, 3, 4, 4, generator=g) * 13).to(dev)
        cx = torch.rand(Bp, generator=g) * 50 + 25
        cy = torch.rand(Bp, generator=g) * 50 + 25
        r = torch.rand(Bp, generator=g) * 20 + 10
        depth = torch.clamp((r[:, None, None] - ((yy - cy[:, None, None]) ** 2 + (xx - cx[:, None, None]) ** 2).sqrt()) / 2 + 0.5, 0, 1)
        depth = depth.unsqueeze(1).to(dev)

        def pstep():
            HR, LRd, _, _ = pm(x, depth)
            loss = torch.nn.functional.mse_loss(x[:, 2:3], LRd)
            popt.zero_grad()
            loss.backward()
            popt.step()
        ms = timeit(pstep, 5)
        tag = "" if Bp == 256 else f"_b{Bp}"
        out[f"tpsf_train_samples_per_s{tag}"] = Bp / (ms * 1e-3)
        with torch.no_grad():
            ms = timeit(lambda: pm(x, depth), 5)
        out[f"tpsf_fwd_samples_per_s{tag}"] = Bp / (ms * 1e-3)
        if Bp != 256:      # the same step with the MLP pinned to the fp32 kernels (<= 1e-5 mode; in the 16-bit modes the two
            pm.precision = "fp32"   # wide layers run on the tcgen05 conv kernels from B = 1024 up)
            ms = timeit(pstep, 5)
            out[f"tpsf_train_fp32_mode_samples_per_s{tag}"] = Bp / (ms * 1e-3)
            pm.precision = None
        if Bp == 256:      # the same step through Trainer_tPSF with the iteration captured in a CUDA graph
            from tactilesr_b200.train.tPSFNet_train import Trainer_tPSF
            pdata = [(x * 100, depth[:, 0])] * 2

            class LP:
                def __len__(self): return 2
                def __iter__(self):
                    while True:
                        yield from pdata
            pm2 = tPSFNet(gama=1.4, perception_scale=None, device=dev).to(dev)
            po2 = FusedAdam(pm2.parameters(), lr=1e-4, weight_decay=1e-5)
            tp = Trainer_tPSF(100, model=pm2, optimizer=po2, lr_scheduler=torch.optim.lr_scheduler.StepLR(po2, 1, 0.9),
                              data_loader=LP(), max_iters=10 ** 9, log_period=10 ** 9, device=dev, cuda_graph=True)
            for _ in range(4):
                tp.train_one_iter()
            ms = timeit(tp.train_one_iter, 20)
            out["tpsf_train_cuda_graph_samples_per_s"] = Bp / (ms * 1e-3)
            del tp, pm2, po2
        if Bp != 256:
            # the two tcgen05 PSF kernels alone (HBM-bound: compulsory 119 472 B / sample forward,
            # 40 000 (depth) + 4 848 (row statistics) + 76 B backward), against the measured HBM copy peak
            from tactilesr_b200 import _lib
            st = torch.cuda.current_stream().cuda_stream
            ab = torch.rand(Bp, 3, device=dev) * 0.5 + 0.5
            d3 = depth.reshape(Bp, 100, 100).contiguous()
            HR = torch.empty(Bp, 100, 100, device=dev); LRd = torch.empty(Bp, 16, device=dev); psf = torch.empty(Bp, 99, 99, device=dev)
            aux = torch.empty(Bp, int(_lib.lib().tsr_psf_aux_floats()), device=dev)
            dL = torch.rand(Bp, 16, device=dev); dab = torch.empty(Bp, 3, device=dev)
            ms_f = timeit(lambda: _lib.call("tsr_psf_forward_tc", ab.data_ptr(), d3.data_ptr(), HR.data_ptr(), LRd.data_ptr(),
                                            psf.data_ptr(), 0, Bp, st), 10)
            ms_b = timeit(lambda: _lib.call("tsr_psf_backward_tc", ab.data_ptr(), d3.data_ptr(), aux.data_ptr(), dL.data_ptr(),
                                            dab.data_ptr(), Bp, st), 10)
            out["psf_fwd_kernel"] = {"samples_per_s": Bp / (ms_f * 1e-3), "GBps_compulsory": Bp / (ms_f * 1e-3) * 119472 / 1e9,
                                     "hbm_frac": Bp / (ms_f * 1e-3) * 119472 / 1e9 / peaks()["hbm"], "batch": Bp}
            out["psf_bwd_kernel"] = {"samples_per_s": Bp / (ms_b * 1e-3), "GBps_compulsory": Bp / (ms_b * 1e-3) * 44924 / 1e9,
                                     "hbm_frac": Bp / (ms_b * 1e-3) * 44924 / 1e9 / peaks()["hbm"], "batch": Bp,
                                     "tensor_TFLOPs": Bp / (ms_b * 1e-3) * 4 * 3 * 2 * 128 * 112 * 112 / 1e12}
            # the single-pass fp16 variants the 16-bit precision modes use (same compulsory bytes, a third of the MMAs)
            ms_f = timeit(lambda: _lib.call("tsr_psf_forward_tc_f16", ab.data_ptr(), d3.data_ptr(), HR.data_ptr(), LRd.data_ptr(),
                                            psf.data_ptr(), 0, Bp, st), 10)
            ms_b = timeit(lambda: _lib.call("tsr_psf_backward_tc_f16", ab.data_ptr(), d3.data_ptr(), aux.data_ptr(), dL.data_ptr(),
                                            dab.data_ptr(), Bp, st), 10)
            out["psf_fwd_kernel_f16"] = {"samples_per_s": Bp / (ms_f * 1e-3), "GBps_compulsory": Bp / (ms_f * 1e-3) * 119472 / 1e9,
                                         "hbm_frac": Bp / (ms_f * 1e-3) * 119472 / 1e9 / peaks()["hbm"], "batch": Bp}
            out["psf_bwd_kernel_f16"] = {"samples_per_s": Bp / (ms_b * 1e-3), "GBps_compulsory": Bp / (ms_b * 1e-3) * 44924 / 1e9,
                                         "hbm_frac": Bp / (ms_b * 1e-3) * 44924 / 1e9 / peaks()["hbm"], "batch": Bp,
                                         "tensor_TFLOPs": Bp / (ms_b * 1e-3) * 4 * 2 * 128 * 112 * 112 / 1e12}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import tactilesr_b200 as tb
    from tactilesr_b200 import _lib
    from tactilesr_b200.cpu import distributed as D
    from tactilesr_b200.train.tactileSR_train import Trainer_tactileSR, build_model_and_optimizer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; tactilesr_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    rank, local_rank, world = D.init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.check(_lib.lib().tsr_check_device(), "device check")
    tb.set_precision(args.precision)
    B = args.batch
    if args.global_batch:                                   # strong scaling: the global batch is fixed, the per-GPU batch shrinks
        assert args.global_batch % world == 0
        B = args.global_batch // world
    S = args.seqs                                           # 1 = the headline workload; 7 = the tactileSRSeqs model (C4)
    global FLOP_PER_SAMPLE_TRAIN, FLOP_PER_SAMPLE_FWD
    if S == 7:
        FLOP_PER_SAMPLE_TRAIN, FLOP_PER_SAMPLE_FWD = 48.229e9, 16.091e9
    sr_config = dict(SR_CONFIG, seqsCnt=S)
    torch.manual_seed(42)                                   # identical weights on every rank
    model, opt = build_model_and_optimizer(sr_config, dev)
    nb = 4                                                  # rotating distinct batches
    host = synthetic_batches(nb, B, 1000 + rank, pin=True, S=S)
    devb = [(a.to(dev), b.to(dev)) for a, b in host]

    class Loader:
        def __init__(self, items): self.items = items
        def __len__(self): return len(self.items)
        def __iter__(self):
            while True:
                for it in self.items:
                    yield it

    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.8)
    tr = Trainer_tactileSR(sr_config, model=model, optimizer=opt, lr_scheduler=sched, data_loader=Loader(devb),
                           max_iters=10 ** 9, log_period=10 ** 9, device=dev, cuda_graph=args.cuda_graph)
    tr._setup_dp()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(tr, steps, sync_loss):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            tr.train_one_iter()
            if sync_loss:
                # end-to-end: the user reads the loss of every step back to the host (reference cpu/trainer.py:259)
                _ = float(tr._loss_acc.item())
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _lib.launch_count() - l0

    for _ in range(max(args.warmup, 3)):
        tr.train_one_iter()
    with ClockSampler(local_rank) as cs:
        ms_dev, launches = timed(tr, args.steps, False)
    clocks = cs.summary()
    # same trainer, batches now start in pinned host memory; the next batch's H2D copy runs on a side stream
    from tactilesr_b200.data import DevicePrefetcher
    tr._data_iter = iter(DevicePrefetcher(Loader(host), dev))
    for _ in range(2):
        tr.train_one_iter()
    ms_e2e, _ = timed(tr, args.steps, True)

    # per-kernel-class breakdown of one device-resident step (CUDA events around every op of the layer program)
    breakdown = None
    if world == 1:
        from tactilesr_b200 import engine as E
        tr._data_iter = iter(Loader(devb))
        use_graph, tr._use_graph = tr._use_graph, False      # (the breakdown times the ops of an eager iteration)
        tr.train_one_iter()
        torch.cuda.synchronize()
        E.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr.train_one_iter()
        e1.record()
        cls = E.profile_end()
        breakdown = {k: round(v, 3) for k, v in sorted(cls.items(), key=lambda kv: -kv[1])}
        breakdown["step_total_with_event_overhead"] = round(e0.elapsed_time(e1), 3)
        tr._use_graph = use_graph

    extras = {}
    if world == 1 and not args.no_extras:
        try:                                  # reported extras must never take the headline line down with them
            extras = measure_extras(dev, model, B)
        except Exception as e:                # noqa: BLE001
            extras = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.synchronize()
        if args.precision != "bf16":          # the all-bf16 mode on the same trainer, for comparison
            tb.set_precision("bf16")
            tr._data_iter = iter(Loader(devb))
            for _ in range(3):
                tr.train_one_iter()
            ms_b, _ = timed(tr, min(args.steps, 5), False)
            extras["bf16_mode_train_samples_per_s"] = B * min(args.steps, 5) / (ms_b * 1e-3)
            tb.set_precision(args.precision)

    total = B * world * args.steps
    value = total / (ms_dev * 1e-3)
    e2e = total / (ms_e2e * 1e-3)
    if rank != 0:
        return
    pk = peaks()
    roof = None
    if args.precision in ("bf16", "fp16"):
        tf, kms, kflops = time_dominant_kernel(B, f16=args.precision == "fp16")
        traffic, traffic_src = committed_traffic(B)
        roof = {"bound": "tensor", "kernel": "conv_tc2_kernel<128> (cta_group::2) 5x5 128->128, MSRB conv_5_2 forward shape",
                "achieved": tf, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": tf / pk["tf_burst"],
                "peak_source": pk["src"] + " bf16 burst (kernel timed alone)", "ms_per_launch": kms, "flops_per_launch": kflops,
                "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": B * 1600 * (128 + 128) * 2 + 25 * 128 * 128 * 2}
    else:
        roof = {"bound": "tensor", "kernel": "whole fp32 step (FFMA implicit GEMM; fp32-accurate parity mode)",
                "achieved": value / world * FLOP_PER_SAMPLE_TRAIN / 1e12, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": value / world * FLOP_PER_SAMPLE_TRAIN / 1e12 / pk["tf_sust"], "peak_source": pk["src"] + " bf16 sustained", "traffic": None}
    cpu = None
    if world == 1 and not args.no_cpu_baseline and S == 1:
        v, cores, kind = cpu_reference_steps(32, 12, 1)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": kind,
               "sample": "12 train steps of B=32 (reference batch size, config/default.py:46) after 1 warm-up; "
                         + ("unmodified reference modules" if kind == "reference" else "oracle port of the reference CPU path")}
    line = {
        "metric": "SR train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.global_batch else "weak",
        "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "fp16", "fp32": "f32"}[args.precision], "data": "synthetic",
        "config": {"workload": f"TactileSR(seqsCnt={S}) train step: fwd + fused HR-prep/MSE + bwd + fused Adam(lr 1e-3, wd 1e-2)"
                               + (" + NCCL gradient all-reduce" if world > 1 else ""),
                   "per_gpu_batch": B, "global_batch": B * world, "input": f"LR (B,{3 * S},4,4), HR (B,1,100,100)",
                   "precision_mode": args.precision + {"fp16": " (fp16 activations / forward weights, bf16 gradients, fp32 accumulation and statistics)",
                                                       "bf16": " (bf16 storage, fp32 accumulation and statistics)", "fp32": ""}[args.precision],
                   "parallelism": f"dp{world}", "cuda_graph": bool(args.cuda_graph),
                   "grad_allreduce": ("bucketed, overlapped with backward" if os.environ.get("TSR_DP_OVERLAP", "0") == "1" else "one NCCL call after backward") if world > 1 else None,
                   "l2": "per-step working set (saved activations, B x ~26-52 MB) >> 126 MB L2; 4 rotating input batches"},
        "tensor_roofline_frac_step": value / world * FLOP_PER_SAMPLE_TRAIN / 1e12 / pk["tf_sust"],
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": B * (3 * S * 16 + 100 * 100) * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks, "step_breakdown_ms": breakdown, "extra": extras,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # default = the tensor-core mode that meets north_star's <= 1e-2 bound on the SR output (fp16 activations, fp32
    # accumulation, bf16 gradients); "bf16" is ~5 % faster and 8x less accurate, "fp32" is the <= 1e-5 parity mode
    ap.add_argument("--precision", default=os.environ.get("TSR_BENCH_PRECISION", "fp16"), choices=["fp32", "bf16", "fp16"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("TSR_BENCH_BATCH", "1024")))
    ap.add_argument("--seqs", type=int, default=int(os.environ.get("TSR_BENCH_SEQS", "1")), choices=[1, 7],
                    help="frames per sample: 1 = headline workload, 7 = tactileSRSeqs model (BASELINE.json configs[3])")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: fixed global batch, per-GPU batch = this / N")
    ap.add_argument("--cuda-graph", action="store_true", help="Trainer(cuda_graph=True): replay the iteration as a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
