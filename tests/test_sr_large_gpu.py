"""Module-level parity at the configurations the bench and BASELINE.json actually run:

  * C1 exactly -- the reference's own seed-42 construction, train mode, B = 32, LR = rand * 8, HR = rand * 250 -- in the
    fp32 and fp16 modes, against the fp32 / fp64 runs of the UNMODIFIED reference (tests/golden/tactilesr_c1_b32.npz);
  * one training step at B = 256 (every persistent tensor-core kernel loops ~9 times per CTA pair: both TMEM accumulator
    buffers, ring wrap across blocks, BatchNorm partials accumulated across blocks, the fused data-gradient epilogues) in
    all three modes against the fp64 reference run (tests/golden/tactilesr_b256.npz);
  * TactileSRCNN forward + backward (reference model/tactileSR_model.py:101-153).

Fixtures are minted by oracle/make_golden.py --round2 from /root/reference.  Bounds: fp32 mode 2e-5 on taps / statistics
(north_star 1e-5 is on the final output vs the fp32 reference run), fp16 mode 1e-2 (north_star's tensor-core bound);
gradients are compared through their l2 norm and 16 sampled elements against max(bound, 6 x the reference's own
fp32-vs-fp64 deviation)."""
import copy

import numpy as np
import pytest
import torch

from tests.util import load_golden, rel_l2, sr_inputs, summarize, summary_close

pytestmark = pytest.mark.gpu

TAPS = ["inputContact", "force", "output0"] + [f"msrb{i}" for i in range(6)]


def _tap(c, view):
    t = c.bufs[view.buf][:, view.c0:view.c0 + view.C].float()
    return t.view(c.B, c.H, c.W, view.C).permute(0, 3, 1, 2).contiguous()


def _check_step(m, g, LR, HR_raw, mode, tol_tap, tol_loss, tol_grad, tol_bn):
    import tactilesr_b200 as tb
    from tactilesr_b200 import engine as E
    from tactilesr_b200.functional import mse_hr_loss
    tb.set_precision(mode)
    try:
        LR, HR_raw = LR.cuda(), HR_raw.cuda()
        m2 = copy.deepcopy(m)
        prog = m2._program()
        _, c = E.run_forward(prog, LR, True, False, mode, keep_taps=True)
        worst_tap = 0.0
        for k in TAPS:
            ok, err = summary_close(summarize(_tap(c, prog.taps[k])), g[f"f64/tap/{k}"], tol_tap)
            assert ok, (mode, k, err)
            worst_tap = max(worst_tap, max(err))
        del c, m2
        out = m(LR)
        loss = mse_hr_loss(out, HR_raw, 10.0)
        assert abs(loss.item() - float(g["f64/loss"])) / float(g["f64/loss"]) < tol_loss, (mode, loss.item(), float(g["f64/loss"]))
        ok, err = summary_close(summarize(out), g["f64/out_summary"], max(tol_tap, 6 * summary_close(g["f32/out_summary"], g["f64/out_summary"], 1.0)[1][0]))
        assert ok, (mode, "out", err)
        assert abs(float((out > 0).double().mean()) - float(g["f64/out_nonzero"])) < 2e-3
        loss.backward()
        names = [str(x) for x in g["param_names"]]
        assert [n for n, _ in m.named_parameters()] == names
        worst_g = 0.0
        for (n, p), want, ref32 in zip(m.named_parameters(), g["f64/grad_summary"], g["f32/grad_summary"]):
            got = summarize(p.grad)
            if want[0] < 1e-9 * max(1.0, float(g["f64/loss"])):
                continue                                   # conv biases before a train-mode BatchNorm: true gradient 0
            _, ref_err = summary_close(ref32, want, 1.0)
            _, err = summary_close(got, want, 1.0)
            assert err[0] < max(tol_grad, 6 * ref_err[0]), (mode, n, err, ref_err)
            assert err[1] < max(4 * tol_grad, 6 * ref_err[1]), (mode, n, err, ref_err)
            worst_g = max(worst_g, err[0])
        sd = m.state_dict()
        for n, want in zip([str(x) for x in g["bn_names"]], g["f64/bn_summary"]):
            ok, err = summary_close(summarize(sd[n]), want, tol_bn)
            assert ok, (mode, n, err)
        print(f"{mode}: worst tap error {worst_tap:.2e}, worst gradient-norm error {worst_g:.2e}")
    finally:
        tb.set_precision("fp32")


@pytest.mark.parametrize("mode,tol_tap,tol_loss,tol_grad,tol_bn", [("fp32", 2e-5, 1e-5, 5e-3, 2e-5), ("fp16", 1e-2, 1e-2, 5e-2, 1e-2)])
def test_c1_reference_configuration(mode, tol_tap, tol_loss, tol_grad, tol_bn):
    from tactilesr_b200.model import TactileSR
    g = load_golden("tactilesr_c1_b32.npz")
    torch.manual_seed(int(g["seed_init"]))
    m = TactileSR().cuda().train()                      # the reference's own seed-42 initialisation (config/default.py:10)
    LR, HR_raw = sr_inputs(int(g["B"]), 1, int(g["seed_x"]))
    _check_step(m, g, LR, HR_raw, mode, tol_tap, tol_loss, tol_grad, tol_bn)
    if mode == "fp32":
        out = m.eval()  # noqa: F841  (module left in a defined state)


@pytest.mark.parametrize("mode,tol_tap,tol_loss,tol_grad,tol_bn", [("fp32", 2e-5, 1e-5, 5e-3, 2e-5), ("fp16", 1e-2, 1e-2, 5e-2, 1e-2),
                                                                   ("bf16", 6e-2, 6e-2, 1.5e-1, 6e-2)])
def test_b256_training_step(mode, tol_tap, tol_loss, tol_grad, tol_bn):
    """bf16 is reported, not claimed: its bound is the 5e-2 class of stock autocast(bf16) on these weights."""
    from oracle import tactilesr_oracle as so
    from tactilesr_b200.model import TactileSR
    g = load_golden("tactilesr_b256.npz")
    m = TactileSR()
    m.load_state_dict(so.make_state(so.tactilesr_layout(1), int(g["seed_w"])), strict=True)
    m = m.cuda().train()
    LR, HR_raw = sr_inputs(int(g["B"]), 1, int(g["seed_x"]))
    _check_step(m, g, LR, HR_raw, mode, tol_tap, tol_loss, tol_grad, tol_bn)


@pytest.mark.parametrize("mode,tol_out,tol_grad", [("fp32", 2e-5, 5e-3), ("fp16", 1e-2, 5e-2)])
def test_srcnn_forward_backward(mode, tol_out, tol_grad):
    import tactilesr_b200 as tb
    from oracle import tactilesr_oracle as so
    from tactilesr_b200.functional import mse_hr_loss
    from tactilesr_b200.model import TactileSRCNN
    g = load_golden("tactilesrcnn_bwd.npz")
    tb.set_precision(mode)
    try:
        m = TactileSRCNN()
        m.load_state_dict(so.make_state(so.tactilesrcnn_layout(), int(g["seed_w"])), strict=True)
        m = m.cuda().train()
        LR, HR_raw = sr_inputs(int(g["B"]), 1, int(g["seed_x"]))
        out = m(LR.cuda())
        assert rel_l2(out, g["f64/out"]) < tol_out
        loss = mse_hr_loss(out, HR_raw.cuda(), 10.0)
        assert abs(loss.item() - float(g["f64/loss"])) / float(g["f64/loss"]) < max(tol_out, 1e-5)
        loss.backward()
        for (n, p), want, ref32 in zip(m.named_parameters(), g["f64/grad_summary"], g["f32/grad_summary"]):
            if want[0] < 1e-9:
                continue
            _, ref_err = summary_close(ref32, want, 1.0)
            _, err = summary_close(summarize(p.grad), want, 1.0)
            assert err[0] < max(tol_grad, 6 * ref_err[0]), (n, err, ref_err)
        # two whole gradient tensors: the last layer and the first (three BatchNorm backward passes and six MSRBs upstream).
        # In the fp16 mode gradients are bf16 tensors; the yardstick there is what stock PyTorch autocast(bf16) -- the
        # reference stack's own reduced-precision mode -- gets on the same weights on this GPU, as in tests/test_sr_bf16_gpu.py
        yard = {}
        if mode != "fp32":
            sd = so.make_state(so.tactilesrcnn_layout(), int(g["seed_w"]))
            keys = set(so.param_keys(sd))
            leaf = {k: v.cuda().requires_grad_(k in keys) for k, v in sd.items()}
            with torch.autocast("cuda", dtype=torch.bfloat16):
                o = so.tactilesrcnn_forward(leaf, LR.cuda(), True)
            l = torch.mean((o.float() - so.prep_hr(HR_raw.cuda(), 10.0, 40)) ** 2)
            l.backward()
            yard = {"output.0.weight": rel_l2(leaf["output.0.weight"].grad, g["f64/grad_output_w"]),
                    "input_zyx.0.weight": rel_l2(leaf["input_zyx.0.weight"].grad, g["f64/grad_in0_w"])}
        for name, key in (("output.0.weight", "grad_output_w"), ("input_zyx.0.weight", "grad_in0_w")):
            got = dict(m.named_parameters())[name].grad
            ref_err = rel_l2(g[f"f32/{key}"], g[f"f64/{key}"])
            e = rel_l2(got, g[f"f64/{key}"])
            print(f"TactileSRCNN {mode} {name}: grad rel-L2 {e:.3e} (reference fp32 {ref_err:.1e}, autocast-bf16 yardstick {yard.get(name, float('nan')):.3e})")
            assert e < max(tol_grad, 6 * ref_err, 1.5 * yard.get(name, 0.0)), (name, e, ref_err, yard)
    finally:
        tb.set_precision("fp32")


def test_wrong_taxel_shape_is_rejected():
    from tactilesr_b200 import TsrError
    from tactilesr_b200.model import TactileSR
    m = TactileSR().cuda()
    with pytest.raises(TsrError):
        m(torch.zeros(2, 3, 8, 8, device="cuda"))
