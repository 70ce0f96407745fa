"""Host logic of the layer-program engine WITHOUT a GPU: every kernel launch of the C ABI is replaced by a recorder (the
library's pure-host query entry points -- workspace sizes, table rows -- still run), tensors live on the CPU, and the test
checks WHICH launches a TactileSR training step plans in each precision mode.  The expected counts are those of the ncu
launch list of the same step on a B200 (profiles/r02_launches_fp16_b1024.txt): they pin the fusions the engine is
responsible for (dual-branch forward, K-concatenated data gradients with ReLU / BatchNorm-backward epilogues, one
weight-pack launch per step, statistics out of the conv epilogues) against silent regressions, and the packed-weight cache
(reference model/tactileSR_model.py:67-84, 196-206 is what the program lowers).  Nothing here computes anything: numerical
parity lives in the -m gpu tests."""
import collections

import pytest
import torch


@pytest.fixture
def recorder(monkeypatch):
    import tactilesr_b200 as tb
    from tactilesr_b200 import _lib, engine
    calls = collections.Counter()
    _lib.lib()                                            # the real library must load (its query functions are used)
    monkeypatch.setattr(_lib, "call", lambda name, *a: calls.update([name]))
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
    monkeypatch.setattr(engine, "_check_device", lambda x: None)
    prev = tb.get_precision()
    yield calls
    tb.set_precision(prev)


def _step(mode, calls, S=1, B=4):
    import tactilesr_b200 as tb
    from tactilesr_b200.model import TactileSR
    tb.set_precision(mode)
    torch.manual_seed(0)
    m = TactileSR(seqsCnt=S).train()
    x = torch.rand(B, 3 * S, 4, 4) * 8
    calls.clear()
    out = m(x)
    fwd = dict(calls)
    calls.clear()
    out.backward(torch.ones_like(out))
    return m, x, fwd, dict(calls)


@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_tensor_core_step_plans_the_fused_launches(recorder, mode):
    _, _, fwd, bwd = _step(mode, recorder)
    # forward: 6 dual-branch 64-channel stages + 12 128-channel convs + 6 confusion convs + contact / force stacks = 29
    # tcgen05 launches, BatchNorm statistics from their epilogues (one separate statistics pass: the head output), ONE
    # weight-pack launch, no standalone ReLU / residual / concat launches
    assert fwd["tsr_conv2d_tc2"] == 29
    assert fwd["tsr_bn_train_stats"] == 1 and fwd["tsr_bn_apply"] == 15
    assert fwd["tsr_pack_conv_weights_multi"] == 1
    assert fwd["tsr_head_fwd"] == 2 and fwd["tsr_tail_fwd"] == 1
    assert fwd.get("tsr_copy_channels", 0) == (1 if mode == "fp16" else 0)        # bf16 shadow of the head output only
    assert not any(k in fwd for k in ("tsr_relu_backward", "tsr_conv2d_f32", "tsr_conv2d_tc"))
    # backward: one weight gradient per conv (35), 23 data-gradient launches for 35 convs (the two branches of an MSRB stage
    # share one K-concatenated launch; the first layers need none), ReLU backward fused everywhere (no standalone launch),
    # BatchNorm backward level 1 from the data-gradient epilogues except for the 12 halves behind the 1x1 data gradient
    assert bwd["tsr_conv2d_wgrad_tc_x"] == 35
    assert bwd["tsr_conv2d_tc2"] == 23
    assert "tsr_relu_backward" not in bwd
    assert bwd["tsr_bn_backward"] == 12 and bwd["tsr_bn_backward_apply"] == 9 and bwd["tsr_bn_bwd_finalize_partials"] == 15
    assert bwd["tsr_tail_dgrad_masked"] == 1 and "tsr_tail_dgrad" not in bwd
    assert bwd["tsr_head_wgrad"] == 2 and bwd["tsr_tail_wgrad"] == 1
    assert bwd["tsr_colsum"] == 8                        # only the biases that do not feed a batch-statistics BatchNorm
    # 52 tcgen05 conv launches + 35 weight gradients per step: the numbers of profiles/r02_launches_fp16_b1024.txt
    assert fwd["tsr_conv2d_tc2"] + bwd["tsr_conv2d_tc2"] == 26 + 20 + 6


def test_fp32_step_uses_only_the_fp32_kernels(recorder):
    _, _, fwd, bwd = _step("fp32", recorder)
    assert fwd["tsr_conv2d_f32"] == 35 and bwd["tsr_conv2d_f32"] == 35 and bwd["tsr_conv2d_wgrad_f32"] == 35
    assert not any(k.startswith(("tsr_conv2d_tc", "tsr_conv2d_wgrad_tc", "tsr_pack_conv_weights_multi")) for k in {**fwd, **bwd})


def test_sequence_model_plans_one_head_per_frame(recorder):
    _, _, fwd, bwd = _step("fp16", recorder, S=7)
    assert fwd["tsr_head_fwd"] == 8 and bwd["tsr_head_wgrad"] == 8       # 7 pattern frames + the force branch
    assert bwd["tsr_conv2d_wgrad_tc_x"] == 35 + 6                        # one more 64-channel conv per additional frame


def test_packed_weight_cache_and_invalidation(recorder):
    """Packed copies are rebuilt only when a weight changed: not on a second forward, again after an in-place update that
    bumps Tensor._version, and -- for writes through .data, which bump nothing -- after invalidate_packed_weights()."""
    import tactilesr_b200 as tb
    m, x, fwd, _ = _step("fp32", recorder)
    assert fwd["tsr_pack_conv_weight_f32"] == 35
    m.eval()
    with torch.no_grad():
        recorder.clear(); m(x); first = recorder.get("tsr_pack_conv_weight_f32", 0) + recorder.get("tsr_pack_conv_weight_folded", 0)
        recorder.clear(); m(x)
        assert recorder.get("tsr_pack_conv_weight_f32", 0) + recorder.get("tsr_pack_conv_weight_folded", 0) == 0
        w = m.patternFeatureExtra_layer[0].conv_3_2[0].weight
        w.mul_(0.5)                                       # bumps _version: exactly this conv is repacked
        recorder.clear(); m(x)
        assert recorder.get("tsr_pack_conv_weight_f32", 0) + recorder.get("tsr_pack_conv_weight_folded", 0) == 1
        w.data.mul_(2.0)                                  # bumps nothing: stale copy until the cache is invalidated
        recorder.clear(); m(x)
        assert recorder.get("tsr_pack_conv_weight_f32", 0) + recorder.get("tsr_pack_conv_weight_folded", 0) == 0
        tb.invalidate_packed_weights()
        recorder.clear(); m(x)
        assert recorder.get("tsr_pack_conv_weight_f32", 0) + recorder.get("tsr_pack_conv_weight_folded", 0) >= 1
    assert first >= 0


def test_parameter_hook_is_rejected_before_any_launch(recorder):
    import tactilesr_b200 as tb
    from tactilesr_b200.model import TactileSR
    m = TactileSR().train()
    m.output_layer[0].weight.register_hook(lambda g: g)
    recorder.clear()
    with pytest.raises(tb.TsrError):
        m(torch.rand(2, 3, 4, 4))
    assert sum(recorder.values()) == 0
