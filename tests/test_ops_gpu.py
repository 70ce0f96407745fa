"""torch.ops.tactilesr.*: schema / fake-tensor / autograd-registration checks (torch.library.opcheck) of the custom ops
that wrap the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_custom_ops_pass_opcheck():
    import tactilesr_b200.ops  # noqa: F401
    from oracle import tpsf_oracle as po
    tests = ("test_schema", "test_faketensor", "test_autograd_registration")
    out = torch.rand(3, 1, 40, 40, device="cuda", requires_grad=True)
    hr = torch.rand(3, 1, 100, 100, device="cuda") * 250
    torch.library.opcheck(torch.ops.tactilesr.mse_hr_loss, (out, hr, 10.0), test_utils=tests)
    torch.library.opcheck(torch.ops.tactilesr.eval_metrics, (out.detach(), hr, 10.0, 250.0, 1e-4, 9e-4), test_utils=tests)
    ab = (torch.rand(3, 3, device="cuda") + 0.5).requires_grad_(True)
    depth = po.synthetic_depth(3, 2).cuda().unsqueeze(1)
    for want_aux in (True, False):
        torch.library.opcheck(torch.ops.tactilesr.psf_model, (ab, depth, want_aux), test_utils=tests)
    # gradient of the op against the FFMA backward through a second route (aux-less => general backward)
    HR, LRd, psf, aux = torch.ops.tactilesr.psf_model(ab, depth, True)
    g1, = torch.autograd.grad(LRd.square().sum(), ab)
    HR2, LRd2, _, _ = torch.ops.tactilesr.psf_model(ab, depth, False)
    g2, = torch.autograd.grad(LRd2.square().sum(), ab)
    assert torch.equal(LRd, LRd2) and ((g1 - g2).norm() / g2.norm()).item() < 1e-4
