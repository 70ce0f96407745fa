import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import tpsf_oracle as po
from tactilesr_b200 import _lib
B = 5
g = torch.Generator().manual_seed(B)
ab = torch.stack([torch.rand(B, generator=g) * 2 + 0.2, torch.rand(B, generator=g) * 3 + 0.25,
                  torch.rand(B, generator=g) * 40 + 0.4], 1).cuda().contiguous()
depth = po.synthetic_depth(min(B, 16), 5).repeat((B + 15) // 16, 1, 1)[:B].clone()
depth[1] *= 3.0
depth[2] = 0.0
depth = depth.cuda().contiguous()
st = torch.cuda.current_stream().cuda_stream
outs = {}
for name in ("tsr_psf_forward_ffma", "tsr_psf_forward_tc"):
    HR = torch.empty(B, 100, 100, device="cuda"); LRd = torch.empty(B, 16, device="cuda"); psf = torch.empty(B, 99, 99, device="cuda")
    _lib.call(name, ab.data_ptr(), depth.data_ptr(), HR.data_ptr(), LRd.data_ptr(), psf.data_ptr(), B, st)
    outs[name] = (HR, LRd, psf)
(h0, l0, p0), (h1, l1, p1) = outs["tsr_psf_forward_ffma"], outs["tsr_psf_forward_tc"]
# fp64 reference of the separable form
for b in range(B):
    a, be, ga = ab[b].double().tolist()
    t = torch.arange(100, dtype=torch.float64, device="cuda")
    E = torch.exp(-(100.0 / 4802.0) * (t[:, None] - t[None, :]) ** 2 / be ** 2) * ((t[:, None] - t[None, :]).abs() <= 49)
    c = a * E @ depth[b].double() @ E
    mask = depth[b] > depth[b].max() - 1e-3
    m2 = torch.where(mask, torch.zeros_like(c), c).max().clamp_min(0)
    ref = torch.where(mask, m2, c)
    s = ref.abs().max().clamp_min(1e-20)
    d0 = (h0[b].double() - ref).abs(); d1 = (h1[b].double() - ref).abs()
    i1 = d1.argmax().item()
    print(b, "ab", [round(x, 3) for x in (a, be, ga)], "max", s.item(), "ffma err", (d0.max() / s).item(), "tc err", (d1.max() / s).item(),
          "at", divmod(i1, 100), "contact there", bool(mask.flatten()[i1]), "m2", m2.item(), "tc m2?", h1[b].flatten()[i1].item(), "ref", ref.flatten()[i1].item())
