"""Per-tap error of the bf16 mode vs the fp64 golden, next to stock PyTorch autocast(bf16) on the same GPU (yardstick)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tactilesr_b200 as tb
from tactilesr_b200 import engine as E
from tactilesr_b200.model import TactileSR
from oracle import tactilesr_oracle as so
from tests.util import load_golden, sr_inputs, summarize, rel_l2

for S in (1,):
    g = load_golden(f"tactilesr_fwdbwd_s{S}.npz")
    sd = so.make_state(so.tactilesr_layout(S), int(g["seed_w"]))
    LR, HR = sr_inputs(int(g["B"]), S, int(g["seed_x"]))
    # oracle fp64 taps on CPU (full tensors)
    taps64 = {}
    out64 = so.tactilesr_forward({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, LR.double(), True, taps=taps64)
    # torch autocast bf16 on GPU
    sdg = {k: v.cuda() for k, v in sd.items()}
    tapsa = {}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outa = so.tactilesr_forward(sdg, LR.cuda(), True, taps=tapsa)
    for mode in ("fp32", "bf16"):
        tb.set_precision(mode)
        m = TactileSR(seqsCnt=S); m.load_state_dict(sd); m = m.cuda().train()
        prog = m._program()
        out, c = E.run_forward(prog, LR.cuda(), True, False, mode, keep_taps=True)
        print(f"== mode {mode}: out rel-L2 vs f64 {rel_l2(out, out64):.3e}   (autocast-bf16 yardstick {rel_l2(outa.float(), out64):.3e})")
        for k in ["inputContact"] + [f"msrb{i}" for i in range(6)] + ["force", "output0"]:
            v = prog.taps[k]
            t = c.bufs[v.buf][:, v.c0:v.c0 + v.C].float().view(c.B, c.H, c.W, v.C).permute(0, 3, 1, 2)
            print(f"   {k:14s} ours {rel_l2(t, taps64[k]):.3e}   autocast {rel_l2(tapsa[k].float(), taps64[k]):.3e}")
# default init, larger batch: features of the last MSRB
torch.manual_seed(42)
m0 = TactileSR()
sd0 = m0.state_dict()
LR, _ = sr_inputs(64, 1, 5)
res = {}
for mode in ("fp32", "bf16"):
    tb.set_precision(mode)
    m = TactileSR(); m.load_state_dict(sd0); m = m.cuda().train()
    prog = m._program()
    out, c = E.run_forward(prog, LR.cuda(), True, False, mode, keep_taps=True)
    v = prog.taps["msrb5"]
    res[mode] = (c.bufs[v.buf][:, v.c0:v.c0 + v.C].float().clone(), out.clone())
print("default init B=64: msrb5 rel-L2 bf16 vs fp32:", rel_l2(res["bf16"][0], res["fp32"][0]), " out:", rel_l2(res["bf16"][1], res["fp32"][1]), "nonzero frac", (res["fp32"][1] > 0).float().mean().item())
