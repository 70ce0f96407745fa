"""Host-side cost of one training iteration WITHOUT a GPU: every kernel launch is stubbed out (the workspace / query entry
points of the library are pure host code and still run), tensors live on the CPU, so what remains is exactly the Python +
ctypes-marshalling work the engine does per iteration.  python tools/host_profile.py [B] [mode] [--prof]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tactilesr_b200 as tb
from tactilesr_b200 import _lib, engine

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
mode = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else "fp16"
ncalls = [0]
real_lib = _lib.lib()


def fake_call(name, *args):
    ncalls[0] += 1
    fn = getattr(real_lib, name)        # same attribute lookup + argument marshalling cost class as the real call
    return None


_lib.call = fake_call
_lib.stream_ptr = lambda: 0
engine._check_device = lambda x: None
tb.set_precision(mode)
from tactilesr_b200.functional import mse_hr_loss
from tactilesr_b200.model import TactileSR
from tactilesr_b200.optim import FusedAdam
import inspect, textwrap
import tactilesr_b200.optim as optim_mod
# FusedAdam refuses CPU parameters (no fallback): for this host-only measurement re-create the method without that check
src = textwrap.dedent(inspect.getsource(FusedAdam._flatten_group)).replace('if dev.type != "cuda":', 'if False:')
ns = {}
exec(src, vars(optim_mod), ns)
FusedAdam._flatten_group = ns["_flatten_group"]

torch.manual_seed(0)
m = TactileSR().train()
opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-2)
LR = torch.rand(B, 3, 4, 4) * 8
G = torch.ones(B, 1, 40, 40)


def step():
    out = m(LR)
    opt.zero_grad()
    out.backward(G)              # (the fused loss is a cuda-only custom op: one more library call, not profiled here)
    opt.step()


for _ in range(3):
    step()
n0 = ncalls[0]
t0 = time.perf_counter()
N = 20
for _ in range(N):
    step()
dt = (time.perf_counter() - t0) / N
print(f"host time per iteration (B={B}, {mode}): {dt*1e3:.2f} ms, {(ncalls[0]-n0)//N} library calls")
if "--prof" in sys.argv:
    pr = cProfile.Profile(); pr.enable()
    for _ in range(10):
        step()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
