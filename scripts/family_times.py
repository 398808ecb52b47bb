"""Per-C-ABI-family device time of one training step (CUDA events around every call), with the weight-gradient side stream
off so that nothing overlaps: the complete list, sorted, plus what is NOT inside any C-ABI call (torch glue, gaps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import autovc_b200
from autovc_b200 import _lib, ops, solver

dev = torch.device("cuda", 0)
prec = sys.argv[1] if len(sys.argv) > 1 else "half"
torch.manual_seed(0)
G = autovc_b200.Generator(16, 256, 512, 16, precision=prec).to(dev).train()
opt = autovc_b200.FusedAdam(G.parameters(), 1e-4)
g = torch.Generator().manual_seed(1)
x = torch.rand(256, 128, 80, generator=g).to(dev)
e = (F.normalize(torch.randn(256, 256, generator=g), dim=-1) * 0.8).to(dev)
ops._WGRAD["on"] = False
for _ in range(5):
    solver.train_step(G, opt, x, e, sync_losses=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    solver.train_step(G, opt, x, e, sync_losses=False)
e1.record()
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / 20
_lib.enable_timing(True)
n = 3
for _ in range(n):
    solver.train_step(G, opt, x, e, sync_losses=False)
torch.cuda.synchronize()
rec = _lib.collect_timing()
_lib.enable_timing(False)
fam = {}
for name, _, ms in rec:
    f = fam.setdefault(name, [0.0, 0])
    f[0] += ms / n
    f[1] += 1 / n
tot = sum(v[0] for v in fam.values())
print(f"{prec}: step {step_ms:.3f} ms (side stream off); inside C-ABI calls {tot:.3f} ms; outside {step_ms - tot:.3f} ms")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:32s} {v[0]:7.3f} ms  x{v[1]:.0f}")
