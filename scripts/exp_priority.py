"""Experiment: does running the training loop on a HIGH-priority stream (the weight-gradient side stream stays at normal
priority) let the side-stream GEMMs fill gaps instead of blocking the critical path?  Prints ms/step for the settings."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import autovc_b200
from autovc_b200 import ops, solver

dev = torch.device("cuda", 0)
torch.manual_seed(0)
G = autovc_b200.Generator(16, 256, 512, 16, precision="half").to(dev).train()
opt = autovc_b200.FusedAdam(G.parameters(), 1e-4)
g = torch.Generator().manual_seed(1)
x = torch.rand(256, 128, 80, generator=g).to(dev)
e = (F.normalize(torch.randn(256, 256, generator=g), dim=-1) * 0.8).to(dev)


def run(n):
    for _ in range(n):
        solver.train_step(G, opt, x, e, sync_losses=False)


def timed(tag, stream=None, side_on=True, n=40):
    ops._WGRAD["on"] = side_on
    ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
    with ctx:
        run(8)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(n)
        e1.record()
        torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) / n:.3f} ms/step", flush=True)


print("priority range", torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else "n/a")
timed("default stream, side stream on")
timed("default stream, side stream off", side_on=False)
hi = torch.cuda.Stream(device=dev, priority=-1)
timed("high-priority main stream (-1), side on", stream=hi)
try:
    hi5 = torch.cuda.Stream(device=dev, priority=-5)
    timed("high-priority main stream (-5), side on", stream=hi5)
except Exception as ex:
    print("priority -5:", ex)
timed("default stream, side stream on (again)")
