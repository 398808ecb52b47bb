#!/usr/bin/env python
"""Numeric check of the NCCL data-parallel path on real GPUs (torchrun --nproc-per-node 2): after a step through
solver.GradBucketReducer every rank's param.grad must equal the mean over ranks of the LOCAL gradients, computed here
independently (a plain backward without the reducer, then dist.all_reduce of the clones).  Exercises the bucket gather on
the weight-gradient side stream, the communication stream and FusedAdam on flat-bucket gradient views."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F

import autovc_b200
from autovc_b200 import solver

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("AVC_CHECK_NO_ENV_CAP") != "1":      # "1": prove that the reducer's own communicator carries the CTA cap
    solver.nccl_env_defaults()
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()

torch.manual_seed(0)
G = autovc_b200.Generator(16, 256, 512, 16, precision="half").to(dev).train()
solver.broadcast_parameters(G)
g = torch.Generator().manual_seed(100 + rank)
x = torch.rand(64, 128, 80, generator=g).to(dev)
e = (F.normalize(torch.randn(64, 256, generator=g), dim=-1) * 0.8).to(dev)

# (1) local gradients, no reducer
loss, _, _ = solver.generator_losses(G, x, e)
loss.backward()
torch.cuda.synchronize()
want = {}
for n, p in G.named_parameters():
    t = p.grad.detach().clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    want[n] = t / world
    p.grad = None

# (2) the same step through the reducer (three times: steady state of the flat buckets / pointer tables)
red = solver.GradBucketReducer(G.parameters(), bucket_mb=25.0)
worst = 0.0
for it in range(3):
    loss, _, _ = solver.generator_losses(G, x, e)
    for p in G.parameters():
        p.grad = None
    red.begin_backward()
    loss.backward()
    red.finish()
    torch.cuda.synchronize()
    for n, p in G.named_parameters():
        a, b = p.grad.double(), want[n].double()
        err = float((a - b).norm() / (b.norm() + 1e-30))
        worst = max(worst, err)
        assert err < 1e-5 or float((a - b).abs().max()) < 1e-9, (it, n, err)
# every rank holds the same averaged gradients
chk = torch.stack([p.grad.double().sum() for p in G.parameters()])
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi), "ranks disagree on the reduced gradients"

# (3) FusedAdam on the flat-bucket views keeps the replicas identical
opt = autovc_b200.FusedAdam(G.parameters(), 1e-4)
for it in range(2):
    solver.train_step(G, opt, x, e, reducer=red, sync_losses=False)
torch.cuda.synchronize()
ps = torch.stack([p.detach().double().sum() for p in G.parameters()])
lo, hi = ps.clone(), ps.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi), "replicas diverged after FusedAdam steps"
# (4) step time at the benched shape with the reducer's communicator (a convoy with the cooperative recurrences would show here)
xb = torch.rand(256, 128, 80, generator=g).to(dev)
eb = (F.normalize(torch.randn(256, 256, generator=g), dim=-1) * 0.8).to(dev)
for _ in range(5):
    solver.train_step(G, opt, xb, eb, reducer=red, sync_losses=False)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    solver.train_step(G, opt, xb, eb, reducer=red, sync_losses=False)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"check_dp_nccl: NCCL CTA cap: {red.cta_cap}; NCCL_MAX_CTAS env: {os.environ.get('NCCL_MAX_CTAS')}; "
          f"{float(ms):.2f} ms per 256-crop step on {world} GPUs")
    print(f"check_dp_nccl: OK on {world} GPUs (worst relative deviation of a reduced gradient {worst:.2e}; replicas identical after 2 optimizer steps)")
dist.destroy_process_group()
