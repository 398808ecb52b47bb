"""World-size-2 gloo test (CPU) of the data-parallel host logic: GradBucketReducer must give every
rank the mean of the per-rank gradients, bucket in reverse registration order, fire each bucket once
even when a parameter is used twice in the step (the two encoder passes), and leave BN-like buffers
rank-local."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = torch.nn.Linear(8, 8)
        self.bn = torch.nn.BatchNorm1d(8)
        self.dec = torch.nn.Linear(8, 8)
        self.post = torch.nn.Linear(8, 8)

    def forward(self, x):
        return self.post(torch.tanh(self.dec(self.bn(self.enc(x)))))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from autovc_b200 import solver
    torch.manual_seed(0)
    m = Tiny()
    solver.broadcast_parameters(m)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(16, 8, generator=g)

    def loss_fn(mod):
        y = mod(x)
        z = mod.enc(y)               # second use of the encoder weights in the same step
        return (y ** 2).mean() + z.abs().mean()

    # local (un-reduced) gradients for the expected value, taken before the reducer installs its hooks
    local = [t.clone() for t in torch.autograd.grad(loss_fn(m), list(m.parameters()))]
    red = solver.GradBucketReducer(m.parameters(), bucket_mb=8 * 8 * 4 * 1.5 / (1024 * 1024))   # ~1.5 weight matrices per bucket
    rm_local = m.bn.running_mean.clone()
    gathered = [[torch.zeros_like(t) for _ in range(world)] for t in local]
    for t, outl in zip(local, gathered):
        dist.all_gather(outl, t)
    expect = [torch.stack(o).mean(0) for o in gathered]

    m.bn.running_mean.zero_(); m.bn.running_var.fill_(1.0); m.bn.num_batches_tracked.zero_()
    for p in m.parameters():
        p.grad = None
    red.begin_backward()
    loss_fn(m).backward()
    red.finish()
    ok = all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(m.parameters(), expect))
    n_buckets = len(red.buckets)
    order = list(red.launch_order)
    first_bucket_params = [id(p) for p in red.buckets[0]["params"]]
    last_param_first = first_bucket_params[0] == id(list(m.parameters())[-1])
    # buffers stay rank-local: ranks see different batches, so running_mean differs across ranks
    rms = [torch.zeros_like(rm_local) for _ in range(world)]
    dist.all_gather(rms, m.bn.running_mean)
    buffers_local = not torch.allclose(rms[0], rms[1])
    # a second step reuses the flat buffers
    for p in m.parameters():
        p.grad = None
    red.begin_backward()
    loss_fn(m).backward()
    red.finish()
    ok2 = all(torch.allclose(p.grad, e, atol=1e-5) for p, e in zip(m.parameters(), expect))
    q.put((rank, ok, n_buckets, order, last_param_first, buffers_local, ok2))
    dist.destroy_process_group()


def test_grad_bucket_reducer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, n_buckets, order, last_first, buffers_local, ok2 in res:
        assert ok, f"rank {rank}: reduced gradients != mean of per-rank gradients"
        assert n_buckets >= 3
        assert sorted(order) == list(range(n_buckets)), "every bucket must be all-reduced exactly once"
        assert order[0] == 0, "the bucket holding the last-registered parameters fills first"
        assert last_first and buffers_local and ok2


def test_reducer_single_process_is_identity():
    from autovc_b200 import solver
    m = Tiny()
    red = solver.GradBucketReducer(m.parameters(), bucket_mb=1.0)
    x = torch.randn(4, 8)
    m(x).sum().backward()
    ref = [p.grad.clone() for p in m.parameters()]
    for p in m.parameters():
        p.grad = None
    red.begin_backward()
    m(x).sum().backward()
    red.finish()
    for p, r in zip(m.parameters(), ref):
        assert torch.allclose(p.grad, r)
