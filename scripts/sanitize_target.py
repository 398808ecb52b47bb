#!/usr/bin/env python
"""Small-shape exercise of the kernels that synchronise across CTAs (persistent recurrences: global release counters, TMA
multicast, DSMEM reduce-scatter; CTA-pair GEMMs: cta_group::2 MMAs, remote mbarrier arrives; chunked 3xTF32 GEMMs) for
`compute-sanitizer --tool memcheck|racecheck|synccheck` (scripts/sanitize.sh).  Every result is also compared with torch
fp64 so that a sanitizer-clean but wrong run cannot pass.  Usage: sanitize_target.py [lstm|gemm|x3|step]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import autovc_b200
from autovc_b200 import ops, solver
from autovc_b200._lib import ACT_CODES, PREC_FP32X3

dev = "cuda"
what = sys.argv[1] if len(sys.argv) > 1 else "lstm"


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


if what == "lstm":
    # persistent forward + BPTT kernels: one batch tile (B=128) and two (B=200), H=128 (2 column tiles) and H=256
    for B, T, I, H in ((128, 3, 32, 128), (200, 3, 32, 256)):
        torch.manual_seed(1)
        lstm = torch.nn.LSTM(I, H, 1, batch_first=True).to(dev)
        ws = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]
        x = (0.5 * torch.randn(B, T, I, device=dev)).requires_grad_(True)
        out, _, _ = ops.LstmLayerH.apply(x, None, None, *ws)
        go = torch.randn(B, T, H, device=dev) / (B * T) ** 0.5
        grads = torch.autograd.grad(out, [x] + ws, go)
        ref = torch.nn.LSTM(I, H, 1, batch_first=True).to(dev).double()
        ref.load_state_dict({k: v.double() for k, v in lstm.state_dict().items()})
        xd = x.detach().double().requires_grad_(True)
        ro, _ = ref(xd)
        rg = torch.autograd.grad(ro, [xd, ref.weight_ih_l0, ref.weight_hh_l0], go.double())
        torch.cuda.synchronize()
        print("lstm", B, T, I, H, "out", rel(out, ro), "dx", rel(grads[0], rg[0]), "dw_hh", rel(grads[2], rg[2]))
        assert rel(out, ro) < 2e-2 and rel(grads[0], rg[0]) < 5e-2
elif what == "gemm":
    # CTA-pair NT / TN kernels (N, K multiples of 256, two batch rows of 128 frames) through the half-mode conv block
    torch.manual_seed(2)
    B, T, C = 2, 128, 256
    conv = torch.nn.Conv1d(C, C, 5, padding=2).to(dev)
    bn = torch.nn.BatchNorm1d(C).to(dev)
    x = torch.randn(B, T, C, device=dev).requires_grad_(True)
    z, z16, z16b = ops.ConvBnActH.apply(x, None, None, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean,
                                        bn.running_var, None, ACT_CODES["relu"], True, True, False)
    go = torch.randn(B, T, C, device=dev)
    gx, gw = torch.autograd.grad(z, [x, conv.weight], go)
    xr = x.detach().double().requires_grad_(True)
    cw = conv.weight.detach().double().requires_grad_(True)
    y = F.relu(F.batch_norm(F.conv1d(xr.transpose(1, 2), cw, conv.bias.detach().double(), padding=2), None, None,
                            bn.weight.detach().double(), bn.bias.detach().double(), True, 0.1, 1e-5)).transpose(1, 2)
    rgx, rgw = torch.autograd.grad(y, [xr, cw], go.double())
    torch.cuda.synchronize()
    print("gemm z", rel(z, y), "dx", rel(gx, rgx), "dw", rel(gw, rgw))
    assert rel(z, y) < 1e-2 and rel(gx, rgx) < 3e-2 and rel(gw, rgw) < 3e-2
elif what == "x3":
    torch.manual_seed(3)
    B, T, Cin, Cout = 2, 70, 72, 136
    conv = torch.nn.Conv1d(Cin, Cout, 5, padding=2).to(dev)
    bn = torch.nn.BatchNorm1d(Cout).to(dev)
    x = torch.randn(B, T, Cin, device=dev).requires_grad_(True)
    z = ops.ConvBnAct.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, None,
                            ACT_CODES["tanh"], True, PREC_FP32X3)
    go = torch.randn(B, T, Cout, device=dev)
    gx, gw = torch.autograd.grad(z, [x, conv.weight], go)
    lstm = torch.nn.LSTM(Cin, 128, 1, batch_first=True).to(dev)
    ws = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]
    out = ops.LstmLayer.apply(x, PREC_FP32X3, *ws)
    gl = torch.autograd.grad(out, [x] + ws, torch.randn_like(out))
    torch.cuda.synchronize()
    print("x3 ok", float(z.abs().mean()), float(gx.abs().mean()), float(out.abs().mean()), float(gl[0].abs().mean()))
else:
    # one whole half-mode training step at the smallest shape that takes the persistent kernels (B = 128 rows per batch tile)
    torch.manual_seed(0)
    G = autovc_b200.Generator(16, 256, 512, 16, precision="half").to(dev).train()
    g = torch.Generator().manual_seed(1)
    x = torch.rand(128, 16, 80, generator=g).to(dev)
    e = (F.normalize(torch.randn(128, 256, generator=g), dim=-1) * 0.8).to(dev)
    out = solver.train_step(G, autovc_b200.FusedAdam(G.parameters(), 1e-4), x, e)
    torch.cuda.synchronize()
    print("step", out)
