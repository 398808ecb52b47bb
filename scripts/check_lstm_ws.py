#!/usr/bin/env python
"""Weight-stationary forward recurrence (lstm_tc_fwd_ws_kernel) against the ring kernel (AVC_LSTM_FWD_WS=0) on identical
inputs: every output tensor, several shapes incl. ragged batches, the reverse direction and inference mode (nothing saved
for BPTT); then the per-step latency of both.  Run under `timeout`: a protocol bug in a persistent kernel is a hang."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from autovc_b200 import _lib
from autovc_b200.ops import _p, _stream, _ws

dev = "cuda"


def run(B, T, H, reverse, save, fmt16, ws_on, P, Wb, reps=0):
    os.environ["AVC_LSTM_FWD_WS"] = "1" if ws_on else "0"
    G4 = 4 * H
    h = torch.full((B, T, H), float("nan"), device=dev)
    gates = torch.full((B, T, G4), float("nan"), device=dev) if save else None
    c = torch.full((B, T, H), float("nan"), device=dev) if save else None
    h16 = torch.zeros(B, T, H, device=dev, dtype=torch.float16 if fmt16 == 2 else torch.bfloat16)
    h16b = torch.zeros(B, T, H, device=dev, dtype=torch.bfloat16)
    nf = _lib.query("avc_lstm_fwd_workspace_bytes", B, T, H, _lib.PREC_BF16)
    wf = _ws(nf, dev)

    def call():
        _lib.call("avc_lstm_seq_fwd_h", _p(P), _p(Wb), 1, _p(h), H, _p(gates) if save else None, _p(c) if save else None, _p(h16), fmt16,
                  _p(h16b), B, T, H, reverse, _p(wf), nf, _stream())

    call()
    torch.cuda.synchronize()
    us = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3 / T
    return dict(h=h, gates=gates, c=c, h16=h16, h16b=h16b), us


def run_bwd(B, T, H, reverse, ws_on, gates, c, dH, WTb, fp32_out, reps=0):
    os.environ["AVC_LSTM_BWD_WS"] = "1" if ws_on else "0"
    G4 = 4 * H
    dP = torch.full((B, T, G4), float("nan"), device=dev) if fp32_out else None
    dP16 = torch.zeros(B, T, G4, device=dev, dtype=torch.bfloat16)
    nb = _lib.query("avc_lstm_bwd_workspace_bytes", B, T, H, _lib.PREC_BF16)
    wb = _ws(nb, dev)

    def call():
        _lib.call("avc_lstm_seq_bwd_h", _p(dH), H, _p(WTb), 1, _p(gates), _p(c), _p(dP) if fp32_out else None, _p(dP16), B, T, H, reverse,
                  _p(wb), nb, _stream())

    call()
    torch.cuda.synchronize()
    us = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3 / T
    return dict(dP=dP, dP16=dP16), us


def main_bwd(argv):
    shapes = [(4, 6, 128, 0), (130, 9, 512, 1), (256, 16, 1024, 0), (37, 12, 768, 0), (300, 10, 256, 1), (600, 5, 1024, 1),
              (256, 128, 1024, 0), (256, 128, 512, 1), (128, 256, 1024, 0), (256, 128, 768, 0)]
    if argv and argv[0] == "big":
        shapes = [sh for sh in shapes if sh[1] >= 64]
    elif argv:
        shapes = shapes[:int(argv[0])]
    fails = []
    for (B, T, H, reverse) in shapes:
        g = torch.Generator().manual_seed(B * 131 + H)
        P = (0.5 * torch.randn(B, T, 4 * H, generator=g)).to(dev)
        W = torch.randn(4 * H, H, generator=g) / H ** 0.5
        Wb = W.to(dev).bfloat16()
        WTb = Wb.t().contiguous()
        dH = (0.1 * torch.randn(B, T, H, generator=g)).to(dev)
        fw, _ = run(B, T, H, reverse, True, 2, True, P, Wb)
        big = T >= 64
        a, us_ks = run_bwd(B, T, H, reverse, False, fw["gates"], fw["c"], dH, WTb, not big, reps=5 if big else 0)
        b, us_ws = run_bwd(B, T, H, reverse, True, fw["gates"], fw["c"], dH, WTb, not big, reps=5 if big else 0)
        res = {"bwd": 1, "B": B, "T": T, "H": H, "reverse": reverse}
        for k in a:
            if a[k] is None:
                continue
            x, y = a[k].float(), b[k].float()
            res[k] = float((x - y).abs().max())
            res[k + "_rel"] = float((x - y).norm() / x.norm().clamp_min(1e-30))
            res[k + "_nan"] = int(torch.isnan(y).sum())
        if big:
            res["us_ks"], res["us_ws"] = round(us_ks, 3), round(us_ws, 3)
        print(json.dumps(res), flush=True)
        bad = [k for k in a if a[k] is not None and (res[k + "_rel"] > 1e-3 or res[k + "_nan"] > 0)]
        if bad:
            fails.append(((B, T, H), bad))
    print("FAILS:", fails)
    sys.exit(1 if fails else 0)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "bwd":
        main_bwd(sys.argv[2:])
    shapes = [(4, 6, 128, 0, True), (130, 9, 512, 1, True), (256, 16, 1024, 0, True), (37, 12, 768, 0, False), (300, 10, 256, 1, True),
              (256, 128, 1024, 0, True), (256, 128, 512, 1, True), (128, 256, 1024, 0, True), (256, 128, 768, 0, False)]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        shapes = [sh for sh in shapes if sh[1] >= 64]
    elif len(sys.argv) > 1:
        shapes = shapes[:int(sys.argv[1])]
    fails = []
    for (B, T, H, reverse, save) in shapes:
        g = torch.Generator().manual_seed(B * 131 + H)
        P = (0.5 * torch.randn(B, T, 4 * H, generator=g)).to(dev)
        Wb = (torch.randn(4 * H, H, generator=g) / H ** 0.5).to(dev).bfloat16()
        big = T >= 64
        a, us_ring = run(B, T, H, reverse, save, 2, False, P, Wb, reps=5 if big else 0)
        b, us_ws = run(B, T, H, reverse, save, 2, True, P, Wb, reps=5 if big else 0)
        res = {"B": B, "T": T, "H": H, "reverse": reverse, "save": save}
        for k in a:
            if a[k] is None:
                continue
            x, y = a[k].float(), b[k].float()
            res[k] = float((x - y).abs().max())
            res[k + "_nan"] = int(torch.isnan(y).sum())
        if big:
            res["us_ring"], res["us_ws"] = round(us_ring, 3), round(us_ws, 3)
        print(json.dumps(res), flush=True)
        bad = [k for k in a if a[k] is not None and (res[k] > 2e-2 or res[k + "_nan"] > 0)]
        if bad:
            fails.append(((B, T, H), bad))
    print("FAILS:", fails)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
