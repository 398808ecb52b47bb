#!/usr/bin/env python
"""Quick numeric + latency check of the persistent recurrences (lstm_tc.cu) against torch.nn.LSTM in fp64, and the
per-step trace of CTA 0 (`trace` argument).  Run under `timeout`: a protocol bug in a persistent kernel shows up as a hang."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from autovc_b200 import _lib, ops
from autovc_b200.ops import _p, _stream, _ws

dev = "cuda"


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def check(B, T, I, H, bwd=True, reps=5):
    torch.manual_seed(5)
    lstm = torch.nn.LSTM(I, H, 1, batch_first=True).to(dev)
    ws = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]
    g = torch.Generator().manual_seed(8)
    x = (0.5 * torch.randn(B, T, I, generator=g)).to(dev).requires_grad_(True)
    go = (torch.randn(B, T, H, generator=g) / (B * T) ** 0.5).to(dev)
    out, h16, h16b = ops.LstmLayerH.apply(x, None, None, *ws)
    torch.cuda.synchronize()
    ref = torch.nn.LSTM(I, H, 1, batch_first=True).to(dev).double()
    ref.load_state_dict({k: v.double() for k, v in lstm.state_dict().items()})
    xd = x.detach().double().requires_grad_(True)
    rout, _ = ref(xd)
    res = {"B": B, "T": T, "I": I, "H": H, "out": rel(out, rout), "h16": rel(h16.float(), rout), "h16b": rel(h16b.float(), rout)}
    if bwd:
        got = torch.autograd.grad(out, [x] + ws, go)
        rg = torch.autograd.grad(rout, [xd, ref.weight_ih_l0, ref.weight_hh_l0, ref.bias_ih_l0, ref.bias_hh_l0], go.double())
        for a, b, n in zip(got, rg, ["dx", "dw_ih", "dw_hh", "db_ih", "db_hh"]):
            res[n] = rel(a, b)
    # latency of the recurrence launches alone
    G4 = 4 * H
    P = torch.randn(B, T, G4, device=dev) * 0.5
    Wb = (torch.randn(G4, H, device=dev) / H ** 0.5).bfloat16()
    h = torch.empty(B, T, H, device=dev)
    gates = torch.empty(B, T, G4, device=dev)
    c = torch.empty(B, T, H, device=dev)
    h16 = torch.empty(B, T, H, device=dev, dtype=torch.float16)
    h16b = torch.empty(B, T, H, device=dev, dtype=torch.bfloat16)
    dH = torch.randn(B, T, H, device=dev) * 0.1
    dP16 = torch.empty(B, T, G4, device=dev, dtype=torch.bfloat16)
    nf = _lib.query("avc_lstm_fwd_workspace_bytes", B, T, H, _lib.PREC_BF16)
    wf = _ws(nf, dev)
    nb = _lib.query("avc_lstm_bwd_workspace_bytes", B, T, H, _lib.PREC_BF16)
    wb = _ws(nb, dev)
    WTb = Wb.t().contiguous()

    def fwd_h():
        _lib.call("avc_lstm_seq_fwd_h", _p(P), _p(Wb), 1, _p(h), H, _p(gates), _p(c), _p(h16), 2, _p(h16b), B, T, H, 0, _p(wf), nf, _stream())

    def bwd_h():
        _lib.call("avc_lstm_seq_bwd_h", _p(dH), H, _p(WTb), 1, _p(gates), _p(c), None, _p(dP16), B, T, H, 0, _p(wb), nb, _stream())
    for name, fn in (("fwd_us_per_step", fwd_h), ("bwd_us_per_step", bwd_h)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / reps * 1e3 / T
    print(json.dumps({k: (round(v, 5) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
    bad = [k for k in ("out", "h16") if res[k] > 1e-2] + [k for k in ("dx", "dw_ih", "dw_hh", "db_ih", "db_hh") if res.get(k, 0) > 2e-2]
    return bad


def trace(B, T, H, which="fwd"):
    """Per-step stamps of pair 0 / leader / batch tile 0 (LP_TRACE slots), medians over steps 20..T-20 in us after the
    previous step's publish."""
    import ctypes
    import numpy as np
    G4 = 4 * H
    P = torch.randn(B, T, G4, device=dev) * 0.5
    Wb = (torch.randn(G4, H, device=dev) / H ** 0.5).bfloat16()
    h = torch.empty(B, T, H, device=dev)
    gates = torch.empty(B, T, G4, device=dev)
    c = torch.empty(B, T, H, device=dev)
    h16 = torch.empty(B, T, H, device=dev, dtype=torch.float16)
    h16b = torch.empty(B, T, H, device=dev, dtype=torch.bfloat16)
    dH = torch.randn(B, T, H, device=dev) * 0.1
    dP16 = torch.empty(B, T, G4, device=dev, dtype=torch.bfloat16)
    nf = _lib.query("avc_lstm_fwd_workspace_bytes", B, T, H, _lib.PREC_BF16)
    wf = _ws(nf, dev)
    nb = _lib.query("avc_lstm_bwd_workspace_bytes", B, T, H, _lib.PREC_BF16)
    wb = _ws(nb, dev)
    WTb = Wb.t().contiguous()
    _lib.call("avc_lstm_seq_fwd_h", _p(P), _p(Wb), 1, _p(h), H, _p(gates), _p(c), _p(h16), 2, _p(h16b), B, T, H, 0, _p(wf), nf, _stream())
    tr = torch.zeros(16 * T + 16384, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    _lib.load().avc_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
    if which == "fwd":
        _lib.call("avc_lstm_seq_fwd_h", _p(P), _p(Wb), 1, _p(h), H, _p(gates), _p(c), _p(h16), 2, _p(h16b), B, T, H, 0, _p(wf), nf, _stream())
    else:
        _lib.call("avc_lstm_seq_bwd_h", _p(dH), H, _p(WTb), 1, _p(gates), _p(c), None, _p(dP16), B, T, H, 0, _p(wb), nb, _stream())
    torch.cuda.synchronize()
    _lib.load().avc_debug_set_trace(ctypes.c_void_p(0))
    t = tr[:16 * T].view(T, 16).double().cpu().numpy()
    rows = []
    for s in range(20, T - 20):
        rows.append((t[s, :12] - t[s - 1, 8]) / 1e3)
    med = np.median(np.array(rows), axis=0)
    names = ["counter_seen", "tma_first", "tma_last", "land_first", "land_last", "mma_issued", "epi_woke", "math_done", "published",
             "rs_sent", "rs_received", "bar_passed"]
    print(which, (B, T, H), "period us:", round(float(np.median(np.diff(t[20:T - 20, 8])) / 1e3), 2),
          {n: round(float(v), 2) for n, v in zip(names, med) if v > -1e6}, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "trace":
        for sh in [(256, 128, 1024), (128, 128, 1024), (256, 128, 512)]:
            trace(*sh, which="fwd")
            if len(sys.argv) > 2:
                trace(*sh, which="bwd")
        sys.exit(0)
    shapes = [(4, 6, 48, 128), (130, 9, 32, 512), (256, 16, 64, 1024), (256, 128, 512, 1024), (130, 128, 288, 512), (128, 256, 512, 1024),
              (40, 96, 80, 768), (300, 20, 64, 256)]
    if len(sys.argv) > 1:
        shapes = shapes[:int(sys.argv[1])]
    fails = []
    for sh in shapes:
        bad = check(*sh)
        if bad:
            fails.append((sh, bad))
    print("FAILS:", fails)
    sys.exit(1 if fails else 0)
