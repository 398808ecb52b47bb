#!/usr/bin/env python
"""Per-timestep latency of the persistent LSTM recurrence kernels (north_star: 'per-timestep latency
for the recurrence').  Times avc_lstm_seq_fwd / avc_lstm_seq_bwd alone with CUDA events."""
import ctypes
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autovc_b200 import _lib
from autovc_b200.ops import _p, _stream, _ws

dev = "cuda"
res = []
for (B, T, H) in [(256, 128, 1024), (256, 128, 512), (128, 256, 1024)]:
    G = 4 * H
    prec = _lib.PREC_BF16
    P = torch.randn(B, T, G, device=dev) * 0.5
    W = torch.randn(G, H, device=dev) * (1.0 / H ** 0.5)
    WT = W.t().contiguous()
    h = torch.empty(B, T, H, device=dev); gates = torch.empty(B, T, G, device=dev); c = torch.empty(B, T, H, device=dev)
    dH = torch.randn(B, T, H, device=dev) * 0.1
    dP = torch.empty(B, T, G, device=dev)
    nf = _lib.query("avc_lstm_fwd_workspace_bytes", B, T, H, prec); wf = _ws(nf, dev)
    nb = _lib.query("avc_lstm_bwd_workspace_bytes", B, T, H, prec); wb = _ws(nb, dev)
    def fwd():
        _lib.call("avc_lstm_seq_fwd", _p(P), _p(W), _p(h), H, _p(gates), _p(c), B, T, H, 0, prec, _p(wf), nf, _stream())
    def bwd():
        _lib.call("avc_lstm_seq_bwd", _p(dH), H, _p(W), _p(WT), _p(gates), _p(c), _p(dP), B, T, H, 0, prec, _p(wb), nb, _stream())
    h16 = torch.empty(B, T, H, device=dev, dtype=torch.float16)
    dP16 = torch.empty(B, T, G, device=dev, dtype=torch.bfloat16)
    Wb, WTb = W.bfloat16(), WT.bfloat16()
    def fwd_h():     # half mode: pre-packed bf16 W_hh, fp16 copy of h
        _lib.call("avc_lstm_seq_fwd_h", _p(P), _p(Wb), 1, _p(h), H, _p(gates), _p(c), _p(h16), 2, None, B, T, H, 0, _p(wf), nf, _stream())
    def bwd_h():     # half mode: gate gradient as bf16 only
        _lib.call("avc_lstm_seq_bwd_h", _p(dH), H, _p(WTb), 1, _p(gates), _p(c), None, _p(dP16), B, T, H, 0, _p(wb), nb, _stream())
    for name, fn in (("fwd", fwd), ("bwd", bwd), ("fwd_h", fwd_h), ("bwd_h", bwd_h)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        flops = 2.0 * B * T * G * H
        res.append({"kernel": f"lstm_seq_{name}", "B": B, "T": T, "H": H, "ms": ms, "us_per_step": ms * 1e3 / T,
                    "tflops": flops / (ms * 1e-3) / 1e12})
        print(json.dumps(res[-1]), flush=True)

# ---- per-step event trace of CTA 0 (one forward and one backward launch, H=1024) ----
B, T, H = 256, 128, 1024
G = 4 * H
prec = _lib.PREC_BF16
P = torch.randn(B, T, G, device=dev) * 0.5
W = torch.randn(G, H, device=dev) * (1.0 / H ** 0.5)
WT = W.t().contiguous()
h = torch.empty(B, T, H, device=dev); gates = torch.empty(B, T, G, device=dev); c = torch.empty(B, T, H, device=dev)
dH = torch.randn(B, T, H, device=dev) * 0.1
dP = torch.empty(B, T, G, device=dev)
nf = _lib.query("avc_lstm_fwd_workspace_bytes", B, T, H, prec); wf = _ws(nf, dev)
nb = _lib.query("avc_lstm_bwd_workspace_bytes", B, T, H, prec); wb = _ws(nb, dev)
names = ["barrier", "tma0", "tmaN", "land0", "landN", "mma_issued", "epi_wake", "math_done", "published", "rs_sent", "rs_done", "bar_passed"]
for which in ("fwd", "bwd"):
    trace = torch.zeros(16 * T, dtype=torch.int64, device=dev)
    _lib.load().avc_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
    if which == "fwd":
        _lib.call("avc_lstm_seq_fwd", _p(P), _p(W), _p(h), H, _p(gates), _p(c), B, T, H, 0, prec, _p(wf), nf, _stream())
    else:
        _lib.call("avc_lstm_seq_bwd", _p(dH), H, _p(W), _p(WT), _p(gates), _p(c), _p(dP), B, T, H, 0, prec, _p(wb), nb, _stream())
    torch.cuda.synchronize()
    _lib.load().avc_debug_set_trace(ctypes.c_void_p(0))
    full = trace.cpu()
    tr = full[:16 * T].view(T, 16)[:, :12].double()
    # steps 20..100: offsets of each event relative to the previous step's 'published'
    rows = []
    for s in range(20, 100):
        base = tr[s - 1, 8]
        rows.append(((tr[s] - base) / 1e3).tolist())
    import numpy as np
    med = np.median(np.array(rows), axis=0)
    print(which, "median us after previous publish:", {n: round(float(v), 2) for n, v in zip(names, med) if v > -1e6}, flush=True)
