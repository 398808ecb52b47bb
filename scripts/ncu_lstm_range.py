"""ONE forward or ONE BPTT launch of the persistent recurrence (B=256, T=128, H=AVC_NCU_H) between cudaProfilerStart/Stop, for
`ncu --replay-mode app-range`: the kernels spin on counters published by other CTAs, so kernel replay cannot profile them
(nan + abort), `--replay-mode range` refuses the capture (cuTensorMapEncodeTiled inside the range), but app-range re-runs the
whole application once per counter pass and measures only the range -- that works (profiles/r02_lstm_counters.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autovc_b200 import _lib
from autovc_b200.ops import _p, _stream, _ws

dev = "cuda"
B, T, H = 256, int(os.environ.get("AVC_NCU_T", "128")), int(os.environ.get("AVC_NCU_H", "1024"))
G = 4 * H
prec = _lib.PREC_BF16
P = torch.randn(B, T, G, device=dev) * 0.5
W = (torch.randn(G, H, device=dev) * (1.0 / H ** 0.5)).bfloat16()
WT = W.t().contiguous()
h = torch.empty(B, T, H, device=dev); gates = torch.empty(B, T, G, device=dev); c = torch.empty(B, T, H, device=dev)
h16 = torch.empty(B, T, H, device=dev, dtype=torch.float16); h16b = torch.empty(B, T, H, device=dev, dtype=torch.bfloat16)
dH = torch.randn(B, T, H, device=dev) * 0.1
dP16 = torch.empty(B, T, G, device=dev, dtype=torch.bfloat16)
nf = _lib.query("avc_lstm_fwd_workspace_bytes", B, T, H, prec); wf = _ws(nf, dev)
nb = _lib.query("avc_lstm_bwd_workspace_bytes", B, T, H, prec); wb = _ws(nb, dev)


def fwd():
    _lib.call("avc_lstm_seq_fwd_h", _p(P), _p(W), 1, _p(h), H, _p(gates), _p(c), _p(h16), 2, _p(h16b), B, T, H, 0, _p(wf), nf, _stream())


def bwd():
    _lib.call("avc_lstm_seq_bwd_h", _p(dH), H, _p(WT), 1, _p(gates), _p(c), None, _p(dP16), B, T, H, 0, _p(wb), nb, _stream())


fwd(); bwd()
torch.cuda.synchronize()
which = os.environ.get("AVC_NCU_WHICH", "fwd")
torch.cuda.profiler.start()
(fwd if which == "fwd" else bwd)()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", which)
