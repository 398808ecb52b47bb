"""Per-tile %globaltimer trace of pair 0's leader CTA in the CTA-pair NT GEMM (avc_debug_set_trace)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autovc_b200 import _lib, ops  # noqa: E402
from autovc_b200._lib import FMT_FP16  # noqa: E402

B, T = 256, 128
M = B * T
N, K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 64
TAPS = int(sys.argv[3]) if len(sys.argv) > 3 else 1
A = torch.randn(M, K, device="cuda").half()
W = (torch.randn(TAPS, N, K, device="cuda") * 0.05).half()
C = torch.empty(M, N, device="cuda")
for _ in range(3):
    ops.gemm_nt_taps_hw(A, FMT_FP16, K, W, FMT_FP16, K, None, C, N, B, T, N, K, TAPS, -(TAPS // 2))
trace = torch.zeros(1024, dtype=torch.int64, device="cuda")
_lib.load().avc_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
ops.gemm_nt_taps_hw(A, FMT_FP16, K, W, FMT_FP16, K, None, C, N, B, T, N, K, TAPS, -(TAPS // 2))
torch.cuda.synchronize()
_lib.load().avc_debug_set_trace(ctypes.c_void_p(0))
full = trace.cpu()
tr = full[:512].view(64, 8).double()
t0 = tr[tr > 0].min()
names = ["slot free", "loads issued", "acc free", "operands in", "MMAs issued", "epi woke", "chunk0 done", "acc released"]
print("tile " + " ".join(f"{n:>13s}" for n in names))
for i in range(28):
    if tr[i].max() == 0:
        break
    print(f"{i:4d} " + " ".join(f"{(x - t0) / 1e3:13.2f}" if x > 0 else f"{'-':>13s}" for x in tr[i].tolist()))
ct = full[512:544].view(8, 4).double()
c0 = ct[ct > 0].min()
print("chunk-level stamps of tile 4 (us): TMEM read | bias/mask | staged | store issued")
for c in range(8):
    print(f"{c:4d} " + " ".join(f"{(x - c0) / 1e3:10.2f}" for x in ct[c].tolist()))
