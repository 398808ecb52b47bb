"""A handful of representative half-mode GEMM launches for one `ncu --set full` capture (see gpu_profile.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autovc_b200 import ops  # noqa: E402
from autovc_b200._lib import FMT_BF16, FMT_FP16  # noqa: E402

DEV = "cuda"
B, T = 256, 128
M = B * T


def nt(N, K, ntaps, stats):
    A = torch.randn(M, K, device=DEV).half()
    W = torch.randn(ntaps, N, K, device=DEV) * 0.05
    bias = torch.randn(N, device=DEV)
    C = torch.empty(M, N, device=DEV)
    st = torch.zeros(2 * N, dtype=torch.double, device=DEV) if stats else None
    ops.gemm_nt_taps_h(A, FMT_FP16, K, W, bias, C, N, B, T, N, K, ntaps, -(ntaps // 2), FMT_FP16, stats=st)
    torch.cuda.synchronize()


def tn(N, K, ntaps):
    dY = torch.randn(M, N, device=DEV).bfloat16()
    X = torch.randn(M, K, device=DEV).bfloat16()
    dW = torch.empty(ntaps, N, K, device=DEV)
    ops.gemm_tn_taps_h(dY, FMT_BF16, N, X, FMT_BF16, K, dW, B, T, N, K, ntaps, -(ntaps // 2), 0)
    torch.cuda.synchronize()


nt(512, 512, 5, False)      # conv dgrad
nt(512, 512, 5, True)       # conv fwd + BN sums
nt(4096, 512, 1, False)     # LSTM input projection (store-heavy)
tn(512, 512, 5)             # conv wgrad
tn(4096, 1024, 1)           # dW_hh
