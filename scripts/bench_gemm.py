"""Per-shape timing of the half-mode GEMMs of one training step (config 2: B=256, T=128, dims 16/16).

    python scripts/bench_gemm.py            # NT (forward / data-gradient) and TN (weight-gradient) shapes

Each line: shape, launches of that shape per step, us per call, TFLOP/s, and the effective HBM rate of the
compulsory traffic (16-bit operands in + fp32 result out).  Operands are pre-staged 16-bit like in the model.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autovc_b200 import ops  # noqa: E402
from autovc_b200._lib import FMT_BF16, FMT_FP16  # noqa: E402

DEV = "cuda"
B, T = 256, 128
M = B * T

# (name, N, K, ntaps, calls per step, stats)
NT = [
    ("enc.conv1 336->512 k5", 512, 336, 5, 2, True),
    ("conv 512->512 k5 (fwd)", 512, 512, 5, 10, True),
    ("conv 512->512 k5 (dgrad)", 512, 512, 5, 10, False),
    ("enc.lstm proj 512->128", 128, 512, 1, 2, False),
    ("dec.lstm1 proj 288->2048", 2048, 288, 1, 1, False),
    ("dec.lstm1 dX 2048->288", 288, 2048, 1, 1, False),
    ("dec.lstm2.0 proj 512->4096", 4096, 512, 1, 1, False),
    ("dec.lstm2.0 dX 4096->512", 512, 4096, 1, 1, False),
    ("dec.lstm2.1 proj 1024->4096", 4096, 1024, 1, 1, False),
    ("dec.lstm2.1 dX 4096->1024", 1024, 4096, 1, 1, False),
    ("linear 1024->80", 80, 1024, 1, 1, False),
    ("postnet 80->512 k5", 512, 80, 5, 1, True),
    ("postnet 512->80 k5", 80, 512, 5, 1, True),
]
# (name, N(dY channels), K(X channels), ntaps, calls)
TN = [
    ("conv 512x512 k5 wgrad", 512, 512, 5, 10),
    ("conv 512x336 k5 wgrad", 512, 336, 5, 2),
    ("lstm1 dW_ih 2048x288", 2048, 288, 1, 1),
    ("lstm1 dW_hh 2048x512", 2048, 512, 1, 1),
    ("lstm2.0 dW_ih 4096x512", 4096, 512, 1, 1),
    ("lstm2.x dW_hh 4096x1024", 4096, 1024, 1, 3),
    ("postnet 512x80 k5 wgrad", 512, 80, 5, 1),
    ("postnet 80x512 k5 wgrad", 80, 512, 5, 1),
]


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    torch.manual_seed(0)
    tot = 0.0
    print("-- NT (fp16 A in place, fp32 W staged per call, fp32 C out)")
    for name, N, K, ntaps, calls, st in NT:
        Kp = (K + 7) // 8 * 8
        A = torch.randn(M, Kp, device=DEV).half()
        W = torch.randn(ntaps, N, K, device=DEV) * 0.05
        bias = torch.randn(N, device=DEV)
        C = torch.empty(M, N, device=DEV)
        stats = torch.zeros(2 * N, dtype=torch.double, device=DEV) if st else None
        us = timed(lambda: ops.gemm_nt_taps_h(A, FMT_FP16, Kp, W, bias, C, N, B, T, N, K, ntaps, -(ntaps // 2), FMT_FP16, stats=stats))
        fl = 2.0 * M * N * K * ntaps
        by = M * K * 2 + M * N * 4
        tot += us * calls
        print(f"{name:32s} x{calls:2d}  {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s  {by / us / 1e3:7.1f} GB/s")
    print(f"NT total per step: {tot / 1e3:.3f} ms")
    tot = 0.0
    print("-- TN (bf16 dY, bf16 X in place, fp32 dW out)")
    for name, N, K, ntaps, calls in TN:
        Np, Kp = (N + 7) // 8 * 8, (K + 7) // 8 * 8
        dY = torch.randn(M, Np, device=DEV).bfloat16()
        X = torch.randn(M, Kp, device=DEV).bfloat16()
        dW = torch.empty(ntaps, N, K, device=DEV)
        us = timed(lambda: ops.gemm_tn_taps_h(dY, FMT_BF16, Np, X, FMT_BF16, Kp, dW, B, T, N, K, ntaps, -(ntaps // 2), 0))
        fl = 2.0 * M * N * K * ntaps
        by = M * (N + K) * 2
        tot += us * calls
        print(f"{name:32s} x{calls:2d}  {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s  {by / us / 1e3:7.1f} GB/s")
    print(f"TN total per step: {tot / 1e3:.3f} ms")


if __name__ == "__main__":
    main()
