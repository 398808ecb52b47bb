import torch, sys, os
sys.path.insert(0, "/root/repo")
from autovc_b200 import ops
from autovc_b200._lib import FMT_FP16
x = torch.empty(1 << 28, device="cuda")   # 1 GiB fp32
def timed(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
us = timed(lambda: x.zero_())
print(f"zero_ 1 GiB: {us:.1f} us  {x.numel()*4/us/1e3:.0f} GB/s")
y = torch.empty_like(x)
us = timed(lambda: y.copy_(x))
print(f"copy 1 GiB: {us:.1f} us  {2*x.numel()*4/us/1e3:.0f} GB/s (r+w)")
B, T = 256, 128
M = B * T
for N, K in [(4096, 64), (4096, 512), (4096, 1024), (2048, 64), (512, 64)]:
    A = torch.randn(M, K, device="cuda").half()
    W = (torch.randn(1, N, K, device="cuda") * 0.05).half()
    C = torch.empty(M, N, device="cuda")
    us = timed(lambda: ops.gemm_nt_taps_hw(A, FMT_FP16, K, W, FMT_FP16, K, None, C, N, B, T, N, K, 1, 0))
    print(f"N={N} K={K}: {us:.1f} us  write {M*N*4/us/1e3:.0f} GB/s  {2.0*M*N*K/us/1e6:.0f} TF/s")
