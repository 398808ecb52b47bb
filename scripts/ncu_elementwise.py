"""One 512-channel conv block and one LSTM(1024) layer, forward + backward in half mode at config 2 sizes: every
non-GEMM kernel family of the training step once, for an `ncu --set full` capture of the memory-bound kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autovc_b200 import ops  # noqa: E402
from autovc_b200.model_vc_mel import Generator  # noqa: E402

torch.manual_seed(0)
B, T = 256, 128
G = Generator(16, 256, 512, 16, precision="half").cuda()
x = torch.randn(B, T, 80, device="cuda")
emb = torch.randn(B, 256, device="cuda")
for _ in range(int(os.environ.get("ITERS", "1"))):
    out = G(x, emb, emb)
    loss = sum(o.float().pow(2).mean() for o in out[:2])
    loss.backward()
torch.cuda.synchronize()
