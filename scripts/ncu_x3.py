"""Representative fp32-mode (3xTF32, chunked accumulation) launches at config-2 sizes for one `ncu --set full` capture:
conv forward (NT, 5 taps, BN sums), conv weight gradient (TN), one recurrence step's split-reduction product."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autovc_b200 import ops  # noqa: E402
from autovc_b200._lib import PREC_FP32X3  # noqa: E402

DEV = "cuda"
B, T = 256, 128
M = B * T
A = torch.randn(M, 512, device=DEV)
W = torch.randn(5, 512, 512, device=DEV) * 0.05
bias = torch.randn(512, device=DEV)
C = torch.empty(M, 512, device=DEV)
st = torch.zeros(1024, dtype=torch.double, device=DEV)
ops.gemm_nt_taps(A, 512, W, bias, C, 512, B, T, 512, 512, 5, -2, stats=st, prec=PREC_FP32X3)
dW = torch.empty(512, 512, 5, device=DEV)
ops.gemm_tn_taps(C, 512, A, 512, dW, B, T, 512, 512, 5, -2, out_mode=1, prec=PREC_FP32X3)
lstm = torch.nn.LSTM(512, 1024, 1, batch_first=True).to(DEV)
x = torch.randn(B, 3, 512, device=DEV)
with torch.no_grad():
    ops.LstmLayer.apply(x, PREC_FP32X3, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0)
torch.cuda.synchronize()
print("done")
