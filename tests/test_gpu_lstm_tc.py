"""Persistent tensor-core LSTM recurrence (AVC_PREC_BF16, lstm_tc.cu) against the fp32 kernels
and torch.nn.LSTM in fp64.  bf16 operands + tanh.approx gate math: tolerances are relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from autovc_b200 import ops
    from autovc_b200._lib import PREC_BF16, PREC_FP32

DEV = "cuda"


def _rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("B,T,I,H", [(4, 12, 48, 128), (130, 9, 32, 512), (256, 16, 64, 1024), (3, 40, 40, 256)])
def test_persistent_lstm_matches_fp32(B, T, I, H):
    torch.manual_seed(3)
    lstm = torch.nn.LSTM(I, H, 1, batch_first=True).to(DEV)
    x = _rand(B, T, I, seed=4).requires_grad_(True)
    ws = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]
    go = _rand(B, T, H, seed=6)
    out16 = ops.LstmLayer.apply(x, PREC_BF16, *ws)
    g16 = torch.autograd.grad(out16, [x] + ws, go)
    out32 = ops.LstmLayer.apply(x, PREC_FP32, *ws)
    g32 = torch.autograd.grad(out32, [x] + ws, go)
    assert torch.isfinite(out16).all()
    assert _rel(out16, out32) < 1e-2, _rel(out16, out32)
    for a, b, n in zip(g16, g32, ["dx", "dw_ih", "dw_hh", "db_ih", "db_hh"]):
        assert torch.isfinite(a).all(), n
        assert _rel(a, b) < 2e-2, (n, _rel(a, b))
