"""Race detection without compute-sanitizer (closed on this GPU pool, see profiles/r02_sanitizer_unavailable.log): the kernels
that synchronise across CTAs -- persistent recurrences (global release counters, TMA multicast, DSMEM reduce-scatter),
CTA-pair GEMMs (cta_group::2 MMAs, remote mbarrier arrives), chunked / split-reduction 3xTF32 GEMMs -- are run repeatedly on
the same inputs while other work keeps the SMs and the L2 busy; a missing fence or a barrier-phase slip shows up as a result
that is not BIT-IDENTICAL from run to run (all of these paths reduce in a fixed order).  Values are checked against torch
fp64 elsewhere (test_gpu_lstm_tc.py, test_gpu_tc_gemm.py, test_gpu_kernels.py); a hang is caught by the timeout."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

if torch.cuda.is_available():
    from autovc_b200 import ops
    from autovc_b200._lib import ACT_CODES, PREC_FP32X3

DEV = "cuda"
REPS = 12


def _noise(stream, buf):
    """Unrelated traffic on a second stream: changes the timing of everything between repetitions."""
    with torch.cuda.stream(stream):
        for _ in range(4):
            buf.mul_(1.0001).add_(1e-3)


@pytest.mark.parametrize("B,T,I,H", [(128, 6, 64, 128), (256, 5, 64, 512), (256, 4, 64, 1024), (200, 7, 32, 256)])
def test_persistent_recurrences_are_bit_reproducible(B, T, I, H):
    torch.manual_seed(1)
    lstm = torch.nn.LSTM(I, H, 1, batch_first=True).to(DEV)
    ws = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]
    x = (0.5 * torch.randn(B, T, I, device=DEV)).requires_grad_(True)
    go = torch.randn(B, T, H, device=DEV) / (B * T) ** 0.5
    side, buf = torch.cuda.Stream(), torch.randn(64 << 20, device=DEV)
    ref = None
    for r in range(REPS):
        if r % 2:
            _noise(side, buf)
        out, h16, _ = ops.LstmLayerH.apply(x, None, None, *ws)
        grads = torch.autograd.grad(out, [x] + ws, go)
        cur = [out.detach().clone(), h16.detach().clone()] + [g.detach().clone() for g in grads]
        torch.cuda.synchronize()
        assert all(torch.isfinite(t.float()).all() for t in cur)
        if ref is None:
            ref = cur
        else:
            for i, (a, b) in enumerate(zip(cur, ref)):
                assert torch.equal(a, b), (r, i, float((a.float() - b.float()).abs().max()))


@pytest.mark.parametrize("B,T,Cin,Cout", [(2, 128, 256, 256), (4, 256, 512, 512), (3, 100, 512, 256)])
def test_pair_gemms_are_bit_reproducible(B, T, Cin, Cout):
    torch.manual_seed(2)
    conv = torch.nn.Conv1d(Cin, Cout, 5, padding=2).to(DEV)
    bn = torch.nn.BatchNorm1d(Cout).to(DEV)
    x = torch.randn(B, T, Cin, device=DEV).requires_grad_(True)
    go = torch.randn(B, T, Cout, device=DEV)
    side, buf = torch.cuda.Stream(), torch.randn(64 << 20, device=DEV)
    ref = None
    for r in range(REPS):
        if r % 2:
            _noise(side, buf)
        z, z16, z16b = ops.ConvBnActH.apply(x, None, None, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean.clone(),
                                            bn.running_var.clone(), None, ACT_CODES["relu"], True, True, False)
        gx, gw = torch.autograd.grad(z, [x, conv.weight], go)
        # the BatchNorm statistics are accumulated with fp64 atomics (order-dependent in the last bit of a double): compare the
        # tensors that do not pass through them bit for bit (weight gradient of the raw conv output is downstream of the
        # statistics, so it gets a 1e-6 relative bound instead)
        cur = [gx.detach().clone(), gw.detach().clone(), z.detach().clone()]
        torch.cuda.synchronize()
        if ref is None:
            ref = cur
        else:
            for i, (a, b) in enumerate(zip(cur, ref)):
                err = float((a - b).abs().max() / b.abs().max())
                assert err < 1e-6, (r, i, err)


def test_split_product_gemms_and_recurrences_are_bit_reproducible():
    torch.manual_seed(3)
    B, T, I, H = 256, 5, 96, 256
    lstm = torch.nn.LSTM(I, H, 1, batch_first=True).to(DEV)
    ws = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]
    x = (0.5 * torch.randn(B, T, I, device=DEV)).requires_grad_(True)
    go = torch.randn(B, T, H, device=DEV)
    ref = None
    for r in range(REPS):
        out = ops.LstmLayer.apply(x, PREC_FP32X3, *ws)
        grads = torch.autograd.grad(out, [x] + ws, go)
        cur = [out.detach().clone()] + [g.detach().clone() for g in grads]
        torch.cuda.synchronize()
        if ref is None:
            ref = cur
        else:
            for i, (a, b) in enumerate(zip(cur, ref)):
                assert torch.equal(a, b), (r, i, float((a - b).abs().max()))
