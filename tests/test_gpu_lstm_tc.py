"""Persistent tensor-core LSTM recurrences (bf16 operands, fp32 state) against (i) this library's fp32 kernels
(which tests/test_gpu_kernels.py ties to torch.nn.LSTM) at short sequences and (ii) torch.nn.LSTM in fp64 directly, at
the benched lengths (T=128 / 256) through the half-mode layer the Generator uses.  bf16 operands + tanh.approx gate
math: tolerances are relative."""
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


# AVC_LSTM_FWD_WS / AVC_LSTM_BWD_WS: "1" (default) = weight-stationary kernels (W_hh slice resident in tensor memory, batch as the
# N dimension), "0" = the ring (forward) / shared-memory K-split (BPTT) kernels, which stay the path for shapes the first do
# not take.  The switches are read per call.
@pytest.mark.parametrize("fwd_ws", ["1", "0"])
@pytest.mark.parametrize("B,T,I,H", [(4, 12, 48, 128), (130, 9, 32, 512), (256, 16, 64, 1024), (3, 40, 40, 256)])
def test_persistent_lstm_matches_fp32(B, T, I, H, fwd_ws, monkeypatch):
    monkeypatch.setenv("AVC_LSTM_FWD_WS", fwd_ws)
    monkeypatch.setenv("AVC_LSTM_BWD_WS", fwd_ws)
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


@pytest.mark.parametrize("fwd_ws", ["1", "0"])
@pytest.mark.parametrize("B,T,I,H", [(256, 128, 512, 1024), (130, 128, 288, 512), (128, 256, 512, 1024), (40, 96, 80, 768)])
def test_half_mode_layer_matches_nn_lstm_fp64(B, T, I, H, fwd_ws, monkeypatch):
    """ops.LstmLayerH (input projection on fp16 operands, persistent bf16 recurrence forward and BPTT, bf16 gradient GEMMs)
    against torch.nn.LSTM evaluated in float64 (model_vc_mel.py:90/:104 semantics: gate order i,f,g,o, zero initial state),
    at the sequence lengths and batch sizes of BASELINE.json configs[1]/[2] (two batch tiles, a ragged second tile, H=512/768/1024)."""
    monkeypatch.setenv("AVC_LSTM_FWD_WS", fwd_ws)
    monkeypatch.setenv("AVC_LSTM_BWD_WS", fwd_ws)
    torch.manual_seed(5)
    lstm = torch.nn.LSTM(I, H, 1, batch_first=True).to(DEV)
    ws = [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0]
    x = (0.5 * _rand(B, T, I, seed=8)).requires_grad_(True)
    go = _rand(B, T, H, seed=9) / (B * T) ** 0.5
    out, h16, h16b = ops.LstmLayerH.apply(x, None, None, *ws)
    got = torch.autograd.grad(out, [x] + ws, go)
    ref = torch.nn.LSTM(I, H, 1, batch_first=True).to(DEV).double()
    ref.load_state_dict({k: v.double() for k, v in lstm.state_dict().items()})
    xd = x.detach().double().requires_grad_(True)
    rout, _ = ref(xd)
    rg = torch.autograd.grad(rout, [xd, ref.weight_ih_l0, ref.weight_hh_l0, ref.bias_ih_l0, ref.bias_hh_l0], go.double())
    assert torch.isfinite(out).all()
    assert _rel(out, rout) < 1e-2, _rel(out, rout)
    assert _rel(h16.float(), rout) < 1e-2 and _rel(h16b.float(), rout) < 1.5e-2
    for a, b, n in zip(got, rg, ["dx", "dw_ih", "dw_hh", "db_ih", "db_hh"]):
        assert torch.isfinite(a).all(), n
        assert _rel(a, b) < 2e-2, (n, _rel(a, b))


@pytest.mark.parametrize("B,T,H,reverse,save", [(4, 6, 128, 0, True), (130, 9, 512, 1, True), (256, 16, 1024, 0, True), (37, 12, 768, 0, False),
                                                (300, 10, 256, 1, True), (256, 128, 1024, 0, True), (600, 5, 1024, 1, True),
                                                (20, 7, 192, 0, True), (70, 5, 960, 1, True), (9, 1, 512, 0, True), (33, 2, 1024, 1, False)])
def test_weight_stationary_forward_equals_ring_kernel(B, T, H, reverse, save, monkeypatch):
    """lstm_tc_fwd_ws_kernel (A operand from tensor memory, batch as N) and lstm_tc_fwd_kernel (activation ring) accumulate the
    same products in the same k order: every output -- h, the 16-bit copies, the gates and cell states saved for BPTT -- is
    bit-identical, on ragged batches, both directions, inference mode (nothing saved), and a batch that needs several launches."""
    from autovc_b200 import _lib
    from autovc_b200.ops import _p, _stream, _ws
    g = torch.Generator().manual_seed(B * 131 + H)
    P = (0.5 * torch.randn(B, T, 4 * H, generator=g)).to(DEV)
    Wb = (torch.randn(4 * H, H, generator=g) / H ** 0.5).to(DEV).bfloat16()
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("AVC_LSTM_FWD_WS", flag)
        h = torch.full((B, T, H), float("nan"), device=DEV)
        gates = torch.full((B, T, 4 * H), float("nan"), device=DEV) if save else None
        c = torch.full((B, T, H), float("nan"), device=DEV) if save else None
        h16 = torch.zeros(B, T, H, device=DEV, dtype=torch.float16)
        h16b = torch.zeros(B, T, H, device=DEV, dtype=torch.bfloat16)
        nf = _lib.query("avc_lstm_fwd_workspace_bytes", B, T, H, PREC_BF16)
        wf = _ws(nf, DEV)
        _lib.call("avc_lstm_seq_fwd_h", _p(P), _p(Wb), 1, _p(h), H, _p(gates) if save else None, _p(c) if save else None, _p(h16), 2,
                  _p(h16b), B, T, H, reverse, _p(wf), nf, _stream())
        torch.cuda.synchronize()
        outs[flag] = [h, gates, c, h16, h16b]
    for a, b, n in zip(outs["0"], outs["1"], ["h", "gates", "c", "h16", "h16b"]):
        if a is None:
            continue
        assert torch.isfinite(b.float()).all(), n
        assert torch.equal(a, b), (n, float((a.float() - b.float()).abs().max()))


@pytest.mark.parametrize("B,T,H,reverse,fp32_out", [(4, 6, 128, 0, True), (130, 9, 512, 1, True), (256, 16, 1024, 0, False), (37, 12, 768, 0, True),
                                                    (300, 10, 256, 1, True), (256, 128, 1024, 0, False), (600, 5, 1024, 1, True),
                                                    (70, 5, 896, 1, True), (20, 7, 320, 0, True), (9, 1, 512, 0, True), (33, 2, 1024, 1, False)])
def test_weight_stationary_bptt_equals_ksplit_kernel(B, T, H, reverse, fp32_out, monkeypatch):
    """lstm_tc_bwd_ws_kernel (W_hh^T slice in tensor memory, 128 units x NB utterances per CTA, K split over a cluster of 4) adds the
    four partial sums of a unit in the order lstm_tc_bwd_ks_kernel uses: the gate gradients (fp32 and bf16) are bit-identical."""
    from autovc_b200 import _lib
    from autovc_b200.ops import _p, _stream, _ws
    g = torch.Generator().manual_seed(B * 131 + H)
    P = (0.5 * torch.randn(B, T, 4 * H, generator=g)).to(DEV)
    Wb = (torch.randn(4 * H, H, generator=g) / H ** 0.5).to(DEV).bfloat16()
    WTb = Wb.t().contiguous()
    dH = (0.1 * torch.randn(B, T, H, generator=g)).to(DEV)
    h = torch.empty(B, T, H, device=DEV)
    gates = torch.empty(B, T, 4 * H, device=DEV)
    c = torch.empty(B, T, H, device=DEV)
    h16 = torch.empty(B, T, H, device=DEV, dtype=torch.float16)
    nf = _lib.query("avc_lstm_fwd_workspace_bytes", B, T, H, PREC_BF16)
    wf = _ws(nf, DEV)
    _lib.call("avc_lstm_seq_fwd_h", _p(P), _p(Wb), 1, _p(h), H, _p(gates), _p(c), _p(h16), 2, None, B, T, H, reverse, _p(wf), nf, _stream())
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("AVC_LSTM_BWD_WS", flag)
        dP = torch.full((B, T, 4 * H), float("nan"), device=DEV) if fp32_out else None
        dP16 = torch.zeros(B, T, 4 * H, device=DEV, dtype=torch.bfloat16)
        nb = _lib.query("avc_lstm_bwd_workspace_bytes", B, T, H, PREC_BF16)
        wb = _ws(nb, DEV)
        _lib.call("avc_lstm_seq_bwd_h", _p(dH), H, _p(WTb), 1, _p(gates), _p(c), _p(dP) if fp32_out else None, _p(dP16), B, T, H, reverse,
                  _p(wb), nb, _stream())
        torch.cuda.synchronize()
        outs[flag] = [dP, dP16]
    for a, b, n in zip(outs["0"], outs["1"], ["dP", "dP16"]):
        if a is None:
            continue
        assert torch.isfinite(b.float()).all(), n
        assert float(b.float().abs().max()) > 0, n
        assert torch.equal(a, b), (n, float((a.float() - b.float()).abs().max()))
