"""tcgen05/TMA GEMM-with-taps (AVC_PREC_BF16) against an fp64 product of the bf16-rounded operands:
the only difference allowed is fp32 accumulation order."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from autovc_b200 import ops
    from autovc_b200._lib import PREC_BF16, PREC_FP32, PREC_TF32

DEV = "cuda"


def _rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV)


def _bf(x, prec=1):
    """Operand rounding of the tensor-core path: bf16 (RN), or tf32 (RN to 10 mantissa bits, done by the TMA unit)."""
    if prec == 1:
        return x.to(torch.bfloat16).double()
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32).double()


def _ref_nt(A, W, bias, nB, T, ntaps, shift0, prec=1):
    """A (nB*T, K), W (ntaps, N, K) -> (nB*T, N) in fp64 from bf16-rounded operands."""
    K = A.shape[1]
    A3 = _bf(A, prec).view(nB, T, K)
    out = torch.zeros(nB, T, W.shape[1], dtype=torch.double, device=DEV)
    for tap in range(ntaps):
        s = shift0 + tap
        lo, hi = max(0, -s), min(T, T - s)
        if hi > lo:
            out[:, lo:hi] += A3[:, lo + s:hi + s] @ _bf(W[tap], prec).t()
    if bias is not None:
        out += bias.double()
    return out.view(nB * T, -1)


@pytest.mark.parametrize("nB,T,N,K,ntaps", [(2, 128, 512, 336, 5), (3, 128, 128, 64, 1), (2, 48, 130, 769, 5),
                                            (4, 256, 80, 1024, 1), (1, 20, 64, 512, 1), (5, 64, 2048, 288, 1),
                                            # enough tiles for the 256-wide CTA-pair kernel: odd tile counts, ragged N / K / T
                                            (40, 128, 640, 96, 5), (37, 100, 520, 72, 3), (75, 128, 256, 64, 1)])
@pytest.mark.parametrize("prec", [1, 2])
def test_nt_taps_tensor(nB, T, N, K, ntaps, prec):
    A = _rand(nB * T, K, seed=1)
    W = _rand(ntaps, N, K, seed=2) * 0.05
    bias = _rand(N, seed=3)
    C = torch.empty(nB * T, N, device=DEV)
    stats = torch.zeros(2 * N, dtype=torch.double, device=DEV)
    shift0 = -(ntaps // 2)
    ops.gemm_nt_taps(A, K, W, bias, C, N, nB, T, N, K, ntaps, shift0, stats=stats, prec=prec)
    ref = _ref_nt(A, W, bias, nB, T, ntaps, shift0, prec)
    scale = float(ref.abs().max())
    tol = (2e-5 if prec == 1 else 2e-4) * scale * max(1.0, (K * ntaps / 256) ** 0.5)   # tf32: emulated RN vs the TMA unit's rounding
    assert float((C.double() - ref).abs().max()) < tol
    torch.testing.assert_close(stats[:N], ref.sum(0), rtol=1e-4, atol=1e-3 * scale)
    torch.testing.assert_close(stats[N:], (ref * ref).sum(0), rtol=1e-4, atol=1e-3 * scale * scale)
    # accumulate mode, no bias
    C2 = C.clone()
    ops.gemm_nt_taps(A, K, W, None, C2, N, nB, T, N, K, ntaps, shift0, accumulate=True, prec=prec)
    ref2 = C.double() + _ref_nt(A, W, None, nB, T, ntaps, shift0, prec)
    assert float((C2.double() - ref2).abs().max()) < 2 * tol


@pytest.mark.parametrize("nB,T,N,K,ntaps,shift0,mode", [(2, 128, 512, 336, 5, -2, 1), (3, 64, 64, 512, 1, 0, 2),
                                                        (2, 48, 130, 769, 5, -2, 1), (4, 128, 4096, 1024, 1, -1, 2),
                                                        (4, 128, 2048, 512, 1, 1, 2), (3, 100, 80, 1024, 1, 0, 0),
                                                        (8, 128, 512, 512, 5, -2, 1), (5, 96, 256, 768, 3, -1, 0)])
@pytest.mark.parametrize("prec", [1, 2])
def test_tn_taps_tensor(nB, T, N, K, ntaps, shift0, mode, prec):
    dY = _rand(nB * T, N, seed=4)
    X = _rand(nB * T, K, seed=5)
    shape = {0: (ntaps, N, K), 1: (N, K, ntaps), 2: (N, K)}[mode]
    dW = torch.empty(*shape, device=DEV)
    ops.gemm_tn_taps(dY, N, X, K, dW, nB, T, N, K, ntaps, shift0, out_mode=mode, prec=prec)
    Y3, X3 = _bf(dY, prec).view(nB, T, N), _bf(X, prec).view(nB, T, K)
    ref = torch.zeros(ntaps, N, K, dtype=torch.double, device=DEV)
    for tap in range(ntaps):
        s = shift0 + tap
        lo, hi = max(0, -s), min(T, T - s)
        if hi > lo:
            ref[tap] = torch.einsum("btn,btk->nk", Y3[:, lo:hi], X3[:, lo + s:hi + s])
    if mode == 1:
        ref = ref.permute(1, 2, 0)
    elif mode == 2:
        H = N // 4
        ref = ref[0].view(H, 4, K).permute(1, 0, 2).reshape(N, K)     # packed row u*4+g -> g*H+u
    scale = float(ref.abs().max())
    assert float((dW.double() - ref).abs().max()) < (3e-5 if prec == 1 else 3e-4) * scale * max(1.0, (nB * T / 256) ** 0.5)
    # the fp32 path computes the same contraction from unrounded operands
    dW32 = torch.empty_like(dW)
    ops.gemm_tn_taps(dY, N, X, K, dW32, nB, T, N, K, ntaps, shift0, out_mode=mode, prec=PREC_FP32)
    assert float((dW32 - dW).abs().max()) < 2e-2 * scale


def _r16(x, dt):
    return x.to(dt).double()


@pytest.mark.parametrize("a_dt,w_fmt", [(torch.float16, 2), (torch.bfloat16, 1)])
def test_nt_taps_prestaged_16bit_operands(a_dt, w_fmt):
    """'half' mode GEMM: A already 16-bit in HBM (read in place by TMA), W staged to w_fmt; includes the mixed
    bf16 x fp16 product."""
    from autovc_b200._lib import FMT_BF16, FMT_FP16
    nB, T, N, K, ntaps = 3, 128, 512, 336, 5
    A = _rand(nB * T, K, seed=1)
    W = _rand(ntaps, N, K, seed=2) * 0.05
    bias = _rand(N, seed=3)
    A16 = A.to(a_dt).contiguous()
    C = torch.empty(nB * T, N, device=DEV)
    stats = torch.zeros(2 * N, dtype=torch.double, device=DEV)
    ops.gemm_nt_taps_h(A16, FMT_FP16 if a_dt == torch.float16 else FMT_BF16, K, W, bias, C, N, nB, T, N, K, ntaps, -2, w_fmt, stats=stats)
    w_dt = torch.float16 if w_fmt == 2 else torch.bfloat16
    A3 = A16.double().view(nB, T, K)
    ref = torch.zeros(nB, T, N, dtype=torch.double, device=DEV)
    for tap in range(ntaps):
        s = tap - 2
        lo, hi = max(0, -s), min(T, T - s)
        ref[:, lo:hi] += A3[:, lo + s:hi + s] @ _r16(W[tap], w_dt).t()
    ref = (ref + bias.double()).view(nB * T, N)
    scale = float(ref.abs().max())
    assert float((C.double() - ref).abs().max()) < 3e-5 * scale * (K * ntaps / 256) ** 0.5
    torch.testing.assert_close(stats[:N], ref.sum(0), rtol=1e-4, atol=1e-3 * scale)


@pytest.mark.parametrize("y_dt,x_dt", [(torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16)])
def test_tn_taps_prestaged_16bit_operands(y_dt, x_dt):
    from autovc_b200._lib import FMT_BF16, FMT_FP16
    nB, T, N, K, ntaps, shift0 = 2, 128, 512, 512, 5, -2
    dY = _rand(nB * T, N, seed=4).to(y_dt).contiguous()
    X = _rand(nB * T, K, seed=5).to(x_dt).contiguous()
    dW = torch.empty(N, K, ntaps, device=DEV)
    f = lambda dt: FMT_FP16 if dt == torch.float16 else FMT_BF16
    ops.gemm_tn_taps_h(dY, f(y_dt), N, X, f(x_dt), K, dW, nB, T, N, K, ntaps, shift0, out_mode=1)
    Y3, X3 = dY.double().view(nB, T, N), X.double().view(nB, T, K)
    ref = torch.zeros(ntaps, N, K, dtype=torch.double, device=DEV)
    for tap in range(ntaps):
        s = shift0 + tap
        lo, hi = max(0, -s), min(T, T - s)
        ref[tap] = torch.einsum("btn,btk->nk", Y3[:, lo:hi], X3[:, lo + s:hi + s])
    ref = ref.permute(1, 2, 0)
    scale = float(ref.abs().max())
    assert float((dW.double() - ref).abs().max()) < 3e-5 * scale


@pytest.mark.gpu
@pytest.mark.parametrize("Cout,Cin,k", [(512, 336, 5), (80, 512, 5), (40, 30, 3), (512, 80, 1)])
def test_pack_conv_weight_h_layouts(Cout, Cin, k):
    """avc_pack_conv_weight(_h): fwd [tap][Cout][ld] / dgrad [k-1-tap][Cin][ld], fp32 and 16-bit, zero-filled tails."""
    from autovc_b200 import _lib
    from autovc_b200.ops import _p, _stream
    w = _rand(Cout, Cin, k, seed=5)
    wf32 = torch.empty(k, Cout, Cin, device=DEV)
    wd32 = torch.empty(k, Cin, Cout, device=DEV)
    _lib.call("avc_pack_conv_weight", _p(w), _p(wf32), _p(wd32), Cout, Cin, k, _stream())
    assert torch.equal(wf32, w.permute(2, 0, 1).contiguous())
    assert torch.equal(wd32, w.flip(2).permute(2, 1, 0).contiguous())
    ldf, ldd = (Cin + 7) // 8 * 8, (Cout + 7) // 8 * 8
    wf = torch.full((k, Cout, ldf), 7.0, device=DEV, dtype=torch.float16)
    wd = torch.full((k, Cin, ldd), 7.0, device=DEV, dtype=torch.bfloat16)
    _lib.call("avc_pack_conv_weight_h", _p(w), _p(wf), ldf, 2, _p(wd), ldd, 1, Cout, Cin, k, _stream())
    assert torch.equal(wf[..., :Cin], wf32.half()) and float(wf[..., Cin:].abs().sum()) == 0.0
    assert torch.equal(wd[..., :Cout], wd32.bfloat16()) and float(wd[..., Cout:].abs().sum()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("H,I", [(1024, 512), (16, 30), (512, 288)])
def test_pack_lstm_weight_h_layouts(H, I):
    from autovc_b200 import _lib
    from autovc_b200.ops import _p, _stream
    w = _rand(4 * H, I, seed=6)
    ref = w.view(4, H, I).permute(1, 0, 2).reshape(4 * H, I).contiguous()       # row u*4+g
    p32, pT32 = torch.empty(4 * H, I, device=DEV), torch.empty(I, 4 * H, device=DEV)
    _lib.call("avc_pack_lstm_weight", _p(w), _p(p32), _p(pT32), H, I, _stream())
    assert torch.equal(p32, ref) and torch.equal(pT32, ref.t().contiguous())
    ldp = (I + 7) // 8 * 8
    p16 = torch.full((4 * H, ldp), 7.0, device=DEV, dtype=torch.float16)
    pT16 = torch.full((I, 4 * H), 7.0, device=DEV, dtype=torch.bfloat16)
    _lib.call("avc_pack_lstm_weight_h", _p(w), _p(p16), ldp, 2, _p(pT16), 4 * H, 1, H, I, _stream())
    assert torch.equal(p16[:, :I], ref.half()) and float(p16[:, I:].abs().sum()) == 0.0
    assert torch.equal(pT16, ref.t().contiguous().bfloat16())


@pytest.mark.gpu
@pytest.mark.parametrize("nB,T,N,K,ntaps", [(40, 128, 640, 96, 5), (3, 100, 130, 200, 3), (4, 128, 2048, 288, 1)])
def test_nt_taps_prepacked_weight_equals_staged(nB, T, N, K, ntaps):
    """avc_gemm_nt_taps_hw (W packed once as 16-bit) == avc_gemm_nt_taps_h (W staged per call), bit for bit."""
    A = _rand(nB * T, K, seed=1).half()
    W = _rand(ntaps, N, K, seed=2) * 0.05
    bias = _rand(N, seed=3)
    shift0 = -(ntaps // 2)
    C1 = torch.empty(nB * T, N, device=DEV)
    C2 = torch.empty_like(C1)
    s1 = torch.zeros(2 * N, dtype=torch.double, device=DEV)
    s2 = torch.zeros_like(s1)
    ops.gemm_nt_taps_h(A, 2, K, W, bias, C1, N, nB, T, N, K, ntaps, shift0, 2, stats=s1)
    ldw = (K + 7) // 8 * 8
    W16 = torch.zeros(ntaps, N, ldw, device=DEV, dtype=torch.float16)
    W16[..., :K] = W.half()
    ops.gemm_nt_taps_hw(A, 2, K, W16, 2, ldw, bias, C2, N, nB, T, N, K, ntaps, shift0, stats=s2)
    assert torch.equal(C1, C2)
    torch.testing.assert_close(s1, s2, rtol=1e-12, atol=1e-9)
    # an fp32 A is staged to W's format
    C3 = torch.empty_like(C1)
    ops.gemm_nt_taps_hw(A.float(), 0, K, W16, 2, ldw, bias, C3, N, nB, T, N, K, ntaps, shift0)
    assert torch.equal(C1, C3)
    # mixed 16-bit formats are refused (tcgen05 kind::f16 would trap)
    with pytest.raises(Exception):
        ops.gemm_nt_taps_hw(A.bfloat16(), 1, K, W16, 2, ldw, bias, C3, N, nB, T, N, K, ntaps, shift0)


@pytest.mark.parametrize("prec", [1, 2])
def test_single_cta_fallback_switch_matches_pair_kernels(prec, monkeypatch):
    """AVC_GEMM_2CTA=0 (the one selection switch the library keeps: it forces the single-CTA kernels where a CTA pair would
    run) computes the same contraction: both against the fp64 product, and against each other within accumulation order."""
    nB, T, N, K, ntaps = 40, 128, 512, 256, 5
    A = _rand(nB * T, K, seed=11)
    W = _rand(ntaps, N, K, seed=12) * 0.05
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("AVC_GEMM_2CTA", flag)
        C = torch.empty(nB * T, N, device=DEV)
        ops.gemm_nt_taps(A, K, W, None, C, N, nB, T, N, K, ntaps, -2, prec=prec)
        dW = torch.empty(N, K, ntaps, device=DEV)
        ops.gemm_tn_taps(C, N, A, K, dW, nB, T, N, K, ntaps, -2, out_mode=1, prec=prec)
        torch.cuda.synchronize()
        out[flag] = (C, dW)
    ref = _ref_nt(A, W, None, nB, T, ntaps, -2, prec)
    scale = float(ref.abs().max())
    tol = (2e-5 if prec == 1 else 2e-4) * scale * (K * ntaps / 256) ** 0.5
    for flag in out:
        assert float((out[flag][0].double() - ref).abs().max()) < tol, flag
    assert float((out["0"][1] - out["1"][1]).abs().max()) <= 2e-3 * float(out["1"][1].abs().max())
