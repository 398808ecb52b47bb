"""GPU unit parity: each C-ABI kernel family against the plain PyTorch fp32 op it replaces."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from autovc_b200 import ops
    from autovc_b200._lib import ACT_CODES, PREC_FP32, PREC_FP32X3

# both implementations of the fp32 parity contract: CUDA-core FFMA ("simt") and 3xTF32 split products on the tensor cores ("x3")
FP32_IMPLS = ["simt", "x3"]


def _prec(impl):
    return PREC_FP32 if impl == "simt" else PREC_FP32X3

DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("B,T,Cin,Cout", [(3, 32, 24, 40), (2, 128, 336, 512), (2, 48, 769, 130), (1, 5, 8, 8)])
@pytest.mark.parametrize("act", ["relu", "tanh", "none"])
@pytest.mark.parametrize("impl", FP32_IMPLS)
def test_conv_bn_act_fwd_bwd(B, T, Cin, Cout, act, impl):
    x = _rand(B, T, Cin, seed=1).requires_grad_(True)
    conv = torch.nn.Conv1d(Cin, Cout, 5, padding=2).to(DEV)
    bn = torch.nn.BatchNorm1d(Cout).to(DEV)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    bn_ref = torch.nn.BatchNorm1d(Cout).to(DEV)
    bn_ref.load_state_dict(bn.state_dict())
    res = _rand(B, T, Cout, seed=5).requires_grad_(True) if act == "none" else None
    z = ops.ConvBnAct.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, res,
                            ACT_CODES[act], True, _prec(impl))
    go = _rand(B, T, Cout, seed=2)
    gx, gw, gb, gg, gbeta = torch.autograd.grad(z, [x, conv.weight, conv.bias, bn.weight, bn.bias], go,
                                                retain_graph=res is not None)
    # reference: torch ops in fp64 on the same device
    xr = x.detach().double().requires_grad_(True)
    cw = conv.weight.detach().double().requires_grad_(True)
    cb = conv.bias.detach().double().requires_grad_(True)
    g_ = bn.weight.detach().double().requires_grad_(True)
    b_ = bn.bias.detach().double().requires_grad_(True)
    y = F.conv1d(xr.transpose(1, 2), cw, cb, padding=2)
    rm = torch.zeros(Cout, dtype=torch.double, device=DEV)
    rv = torch.ones(Cout, dtype=torch.double, device=DEV)
    y = F.batch_norm(y, rm, rv, g_, b_, True, 0.1, 1e-5)
    y = {"relu": F.relu, "tanh": torch.tanh, "none": lambda v: v}[act](y).transpose(1, 2)
    if res is not None:
        y = y + res.detach().double()
    rgx, rgw, rgb, rgg, rgbeta = torch.autograd.grad(y, [xr, cw, cb, g_, b_], go.double())
    tol = dict(rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(z.double(), y, **tol)
    torch.testing.assert_close(gx.double(), rgx, **tol)
    torch.testing.assert_close(gw.double(), rgw, rtol=2e-4, atol=2e-4 * max(1.0, float(rgw.abs().max())))
    torch.testing.assert_close(gg.double(), rgg, rtol=2e-4, atol=1e-3)
    torch.testing.assert_close(gbeta.double(), rgbeta, rtol=2e-4, atol=1e-3)
    assert float(gb.abs().max()) == 0.0 and float(rgb.abs().max()) < 1e-6
    torch.testing.assert_close(bn.running_mean.double(), rm, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn.running_var.double(), rv, rtol=1e-5, atol=1e-6)
    if res is not None:
        (gr,) = torch.autograd.grad(z, [res], go)
        torch.testing.assert_close(gr, go)


@pytest.mark.parametrize("B,T,I,H,bidir", [(5, 24, 40, 16, True), (3, 16, 32, 32, True), (4, 12, 48, 128, False),
                                           (130, 6, 20, 72, False), (2, 9, 12, 20, True), (260, 7, 36, 256, False)])
@pytest.mark.parametrize("impl", FP32_IMPLS)
def test_lstm_layer_fwd_bwd(B, T, I, H, bidir, impl):
    torch.manual_seed(3)
    lstm = torch.nn.LSTM(I, H, 1, batch_first=True, bidirectional=bidir).to(DEV)
    x = _rand(B, T, I, seed=4).requires_grad_(True)
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    ws = [getattr(lstm, n) for n in names]
    if bidir:
        ws += [getattr(lstm, n + "_reverse") for n in names]
    out = ops.LstmLayer.apply(x, _prec(impl), *ws)
    go = _rand(*out.shape, seed=6)
    grads = torch.autograd.grad(out, [x] + ws, go)
    ref = torch.nn.LSTM(I, H, 1, batch_first=True, bidirectional=bidir).to(DEV).double()
    ref.load_state_dict({k: v.double() for k, v in lstm.state_dict().items()})
    xr = x.detach().double().requires_grad_(True)
    with torch.backends.cudnn.flags(enabled=False):
        ro, _ = ref(xr)
        rws = [getattr(ref, n) for n in names] + ([getattr(ref, n + "_reverse") for n in names] if bidir else [])
        rgrads = torch.autograd.grad(ro, [xr] + rws, go.double())
    torch.testing.assert_close(out.double(), ro, rtol=1e-4, atol=1e-5)
    for a, b in zip(grads, rgrads):
        torch.testing.assert_close(a.double(), b, rtol=2e-4, atol=2e-5 * max(1.0, float(b.abs().max())))


@pytest.mark.parametrize("impl", FP32_IMPLS)
def test_linear_and_glue_and_losses(impl):
    B, T, K, N = 3, 20, 64, 80
    x = _rand(B, T, K, seed=1).requires_grad_(True)
    lin = torch.nn.Linear(K, N).to(DEV)
    y = ops.Linear.apply(x, lin.weight, lin.bias, _prec(impl))
    go = _rand(B, T, N, seed=2)
    g = torch.autograd.grad(y, [x, lin.weight, lin.bias], go)
    xr = x.detach().double().requires_grad_(True)
    wr, br = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    yr = F.linear(xr, wr, br)
    gr = torch.autograd.grad(yr, [xr, wr, br], go.double())
    torch.testing.assert_close(y.double(), yr, rtol=1e-5, atol=1e-5)
    for a, b in zip(g, gr):
        torch.testing.assert_close(a.double(), b, rtol=1e-4, atol=1e-4)

    # concat + codes + upsample
    n, f, E = 4, 5, 6
    xm = _rand(B, T, 10, seed=3).requires_grad_(True)
    e = _rand(B, E, seed=4)
    cat = ops.ConcatEmb.apply(xm, e)
    ref = torch.cat([xm, e.unsqueeze(1).expand(-1, T, -1)], -1)
    torch.testing.assert_close(cat, ref)
    gcat = _rand(B, T, 10 + E, seed=5)
    torch.testing.assert_close(torch.autograd.grad(cat, xm, gcat)[0], gcat[..., :10].contiguous())
    enc = _rand(B, T, 2 * n, seed=6).requires_grad_(True)
    codes = ops.Codes.apply(enc, n, f)
    rc = torch.stack([torch.cat([enc[:, i + f - 1, :n], enc[:, i, n:]], -1) for i in range(0, T, f)], 1)
    torch.testing.assert_close(codes, rc)
    gc = _rand(*codes.shape, seed=7)
    torch.testing.assert_close(torch.autograd.grad(codes, enc, gc)[0], torch.autograd.grad(rc, enc, gc)[0])
    cd = codes.detach().requires_grad_(True)
    up = ops.UpsampleConcat.apply(cd, e, T)
    rup = torch.cat([cd.repeat_interleave(f, dim=1), e.unsqueeze(1).expand(-1, T, -1)], -1)
    torch.testing.assert_close(up, rup)
    gu = _rand(*up.shape, seed=8)
    torch.testing.assert_close(torch.autograd.grad(up, cd, gu)[0], torch.autograd.grad(rup, cd, gu)[0], rtol=1e-5, atol=1e-5)

    # losses
    a = _rand(7, 33, seed=9).requires_grad_(True)
    b = _rand(7, 33, seed=10).requires_grad_(True)
    for mine, theirs in ((ops.mse_loss, F.mse_loss), (ops.l1_loss, F.l1_loss)):
        l, lr = mine(a, b), theirs(a, b)
        torch.testing.assert_close(l, lr, rtol=1e-5, atol=1e-6)
        ga, gb = torch.autograd.grad(l * 3.0, [a, b])
        ra, rb = torch.autograd.grad(lr * 3.0, [a, b])
        torch.testing.assert_close(ga, ra, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(gb, rb, rtol=1e-5, atol=1e-7)


def test_no_cpu_fallback():
    from autovc_b200 import AvcError
    with pytest.raises(AvcError):
        ops.mse_loss(torch.zeros(4), torch.zeros(4))


@pytest.mark.gpu
@pytest.mark.parametrize("M,C", [(4096, 512), (1000, 80), (333, 30), (77, 516)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_bn_backward_recompute_matches_z_read(M, C, act):
    """avc_bn_act_bwd_{reduce,apply}_y (activation recomputed from y) == the z-reading entry points, bit for bit,
    and the column-sum kernels agree with torch in fp64."""
    from autovc_b200 import _lib
    from autovc_b200.ops import _p, _stream, _NULL
    y = _rand(M, C, seed=1)
    dz = _rand(M, C, seed=2)
    gamma, beta = _rand(C, seed=3) + 1.5, _rand(C, seed=4) * 0.3
    stats = torch.zeros(2 * C, dtype=torch.float64, device=DEV)
    _lib.call("avc_channel_stats", _p(y), C, M, C, _p(stats), _stream())
    torch.testing.assert_close(stats[:C], y.double().sum(0), rtol=1e-6, atol=1e-4)
    torch.testing.assert_close(stats[C:], (y.double() ** 2).sum(0), rtol=1e-6, atol=1e-4)
    mean, rstd = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    _lib.call("avc_bn_finalize", _p(stats), M, C, 1e-5, 0.1, _p(mean), _p(rstd), _NULL, _NULL, _stream())
    z = torch.empty_like(y)
    _lib.call("avc_bn_act_fwd", _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _NULL, _p(z), M, C, act, _stream())
    pre = (y - mean) * (rstd * gamma) + beta
    ref_z = {0: pre, 1: pre.clamp_min(0), 2: pre.tanh()}[act]
    torch.testing.assert_close(z, ref_z, rtol=1e-5, atol=1e-5)
    s_old = torch.zeros(2 * C, dtype=torch.float64, device=DEV)
    s_new = torch.zeros_like(s_old)
    _lib.call("avc_bn_act_bwd_reduce", _p(dz), _p(z), _p(y), _p(mean), _p(rstd), _p(s_old), M, C, act, _stream())
    # z may be omitted only on the float4 path
    zarg = _NULL if C % 4 == 0 else _p(z)
    _lib.call("avc_bn_act_bwd_reduce_y", _p(dz), zarg, _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(s_new), M, C, act, _stream())
    torch.testing.assert_close(s_new, s_old, rtol=1e-12, atol=1e-9)     # same values, atomics in another order
    dy_old, dy_new = torch.empty_like(y), torch.empty_like(y)
    dg, db = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    _lib.call("avc_bn_act_bwd_apply", _p(dz), _p(z), _p(y), _p(mean), _p(rstd), _p(gamma), _p(s_old), _p(dy_old), _p(dg), _p(db),
              M, C, act, 0, _stream())
    dy16 = torch.empty(M, C, dtype=torch.bfloat16, device=DEV) if C % 4 == 0 else None
    _lib.call("avc_bn_act_bwd_apply_y", _p(dz), zarg, _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(s_old), _p(dy_new),
              _p(dy16), 1, _p(dg), _p(db), M, C, act, 0, _stream())
    assert torch.equal(dy_new, dy_old)
    if dy16 is not None:
        assert torch.equal(dy16, dy_old.bfloat16())
    if C % 4 != 0:
        with pytest.raises(_lib.AvcError):
            _lib.call("avc_bn_act_bwd_reduce_y", _p(dz), _NULL, _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(s_new), M, C, act, _stream())


@pytest.mark.gpu
@pytest.mark.parametrize("M,C,dt,fmt", [(4096, 4096, torch.bfloat16, 1), (1000, 64, torch.float16, 2), (77, 516, torch.bfloat16, 1)])
def test_colsum16(M, C, dt, fmt):
    """avc_colsum16: column sums of a 16-bit matrix, plain and with the LSTM un-permute (c = u*4+g -> g*H+u)."""
    from autovc_b200 import ops
    x = (_rand(M, C, seed=9) * 0.1).to(dt)
    out = torch.empty(C, device=DEV)
    ops.colsum16(x, fmt, C, M, C, out)
    torch.testing.assert_close(out, x.double().sum(0).float(), rtol=1e-5, atol=1e-5)
    o1, o2 = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    ops.colsum16(x, fmt, C, M, C, o1, o2, out_mode=2)
    ref = x.double().sum(0).view(C // 4, 4).t().reshape(-1).float()
    torch.testing.assert_close(o1, ref, rtol=1e-5, atol=1e-5)
    assert torch.equal(o1, o2)
