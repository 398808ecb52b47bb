"""autograd Functions of the waveform variant (model_vc_wav.py:11-102; solver_encoder.py:264-290) over the C-ABI.

The two filterbank layers and the k=3 (transposed) convolutions run on the same taps-GEMM kernels as the mel model:

* ``Conv1d(1 -> N, k = taps*S, stride = S)`` over a waveform (B, L) is a ``taps``-tap convolution over the view
  (B, L/S, S): frame t of the output reads waveform blocks t .. t+taps-1 (model_vc_wav.py:18: N=512, k=1024, S=256).
* ``ConvTranspose1d(N -> 1, k = taps*S, stride = S)`` (:52) is that convolution's adjoint, i.e. its data-gradient form.
* ``ConvTranspose1d(N -> N, k=3, s=1, p=1)`` (:44) is a k=3 convolution with the weight read as the adjoint layer's.

Activations are channels-last (B, T, C) float32 like everywhere else in the package; no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ACT_NONE, PREC_FP32, call
from .ops import (BN_EPS, BN_MOMENTUM, _GLOBAL_CACHE, _NULL, PackCache, _check, _p, _stream, colsum, gemm_nt_taps,
                  gemm_tn_taps, pack_conv)


def copy_rows3d(src, Tsrc, dst, Tdst, nB, C):
    call("avc_copy_rows3d", _p(src), Tsrc, _p(dst), Tdst, nB, C, _stream())


def pack_frame(weight: torch.Tensor, S: int, cache: PackCache = _GLOBAL_CACHE):
    """Filterbank weight (N, 1, taps*S) -> forward pack [tap][N][S] and adjoint pack [tap][S][N] with the taps reversed."""
    def build():
        N, one, K = weight.shape
        taps = K // S
        w = weight.detach()
        wf = torch.empty(taps, N, S, device=w.device, dtype=torch.float32)
        call("avc_permute021", _p(w), _p(wf), N, taps, S, _stream())
        wd = torch.empty(taps, S, N, device=w.device, dtype=torch.float32)
        for tp in range(taps):
            call("avc_transpose", _p(wf[taps - 1 - tp]), _p(wd[tp]), N, S, _stream())
        return wf, wd
    return cache.get(("frame", S), weight, build)


def _frame_dims(weight, S):
    N, one, K = weight.shape
    if one != 1 or K % S != 0:
        raise _lib.AvcError(f"filterbank layer: weight {tuple(weight.shape)} needs 1 channel and a kernel that is a multiple of the stride {S}")
    return N, K, K // S


class FrameConv(torch.autograd.Function):
    """Conv1d(1 -> N, kernel taps*S, stride S, no padding): (B, L) -> (B, L/S - taps + 1, N)   (model_vc_wav.py:18,:30)."""

    @staticmethod
    def forward(ctx, x, weight, bias, S: int, prec: int):
        x = x.contiguous()
        _check(x, weight, bias)
        N, K, taps = _frame_dims(weight, S)
        B, L = x.shape
        if L % S != 0 or L < K:
            raise _lib.AvcError(f"filterbank layer: waveform length {L} must be a multiple of the stride {S} and >= {K}")
        Tin = L // S
        T = Tin - taps + 1
        wf, wd = pack_frame(weight, S)
        yfull = torch.empty(B, Tin, N, device=x.device, dtype=torch.float32)
        gemm_nt_taps(x, S, wf, bias, yfull, N, B, Tin, N, S, taps, 0, prec=prec)
        y = torch.empty(B, T, N, device=x.device, dtype=torch.float32)
        copy_rows3d(yfull, Tin, y, T, B, N)
        ctx.save_for_backward(x, weight)
        ctx.wd, ctx.S, ctx.prec = wd, S, prec
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        S, prec = ctx.S, ctx.prec
        N, K, taps = _frame_dims(weight, S)
        B, L = x.shape
        Tin = L // S
        T = Tin - taps + 1
        dyf = torch.empty(B, Tin, N, device=x.device, dtype=torch.float32)
        copy_rows3d(dy, T, dyf, Tin, B, N)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, L, device=x.device, dtype=torch.float32)
            gemm_nt_taps(dyf, N, ctx.wd, None, dx, S, B, Tin, S, N, taps, -(taps - 1), prec=prec)
        dwp = torch.empty(taps, N, S, device=x.device, dtype=torch.float32)
        gemm_tn_taps(dyf, N, x, S, dwp, B, Tin, N, S, taps, 0, out_mode=0, prec=prec)
        dw = torch.empty_like(weight)
        call("avc_permute021", _p(dwp), _p(dw), taps, N, S, _stream())
        db = torch.empty(N, device=x.device, dtype=torch.float32)
        colsum(dy, N, B * T, N, db)
        return dx, dw, db, None, None


class FrameConvT(torch.autograd.Function):
    """ConvTranspose1d(N -> 1, kernel taps*S, stride S, no padding): (B, T, N) -> (B, (T + taps - 1)*S)
    (model_vc_wav.py:52,:57)."""

    @staticmethod
    def forward(ctx, x, weight, bias, S: int, prec: int):
        x = x.contiguous()
        _check(x, weight, bias)
        N, K, taps = _frame_dims(weight, S)
        B, T, Nx = x.shape
        if Nx != N:
            raise _lib.AvcError(f"synthesis layer: input has {Nx} channels, weight expects {N}")
        Tout = T + taps - 1
        wf, wd = pack_frame(weight, S)
        xp = torch.empty(B, Tout, N, device=x.device, dtype=torch.float32)
        copy_rows3d(x, T, xp, Tout, B, N)
        out = torch.empty(B, Tout * S, device=x.device, dtype=torch.float32)
        bias_v = bias.detach().expand(S).contiguous()
        gemm_nt_taps(xp, N, wd, bias_v, out, S, B, Tout, S, N, taps, -(taps - 1), prec=prec)
        ctx.save_for_backward(xp, weight)
        ctx.wf, ctx.S, ctx.prec, ctx.T = wf, S, prec, T
        return out

    @staticmethod
    def backward(ctx, dout):
        xp, weight = ctx.saved_tensors
        dout = dout.contiguous()
        S, prec, T = ctx.S, ctx.prec, ctx.T
        N, K, taps = _frame_dims(weight, S)
        B, Tout, _ = xp.shape
        dx = None
        if ctx.needs_input_grad[0]:
            dxf = torch.empty(B, Tout, N, device=xp.device, dtype=torch.float32)
            gemm_nt_taps(dout, S, ctx.wf, None, dxf, N, B, Tout, N, S, taps, 0, prec=prec)
            dx = torch.empty(B, T, N, device=xp.device, dtype=torch.float32)
            copy_rows3d(dxf, Tout, dx, T, B, N)
        dwp = torch.empty(taps, N, S, device=xp.device, dtype=torch.float32)
        gemm_tn_taps(xp, N, dout, S, dwp, B, Tout, N, S, taps, 0, out_mode=0, prec=prec)
        dw = torch.empty_like(weight)
        call("avc_permute021", _p(dwp), _p(dw), taps, N, S, _stream())
        db = torch.empty(1, device=xp.device, dtype=torch.float32)
        scratch = torch.zeros(1, device=xp.device, dtype=torch.float64)
        call("avc_sum_all", _p(dout), dout.numel(), _p(scratch), _p(db), 0, _stream())
        return dx, dw, db, None, None


class ConvPReLUBn(torch.autograd.Function):
    """BatchNorm1d(PReLU(conv(x))) with conv = Conv1d(k, s=1, 'same') or, ``transposed``, ConvTranspose1d(k, s=1, p=k//2)
    (model_vc_wav.py:22-25 / :44-47).  Unlike the mel model's conv -> BN -> act blocks the activation sits BEFORE the
    normalisation, so the batch statistics are taken on the PReLU output and the conv bias does receive a gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, slope, gamma, beta, running_mean, running_var, transposed: bool, training: bool,
                prec: int):
        x = x.contiguous()
        _check(x, weight, bias, slope, gamma, beta, running_mean, running_var)
        if slope.numel() != 1:
            raise _lib.AvcError("PReLU with one shared slope is what the reference builds (nn.PReLU())")
        B, T, Cin = x.shape
        k = weight.shape[2]
        Cout = weight.shape[1] if transposed else weight.shape[0]
        if (weight.shape[0] if transposed else weight.shape[1]) != Cin or k % 2 != 1:
            raise _lib.AvcError(f"conv: input has {Cin} channels, weight is {tuple(weight.shape)}")
        M = B * T
        wa, wb = pack_conv(weight)          # [k][dim0][dim1] and [k][dim1][dim0] with the taps reversed
        w_fwd, w_dgrad = (wb, wa) if transposed else (wa, wb)
        y = torch.empty(B, T, Cout, device=x.device, dtype=torch.float32)
        gemm_nt_taps(x, Cin, w_fwd, bias, y, Cout, B, T, Cout, Cin, k, -(k // 2), prec=prec)
        p = torch.empty_like(y)
        mean = torch.empty(Cout, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        if training:
            stats = torch.zeros(2 * Cout, device=x.device, dtype=torch.float64)
            call("avc_prelu_fwd", _p(y), _p(slope), _p(p), _p(stats), M, Cout, _stream())
            call("avc_bn_finalize", _p(stats), M, Cout, BN_EPS, BN_MOMENTUM, _p(mean), _p(rstd), _p(running_mean),
                 _p(running_var), _stream())
        else:
            call("avc_prelu_fwd", _p(y), _p(slope), _p(p), _NULL, M, Cout, _stream())
            call("avc_bn_eval_stats", _p(running_mean), _p(running_var), Cout, BN_EPS, _p(mean), _p(rstd), _stream())
        z = torch.empty_like(y)
        call("avc_bn_act_fwd", _p(p), _p(mean), _p(rstd), _p(gamma), _p(beta), _NULL, _p(z), M, Cout, ACT_NONE, _stream())
        ctx.save_for_backward(x, weight, slope, gamma, beta, y, p, z, mean, rstd)
        ctx.transposed, ctx.training, ctx.prec, ctx.w_dgrad = transposed, training, prec, w_dgrad
        return z

    @staticmethod
    def backward(ctx, dz):
        x, weight, slope, gamma, beta, y, p, z, mean, rstd = ctx.saved_tensors
        if not ctx.training:
            raise _lib.AvcError("backward through eval-mode BatchNorm is not on the supported path")
        dz = dz.contiguous()
        B, T, Cin = x.shape
        Cout = y.shape[2]
        k = weight.shape[2]
        M = B * T
        prec = ctx.prec
        sums = torch.zeros(2 * Cout, device=x.device, dtype=torch.float64)
        call("avc_bn_act_bwd_reduce_y", _p(dz), _p(z), _p(p), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(sums), M, Cout,
             ACT_NONE, _stream())
        dp = torch.empty_like(y)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        call("avc_bn_act_bwd_apply_y", _p(dz), _p(z), _p(p), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(sums), _p(dp), _NULL, 0,
             _p(dgamma), _p(dbeta), M, Cout, ACT_NONE, 0, _stream())
        dy = torch.empty_like(y)
        dslope = torch.empty_like(slope)
        scratch = torch.zeros(1, device=x.device, dtype=torch.float64)
        call("avc_prelu_bwd", _p(dp), _p(y), _p(slope), _p(dy), _p(dslope), 0, _p(scratch), dy.numel(), _stream())
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            gemm_nt_taps(dy, Cout, ctx.w_dgrad, None, dx, Cin, B, T, Cin, Cout, k, -(k // 2), prec=prec)
        dw = torch.empty_like(weight)
        if ctx.transposed:      # weight (Cin, Cout, k): the adjoint layer's "output gradient" is x, its "input" is dy
            gemm_tn_taps(x, Cin, dy, Cout, dw, B, T, Cin, Cout, k, -(k // 2), out_mode=1, prec=prec)
        else:
            gemm_tn_taps(dy, Cout, x, Cin, dw, B, T, Cout, Cin, k, -(k // 2), out_mode=1, prec=prec)
        db = torch.empty(Cout, device=x.device, dtype=torch.float32)
        colsum(dy, Cout, M, Cout, db)
        return dx, dw, db, dslope, dgamma, dbeta, None, None, None, None, None


class SiSnr(torch.autograd.Function):
    """The SI-SNR term of solver_encoder.py:276-283 on est = x_identic, tgt = x_real, both (B, L) (or (B, L, 1))."""

    @staticmethod
    def forward(ctx, est, tgt):
        if tgt.requires_grad:
            raise _lib.AvcError("SiSnr: the target (x_real) carries no gradient on the reference's path")
        B = est.shape[0]
        ctx.est_shape = est.shape
        est = est.reshape(B, -1).contiguous()
        tgt = tgt.reshape(B, -1).contiguous()
        _check(est, tgt)
        if est.shape != tgt.shape:
            raise _lib.AvcError("SiSnr operands differ in shape")
        L = est.shape[1]
        scratch = torch.zeros(4 * B, device=est.device, dtype=torch.float64)
        saved = torch.empty(4 * B, device=est.device, dtype=torch.float32)
        out = torch.empty((), device=est.device, dtype=torch.float32)
        call("avc_sisnr_fwd", _p(est), _p(tgt), B, L, _p(scratch), _p(saved), _p(out), _stream())
        ctx.save_for_backward(est, tgt, saved)
        return out

    @staticmethod
    def backward(ctx, gout):
        est, tgt, saved = ctx.saved_tensors
        gout = gout.contiguous().float()
        B, L = est.shape
        dest = torch.empty_like(est)
        call("avc_sisnr_bwd", _p(est), _p(tgt), _p(saved), _p(gout), B, L, _p(dest), 0, _stream())
        return dest.view(ctx.est_shape), None


def sisnr_loss(est, tgt):
    return SiSnr.apply(est, tgt)


__all__ = ["FrameConv", "FrameConvT", "ConvPReLUBn", "SiSnr", "sisnr_loss", "PREC_FP32"]
