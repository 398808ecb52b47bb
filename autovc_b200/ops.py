"""torch.autograd.Function wrappers over the C-ABI kernels (include/autovc_b200.h).

PyTorch is plumbing here: it owns device memory (caching allocator), the current stream and
the autograd graph.  All arithmetic on the path happens in libautovc_b200.so.  Activations
are channels-last (B, T, C) float32 contiguous.  No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import ACT_CODES, ACT_NONE, FMT_BF16, FMT_FP16, FMT_FP32, PREC_BF16, PREC_FP32, PREC_HALF, PREC_TF32, call, query

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
_NULL = ctypes.c_void_p(0)


def _p(t: Optional[torch.Tensor]):
    return _NULL if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(*tensors):
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.AvcError("autovc_b200 ops need CUDA tensors (there is no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            # the launchers use the CURRENT device and its current stream: a tensor elsewhere would be touched from the wrong GPU
            raise _lib.AvcError(f"autovc_b200 ops launch on the current CUDA device ({cur}) but got a tensor on {t.device}; "
                                f"use torch.cuda.set_device / `with torch.cuda.device(...)` for modules placed on another GPU")
        if t.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise _lib.AvcError(f"autovc_b200 ops take float32 tensors (or their 16-bit operand copies), got {t.dtype}")
        if not t.is_contiguous():
            raise _lib.AvcError("autovc_b200 ops need contiguous tensors")


def _ws(nbytes: int, device) -> Optional[torch.Tensor]:
    if nbytes <= 0:
        return None
    return torch.empty((nbytes + 15) // 16 * 2, dtype=torch.float64, device=device)


# --------------------------------------------------------------------------------------
# weight-gradient side stream
# --------------------------------------------------------------------------------------
# A weight gradient is a leaf of the backward graph: nothing in backward reads it, only the optimizer (and the
# data-parallel all-reduce) after the pass.  The blocks that own single-use parameters (decoder, postnet) therefore
# launch their weight-gradient GEMMs / bias column sums on a second stream: they then run concurrently with the next
# layer's HBM-bound BatchNorm passes and on the SMs a persistent recurrence leaves idle, instead of in front of them.
# The main stream re-joins the side stream in an end-of-backward engine callback (and in GradBucketReducer's hooks).
# Parameters used more than once per graph (the encoder: two passes per step) stay on the main stream, because autograd
# sums their contributions there.
_WGRAD = {"on": os.environ.get("AUTOVC_B200_WGRAD_STREAM", "1") != "0", "streams": {}, "armed": set(), "fwd_since_bwd": 0,
          "keep": []}


def wgrad_stream(device) -> Optional[torch.cuda.Stream]:
    """The weight-gradient side stream of `device` if one was ever used, else None (GradBucketReducer joins it)."""
    return _WGRAD["streams"].get(torch.device(device).index)


def note_full_forward():
    """Called by Generator.forward: two graph-building forwards before one backward mean a parameter may receive two
    contributions, which autograd sums on the main stream -> the side stream is not used for that backward."""
    if torch.is_grad_enabled():
        _WGRAD["fwd_since_bwd"] += 1
    _WGRAD["armed"].clear()


def keep_until_join(tensors):
    """Main-stream tensors that side-stream work reads stay referenced until the main stream has re-joined the side stream
    (end of backward): freed after that point, their blocks can only be reused by work ordered after the readers.
    (`Tensor.record_stream` would do the same, but it parks every such block behind an event and made the caching
    allocator call cudaMalloc inside steady-state steps -- 7 to 14 calls per 10 steps, with sporadic 30 ms stalls.)"""
    _WGRAD["keep"].extend(t for t in tensors if t is not None)


def _side_join(index: int):
    st = _WGRAD["streams"].get(index)
    if st is not None:
        torch.cuda.current_stream(st.device).wait_stream(st)
    _WGRAD["keep"].clear()
    _WGRAD["armed"].discard(index)
    _WGRAD["fwd_since_bwd"] = 0


class _MainStream:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def _wgrad_side(allowed: bool, device, reads: Sequence[Optional[torch.Tensor]], params: Sequence[torch.Tensor] = ()):
    """Context manager for the weight-gradient launches of one backward node: the side stream (ordered after everything
    queued on the main stream so far) when allowed, else a no-op.  `reads`: tensors the side launches read, which the
    main-stream allocator must not recycle before they ran (kept referenced until the join, see keep_until_join).  `params`: the parameters whose gradients are produced; if
    one already holds a gradient (accumulation over several backward passes) autograd adds to it on the main stream,
    so the node stays there."""
    if not (allowed and _WGRAD["on"]):
        return _MainStream()
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _WGRAD["armed"]:        # once per backward pass: join the side stream and reset the forward counter at its end
        _WGRAD["armed"].add(idx)
        torch.autograd.Variable._execution_engine.queue_callback(lambda: _side_join(idx))
    if _WGRAD["fwd_since_bwd"] > 1 or any(p.grad is not None for p in params):
        return _MainStream()
    st = _WGRAD["streams"].get(idx)
    if st is None:
        st = _WGRAD["streams"][idx] = torch.cuda.Stream(device=device)
    st.wait_stream(torch.cuda.current_stream(device))
    keep_until_join(reads)
    return torch.cuda.stream(st)


# --------------------------------------------------------------------------------------
# thin kernel launchers
# --------------------------------------------------------------------------------------
def gemm_nt_taps(A, lda, W, bias, C, ldc, nB, T, N, K, ntaps, shift0, stats=None, accumulate=False, prec=PREC_FP32):
    nbytes = query("avc_gemm_nt_workspace_bytes", nB, T, N, K, ntaps, prec) if prec != PREC_FP32 else 0
    ws = _ws(nbytes, C.device)
    call("avc_gemm_nt_taps", _p(A), lda, _p(W), _p(bias), _p(C), ldc, nB, T, N, K, ntaps, shift0, _p(stats),
         int(accumulate), prec, _p(ws), nbytes, _stream())


def gemm_tn_taps(dY, ldy, X, ldx, dW, nB, T, N, K, ntaps, shift0, out_mode, accumulate=False, prec=PREC_FP32):
    nbytes = query("avc_gemm_tn_workspace_bytes", nB, T, N, K, ntaps, prec)
    ws = _ws(nbytes, dY.device)
    call("avc_gemm_tn_taps", _p(dY), ldy, _p(X), ldx, _p(dW), nB, T, N, K, ntaps, shift0, out_mode, int(accumulate),
         prec, _p(ws), nbytes, _stream())


def colsum(x, ldx, M, C, out, out2=None, out_mode=0, accumulate=False):
    ws = torch.empty(2 * C, dtype=torch.float64, device=x.device)
    call("avc_colsum", _p(x), ldx, M, C, _p(out), _p(out2), out_mode, int(accumulate), _p(ws), ws.numel() * 8, _stream())


def colsum16(x16, fmt, ldx, M, C, out, out2=None, out_mode=0, accumulate=False):
    ws = torch.empty(2 * C, dtype=torch.float64, device=x16.device)
    call("avc_colsum16", _p(x16), fmt, ldx, M, C, _p(out), _p(out2), out_mode, int(accumulate), _p(ws), ws.numel() * 8, _stream())


class PackCache:
    """Packed copies of the weights so the two encoder passes of one step pack once.

    An entry is valid while the parameter object is alive, its in-place version counter (bumped by
    the optimizer) and its storage address are unchanged.  ``begin_step()`` drops everything; the
    Generator calls it at the start of each full forward, which also covers writers that bypass
    the version counter (``param.data.copy_`` in solver_encoder.py:168-177)."""

    def __init__(self):
        self._d = {}

    def get(self, kind, param: torch.Tensor, builder):
        k = (kind, id(param))
        hit = self._d.get(k)
        if hit is not None and hit[0]() is param and hit[1] == param._version and hit[2] == param.data_ptr():
            return hit[3]
        val = builder()
        self._d[k] = (weakref.ref(param), param._version, param.data_ptr(), val)
        return val

    def begin_step(self):
        self._d.clear()


_GLOBAL_CACHE = PackCache()


class _ZeroArena:
    """float64 zeros handed out in slices: the BatchNorm statistics buffers (2 C doubles per layer, forward and backward) are
    accumulated into by their kernels and must start at zero; one 512 KB fill per ~two training steps replaces 28 tiny fill
    launches per step.  A slice is handed out once and never re-used; an exhausted arena is replaced (the old one lives on
    until its slices are freed).  Everything happens on the calling stream."""
    SIZE = 65536

    def __init__(self):
        self._buf = {}

    def take(self, n: int, device) -> torch.Tensor:
        n2 = (n + 1) // 2 * 2                       # 16-byte granularity
        if n2 > self.SIZE // 4:
            return torch.zeros(n, device=device, dtype=torch.float64)
        key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
        ent = self._buf.get(key)
        if ent is None or ent[1] + n2 > self.SIZE:
            ent = [torch.zeros(self.SIZE, device=device, dtype=torch.float64), 0]
            self._buf[key] = ent
        out = ent[0][ent[1]:ent[1] + n]
        ent[1] += n2
        return out


_ZEROS = _ZeroArena()


def pack_conv(weight: torch.Tensor, cache: PackCache = _GLOBAL_CACHE):
    """(Cout, Cin, 5) -> fwd [5][Cout][Cin], dgrad [5][Cin][Cout] (taps flipped)."""
    def build():
        Cout, Cin, k = weight.shape
        w = weight.detach()
        wf = torch.empty(k, Cout, Cin, device=w.device, dtype=torch.float32)
        wd = torch.empty(k, Cin, Cout, device=w.device, dtype=torch.float32)
        call("avc_pack_conv_weight", _p(w), _p(wf), _p(wd), Cout, Cin, k, _stream())
        return wf, wd
    return cache.get("conv", weight, build)


def pack_lstm_w(weight: torch.Tensor, cache: PackCache = _GLOBAL_CACHE):
    """(4H, I) -> interleaved (4H, I) and its transpose (I, 4H)."""
    def build():
        G, I = weight.shape
        w = weight.detach()
        p = torch.empty(G, I, device=w.device, dtype=torch.float32)
        pT = torch.empty(I, G, device=w.device, dtype=torch.float32)
        call("avc_pack_lstm_weight", _p(w), _p(p), _p(pT), G // 4, I, _stream())
        return p, pT
    return cache.get("lstm_w", weight, build)


def _ld8(n: int) -> int:
    return (n + 7) // 8 * 8


_T16 = {FMT_BF16: torch.bfloat16, FMT_FP16: torch.float16}


def pack_conv_h(weight: torch.Tensor, cache: PackCache = _GLOBAL_CACHE):
    """(Cout, Cin, k) -> fp16 fwd [k][Cout][ld8(Cin)] and bf16 dgrad [k][Cin][ld8(Cout)] (taps flipped), the layouts the
    half-mode GEMMs read in place (forward operands fp16, gradient operands bf16)."""
    def build():
        Cout, Cin, k = weight.shape
        w = weight.detach()
        wf = torch.empty(k, Cout, _ld8(Cin), device=w.device, dtype=torch.float16)
        wd = torch.empty(k, Cin, _ld8(Cout), device=w.device, dtype=torch.bfloat16)
        call("avc_pack_conv_weight_h", _p(w), _p(wf), _ld8(Cin), FMT_FP16, _p(wd), _ld8(Cout), FMT_BF16, Cout, Cin, k, _stream())
        return wf, wd
    return cache.get("conv_h", weight, build)


def pack_lstm_w_h(weight: torch.Tensor, fmt_p: int, cache: PackCache = _GLOBAL_CACHE):
    """(4H, I) -> gate-interleaved (4H, ld8(I)) in `fmt_p` and its bf16 transpose (I, 4H)."""
    def build():
        G, I = weight.shape
        w = weight.detach()
        p = torch.empty(G, _ld8(I), device=w.device, dtype=_T16[fmt_p])
        pT = torch.empty(I, G, device=w.device, dtype=torch.bfloat16)
        call("avc_pack_lstm_weight_h", _p(w), _p(p), _ld8(I), fmt_p, _p(pT), G, FMT_BF16, G // 4, I, _stream())
        return p, pT
    return cache.get(("lstm_w_h", fmt_p), weight, build)


def gemm_nt_taps_hw(A, a_fmt, lda, W16, w_fmt, ldw, bias, C, ldc, nB, T, N, K, ntaps, shift0, stats=None, accumulate=False):
    nbytes = query("avc_gemm_nt_h_workspace_bytes", nB, T, N, K, ntaps, a_fmt) if a_fmt == FMT_FP32 else 0
    ws = _ws(nbytes, C.device) if nbytes else None
    call("avc_gemm_nt_taps_hw", _p(A), a_fmt, lda, _p(W16), w_fmt, ldw, _p(bias), _p(C), ldc, nB, T, N, K, ntaps, shift0,
         _p(stats), int(accumulate), _p(ws), nbytes, _stream())


def pack_lstm_b(b_ih: torch.Tensor, b_hh: torch.Tensor, cache: PackCache = _GLOBAL_CACHE):
    def build():
        out = torch.empty_like(b_ih)
        call("avc_pack_lstm_bias", _p(b_ih.detach()), _p(b_hh.detach()), _p(out), b_ih.numel() // 4, _stream())
        return out
    # both versions matter: key on b_ih, fold b_hh's version into the key
    return cache.get(("lstm_b", id(b_hh), b_hh.data_ptr(), b_hh._version), b_ih, build)


def pack_bilstm(weights, cache: PackCache = _GLOBAL_CACHE):
    """Both directions of a small-H BiLSTM layer stacked for single GEMMs: W_ih (2G, I) [dir 0 rows | dir 1 rows, gate-
    interleaved], its transpose (I, 2G), the folded biases (2G) and W_hh (2, G, H).  Cached per step on the first weight."""
    def build():
        wi, wiT, bs, wh = [], [], [], []
        for d in range(2):
            w_ih, w_hh, b_ih, b_hh = weights[4 * d:4 * d + 4]
            p, pT = pack_lstm_w(w_ih, cache)
            wi.append(p)
            wiT.append(pT)
            wh.append(pack_lstm_w(w_hh, cache)[0])
            bs.append(pack_lstm_b(b_ih, b_hh, cache))
        return torch.cat(wi, 0), torch.cat(wiT, 1).contiguous(), torch.cat(bs, 0), torch.stack(wh, 0)
    key = ("bilstm",) + tuple((id(w), w._version, w.data_ptr()) for w in weights[1:])
    return cache.get(key, weights[0], build)


def transpose2d(w: torch.Tensor, cache: PackCache = _GLOBAL_CACHE):
    def build():
        R, C = w.shape
        out = torch.empty(C, R, device=w.device, dtype=torch.float32)
        call("avc_transpose", _p(w.detach()), _p(out), R, C, _stream())
        return out
    return cache.get("T", w, build)


# --------------------------------------------------------------------------------------
# Conv1d(k=5,p=2) + BatchNorm1d + activation (+ residual)      model_vc_mel.py:28-31,:57,:69,:115,:165-167,:197
# --------------------------------------------------------------------------------------
class ConvBnAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, residual, act: int, training: bool,
                prec: int):
        x = x.contiguous()
        _check(x, weight, bias, gamma, beta, running_mean, running_var, residual)
        if residual is not None and act != ACT_NONE:
            raise _lib.AvcError("residual is only supported with act='none'")
        B, T, Cin = x.shape
        Cout, Cin_w, k = weight.shape
        if Cin_w != Cin:
            raise _lib.AvcError(f"conv: input has {Cin} channels, weight expects {Cin_w}")
        M = B * T
        wf, wd = pack_conv(weight)
        y = torch.empty(B, T, Cout, device=x.device, dtype=torch.float32)
        mean = torch.empty(Cout, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        if training:
            stats = _ZEROS.take(2 * Cout, x.device)
            gemm_nt_taps(x, Cin, wf, bias, y, Cout, B, T, Cout, Cin, k, -(k // 2), stats=stats, prec=prec)
            call("avc_bn_finalize", _p(stats), M, Cout, BN_EPS, BN_MOMENTUM, _p(mean), _p(rstd), _p(running_mean),
                 _p(running_var), _stream())
        else:
            gemm_nt_taps(x, Cin, wf, bias, y, Cout, B, T, Cout, Cin, k, -(k // 2), prec=prec)
            call("avc_bn_eval_stats", _p(running_mean), _p(running_var), Cout, BN_EPS, _p(mean), _p(rstd), _stream())
        z = torch.empty_like(y)
        call("avc_bn_act_fwd", _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(residual), _p(z), M, Cout, act, _stream())
        ctx.save_for_backward(x, weight, gamma, beta, y, z, mean, rstd)
        ctx.act, ctx.training, ctx.prec, ctx.has_res = act, training, prec, residual is not None
        ctx.wd = wd
        return z

    @staticmethod
    def backward(ctx, dz):
        x, weight, gamma, beta, y, z, mean, rstd = ctx.saved_tensors
        if not ctx.training:
            raise _lib.AvcError("backward through eval-mode BatchNorm is not on the supported path")
        dz = dz.contiguous()
        B, T, Cin = x.shape
        Cout, _, k = weight.shape
        M = B * T
        prec = ctx.prec
        sums = _ZEROS.take(2 * Cout, x.device)
        # act'(z) is recomputed from y where the float4 kernels apply (z is then not read)
        call("avc_bn_act_bwd_reduce_y", _p(dz), _p(z), _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(sums), M, Cout,
             ctx.act, _stream())
        dy = torch.empty_like(y)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        call("avc_bn_act_bwd_apply_y", _p(dz), _p(z), _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(sums), _p(dy), _NULL, 0,
             _p(dgamma), _p(dbeta), M, Cout, ctx.act, 0, _stream())
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            gemm_nt_taps(dy, Cout, ctx.wd, None, dx, Cin, B, T, Cin, Cout, k, -(k // 2), prec=prec)
        dw = torch.empty_like(weight)
        gemm_tn_taps(dy, Cout, x, Cin, dw, B, T, Cout, Cin, k, -(k // 2), out_mode=1, prec=prec)
        # d(conv bias) is identically zero under train-mode BatchNorm (sum_m dy[m, c] = 0); the reference
        # produces 1e-8-level rounding noise there.
        db = torch.zeros(Cout, device=x.device, dtype=torch.float32)
        dres = dz if ctx.has_res else None
        return dx, dw, db, dgamma, dbeta, None, None, dres, None, None, None


# --------------------------------------------------------------------------------------
# one nn.LSTM layer (1 or 2 directions)                        model_vc_mel.py:61/:73, :90/:111, :104/:118
# --------------------------------------------------------------------------------------
class LstmLayer(torch.autograd.Function):
    """x (B,T,I) -> (B,T,D*H).  weights = [w_ih, w_hh, b_ih, b_hh] per direction (forward first)."""

    @staticmethod
    def forward(ctx, x, prec: int, *weights):
        x = x.contiguous()
        _check(x, *weights)
        D = len(weights) // 4
        B, T, I = x.shape
        H = weights[1].shape[1]
        G = 4 * H
        out = torch.empty(B, T, D * H, device=x.device, dtype=torch.float32)
        saved = []
        packs = []
        fused = D == 2 and H <= 64 and H % 8 == 0          # both directions of the encoder BiLSTM in one launch
        if fused:
            # ONE input-projection GEMM for both directions (N = 2G: x is read once), one recurrence launch (reverse=3:
            # P rows are [dir 0 gates | dir 1 gates])
            wi2, wiT2, b2, wh2 = pack_bilstm(weights)
            Pre = torch.empty(B, T, 2 * G, device=x.device, dtype=torch.float32)
            gemm_nt_taps(x, I, wi2, b2, Pre, 2 * G, B, T, 2 * G, I, 1, 0, prec=prec)
            gates = torch.empty(2, B, T, G, device=x.device, dtype=torch.float32)
            c_seq = torch.empty(2, B, T, H, device=x.device, dtype=torch.float32)
            call("avc_lstm_seq_fwd", _p(Pre), _p(wh2), _p(out), 2 * H, _p(gates), _p(c_seq), B, T, H, 3, prec, _NULL, 0, _stream())
            saved = [gates, c_seq, wh2]
            packs = [wiT2]
        for d in range(0 if not fused else D, D):
            w_ih, w_hh, b_ih, b_hh = weights[4 * d:4 * d + 4]
            wi_p, wi_pT = pack_lstm_w(w_ih)
            wh_p, wh_pT = pack_lstm_w(w_hh)
            b_p = pack_lstm_b(b_ih, b_hh)
            packs += [wi_pT, wh_p, wh_pT]
            Pre = torch.empty(B, T, G, device=x.device, dtype=torch.float32)
            gemm_nt_taps(x, I, wi_p, b_p, Pre, G, B, T, G, I, 1, 0, prec=prec)
            gates = torch.empty(B, T, G, device=x.device, dtype=torch.float32)
            c_seq = torch.empty(B, T, H, device=x.device, dtype=torch.float32)
            h_view = out.view(-1)[d * H:]
            nbytes = query("avc_lstm_fwd_workspace_bytes", B, T, H, prec)
            ws = _ws(nbytes, x.device)
            call("avc_lstm_seq_fwd", _p(Pre), _p(wh_p), _p(h_view), D * H, _p(gates), _p(c_seq), B, T, H, int(d == 1), prec,
                 _p(ws), nbytes, _stream())
            saved += [gates, c_seq]
        ctx.save_for_backward(x, out, *weights, *saved)
        ctx.D, ctx.prec, ctx.packs, ctx.fused = D, prec, packs, fused
        return out

    @staticmethod
    def backward(ctx, dout):
        D, prec = ctx.D, ctx.prec
        t = ctx.saved_tensors
        x, out = t[0], t[1]
        weights = t[2:2 + 4 * D]
        saved = t[2 + 4 * D:]
        dout = dout.contiguous()
        B, T, I = x.shape
        H = weights[1].shape[1]
        G = 4 * H
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty_like(x) if need_dx else None
        grads: List[Optional[torch.Tensor]] = []
        if ctx.fused:
            gates2, c2, wh2 = saved
            (wiT2,) = ctx.packs
            dP2 = torch.empty(B, T, 2 * G, device=x.device, dtype=torch.float32)          # [dir 0 | dir 1] per row
            call("avc_lstm_seq_bwd", _p(dout), 2 * H, _p(wh2), _p(wh2), _p(gates2), _p(c2), _p(dP2), B, T, H, 3, prec,
                 _NULL, 0, _stream())
            # dW_ih of both directions in one GEMM: the 2G gate columns are un-permuted as ONE virtual layer of 2H units
            # (row g*2H + d*H + u), then split per direction
            dwi = torch.empty(4, 2, H, I, device=x.device, dtype=torch.float32)
            gemm_tn_taps(dP2, 2 * G, x, I, dwi, B, T, 2 * G, I, 1, 0, out_mode=2, prec=prec)
            dwi = dwi.permute(1, 0, 2, 3).contiguous()
            if need_dx:
                gemm_nt_taps(dP2, 2 * G, wiT2, None, dx, I, B, T, I, 2 * G, 1, 0, prec=prec)
            for d in range(2):
                w_ih, w_hh, b_ih, b_hh = weights[4 * d:4 * d + 4]
                dPd = dP2.view(-1)[d * G:]
                dw_hh = torch.empty_like(w_hh)
                gemm_tn_taps(dPd, 2 * G, out.view(-1)[d * H:], 2 * H, dw_hh, B, T, G, H, 1, (+1 if d == 1 else -1), out_mode=2,
                             prec=prec)
                db_ih = torch.empty_like(b_ih)
                db_hh = torch.empty_like(b_hh)
                colsum(dPd, 2 * G, B * T, G, db_ih, db_hh, out_mode=2)
                grads += [dwi[d].view(G, I), dw_hh, db_ih, db_hh]
            return (dx, None, *grads)
        for d in range(D):
            w_ih, w_hh, b_ih, b_hh = weights[4 * d:4 * d + 4]
            wi_pT, wh_p, wh_pT = ctx.packs[3 * d:3 * d + 3]
            rev = int(d == 1)
            gates, c_seq = saved[2 * d:2 * d + 2]
            dP = torch.empty(B, T, G, device=x.device, dtype=torch.float32)
            nbytes = query("avc_lstm_bwd_workspace_bytes", B, T, H, prec)
            ws = _ws(nbytes, x.device)
            call("avc_lstm_seq_bwd", _p(dout.view(-1)[d * H:]), D * H, _p(wh_p), _p(wh_pT), _p(gates), _p(c_seq), _p(dP),
                 B, T, H, rev, prec, _p(ws), nbytes, _stream())
            dw_ih = torch.empty_like(w_ih)
            gemm_tn_taps(dP, G, x, I, dw_ih, B, T, G, I, 1, 0, out_mode=2, prec=prec)
            dw_hh = torch.empty_like(w_hh)
            # dW_hh = sum_t dG_t^T h_{t-1}: the h operand is the output sequence shifted one step back
            # along the walk direction (zero at the first step)
            gemm_tn_taps(dP, G, out.view(-1)[d * H:], D * H, dw_hh, B, T, G, H, 1, (+1 if rev else -1), out_mode=2, prec=prec)
            db_ih = torch.empty_like(b_ih)
            db_hh = torch.empty_like(b_hh)
            colsum(dP, G, B * T, G, db_ih, db_hh, out_mode=2)
            if need_dx:
                gemm_nt_taps(dP, G, wi_pT, None, dx, I, B, T, I, G, 1, 0, accumulate=(d > 0), prec=prec)
            grads += [dw_ih, dw_hh, db_ih, db_hh]
        return (dx, None, *grads)


# --------------------------------------------------------------------------------------
# Linear(1024 -> n_bins)                                        model_vc_mel.py:10,:120
# --------------------------------------------------------------------------------------
class Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, prec: int):
        x = x.contiguous()
        _check(x, weight, bias)
        B, T, K = x.shape
        N = weight.shape[0]
        y = torch.empty(B, T, N, device=x.device, dtype=torch.float32)
        gemm_nt_taps(x, K, weight.detach(), bias, y, N, B, T, N, K, 1, 0, prec=prec)
        ctx.save_for_backward(x, weight)
        ctx.prec = prec
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        B, T, K = x.shape
        N = weight.shape[0]
        dx = None
        if ctx.needs_input_grad[0]:
            wT = transpose2d(weight, PackCache())         # (K, N): [N'=K][K'=N]; tiny, not cached
            dx = torch.empty_like(x)
            gemm_nt_taps(dy, N, wT, None, dx, K, B, T, K, N, 1, 0, prec=ctx.prec)
        dw = torch.empty_like(weight)
        gemm_tn_taps(dy, N, x, K, dw, B, T, N, K, 1, 0, out_mode=0, prec=ctx.prec)
        db = torch.empty(N, device=x.device, dtype=torch.float32)
        colsum(dy, N, B * T, N, db)
        return dx, dw, db, None


# --------------------------------------------------------------------------------------
# glue                                                          model_vc_mel.py:64-66, :74-79, :186-192
# --------------------------------------------------------------------------------------
class ConcatEmb(torch.autograd.Function):
    """[x (B,T,Cx) | e (B,E) broadcast over T] -> (B,T,Cx+E)."""

    @staticmethod
    def forward(ctx, x, e):
        x = x.contiguous()
        e = e.contiguous()
        _check(x, e)
        B, T, Cx = x.shape
        E = e.shape[1]
        out = torch.empty(B, T, Cx + E, device=x.device, dtype=torch.float32)
        call("avc_concat_bcast", _p(x), Cx, _p(e), _p(out), B, T, Cx, E, _stream())
        ctx.shape = (B, T, Cx, E)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, T, Cx, E = ctx.shape
        dx = None
        if ctx.needs_input_grad[0]:
            dout = dout.contiguous()
            dx = torch.empty(B, T, Cx, device=dout.device, dtype=torch.float32)
            call("avc_copy2d", _p(dout), Cx + E, _p(dx), Cx, B * T, Cx, _stream())
        if ctx.needs_input_grad[1]:
            raise _lib.AvcError("gradient w.r.t. the speaker embedding is not on the supported path")
        return dx, None


class Codes(torch.autograd.Function):
    """enc (B,T,2n) -> codes (B,T/f,2n): forward half sampled at the end of each group, backward half
    at its start (model_vc_mel.py:77-79)."""

    @staticmethod
    def forward(ctx, enc, n: int, f: int):
        enc = enc.contiguous()
        _check(enc)
        B, T, n2 = enc.shape
        if T % f != 0:
            raise _lib.AvcError(f"T={T} must be a multiple of freq={f}")
        codes = torch.empty(B, T // f, n2, device=enc.device, dtype=torch.float32)
        call("avc_codes_fwd", _p(enc), _p(codes), B, T, n, f, _stream())
        ctx.meta = (B, T, n, f)
        return codes

    @staticmethod
    def backward(ctx, dcodes):
        B, T, n, f = ctx.meta
        dcodes = dcodes.contiguous()
        denc = torch.zeros(B, T, 2 * n, device=dcodes.device, dtype=torch.float32)
        call("avc_codes_bwd", _p(dcodes), _p(denc), B, T, n, f, _stream())
        return denc, None, None


class UpsampleConcat(torch.autograd.Function):
    """codes (B,J,2n), c_trg (B,E) -> (B,T,2n+E) (model_vc_mel.py:186-192)."""

    @staticmethod
    def forward(ctx, codes, c_trg, T: int):
        codes = codes.contiguous()
        c_trg = c_trg.contiguous()
        _check(codes, c_trg)
        B, J, n2 = codes.shape
        E = c_trg.shape[1]
        f = T // J
        out = torch.empty(B, T, n2 + E, device=codes.device, dtype=torch.float32)
        call("avc_upsample_concat_fwd", _p(codes), _p(c_trg), _p(out), B, T, n2, f, E, _stream())
        ctx.meta = (B, T, J, n2, E, f)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, T, J, n2, E, f = ctx.meta
        dout = dout.contiguous()
        dcodes = torch.empty(B, J, n2, device=dout.device, dtype=torch.float32)
        call("avc_upsample_concat_bwd", _p(dout), n2 + E, _p(dcodes), B, T, n2, f, 0, _stream())
        return dcodes, None, None


# --------------------------------------------------------------------------------------
# losses                                                        solver_encoder.py:230,:233,:236
# --------------------------------------------------------------------------------------
class _Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, is_l1: bool):
        a = a.contiguous()
        b = b.contiguous()
        _check(a, b)
        if a.numel() != b.numel():
            raise _lib.AvcError("loss operands differ in size")
        scratch = torch.zeros(1, device=a.device, dtype=torch.float64)
        out = torch.empty((), device=a.device, dtype=torch.float32)
        call("avc_l1_loss_fwd" if is_l1 else "avc_mse_loss_fwd", _p(a), _p(b), a.numel(), _p(scratch), _p(out), _stream())
        ctx.save_for_backward(a, b)
        ctx.is_l1 = is_l1
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        gout = gout.contiguous().float()
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        if da is not None or db is not None:
            call("avc_loss_bwd", _p(a), _p(b), a.numel(), _p(gout), int(ctx.is_l1), _p(da), _p(db), 0, _stream())
        return da, db, None


def mse_loss(a, b):
    return _Loss.apply(a, b, False)


def l1_loss(a, b):
    return _Loss.apply(a, b, True)


# ======================================================================================
# "half" mode (AVC_PREC_HALF): 16-bit operand copies travel next to the fp32 tensors
# ======================================================================================
_DT16 = {FMT_FP16: torch.float16, FMT_BF16: torch.bfloat16}


def cast16(x: torch.Tensor, fmt: int) -> torch.Tensor:
    """fp32 (..., C) contiguous -> 16-bit copy (fmt: FMT_FP16 forward operands, FMT_BF16 gradient operands)."""
    C = x.shape[-1]
    M = x.numel() // C
    out = torch.empty(x.shape, dtype=_DT16[fmt], device=x.device)
    call("avc_cast16", _p(x), C, _p(out), C, M, C, fmt, _stream())
    return out


def _operand(t32: torch.Tensor, t16: Optional[torch.Tensor], C: int, fmt: int):
    """Pick what the tensor cores read: the 16-bit copy when it exists (or can be made) and is TMA-addressable
    (C % 8 == 0), else the fp32 tensor, which the GEMM stages internally."""
    if C % 8 != 0:
        return t32, FMT_FP32
    if t16 is None:
        t16 = cast16(t32, fmt)
    return t16, (FMT_FP16 if t16.dtype == torch.float16 else FMT_BF16)


def gemm_nt_taps_h(A, a_fmt, lda, W, bias, C, ldc, nB, T, N, K, ntaps, shift0, half_fmt, stats=None, accumulate=False):
    nbytes = query("avc_gemm_nt_h_workspace_bytes", nB, T, N, K, ntaps, a_fmt)
    ws = _ws(nbytes, C.device)
    call("avc_gemm_nt_taps_h", _p(A), a_fmt, lda, _p(W), _p(bias), _p(C), ldc, nB, T, N, K, ntaps, shift0, _p(stats),
         int(accumulate), half_fmt, _p(ws), nbytes, _stream())


def gemm_tn_taps_h(dY, y_fmt, ldy, X, x_fmt, ldx, dW, nB, T, N, K, ntaps, shift0, out_mode, half_fmt=FMT_BF16, accumulate=False):
    nbytes = query("avc_gemm_tn_h_workspace_bytes", nB, T, N, K, ntaps, y_fmt, x_fmt)
    ws = _ws(nbytes, dW.device)
    call("avc_gemm_tn_taps_h", _p(dY), y_fmt, ldy, _p(X), x_fmt, ldx, _p(dW), nB, T, N, K, ntaps, shift0, out_mode,
         int(accumulate), half_fmt, _p(ws), nbytes, _stream())


class ConvBnActH(torch.autograd.Function):
    """ConvBnAct for the half mode: (x, x16, x16b) -> (z, z16, z16b).  Forward operands fp16 (x16 / z16), gradient operands
    bf16 (x16b / z16b: the activation copy the NEXT layer's weight-gradient GEMM reads, emitted here by the producing kernel
    instead of being cast in backward).  With ``need_z32=False`` the fp32 activation is not written at all: ``z`` is then only
    the autograd carrier (allocated, never filled) and every consumer must read the 16-bit copies."""

    @staticmethod
    def forward(ctx, x, x16, x16b, weight, bias, gamma, beta, running_mean, running_var, residual, act: int, training: bool,
                need_z32: bool = True, side_wgrad: bool = False):
        x = x.contiguous()
        _check(x, x16, x16b, weight, bias, gamma, beta, running_mean, running_var, residual)
        if residual is not None and act != ACT_NONE:
            raise _lib.AvcError("residual is only supported with act='none'")
        B, T, Cin = x.shape
        Cout, Cin_w, k = weight.shape
        if Cin_w != Cin:
            raise _lib.AvcError(f"conv: input has {Cin} channels, weight expects {Cin_w}")
        M = B * T
        wf, wd = pack_conv_h(weight)
        A, a_fmt = _operand(x, x16, Cin, FMT_FP16)
        y = torch.empty(B, T, Cout, device=x.device, dtype=torch.float32)
        mean = torch.empty(Cout, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        stats = _ZEROS.take(2 * Cout, x.device) if training else None
        gemm_nt_taps_hw(A, a_fmt, Cin, wf, FMT_FP16, wf.shape[-1], bias, y, Cout, B, T, Cout, Cin, k, -(k // 2), stats=stats)
        if training:
            call("avc_bn_finalize", _p(stats), M, Cout, BN_EPS, BN_MOMENTUM, _p(mean), _p(rstd), _p(running_mean),
                 _p(running_var), _stream())
        else:
            call("avc_bn_eval_stats", _p(running_mean), _p(running_var), Cout, BN_EPS, _p(mean), _p(rstd), _stream())
        z = torch.empty_like(y)
        if Cout % 8 == 0:
            z16 = torch.empty(B, T, Cout, device=x.device, dtype=torch.float16)
            # the bf16 copy only serves the next layer's weight-gradient GEMM: skipped when nothing here needs a gradient
            z16b = torch.empty(B, T, Cout, device=x.device, dtype=torch.bfloat16) if (training and any(ctx.needs_input_grad)) else None
            call("avc_bn_act_fwd_h", _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(residual), _p(z) if need_z32 else _NULL,
                 _p(z16), FMT_FP16, _p(z16b), FMT_BF16, M, Cout, act, _stream())
            if z16b is None:
                z16b = torch.empty(0, device=x.device, dtype=torch.bfloat16)
        else:
            if not need_z32:
                raise _lib.AvcError("need_z32=False needs Cout % 8 == 0 (the 16-bit copies are the only output then)")
            z16 = torch.empty(0, device=x.device, dtype=torch.float16)
            z16b = torch.empty(0, device=x.device, dtype=torch.bfloat16)
            call("avc_bn_act_fwd", _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(residual), _p(z), M, Cout, act, _stream())
        # backward needs x only as the weight-gradient operand: keep the bf16 copy when the producer supplied one
        xb = x16b if (x16b is not None and x16b.numel() == x.numel()) else None
        # act'(z) is recomputed from y in backward where the float4 kernels apply (Cout % 4 == 0): z is then not saved
        ctx.save_for_backward(x if xb is None else None, xb, weight, gamma, beta, y, mean, rstd, z if Cout % 4 != 0 else None)
        ctx.x_shape = tuple(x.shape)
        ctx.act, ctx.training, ctx.has_res, ctx.wd = act, training, residual is not None, wd
        ctx.side_wgrad = side_wgrad
        ctx.mark_non_differentiable(z16, z16b)
        ctx.set_materialize_grads(False)      # no zero tensors for the 16-bit side outputs' (absent) gradients
        return z, z16, z16b

    @staticmethod
    def backward(ctx, dz, _dz16, _dz16b):
        x, xb, weight, gamma, beta, y, mean, rstd, z = ctx.saved_tensors
        if not ctx.training:
            raise _lib.AvcError("backward through eval-mode BatchNorm is not on the supported path")
        dz = dz.contiguous()
        B, T, Cin = ctx.x_shape
        Cout, _, k = weight.shape
        M = B * T
        sums = _ZEROS.take(2 * Cout, dz.device)
        call("avc_bn_act_bwd_reduce_y", _p(dz), _p(z), _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(sums), M, Cout,
             ctx.act, _stream())
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        if Cout % 8 == 0:      # the gradient w.r.t. the conv output is only ever a GEMM operand: emit it as bf16 only
            dy = torch.empty(B, T, Cout, device=dz.device, dtype=torch.bfloat16)
            call("avc_bn_act_bwd_apply_y", _p(dz), _p(z), _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(sums), _NULL, _p(dy),
                 FMT_BF16, _p(dgamma), _p(dbeta), M, Cout, ctx.act, 0, _stream())
            y_fmt = FMT_BF16
        else:
            dy = torch.empty_like(y)
            call("avc_bn_act_bwd_apply_y", _p(dz), _p(z), _p(y), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(sums), _p(dy), _NULL, 0,
                 _p(dgamma), _p(dbeta), M, Cout, ctx.act, 0, _stream())
            y_fmt = FMT_FP32
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, T, Cin, device=dz.device, dtype=torch.float32)
            gemm_nt_taps_hw(dy, y_fmt, Cout, ctx.wd, FMT_BF16, ctx.wd.shape[-1], None, dx, Cin, B, T, Cin, Cout, k, -(k // 2))
        with _wgrad_side(ctx.side_wgrad, dz.device, (dy, x, xb), (weight,)):
            dw = torch.empty_like(weight)
            # both operands of a GEMM must share a 16-bit format: bf16 copy of the activation from its producer, else cast here
            if y_fmt == FMT_BF16:
                X, x_fmt = (xb, FMT_BF16) if xb is not None else _operand(x, None, Cin, FMT_BF16)
            elif x is not None:
                X, x_fmt = x, FMT_FP32
            else:                       # fp32 dy (odd Cout, e.g. 513 bins) is staged to bf16 inside the GEMM: the bf16 copy of x matches it
                X, x_fmt = xb, FMT_BF16
            gemm_tn_taps_h(dy, y_fmt, Cout, X, x_fmt, Cin, dw, B, T, Cout, Cin, k, -(k // 2), out_mode=1)   # fp32 operands are staged to bf16
        db = torch.zeros(Cout, device=dz.device, dtype=torch.float32)
        dres = dz if ctx.has_res else None
        return dx, None, None, dw, db, dgamma, dbeta, None, None, dres, None, None, None, None


class LstmLayerH(torch.autograd.Function):
    """One unidirectional nn.LSTM layer in half mode: (x, x16, x16b) -> (h, h16, h16b).  Persistent recurrence kernels with
    16-bit side outputs (fp16: the next forward GEMM's operand; bf16: the weight-gradient GEMMs' operand); input projection on
    fp16 operands, gradient GEMMs on bf16 (dP) x bf16 (x, h)."""

    @staticmethod
    def forward(ctx, x, x16, x16b, w_ih, w_hh, b_ih, b_hh, side_wgrad: bool = False):
        x = x.contiguous()
        _check(x, x16, x16b, w_ih, w_hh, b_ih, b_hh)
        ctx.side_wgrad = side_wgrad
        B, T, I = x.shape
        H = w_hh.shape[1]
        G = 4 * H
        wi_p, wi_pT = pack_lstm_w_h(w_ih, FMT_FP16)      # projection operands fp16, dX operands bf16
        wh_p, wh_pT = pack_lstm_w_h(w_hh, FMT_BF16)      # the recurrences run on bf16
        b_p = pack_lstm_b(b_ih, b_hh)
        A, a_fmt = _operand(x, x16, I, FMT_FP16)
        Pre = torch.empty(B, T, G, device=x.device, dtype=torch.float32)
        gemm_nt_taps_hw(A, a_fmt, I, wi_p, FMT_FP16, wi_p.shape[-1], b_p, Pre, G, B, T, G, I, 1, 0)
        out = torch.empty(B, T, H, device=x.device, dtype=torch.float32)
        h16 = torch.empty(B, T, H, device=x.device, dtype=torch.float16)
        train = any(ctx.needs_input_grad)          # under no_grad (conversion) nothing is saved for BPTT
        h16b = torch.empty(B, T, H, device=x.device, dtype=torch.bfloat16) if train else None
        gates = torch.empty(B, T, G, device=x.device, dtype=torch.float32) if train else None
        c_seq = torch.empty(B, T, H, device=x.device, dtype=torch.float32) if train else None
        nbytes = query("avc_lstm_fwd_workspace_bytes", B, T, H, PREC_BF16)
        ws = _ws(nbytes, x.device)
        call("avc_lstm_seq_fwd_h", _p(Pre), _p(wh_p), FMT_BF16, _p(out), H, _p(gates), _p(c_seq), _p(h16), FMT_FP16, _p(h16b),
             B, T, H, 0, _p(ws), nbytes, _stream())
        xb = x16b if (x16b is not None and x16b.numel() == x.numel()) else None
        ctx.save_for_backward(x if xb is None else None, xb, h16b, w_ih, w_hh, b_ih, gates, c_seq)
        ctx.x_shape = tuple(x.shape)
        ctx.packs = (wi_pT, wh_pT)
        if h16b is None:
            h16b = torch.empty(0, device=x.device, dtype=torch.bfloat16)
        ctx.mark_non_differentiable(h16, h16b)
        ctx.set_materialize_grads(False)
        return out, h16, h16b

    @staticmethod
    def backward(ctx, dout, _d16, _d16b):
        x, xb, h16b, w_ih, w_hh, b_ih, gates, c_seq = ctx.saved_tensors
        wi_pT, wh_pT = ctx.packs
        dout = dout.contiguous()
        B, T, I = ctx.x_shape
        H = w_hh.shape[1]
        G = 4 * H
        # the gate gradient is only ever a GEMM operand / reduced over rows: emit it as bf16 only (no fp32 dP tensor)
        dP16 = torch.empty(B, T, G, device=dout.device, dtype=torch.bfloat16)
        nbytes = query("avc_lstm_bwd_workspace_bytes", B, T, H, PREC_BF16)
        ws = _ws(nbytes, dout.device)
        call("avc_lstm_seq_bwd_h", _p(dout), H, _p(wh_pT), FMT_BF16, _p(gates), _p(c_seq), _NULL, _p(dP16), B, T, H, 0, _p(ws),
             nbytes, _stream())
        # the data gradient first (the next BPTT / conv backward waits for it), the parameter gradients on the side stream
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, T, I, device=dout.device, dtype=torch.float32)
            gemm_nt_taps_hw(dP16, FMT_BF16, G, wi_pT, FMT_BF16, G, None, dx, I, B, T, I, G, 1, 0)
        with _wgrad_side(ctx.side_wgrad, dout.device, (dP16, x, xb, h16b), (w_ih, w_hh, b_ih)):
            dw_ih = torch.empty_like(w_ih)
            X, x_fmt = (xb, FMT_BF16) if xb is not None else _operand(x, None, I, FMT_BF16)
            gemm_tn_taps_h(dP16, FMT_BF16, G, X, x_fmt, I, dw_ih, B, T, G, I, 1, 0, out_mode=2)
            dw_hh = torch.empty_like(w_hh)
            gemm_tn_taps_h(dP16, FMT_BF16, G, h16b, FMT_BF16, H, dw_hh, B, T, G, H, 1, -1, out_mode=2)
            db_ih = torch.empty_like(b_ih)
            db_hh = torch.empty_like(b_ih)
            colsum16(dP16, FMT_BF16, G, B * T, G, db_ih, db_hh, out_mode=2)
        return dx, None, None, dw_ih, dw_hh, db_ih, db_hh, None


class LinearH(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, x16, x16b, weight, bias, side_wgrad: bool = False):
        x = x.contiguous()
        _check(x, x16, x16b, weight, bias)
        ctx.side_wgrad = side_wgrad
        B, T, K = x.shape
        N = weight.shape[0]
        A, a_fmt = _operand(x, x16, K, FMT_FP16)
        y = torch.empty(B, T, N, device=x.device, dtype=torch.float32)
        gemm_nt_taps_h(A, a_fmt, K, weight.detach(), bias, y, N, B, T, N, K, 1, 0, FMT_FP16)
        xb = x16b if (x16b is not None and x16b.numel() == x.numel()) else None
        ctx.save_for_backward(x if xb is None else None, xb, weight)
        ctx.x_shape = tuple(x.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, xb, weight = ctx.saved_tensors
        dy = dy.contiguous()
        B, T, K = ctx.x_shape
        N = weight.shape[0]
        D, d_fmt = _operand(dy, None, N, FMT_BF16)
        dx = None
        if ctx.needs_input_grad[0]:
            wT = transpose2d(weight, PackCache())
            dx = torch.empty(B, T, K, device=dy.device, dtype=torch.float32)
            gemm_nt_taps_h(D, d_fmt, N, wT, None, dx, K, B, T, K, N, 1, 0, FMT_BF16)
        with _wgrad_side(ctx.side_wgrad, dy.device, (dy, D, x, xb), (weight,)):
            dw = torch.empty_like(weight)
            if d_fmt == FMT_BF16:
                X, x_fmt = (xb, FMT_BF16) if xb is not None else _operand(x, None, K, FMT_BF16)
            else:
                X, x_fmt = (x, FMT_FP32) if x is not None else (xb, FMT_BF16)
            gemm_tn_taps_h(D, d_fmt, N, X, x_fmt, K, dw, B, T, N, K, 1, 0, out_mode=0)
            db = torch.empty(N, device=dy.device, dtype=torch.float32)
            colsum(dy, N, B * T, N, db)
        return dx, None, None, dw, db, None
