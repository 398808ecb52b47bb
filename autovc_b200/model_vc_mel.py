"""Drop-in ``Generator`` for the reference's ``model_vc_mel.Generator`` (model_vc_mel.py:172-203).

Same constructor ``Generator(dim_neck, dim_emb, dim_pre, freq)``, same ``forward(x, c_org, c_trg)``
contract and return values, same sub-module names (``encoder``, ``decoder``, ``postnet``), same
parameter/buffer names, shapes, registration order and initialisers — so ``torch.manual_seed(s)``
followed by construction gives the reference's init bit for bit, ``state_dict()`` /
``load_state_dict()`` / the optimizer-state indexing of a reference checkpoint carry over, and the
reference ``Solver`` loop (solver_encoder.py:182-421) runs unchanged on top of it.

What differs is who does the arithmetic: the torch.nn layers below are used only as *parameter
containers*; every forward/backward computation goes through ``autovc_b200.ops`` into the CUDA
kernels of libautovc_b200.so.  Activations stay channels-last (B, T, C) from end to end, so none of
the reference's transposes (model_vc_mel.py:64,:70,:112,:116,:196-197) exist here.  There is no
CPU path: parameters and inputs must live on a CUDA device.

Extra (non-reference) constructor keywords: ``n_bins`` (80; 513 builds the model_vc_stft layer
shapes) and ``precision`` ("fp32" | "fp32_simt" | "tf32" | "half", also settable later through ``set_precision``).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_CODES, PREC_FP32, PREC_FP32X3, PREC_HALF, PREC_TF32

# "half" is the 16-bit tensor-core mode (fp16 forward operands, bf16 gradient operands and recurrences, fp32 accumulation and
# state).  A mode with bf16 operands everywhere was measured at 2.2e-2 / 4.6e-2 relative L2 against the reference (rounding
# the weights alone to bf16 costs 1.3e-2 in the reference itself): it cannot meet the 1e-2 gate and is not offered.
# "fp32" is the parity mode (<= 1e-4 max-abs of the reference's fp32 path) on the tensor cores: 3xTF32 split products
# (csrc/fp32x3.cu).  "fp32_simt" is the same arithmetic contract on the CUDA cores (FFMA GEMMs, one launch per recurrence
# step), kept as the independent cross-check of the split-precision path.
_PREC = {"fp32": PREC_FP32X3, "fp32_simt": PREC_FP32, "tf32": PREC_TF32, "half": PREC_HALF}


def _default_precision() -> str:
    return os.environ.get("AUTOVC_B200_PRECISION", "fp32")


class LinearNorm(nn.Module):
    """Parameter container mirroring model_vc_mel.py:7-17 (``linear_layer`` + Xavier init)."""

    def __init__(self, in_dim, out_dim, bias=True, w_init_gain="linear"):
        super().__init__()
        self.linear_layer = nn.Linear(in_dim, out_dim, bias=bias)
        nn.init.xavier_uniform_(self.linear_layer.weight, gain=nn.init.calculate_gain(w_init_gain))
        self.prec = PREC_FP32

    def forward(self, x):
        if self.prec == PREC_HALF:
            return ops.LinearH.apply(x, None, None, self.linear_layer.weight, self.linear_layer.bias)
        return ops.Linear.apply(x, self.linear_layer.weight, self.linear_layer.bias, self.prec)


class ConvNorm(nn.Module):
    """Parameter container mirroring model_vc_mel.py:20-38 (``conv`` + Xavier init).  Only the
    kernel_size=5 / stride 1 / 'same' padding / dilation 1 case the Generator uses is supported."""

    def __init__(self, in_channels, out_channels, kernel_size=5, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear"):
        super().__init__()
        if padding is None:
            padding = dilation * (kernel_size - 1) // 2
        if kernel_size % 2 != 1 or stride != 1 or dilation != 1 or padding != (kernel_size - 1) // 2 or not bias:
            raise ValueError("autovc_b200.ConvNorm supports odd kernel, stride 1, dilation 1, 'same' padding, bias=True")
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)
        nn.init.xavier_uniform_(self.conv.weight, gain=nn.init.calculate_gain(w_init_gain))

    def forward(self, signal):
        raise RuntimeError("ConvNorm is fused with its BatchNorm and activation; call the owning block")


def _conv_bn_act(block: nn.Sequential, x, act: str, residual=None, prec=PREC_FP32, x16=None, x16b=None, need_z32=True,
                 side_wgrad=False):
    """act(BN(conv(x))) (+ residual) on channels-last x; bumps num_batches_tracked like nn.BatchNorm1d.
    Returns (z, z16, z16b): the fp16 / bf16 operand copies in half mode, else None.  ``need_z32=False`` (half mode only):
    the caller promises that nothing reads the fp32 activation, which is then not written (z is the autograd carrier)."""
    conv, bn = block[0].conv, block[1]
    training = bn.training
    z16 = z16b = None
    if prec == PREC_HALF:
        Cout = conv.weight.shape[0]
        z, z16, z16b = ops.ConvBnActH.apply(x, x16, x16b, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean,
                                            bn.running_var, residual, ACT_CODES[act], training,
                                            need_z32 or Cout % 8 != 0 or not training, side_wgrad)
        z16 = z16 if z16.numel() else None
        z16b = z16b if z16b.numel() else None
    else:
        z = ops.ConvBnAct.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual,
                                ACT_CODES[act], training, prec)
    if training:
        bn.num_batches_tracked += 1
    return z, z16, z16b


def _lstm(x, lstm: nn.LSTM, prec, x16=None, x16b=None, side_wgrad=False):
    """Run an nn.LSTM's parameters through the LstmLayer kernels, layer by layer.  Returns (h, h16, h16b)."""
    H = lstm.hidden_size
    half = prec == PREC_HALF and not lstm.bidirectional and 128 <= H <= 1024 and H % 64 == 0
    for l in range(lstm.num_layers):
        ws = []
        for suffix in (("", "_reverse") if lstm.bidirectional else ("",)):
            ws += [getattr(lstm, f"{n}_l{l}{suffix}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        if half:
            x, x16, x16b = ops.LstmLayerH.apply(x, x16, x16b, *ws, side_wgrad)
            x16b = x16b if x16b.numel() else None
        else:   # encoder BiLSTM (H = dim_neck): tiny GEMMs, fp32 recurrence; tf32 operands in half mode
            x = ops.LstmLayer.apply(x, PREC_TF32 if prec == PREC_HALF else prec, *ws)
            x16 = x16b = None
    return x, x16, x16b


class Encoder(nn.Module):
    """model_vc_mel.py:41-81."""

    def __init__(self, dim_neck, dim_emb, freq, n_bins=80):
        super().__init__()
        self.dim_neck = dim_neck
        self.freq = freq
        self.prec = PREC_FP32
        convolutions = []
        for i in range(3):
            convolutions.append(nn.Sequential(
                ConvNorm(n_bins + dim_emb if i == 0 else 512, 512, kernel_size=5, stride=1, padding=2, dilation=1,
                         w_init_gain="relu"),
                nn.BatchNorm1d(512)))
        self.convolutions = nn.ModuleList(convolutions)
        self.lstm = nn.LSTM(512, dim_neck, 2, batch_first=True, bidirectional=True)

    def codes(self, x, c_org):
        """(B,T,n_bins) | (B,1,T,n_bins), (B,dim_emb) -> codes (B, T/freq, 2*dim_neck)."""
        if x.dim() == 4:
            x = x.squeeze(1)
        h = ops.ConcatEmb.apply(x, c_org)
        h16 = h16b = None
        n = len(self.convolutions)
        for i, block in enumerate(self.convolutions):
            # the BiLSTM reads the last conv's activation in fp32; the inner ones are only ever 16-bit GEMM operands
            h, h16, h16b = _conv_bn_act(block, h, "relu", prec=self.prec, x16=h16, x16b=h16b, need_z32=(i == n - 1))
        h, _, _ = _lstm(h, self.lstm, self.prec)
        return ops.Codes.apply(h, self.dim_neck, self.freq)

    def forward(self, x, c_org):
        return list(self.codes(x, c_org).unbind(1))       # the reference returns a list (model_vc_mel.py:77-81)


class Decoder(nn.Module):
    """model_vc_mel.py:84-122."""

    def __init__(self, dim_neck, dim_emb, dim_pre, n_bins=80):
        super().__init__()
        self.prec = PREC_FP32
        self.lstm1 = nn.LSTM(dim_neck * 2 + dim_emb, dim_pre, 1, batch_first=True)
        convolutions = []
        for i in range(3):
            convolutions.append(nn.Sequential(
                ConvNorm(dim_pre, dim_pre, kernel_size=5, stride=1, padding=2, dilation=1, w_init_gain="relu"),
                nn.BatchNorm1d(dim_pre)))
        self.convolutions = nn.ModuleList(convolutions)
        self.lstm2 = nn.LSTM(dim_pre, 1024, 2, batch_first=True)
        self.linear_projection = LinearNorm(1024, n_bins)

    def forward(self, x, side_wgrad: bool = False):
        # side_wgrad (set by Generator.forward, where decoder / postnet parameters are used once per graph): the weight
        # gradients are computed on the side stream (ops._wgrad_side)
        h, h16, h16b = _lstm(x, self.lstm1, self.prec, side_wgrad=side_wgrad)
        # half mode: lstm2 (persistent kernels, H = 1024) reads the 16-bit copies only -> no fp32 activation in between
        lstm2_half = self.prec == PREC_HALF and 128 <= self.lstm2.hidden_size <= 1024 and self.lstm2.hidden_size % 64 == 0
        n = len(self.convolutions)
        for i, block in enumerate(self.convolutions):
            h, h16, h16b = _conv_bn_act(block, h, "relu", prec=self.prec, x16=h16, x16b=h16b,
                                        need_z32=(i == n - 1 and not lstm2_half), side_wgrad=side_wgrad)
        h, h16, h16b = _lstm(h, self.lstm2, self.prec, x16=h16, x16b=h16b, side_wgrad=side_wgrad)
        lin = self.linear_projection.linear_layer
        if self.prec == PREC_HALF:
            return ops.LinearH.apply(h, h16, h16b, lin.weight, lin.bias, side_wgrad)
        return ops.Linear.apply(h, lin.weight, lin.bias, self.prec)


class Postnet(nn.Module):
    """model_vc_mel.py:125-169 — five conv(k5)+BN layers, tanh on the first four."""

    def __init__(self, n_bins=80):
        super().__init__()
        self.prec = PREC_FP32
        self.convolutions = nn.ModuleList()
        self.convolutions.append(nn.Sequential(
            ConvNorm(n_bins, 512, kernel_size=5, stride=1, padding=2, dilation=1, w_init_gain="tanh"),
            nn.BatchNorm1d(512)))
        for _ in range(1, 5 - 1):
            self.convolutions.append(nn.Sequential(
                ConvNorm(512, 512, kernel_size=5, stride=1, padding=2, dilation=1, w_init_gain="tanh"),
                nn.BatchNorm1d(512)))
        self.convolutions.append(nn.Sequential(
            ConvNorm(512, n_bins, kernel_size=5, stride=1, padding=2, dilation=1, w_init_gain="linear"),
            nn.BatchNorm1d(n_bins)))

    def channels_last(self, x, residual=None, side_wgrad: bool = False):
        """x (B,T,n_bins) -> postnet(x) (+ residual), channels-last."""
        n = len(self.convolutions)
        x16 = x16b = None
        for i in range(n - 1):
            x, x16, x16b = _conv_bn_act(self.convolutions[i], x, "tanh", prec=self.prec, x16=x16, x16b=x16b, need_z32=False,
                                        side_wgrad=side_wgrad)
        return _conv_bn_act(self.convolutions[-1], x, "none", residual=residual, prec=self.prec, x16=x16, x16b=x16b,
                            side_wgrad=side_wgrad)[0]

    def forward(self, x):
        # reference layout: channel-first (B, n_bins, T) in and out (model_vc_mel.py:163-169, :196)
        return self.channels_last(x.transpose(1, 2).contiguous()).transpose(1, 2)


class Generator(nn.Module):
    """Generator network (model_vc_mel.py:172-203)."""

    def __init__(self, dim_neck, dim_emb, dim_pre, freq, n_bins: int = 80, precision: Optional[str] = None):
        super().__init__()
        self.encoder = Encoder(dim_neck, dim_emb, freq, n_bins=n_bins)
        self.decoder = Decoder(dim_neck, dim_emb, dim_pre, n_bins=n_bins)
        self.postnet = Postnet(n_bins=n_bins)
        self.set_precision(precision or _default_precision())

    def set_precision(self, precision: str):
        if precision not in _PREC:
            raise ValueError(f"precision must be one of {sorted(_PREC)}")
        self.precision = precision
        for m in (self.encoder, self.decoder, self.postnet, self.decoder.linear_projection):
            m.prec = _PREC[precision]
        return self

    def forward(self, x, c_org, c_trg):
        if c_trg is None:                                            # model_vc_mel.py:183-184
            if not torch.is_grad_enabled():
                # stand-alone encoder call (no graph that could still hold packs of this step): never trust packs made
                # before a writer that bypasses the version counters (param.data.copy_, raw-pointer optimizers)
                ops._GLOBAL_CACHE.begin_step()
            codes = self.encoder.codes(x, c_org)
            return codes.reshape(codes.size(0), -1)
        if x.dim() != 3:
            raise ValueError("full forward takes x of shape (B, T, n_bins)")
        ops._GLOBAL_CACHE.begin_step()
        ops.note_full_forward()
        T = x.size(1)
        codes = self.encoder.codes(x, c_org)
        dec_in = ops.UpsampleConcat.apply(codes, c_trg, T)          # :186-192
        x_identic = self.decoder(dec_in, side_wgrad=True)            # :194
        x_identic_psnt = self.postnet.channels_last(x_identic, residual=x_identic, side_wgrad=True)   # :196-197
        code_real = codes.reshape(codes.size(0), -1)                 # :201
        return x_identic.unsqueeze(1), x_identic_psnt.unsqueeze(1), code_real
