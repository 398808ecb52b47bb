"""Drop-in for the reference's ``model_vc_wav.GeneratorWav`` (model_vc_wav.py:11-102): the AutoVC encoder / decoder
between a learned analysis filterbank (``ConvTasNetEncoder``) and its synthesis counterpart (``ConvTasNetDecoder``),
trained on raw waveform crops of (127*256)+1024 = 33536 samples (main.py:59).

Same constructors, sub-module names (``tasEncoder``, ``encoder``, ``decoder``, ``tasDecoder``), parameter / buffer names
and registration order as the reference, so a seeded construction reproduces its initialisation and ``state_dict()``
interchanges.  The torch.nn layers are parameter containers only: the arithmetic goes through ``autovc_b200.ops`` /
``ops_wav`` into libautovc_b200.so (no CPU path).

Return contract of ``forward(x, c_org, c_trg)`` (model_vc_wav.py:74-102): ``x`` is (B, L, 1); with ``c_trg is None`` the
concatenated codes; otherwise ``(x_CTencoder (B,512,T), x_identic (B,L,1), x_decoder (B,512,T), code_real)``.  The two
(B,512,T) tensors are transposed *views* of channels-last storage.

Precision: the filterbank and k=3 layers run as 3xTF32 split products in ``fp32`` mode (CUDA-core fp32 in ``fp32_simt``)
and as plain tf32 on the tensor cores in ``tf32`` / ``half`` mode; the AutoVC encoder / decoder in between follow the mode like the mel model.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops, ops_wav
from ._lib import PREC_FP32, PREC_FP32X3, PREC_TF32
from .model_vc_mel import _PREC, ConvNorm, Decoder, Encoder, LinearNorm, _default_precision

_N, _L, _S = 512, 1024, 256          # model_vc_wav.py:14-16 / :38-40


def _conv_prelu_bn(block: nn.Sequential, x, transposed: bool, prec: int):
    conv, prelu, bn = block[0], block[1], block[2]
    z = ops_wav.ConvPReLUBn.apply(x, conv.weight, conv.bias, prelu.weight, bn.weight, bn.bias, bn.running_mean,
                                  bn.running_var, transposed, bn.training, prec)
    if bn.training:
        bn.num_batches_tracked += 1
    return z


class ConvTasNetEncoder(nn.Module):
    """model_vc_wav.py:11-33."""

    def __init__(self, depth, bias=True):
        super().__init__()
        if not bias:
            raise ValueError("autovc_b200.ConvTasNetEncoder supports bias=True (what GeneratorWav builds)")
        self.prec = PREC_FP32
        self.conv1x1 = nn.Conv1d(1, _N, kernel_size=_L, stride=_S, padding=0, bias=bias)
        self.convD = nn.ModuleList([
            nn.Sequential(nn.Conv1d(_N, _N, kernel_size=3, stride=1, padding=1, bias=bias), nn.PReLU(), nn.BatchNorm1d(_N))
            for _ in range(depth)])

    def channels_last(self, wav):
        """wav (B, L) -> (B, T, 512), T = (L - 1024)/256 + 1."""
        x = ops_wav.FrameConv.apply(wav, self.conv1x1.weight, self.conv1x1.bias, _S, self.prec)
        for block in self.convD:
            x = _conv_prelu_bn(block, x, False, self.prec)
        return x

    def forward(self, x):
        # reference layout: (B, 1, L) -> (B, 512, T)
        return self.channels_last(x.reshape(x.size(0), -1)).transpose(1, 2)


class ConvTasNetDecoder(nn.Module):
    """model_vc_wav.py:36-58."""

    def __init__(self, depth, bias=True):
        super().__init__()
        if not bias:
            raise ValueError("autovc_b200.ConvTasNetDecoder supports bias=True (what GeneratorWav builds)")
        self.prec = PREC_FP32
        self.convTD = nn.ModuleList([
            nn.Sequential(nn.ConvTranspose1d(_N, _N, kernel_size=3, stride=1, padding=1, bias=bias), nn.PReLU(),
                          nn.BatchNorm1d(_N))
            for _ in range(depth)])
        self.convT1x1 = nn.ConvTranspose1d(_N, 1, kernel_size=_L, stride=_S, padding=0, bias=bias)

    def channels_last(self, x):
        """x (B, T, 512) -> waveform (B, (T + 3)*256)."""
        for block in self.convTD:
            x = _conv_prelu_bn(block, x, True, self.prec)
        return ops_wav.FrameConvT.apply(x, self.convT1x1.weight, self.convT1x1.bias, _S, self.prec)

    def forward(self, x):
        # reference layout: (B, 512, T) -> (B, 1, L)
        return self.channels_last(x.transpose(1, 2).contiguous()).unsqueeze(1)


class GeneratorWav(nn.Module):
    """Generator network on raw waveforms (model_vc_wav.py:60-102)."""

    def __init__(self, dim_neck, dim_emb, dim_pre, freq, depth, precision: Optional[str] = None):
        super().__init__()
        # construction order = the reference's (model_vc_wav.py:66-73): it fixes the RNG stream of a seeded initialisation
        self.tasEncoder = ConvTasNetEncoder(depth)
        self.encoder = Encoder(dim_neck, dim_emb, freq)
        self.decoder = Decoder(dim_neck, dim_emb, dim_pre)
        self.encoder.convolutions[0][0] = ConvNorm(_N + dim_emb, 512, kernel_size=5, stride=1, padding=2)
        self.decoder.linear_projection = LinearNorm(in_dim=1024, out_dim=_N)
        self.tasDecoder = ConvTasNetDecoder(depth)
        self.set_precision(precision or _default_precision())

    def set_precision(self, precision: str):
        if precision not in _PREC:
            raise ValueError(f"precision must be one of {sorted(_PREC)}")
        self.precision = precision
        for m in (self.encoder, self.decoder, self.decoder.linear_projection):
            m.prec = _PREC[precision]
        for m in (self.tasEncoder, self.tasDecoder):
            m.prec = {"fp32": PREC_FP32X3, "fp32_simt": PREC_FP32}.get(precision, PREC_TF32)
        return self

    def forward(self, x, c_org, c_trg):
        if x.dim() != 3 or x.size(2) != 1:
            raise ValueError("GeneratorWav takes x of shape (B, L, 1)")        # model_vc_wav.py:75-76
        if c_trg is None:
            if not torch.is_grad_enabled():
                ops._GLOBAL_CACHE.begin_step()
        else:
            ops._GLOBAL_CACHE.begin_step()
        B = x.size(0)
        ct = self.tasEncoder.channels_last(x.reshape(B, -1))                  # :81-83   (B, T, 512)
        codes = self.encoder.codes(ct, c_org)                                  # :87
        if c_trg is None:
            return codes.reshape(B, -1)                                        # :89-90
        T = ct.size(1)
        dec_in = ops.UpsampleConcat.apply(codes, c_trg, T)                     # :91-96
        dec = self.decoder(dec_in)                                             # :97      (B, T, 512)
        wav = self.tasDecoder.channels_last(dec)                               # :100     (B, L)
        code_real = codes.reshape(B, -1)                                       # :101
        return ct.transpose(1, 2), wav.unsqueeze(-1), dec.transpose(1, 2), code_real
