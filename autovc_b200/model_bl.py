"""Drop-in for the reference's speaker encoder ``model_bl.D_VECTOR`` (model_bl.py:5-20): a
``num_layers``-layer ``nn.LSTM(dim_input -> dim_cell)``, ``nn.Linear(dim_cell -> dim_emb)`` on the LAST frame and an L2
normalisation.  make_metadata.py:42 instantiates it as ``D_VECTOR(dim_input=80, dim_cell=768, dim_emb=256)`` on 128-frame
mel crops.  Same constructor, parameter names (``lstm.*``, ``embedding.*``) and ``state_dict`` layout, so ``3000000-BL.ckpt``
loads unchanged; the arithmetic runs in the same LSTM / GEMM kernels as the Generator (persistent tcgen05 recurrences in
``half`` mode) plus ``avc_l2_normalize_rows``.  No CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import PREC_HALF, call
from .model_vc_mel import _PREC, _default_precision, _lstm


class D_VECTOR(nn.Module):
    """d vector speaker embedding."""

    def __init__(self, num_layers=3, dim_input=40, dim_cell=256, dim_emb=64, precision: Optional[str] = None):
        super().__init__()
        self.lstm = nn.LSTM(input_size=dim_input, hidden_size=dim_cell, num_layers=num_layers, batch_first=True)
        self.embedding = nn.Linear(dim_cell, dim_emb)
        self.set_precision(precision or _default_precision())

    def set_precision(self, precision: str):
        if precision not in _PREC:
            raise ValueError(f"precision must be one of {sorted(_PREC)}")
        self.precision, self.prec = precision, _PREC[precision]
        return self

    def forward(self, x):
        """x (B, T, dim_input) float32 CUDA -> (B, dim_emb), rows of unit L2 norm."""
        h, h16, h16b = _lstm(x.contiguous(), self.lstm, self.prec)            # model_bl.py:15
        last = h[:, -1:, :].contiguous()                                      # :16  lstm_out[:, -1, :]
        if self.prec == PREC_HALF:
            emb = ops.LinearH.apply(last, None, None, self.embedding.weight, self.embedding.bias)
        else:
            emb = ops.Linear.apply(last, self.embedding.weight, self.embedding.bias, self.prec)
        return L2Normalize.apply(emb.reshape(emb.size(0), -1))               # :17-19


class L2Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ops._check(x)
        out = torch.empty_like(x)
        call("avc_l2_normalize_rows", ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), x.shape[0], x.shape[1],
             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        return out

    @staticmethod
    def backward(ctx, g):
        raise ops._lib.AvcError("the speaker encoder is an inference component (make_metadata.py:41-81): no backward")
