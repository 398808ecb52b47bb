"""Drop-in for the reference's ``model_vc_stft.GeneratorSTFT`` (model_vc_stft.py:7-53): a mel
``Generator`` whose first encoder conv, output projection and first/last postnet layers are
re-created for 513-bin linear spectrograms.

Construction order (and therefore seeded initialisation and ``state_dict`` keys, all under the
``model.`` prefix) follows model_vc_stft.py:13-29: build the 80-bin Generator, then replace the
four layers.  The reference's own ``forward`` is broken — it calls ``self.decoder`` /
``self.postnet`` which do not exist (model_vc_stft.py:44,46) — so the oracle for
this variant is ``GeneratorSTFT(...).model(x, c_org, c_trg)``; here ``forward`` simply does that.
"""
from __future__ import annotations

import torch.nn as nn

from .model_vc_mel import ConvNorm, Generator, LinearNorm


class GeneratorSTFT(nn.Module):
    def __init__(self, dim_neck, dim_emb, dim_pre, freq, precision=None):
        super().__init__()
        self.model = Generator(dim_neck, dim_emb, dim_pre, freq, precision=precision)
        self.model.encoder.convolutions[0][0] = ConvNorm(513 + dim_emb, 512, kernel_size=5, stride=1, padding=2)
        self.model.decoder.linear_projection = LinearNorm(in_dim=1024, out_dim=513)
        self.model.postnet.convolutions[0][0] = ConvNorm(513, 512, kernel_size=5, stride=1, padding=2, dilation=1,
                                                         w_init_gain="tanh")
        self.model.postnet.convolutions[4] = nn.Sequential(
            ConvNorm(512, 513, kernel_size=5, stride=1, padding=2, dilation=1, w_init_gain="linear"),
            nn.BatchNorm1d(513))
        self.model.set_precision(self.model.precision)

    def forward(self, x, c_org, c_trg):
        return self.model(x, c_org, c_trg)
