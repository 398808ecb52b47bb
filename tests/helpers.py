"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import generator_ref as gref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(t, n=48):
    f = t.detach().double().flatten().cpu()
    step = max(1, f.numel() // n)
    pad = lambda v: torch.cat([v, v.new_zeros(n - v.numel())])
    return torch.cat([torch.stack([f.sum(), f.abs().sum(), f.norm()]), pad(f[:n]), pad(f[::step][:n])]).numpy()


def synth_inputs(B, T, n_bins, dim_emb, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, T, n_bins, generator=g)
    e = F.normalize(torch.randn(B, dim_emb, generator=g), dim=-1) * 0.8
    e2 = F.normalize(torch.randn(B, dim_emb, generator=g), dim=-1) * 0.8
    return x, e, e2


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def seeded_state_dict(dim_neck, freq, n_bins=80, seed=0):
    """Reference init reproduced from the seed (no weights are committed)."""
    torch.manual_seed(seed)
    if n_bins == 80:
        G = gref.build_reference_like_module(dim_neck, 256, 512, freq)
    else:
        G = build_stft_like_module(dim_neck, freq, seed)
    return {k: v.detach().clone() for k, v in G.state_dict().items()}


def build_stft_like_module(dim_neck, freq, seed):
    """GeneratorSTFT(...).model init order (model_vc_stft.py:13-29): a full 80-bin Generator
    is built first (consuming RNG), then four layers are re-created in this order."""
    import torch.nn as nn
    torch.manual_seed(seed)
    G = gref.build_reference_like_module(dim_neck, 256, 512, freq)
    proto = gref.build_reference_like_module.__globals__  # noqa: F841  (keep flake quiet)

    def convnorm(ci, co, gain):
        m = nn.Module()
        m.conv = nn.Conv1d(ci, co, 5, 1, 2)
        nn.init.xavier_uniform_(m.conv.weight, gain=nn.init.calculate_gain(gain))
        return m

    G.encoder.convolutions[0][0] = convnorm(513 + 256, 512, "linear")
    lin = nn.Module()
    lin.linear_layer = nn.Linear(1024, 513)
    nn.init.xavier_uniform_(lin.linear_layer.weight, gain=1.0)
    G.decoder.linear_projection = lin
    G.postnet.convolutions[0][0] = convnorm(513, 512, "tanh")
    G.postnet.convolutions[4] = nn.Sequential(convnorm(512, 513, "linear"), nn.BatchNorm1d(513))
    return G
