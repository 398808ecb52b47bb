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


def loss_curve_corpus(n_utt=64, frames=400, seed=7):
    """Synthetic 'speakers' of the loss-curve gate: smooth band-limited mel-like trajectories in [0,1] plus
    per-speaker embeddings (shared by oracle/gen_loss_curve_ref.py, tests/test_gpu_loss_curve.py and
    scripts/loss_curve.py so that the reference curve and the GPU curves see the same data stream)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(n_utt, frames // 8 + 2, 80, generator=g)
    x = F.interpolate(base.permute(0, 2, 1), size=frames, mode="linear", align_corners=True).permute(0, 2, 1)
    x = (0.7 * x + 0.3 * torch.rand(n_utt, frames, 80, generator=g)).clamp(0, 1)
    e = F.normalize(torch.randn(n_utt, 256, generator=g), dim=-1) * 0.8
    return x, e


def loss_curve_batches(steps, B, T, n_utt=64, frames=400, seed=123):
    """The (utterance index, crop offset) draws of every step of the loss-curve gate."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(steps):
        idx = rs.randint(0, n_utt, size=B)
        off = rs.randint(0, frames - T, size=B)
        out.append((idx, off))
    return out


def movavg(v, w=25):
    c = np.cumsum(np.insert(np.asarray(v, np.float64), 0, 0.0))
    return (c[w:] - c[:-w]) / w
