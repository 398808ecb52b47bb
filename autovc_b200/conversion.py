"""Zero-shot conversion path (conversion.py:40-44 pad_seq, :46-52 model in eval mode, :90-100 forward with a
swapped speaker embedding and removal of the padded frames), batched: waveforms -> log-mel front-end
(make_spect.py:72-83) -> zero-pad frames to a multiple of 32 -> Generator.eval() forward -> trim.

Eval-mode BatchNorm has no cross-sample coupling, so batching many utterances is exact.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .make_spect import Spect


def padded_frames(n_samples: int, hop: int = 256, base: int = 32) -> int:
    """Frames make_spect produces for n_samples (1 + n//hop), rounded up to a multiple of `base`
    (conversion.py:40-44)."""
    f = 1 + n_samples // hop
    return (f + base - 1) // base * base


@torch.no_grad()
def convert(G, spect: Spect, wav: torch.Tensor, dither: torch.Tensor, lengths: Optional[torch.Tensor],
            emb_org: torch.Tensor, emb_trg: torch.Tensor, chunk: int = 256, base: int = 32
            ) -> Tuple[torch.Tensor, torch.Tensor]:
    """wav, dither: (n, L) float32 CUDA; emb_org/emb_trg: (n, dim_emb).  Returns
    (x_identic_psnt (n, 1, Tpad, 80), n_frames (n,)): frames >= n_frames[i] are padding
    (what conversion.py:97-100 drops)."""
    was_training = G.training
    G.eval()
    n, L = wav.shape
    Tpad = padded_frames(L, spect.hop_length, base)
    S = spect.logmel(wav, dither, lengths, max_frames=Tpad)
    outs: List[torch.Tensor] = []
    for i in range(0, n, chunk):
        _, x_identic_psnt, _ = G(S[i:i + chunk], emb_org[i:i + chunk], emb_trg[i:i + chunk])
        outs.append(x_identic_psnt)
    if lengths is None:
        n_frames = torch.full((n,), 1 + L // spect.hop_length, dtype=torch.int32, device=wav.device)
    else:
        n_frames = 1 + lengths.to(torch.int32) // spect.hop_length
    if was_training:
        G.train()
    return torch.cat(outs, 0), n_frames
