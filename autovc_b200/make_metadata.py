"""The speaker-embedding part of the reference's ``make_metadata.Metadata.metadata`` (make_metadata.py:41-81): per speaker,
``num_uttrs`` (10) utterances are drawn without replacement, each is cropped to ``len_crop`` (128) frames at a random offset
(utterances shorter than the crop are replaced by another draw, :68-72), the crops go through ``D_VECTOR`` and the speaker's
embedding is the mean of the 10 d-vectors (:78).

Host logic (the numpy draws, in the reference's order and from the same ``np.random`` stream so that a seeded run picks the
same crops) is separated from the device work: ALL crops of ALL speakers are gathered into one (S*10, 128, 80) batch and run
through ONE ``D_VECTOR`` forward (the reference runs S*10 batch-1 forwards), then averaged per speaker on the device.
File walking, the pickles and the conversion log of :83-133 are control plane and stay with the caller.  No CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch


def draw_crops(n_frames: Sequence[int], num_uttrs: int = 10, len_crop: int = 128, rng=np.random) -> List[Tuple[int, int]]:
    """The random choices of make_metadata.py:62-75 for ONE speaker whose utterances have ``n_frames[i]`` frames:
    returns [(utterance index, left offset)] * num_uttrs, consuming ``rng`` exactly like the reference
    (``choice(len, size, replace=False)``; per utterance, ``choice(candidates)`` while too short; ``randint(0, F - len_crop)``)."""
    n = len(n_frames)
    assert n >= num_uttrs                                                        # :61
    idx_uttrs = rng.choice(n, size=num_uttrs, replace=False)                     # :62
    picks = []
    for i in range(num_uttrs):
        cur = int(idx_uttrs[i])
        candidates = np.delete(np.arange(n), idx_uttrs)                          # :66
        while n_frames[cur] < len_crop:                                          # :68
            idx_alt = rng.choice(candidates)                                     # :69
            cur = int(idx_alt)
            candidates = np.delete(candidates, np.argwhere(candidates == idx_alt))   # :71
        left = int(rng.randint(0, n_frames[cur] - len_crop))                     # :72  (raises for F == len_crop, like the reference)
        picks.append((cur, left))
    return picks


@torch.no_grad()
def speaker_embeddings(C, speakers: "Dict[str, List[torch.Tensor | np.ndarray]]", num_uttrs: int = 10, len_crop: int = 128,
                       rng=np.random, device=None) -> "Dict[str, np.ndarray]":
    """speakers: name -> list of (F_i, 80) mel spectrograms (numpy or tensors).  Speakers are visited in sorted order
    (make_metadata.py:54) so the draws line up with the reference.  Returns name -> (dim_emb,) float32 mean d-vector."""
    names = sorted(speakers)
    if not names:
        return {}
    device = device or next(C.parameters()).device
    crops = []
    for name in names:
        utts = speakers[name]
        for ui, left in draw_crops([int(u.shape[0]) for u in utts], num_uttrs, len_crop, rng):
            u = utts[ui]
            u = torch.from_numpy(np.ascontiguousarray(u)) if isinstance(u, np.ndarray) else u
            crops.append(u[left:left + len_crop].to(device=device, dtype=torch.float32, non_blocking=True))
    batch = torch.stack(crops, 0)                                    # (S * num_uttrs, len_crop, 80)
    embs = C(batch)                                                  # :76  one forward for every crop of every speaker
    mean = embs.view(len(names), num_uttrs, -1).mean(dim=1)          # :78  np.mean(embs, axis=0) per speaker
    out = mean.cpu().numpy()
    return {n: out[i] for i, n in enumerate(names)}
