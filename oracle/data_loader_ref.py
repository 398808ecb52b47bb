"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's crop loader, the step BEFORE the hot path
(SURVEY 8(f) rank 1): data_loader.py:61-80 (`Utterances.__getitem__`) and the default collate of :90-102.

Pinned: oracle/gen_golden_loader.py runs the UNMODIFIED reference class on a seeded synthetic corpus written in its
on-disk format and stores what it returned (tests/golden/loader_synth.npz); tests/test_oracle_loader.py replays this
restatement against it bit for bit.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import oracle/.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def synth_corpus(seed: int = 5, n_spk: int = 5, n_bins: int = 80, dim_emb: int = 256,
                 lens: Sequence[Sequence[int]] = ((40, 128, 300), (129, 127, 500, 64), (128,), (1000, 10), (200, 260, 131))):
    """A small corpus in the reference's in-memory layout: per speaker [id, emb (dim_emb,), utt (F_i, n_bins) float32, ...],
    with utterances shorter than, equal to and longer than len_crop = 128."""
    rs = np.random.RandomState(seed)
    corpus = []
    for s in range(n_spk):
        emb = rs.randn(dim_emb).astype(np.float32)
        emb *= 0.8 / np.linalg.norm(emb)
        uttrs = [rs.rand(F, n_bins).astype(np.float32) for F in lens[s % len(lens)]]
        corpus.append(["p%03d" % s, emb] + uttrs)
    return corpus


def get_item(list_uttrs: list, len_crop: int, rs) -> Tuple[np.ndarray, np.ndarray, int, int]:
    """data_loader.py:61-80.  `rs`: numpy RandomState-like (the reference uses the global np.random).
    Returns (uttr (len_crop, n_bins), emb_org, a, left) -- a/left are the draws (left = -1 when none was made)."""
    emb_org = list_uttrs[1]                                   # :65
    a = rs.randint(2, len(list_uttrs))                        # :68
    tmp = list_uttrs[a]
    left = -1
    if tmp.shape[0] < len_crop:                               # :70-73  zero-pad at the end
        uttr = np.pad(tmp, ((0, len_crop - tmp.shape[0]), (0, 0)), "constant")
    elif tmp.shape[0] > len_crop:                             # :74-76  randint's upper bound is exclusive: the last window is never drawn
        left = rs.randint(tmp.shape[0] - len_crop)
        uttr = tmp[left:left + len_crop, :]
    else:                                                     # :77-78
        uttr = tmp
    return uttr, emb_org, a, left


def get_batch(corpus: list, indices: Sequence[int], len_crop: int, rs):
    """One DataLoader batch for the speaker indices the sampler produced, items drawn in order (default collate = stack)."""
    xs, es, draws = [], [], []
    for i in indices:
        u, e, a, left = get_item(corpus[i], len_crop, rs)
        xs.append(u)
        es.append(np.asarray(e, dtype=np.float32))
        draws.append((a, left))
    return np.stack(xs).astype(np.float32), np.stack(es), draws
