"""Drop-in for the reference's crop loader (data_loader.py:11-102), with the corpus resident in HBM.

`Utterances(data_dir, len_crop, model_type)` reads the same `train.pkl` + `.npy` files (data_loader.py:20-46) -- or takes the
in-memory list `[[speaker, emb, utt, ...], ...]` -- and uploads every utterance ONCE into a ragged `(sum F_i, n_bins)` device
buffer.  A batch is then three small int arrays (utterance, crop offset, speaker) drawn on the host EXACTLY like
`__getitem__` draws them (`np.random.randint(2, len(list_uttrs))`, then `np.random.randint(F - len_crop)` for long
utterances, item after item, data_loader.py:68,:75) and one launch of `avc_crop_batch`, instead of B python-side crops, a
collate and a 10 MB host-to-device copy per step.  `get_loader` mirrors data_loader.py:90-102 (shuffle over speakers,
`drop_last=True`).  No CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
import pickle
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import call


class CorpusIndex:
    """Host-side (numpy) index of a corpus in the reference's in-memory layout `[[speaker, emb, utt, ...], ...]`: the
    per-speaker utterance ranges, the utterance lengths / row offsets of the ragged buffer, and the random draws of a batch in
    the reference's order.  No device needed (the CPU tests replay it against the oracle)."""

    def __init__(self, corpus, len_crop: int):
        self.len_crop = len_crop
        lens, first = [], []
        for spk in corpus:
            if len(spk) < 3:
                raise ValueError(f"speaker {spk[0]!r} has no utterance")
            first.append(len(lens))
            lens += [int(np.asarray(u).shape[0]) for u in spk[2:]]
        self.spk_first = np.asarray(first, dtype=np.int64)                  # global index of a speaker's first utterance
        self.spk_count = np.asarray([len(s) - 2 for s in corpus], dtype=np.int64)
        self.utt_len = np.asarray(lens, dtype=np.int64)
        self.utt_row0 = np.concatenate([[0], np.cumsum(self.utt_len)[:-1]]).astype(np.int64)
        self.num_tokens = len(corpus)                                       # data_loader.py:45

    def draw(self, indices: Sequence[int], rs=np.random) -> np.ndarray:
        """The host-side random choices of one batch, in the reference's order: (3, B) int32 = utterance, left, speaker."""
        sel = np.zeros((3, len(indices)), dtype=np.int32)
        T = self.len_crop
        for k, i in enumerate(indices):
            a = rs.randint(2, self.spk_count[i] + 2)                        # data_loader.py:68
            u = self.spk_first[i] + a - 2
            F = self.utt_len[u]
            sel[0, k] = u
            sel[1, k] = rs.randint(F - T) if F > T else 0                   # :75 (no draw otherwise, like the reference)
            sel[2, k] = i
        return sel


class Utterances:
    """Device-resident counterpart of data_loader.Utterances (one entry per speaker, like the reference)."""

    def __init__(self, data_dir: Optional[str] = None, len_crop: int = 128, model_type: str = "spmel", corpus=None,
                 device="cuda"):
        self.len_crop = len_crop
        if corpus is None:
            root = os.path.join(data_dir, model_type)                       # data_loader.py:16
            with open(os.path.join(root, "train.pkl"), "rb") as f:          # :21-22
                meta = pickle.load(f)
            corpus = [[s[0], s[1]] + [np.load(os.path.join(root, p)) for p in s[2:]] for s in meta]   # :49-57
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.AvcError("autovc_b200.data_loader keeps the corpus in GPU memory (there is no CPU fallback)")
        self.index = CorpusIndex(corpus, len_crop)
        chunks = [np.ascontiguousarray(np.asarray(u, dtype=np.float32)) for spk in corpus for u in spk[2:]]
        self.n_bins = chunks[0].shape[1]
        if any(c.shape[1] != self.n_bins for c in chunks):
            raise ValueError("all utterances must have the same number of bins")
        self.corpus = torch.from_numpy(np.concatenate(chunks, 0)).to(self.device)         # (sum F, n_bins)
        self.utt_row0 = torch.from_numpy(self.index.utt_row0).to(self.device)
        self.utt_len = torch.from_numpy(self.index.utt_len.astype(np.int32)).to(self.device)
        self.emb_table = torch.from_numpy(np.stack([np.asarray(s[1], dtype=np.float32) for s in corpus])).to(self.device)
        self.num_tokens = self.index.num_tokens

    def __len__(self):
        return self.num_tokens

    def draw(self, indices: Sequence[int], rs=np.random) -> np.ndarray:
        return self.index.draw(indices, rs)

    def batch(self, indices: Sequence[int], rs=np.random) -> Tuple[torch.Tensor, torch.Tensor]:
        """(x_real (B, len_crop, n_bins), emb_org (B, dim_emb)) on the device for the given speaker indices."""
        sel = self.draw(indices, rs)
        return self.gather(sel)

    def gather(self, sel: np.ndarray) -> Tuple[torch.Tensor, torch.Tensor]:
        B = sel.shape[1]
        dsel = torch.from_numpy(sel).to(self.device, non_blocking=True)    # pageable source: staged by the driver, no stream sync
        x = torch.empty(B, self.len_crop, self.n_bins, device=self.device, dtype=torch.float32)
        e = torch.empty(B, self.emb_table.shape[1], device=self.device, dtype=torch.float32)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        call("avc_crop_batch", P(self.corpus), P(self.utt_row0), P(self.utt_len), P(dsel[0]), P(dsel[1]), P(self.emb_table),
             P(dsel[2]), P(x), P(e), B, self.len_crop, self.n_bins, self.emb_table.shape[1],
             ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        return x, e


class CropLoader:
    """Iterates like the reference's DataLoader(shuffle=True, drop_last=True) over speakers (data_loader.py:95-101)."""

    def __init__(self, dataset: Utterances, batch_size: int, seed: Optional[int] = None, rs=None):
        self.dataset, self.batch_size = dataset, batch_size
        self.gen = torch.Generator()
        if seed is not None:
            self.gen.manual_seed(seed)
        self.rs = rs if rs is not None else (np.random.RandomState(seed) if seed is not None else np.random)

    def __len__(self):
        return len(self.dataset) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        perm = torch.randperm(len(self.dataset), generator=self.gen).tolist()
        for k in range(len(self)):
            yield self.dataset.batch(perm[k * self.batch_size:(k + 1) * self.batch_size], self.rs)


def get_loader(root_dir, batch_size=16, len_crop=128, model_type="spmel", num_workers=0, corpus=None, device="cuda",
               seed: Optional[int] = None) -> CropLoader:
    """Same call as data_loader.get_loader (data_loader.py:90); `num_workers` is accepted and ignored (nothing to
    parallelise: a batch is one kernel launch).  Under data parallelism pass seed = base + rank."""
    return CropLoader(Utterances(root_dir, len_crop, model_type, corpus=corpus, device=device), batch_size, seed=seed)
