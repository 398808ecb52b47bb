"""Generates tests/golden/loader_synth.npz by running the UNMODIFIED reference `data_loader.Utterances` (imported from
/root/reference) on oracle.data_loader_ref.synth_corpus() written to a temporary directory in the reference's on-disk format
(train.pkl + .npy files).  Run in the build container only:  python oracle/gen_golden_loader.py"""
import os
import pickle
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import data_loader_ref as lref  # noqa: E402


def main():
    import data_loader as ref_dl          # the reference module
    corpus = lref.synth_corpus()
    tmp = tempfile.mkdtemp()
    root = os.path.join(tmp, "spmel")
    os.makedirs(root)
    meta = []
    for spk in corpus:
        os.makedirs(os.path.join(root, spk[0]))
        entry = [spk[0], spk[1]]
        for j, u in enumerate(spk[2:]):
            rel = f"{spk[0]}/{spk[0]}_{j:03d}.npy"
            np.save(os.path.join(root, rel), u)
            entry.append(rel)
        meta.append(entry)
    pickle.dump(meta, open(os.path.join(root, "train.pkl"), "wb"))
    ds = ref_dl.Utterances(tmp, 128, "spmel")
    order = [3, 0, 4, 1, 2, 2, 4, 0, 1, 3, 3, 3, 1, 0, 4, 2] * 3          # speaker indices handed to __getitem__
    np.random.seed(2024)                                                   # the reference draws from the global numpy RNG
    xs, es = [], []
    for i in order:
        u, e = ds[i]
        xs.append(np.asarray(u.cpu().numpy() if hasattr(u, "cpu") else u, dtype=np.float32))
        es.append(np.asarray(e, dtype=np.float32))
    out = os.path.join(ROOT, "tests", "golden", "loader_synth.npz")
    np.savez_compressed(out, order=np.array(order), seed=np.array(2024), len_crop=np.array(128),
                        x=np.stack(xs).astype(np.float16 if False else np.float32)[:, ::8, ::5],   # sub-sampled rows/bins keep it small
                        x_sum=np.stack(xs).astype(np.float64).sum(axis=(1, 2)), e=np.stack(es)[:, :8])
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
