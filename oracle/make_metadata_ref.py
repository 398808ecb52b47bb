"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the speaker-embedding loop of make_metadata.py:54-78 over in-memory
utterances: same statements, same ``np.random`` calls in the same order, ``D_VECTOR`` through oracle/model_bl_ref.py one crop
at a time (batch 1, as the reference does)."""
import numpy as np
import torch

from oracle import model_bl_ref as bref


def speaker_embeddings_ref(sd, speakers, num_uttrs=10, len_crop=128, rng=np.random):
    out = {}
    for speaker in sorted(speakers):                                              # :54
        file_list = speakers[speaker]
        assert len(file_list) >= num_uttrs                                        # :61
        idx_uttrs = rng.choice(len(file_list), size=num_uttrs, replace=False)     # :62
        embs = []
        for i in range(num_uttrs):
            tmp = file_list[idx_uttrs[i]]                                         # :65
            candidates = np.delete(np.arange(len(file_list)), idx_uttrs)          # :66
            while tmp.shape[0] < len_crop:                                        # :68
                idx_alt = rng.choice(candidates)
                tmp = file_list[idx_alt]
                candidates = np.delete(candidates, np.argwhere(candidates == idx_alt))
            left = rng.randint(0, tmp.shape[0] - len_crop)                        # :72
            melsp = torch.from_numpy(tmp[np.newaxis, left:left + len_crop, :])    # :73
            emb = bref.dvector_forward(sd, melsp)                                 # :74
            embs.append(emb.detach().squeeze().numpy())                           # :75
        out[speaker] = np.mean(embs, axis=0)                                      # :78
    return out


def synth_speakers(n_speakers=3, seed=5):
    rs = np.random.RandomState(seed)
    sp = {}
    for s in range(n_speakers):
        lens = rs.randint(90, 260, size=13)
        lens[:2] = [60, 100]                       # some utterances shorter than the crop: exercises the re-draw of :68-71
        sp["p%03d" % (225 + s)] = [rs.rand(int(n), 80).astype(np.float32) for n in lens]
    return sp
