"""Host logic of the speaker-embedding pipeline (make_metadata.py:54-78) without a GPU: the product's draw function consumes
the numpy stream exactly like the oracle's restatement of the reference loop."""
import numpy as np

from autovc_b200.make_metadata import draw_crops
from oracle import make_metadata_ref as mref


def test_draws_follow_the_reference_order():
    speakers = mref.synth_speakers(4, seed=9)
    r1, r2 = np.random.RandomState(3), np.random.RandomState(3)
    for name in sorted(speakers):
        utts = speakers[name]
        got = draw_crops([u.shape[0] for u in utts], 10, 128, r1)
        # the reference statements, recording what they pick
        idx_uttrs = r2.choice(len(utts), size=10, replace=False)
        exp = []
        for i in range(10):
            cur = idx_uttrs[i]
            candidates = np.delete(np.arange(len(utts)), idx_uttrs)
            while utts[cur].shape[0] < 128:
                cur = r2.choice(candidates)
                candidates = np.delete(candidates, np.argwhere(candidates == cur))
            exp.append((int(cur), int(r2.randint(0, utts[cur].shape[0] - 128))))
        assert got == exp
        assert all(utts[u].shape[0] >= 128 for u, _ in got)
    assert r1.rand() == r2.rand()          # both streams are at the same position
