"""Device-resident crop loader (avc_crop_batch + autovc_b200.data_loader) against the oracle restatement of
data_loader.py:61-80 on the same numpy draws: pure data movement -> bit-exact."""
import numpy as np
import pytest
import torch

from oracle import data_loader_ref as lref

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from autovc_b200 import AvcError, data_loader


@pytest.mark.parametrize("n_bins", [80, 513])
def test_batches_match_the_reference_semantics_bit_for_bit(n_bins):
    corpus = lref.synth_corpus(seed=9, n_bins=n_bins)
    ds = data_loader.Utterances(corpus=corpus, len_crop=128)
    order = [3, 0, 4, 1, 2, 2, 4, 0, 1, 3, 3, 3, 1, 0, 4, 2, 0, 0, 1]
    rs_a, rs_b = np.random.RandomState(77), np.random.RandomState(77)
    for rep in range(4):
        x, e = ds.batch(order, rs_a)
        xr, er, draws = lref.get_batch(corpus, order, 128, rs_b)
        assert x.shape == (len(order), 128, n_bins) and e.shape == (len(order), 256)
        np.testing.assert_array_equal(x.cpu().numpy(), xr)
        np.testing.assert_array_equal(e.cpu().numpy(), er)
    assert rs_a.randint(1 << 30) == rs_b.randint(1 << 30)          # both sides consumed the stream identically


def test_loader_epochs_drop_last_and_feed_a_training_step():
    import autovc_b200
    from autovc_b200 import solver
    corpus = lref.synth_corpus(seed=3, n_spk=7)
    loader = data_loader.get_loader(None, batch_size=2, len_crop=64, corpus=corpus, seed=5)
    assert len(loader) == 3                                          # 7 speakers, batch 2, drop_last (data_loader.py:99)
    seen = []
    for x, e in loader:
        assert x.shape == (2, 64, 80) and x.is_cuda and e.shape == (2, 256)
        seen.append(x)
    assert len(seen) == 3
    torch.manual_seed(0)
    G = autovc_b200.Generator(16, 256, 512, 16).cuda().train()
    out = solver.train_step(G, autovc_b200.FusedAdam(G.parameters(), 1e-4), seen[0], e)
    assert np.isfinite(out["g_loss"])


def test_loader_refuses_cpu():
    with pytest.raises(AvcError):
        data_loader.Utterances(corpus=lref.synth_corpus(), device="cpu")
