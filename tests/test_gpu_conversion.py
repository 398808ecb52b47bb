"""Conversion path (waveform -> log-mel -> pad to x32 -> eval forward with swapped embedding) against
the two CPU oracles chained together."""
import numpy as np
import pytest
import torch

from oracle import generator_ref as gref
from oracle import make_spect_ref as fref
from tests.helpers import synth_inputs

pytestmark = pytest.mark.gpu


def test_conversion_matches_chained_oracles():
    import autovc_b200
    from autovc_b200.conversion import convert, padded_frames
    from autovc_b200.make_spect import Spect
    torch.manual_seed(0)
    G = autovc_b200.Generator(32, 256, 512, 32).cuda()
    # make the BN running statistics non-trivial, as a trained checkpoint would have
    x, e, e2 = synth_inputs(3, 64, 80, 256, 11)
    G.train()
    with torch.no_grad():
        G(x.cuda(), e.cuda(), e.cuda())
    sd = {k: v.detach().cpu().clone() for k, v in G.state_dict().items()}
    n, L = 3, 16000 + 700
    wav, dither = fref.synthetic_waveforms(n, L, seed=5)
    lengths = np.array([L, 12000, 9999], np.int32)
    out, n_frames = convert(G, Spect(), torch.from_numpy(wav).cuda(), torch.from_numpy(dither.astype(np.float32)).cuda(),
                            torch.from_numpy(lengths).cuda(), e.cuda(), e2.cuda(), chunk=2)
    Tpad = padded_frames(L)
    assert out.shape == (n, 1, Tpad, 80) and Tpad % 32 == 0
    for i in range(n):
        S = fref.logmel_from_wav(wav[i, :lengths[i]], dither[i, :lengths[i]])
        F = S.shape[0]
        assert int(n_frames[i]) == F
        Town = padded_frames(int(lengths[i]))                          # conversion.py:40-44: each utterance's OWN multiple of 32
        Sp = np.zeros((Town, 80), np.float32)
        Sp[:F] = S                                                      # pad_seq: zero frames
        with torch.no_grad():
            _, ref, _ = gref.generator_forward(sd, torch.from_numpy(Sp)[None], e[i:i + 1], e2[i:i + 1], 32, 32, training=False)
        err = float((out[i, 0, :F].cpu() - ref[0, 0, :F]).abs().max())
        assert err < 2e-4, (i, err)          # 1e-4 front-end + 1e-4 generator budgets


@pytest.mark.parametrize("precision", ["fp32", "half"])
def test_chunks_on_several_streams_change_nothing(precision):
    """convert() alternates chunks over CUDA streams (recurrences of one chunk run beside the GEMMs of another): eval-mode chunks
    are independent, so any (chunk, streams) setting must give the result of one call over the whole batch."""
    import autovc_b200
    from autovc_b200.conversion import convert
    from autovc_b200.make_spect import Spect
    torch.manual_seed(1)
    G = autovc_b200.Generator(32, 256, 512, 32, precision=precision).cuda()
    x, e, e2 = synth_inputs(7, 64, 80, 256, 3)
    G.train()
    with torch.no_grad():
        G(x.cuda(), e.cuda(), e.cuda())
    wav, dither = fref.synthetic_waveforms(7, 9000, seed=2)
    w, d = torch.from_numpy(wav).cuda(), torch.from_numpy(dither.astype(np.float32)).cuda()
    sp = Spect()
    ref, _ = convert(G, sp, w, d, None, e.cuda(), e2.cuda(), chunk=16, streams=1)
    for chunk, streams in ((2, 3), (3, 2), (1, 4)):
        for _ in range(3):
            out, _ = convert(G, sp, w, d, None, e.cuda(), e2.cuda(), chunk=chunk, streams=streams)
            torch.cuda.synchronize()
            assert float((out - ref).abs().max()) < (1e-5 if precision == "fp32" else 2e-3), (chunk, streams)
