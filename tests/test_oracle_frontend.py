"""Pin oracle/make_spect_ref.py against the reference's bundled wav -> spmel goldens."""
import os
import warnings

import numpy as np
import pytest

from oracle import make_spect_ref as fref
from tests.helpers import load_golden

REF = "/root/reference"


def test_mel_filterbank_shape_and_support():
    m = fref.mel_filterbank()
    assert m.shape == (80, 513) and m.dtype == np.float32
    assert int((m != 0).sum()) == 941                      # SURVEY §8 a9 (probe of librosa 0.9.1 output)
    nz = np.nonzero(m.sum(0))[0]
    assert nz[0] == 6 and nz[-1] == 486


def test_butter_coefficients():
    b, a = fref.butter_highpass()
    np.testing.assert_allclose(b, [0.98111838, -4.90559192, 9.81118384, -9.81118384, 4.90559192, -0.98111838], atol=1e-8)
    np.testing.assert_allclose(a, [1, -4.96187604, 9.84822985, -9.77342482, 4.84966429, -0.96259328], atol=1e-8)


def test_frontend_matches_committed_goldens():
    g = load_golden("frontend_bundled")
    for i, (name, off) in enumerate(zip(g["names"].tolist(), g["offsets"].tolist())):
        spk = name.split("/")[0]
        wav = g[f"wav{i}"].astype(np.float32) / 32768.0
        prng = np.random.RandomState(int(spk[1:]))
        prng.rand(off)                                       # skip earlier files of this speaker
        S = fref.logmel_from_wav(wav, prng.rand(len(wav)))
        ref = g[f"spmel{i}"]
        assert S.shape == ref.shape and S.dtype == np.float32
        assert np.abs(S - ref).max() < 2e-6, name


@pytest.mark.skipif(not os.path.isdir(REF + "/wavs"), reason="reference tree only exists in the build container")
def test_frontend_matches_all_71_reference_goldens():
    from scipy.io import wavfile
    n = 0
    worst = 0.0
    for spk in sorted(os.listdir(REF + "/wavs")):
        prng = np.random.RandomState(int(spk[1:]))
        for fn in sorted(os.listdir(f"{REF}/wavs/{spk}")):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                fs, w = wavfile.read(f"{REF}/wavs/{spk}/{fn}")
            d = prng.rand(len(w))
            npy = f"{REF}/spmel/{spk}/{fn[:-4]}.npy"
            if not os.path.exists(npy):
                continue
            S = fref.logmel_from_wav(w.astype(np.float32) / 32768.0, d)
            worst = max(worst, float(np.abs(S - np.load(npy)).max()))
            n += 1
    assert n == 71 and worst < 2e-6, (n, worst)


def test_product_mel_basis_equals_oracle_filterbank():
    """Host-side constant of the product (autovc_b200.make_spect.mel_basis) vs the oracle's
    restatement of librosa.filters.mel."""
    from autovc_b200.make_spect import Spect, mel_basis
    np.testing.assert_allclose(mel_basis(), fref.mel_filterbank().T, rtol=3e-7, atol=0)   # 1 float32 ulp
    assert int((mel_basis() != 0).sum()) == 941
    b, a = Spect().butter_highpass()
    rb, ra = fref.butter_highpass()
    np.testing.assert_array_equal(b, rb)
    np.testing.assert_array_equal(a, ra)
