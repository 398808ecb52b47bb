"""The waveform-variant oracle (oracle/generator_wav_ref.py) against the goldens produced by the UNMODIFIED reference
(oracle/gen_golden_wav.py -> tests/golden/wav_*.npz).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import generator_wav_ref as wref
from tests.helpers import digest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    dim_neck, freq, depth, B, L, wseed, iseed, steps = g["meta"].tolist()
    return g, dim_neck, freq, depth, B, L, wseed, iseed


@pytest.mark.parametrize("name", ["wav_16_16_d1_b2"])
def test_wav_oracle_matches_reference_golden(name):
    g, dim_neck, freq, depth, B, L, wseed, iseed = _load(name)
    torch.manual_seed(wseed)
    M = wref.build_wav_module(dim_neck, 256, 512, freq, depth)
    assert [k for k, _ in M.named_parameters()] == g["param_names"].tolist()
    assert list(M.state_dict().keys()) == g["state_dict_keys"].tolist()
    got = np.stack([digest(p) for p in M.parameters()])
    assert np.array_equal(got, g["param_digest0"]), "seeded init differs from the reference's"
    sd = {k: v.detach().clone() for k, v in M.state_dict().items()}
    x, e = wref.synth_wav_inputs(B, L, 256, iseed)
    losses, outs, grads = wref.wav_train_step(sd, x, e, dim_neck, freq)
    ref = g["s0_losses"]
    for i, k in enumerate(("g_loss", "L_id", "L_gen", "L_cd", "L_SISNR")):
        assert abs(float(losses[k]) - ref[i]) <= 2e-5 * max(1.0, abs(ref[i])), (k, float(losses[k]), ref[i])
    for k in ("x_convtas", "x_identic", "gen_outputs", "code_real", "code_reconst"):
        assert np.abs(outs[k].numpy() - g["s0_" + k]).max() < 1e-4, k
    gd = np.stack([digest(v) for v in grads.values()])
    ref_gd = g["s0_grad_digest"]
    # l2 norms of every gradient tensor (column 2 of the digest)
    rel = np.abs(gd[:, 2] - ref_gd[:, 2]) / np.maximum(ref_gd[:, 2], 1e-6)
    # conv biases in front of a train-mode BatchNorm have an identically-zero gradient; both sides hold rounding noise there
    keep = np.array([not k.endswith(".conv.bias") for k in grads.keys()])
    assert rel[keep].max() < 2e-3, (rel.argmax(), rel.max())
    for k in g.files:
        if k.startswith("s0_buf/"):
            assert np.allclose(sd[k[7:]].numpy(), g[k], rtol=1e-5, atol=1e-6), k
