"""Pin the CPU oracle (oracle/generator_ref.py) to the golden vectors produced by the
unmodified reference (oracle/gen_golden.py -> tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from oracle import generator_ref as gref
from tests.helpers import digest, load_golden, seeded_state_dict, synth_inputs


@pytest.mark.parametrize("name", ["train_16_16_b2_t128", "train_32_32_b3_t64", "train_stft_16_16_b2_t32"])
def test_seeded_init_matches_reference(name):
    g = load_golden(name)
    dim_neck, freq, B, T, n_bins, wseed, iseed, steps = g["meta"].tolist()
    sd = seeded_state_dict(dim_neck, freq, n_bins, wseed)
    assert list(sd.keys()) == g["state_dict_keys"].tolist()
    params, _ = gref.split_state_dict(sd)
    assert list(params.keys()) == g["param_names"].tolist()
    got = np.stack([digest(p) for p in params.values()])
    np.testing.assert_array_equal(got[:, 3:], g["param_digest0"][:, 3:])          # sampled values: bit-exact init
    np.testing.assert_allclose(got[:, :3], g["param_digest0"][:, :3], rtol=1e-12)  # reductions: summation order only


@pytest.mark.parametrize("name", ["train_16_16_b2_t128", "train_32_32_b3_t64", "train_stft_16_16_b2_t32"])
def test_oracle_train_step_matches_reference(name):
    g = load_golden(name)
    dim_neck, freq, B, T, n_bins, wseed, iseed, steps = g["meta"].tolist()
    sd = seeded_state_dict(dim_neck, freq, n_bins, wseed)
    x, e, _ = synth_inputs(B, T, n_bins, 256, iseed)
    adam = {}
    losses, outs, grads = gref.train_step(sd, x, e, dim_neck, freq, 1.0, adam_state=adam, lr=1e-4)
    ref_l = g["s0_losses"]
    got_l = np.array([losses[k].item() for k in ("g_loss", "L_id", "L_id_psnt", "L_cd")])
    np.testing.assert_allclose(got_l, ref_l, rtol=0, atol=2e-6)
    for k in ("x_identic", "x_identic_psnt", "code_real", "code_reconst"):
        err = np.abs(outs[k].numpy() - g["s0_" + k]).max()
        assert err < 2e-5, (k, err)
    # gradients: digests (sum, |sum|, l2, samples) vs the reference's autograd
    gd = np.stack([digest(v) for v in grads.values()])
    ref = g["s0_grad_digest"]
    names = g["param_names"].tolist()
    for i, n in enumerate(names):
        if ".conv.bias" in n:        # mathematically zero, rounding noise only
            assert np.abs(gd[i][3:]).max() < 1e-6
            continue
        scale = max(ref[i][2] / np.sqrt(max(1, grads[n].numel())), 1e-12)   # rms of the tensor
        assert abs(gd[i][2] - ref[i][2]) <= 2e-3 * ref[i][2] + 1e-9, (n, gd[i][2], ref[i][2])
        assert np.abs(gd[i][3:] - ref[i][3:]).max() <= 5e-2 * scale + 1e-7, n
    # BN buffers after the two encoder passes + one decoder/postnet pass
    for k in g.files:
        if k.startswith("s0_buf/"):
            np.testing.assert_allclose(sd[k[7:]].numpy(), g[k], rtol=1e-5, atol=1e-6, err_msg=k)
    assert int(sd["encoder.convolutions.0.1.num_batches_tracked"]) == 2
    assert int(sd["decoder.convolutions.0.1.num_batches_tracked"]) == 1
    # Adam step (solver_encoder.py:130,:300)
    params, _ = gref.split_state_dict(sd)
    pd = np.stack([digest(p) for p in params.values()])
    refp = g["s0_param_digest"]
    for i, n in enumerate(names):
        if ".conv.bias" in n:
            continue
        np.testing.assert_allclose(pd[i][2], refp[i][2], rtol=1e-3, err_msg=n)   # Adam turns ~0 grads into +-lr


def test_oracle_eval_conversion_matches_reference():
    g = load_golden("eval_32_32_b2_t96")
    dim_neck, freq, B, T, n_bins, wseed, iseed = g["meta"].tolist()
    sd = seeded_state_dict(dim_neck, freq, n_bins, wseed)
    x, e, e2 = synth_inputs(B, T, n_bins, 256, iseed)
    with torch.no_grad():
        gref.generator_forward(sd, x, e, e, dim_neck, freq, training=True)
        gref.generator_forward(sd, x.flip(0), e2, e, dim_neck, freq, training=True)
        xi, xp, codes = gref.generator_forward(sd, x, e, e2, dim_neck, freq, training=False)
    for k in g.files:
        if k.startswith("buf/"):
            np.testing.assert_allclose(sd[k[4:]].numpy(), g[k], rtol=1e-5, atol=1e-6, err_msg=k)
    assert np.abs(xi.numpy() - g["x_identic"]).max() < 2e-5
    assert np.abs(xp.numpy() - g["x_identic_psnt"]).max() < 2e-5
    assert np.abs(codes.numpy() - g["codes"]).max() < 2e-5


def test_fast_module_equals_functional_oracle():
    """The nn.Module used for the timed CPU baseline computes the same thing as the
    functional oracle."""
    torch.manual_seed(0)
    G = gref.build_reference_like_module(16, 256, 512, 16)
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    x, e, _ = synth_inputs(2, 32, 80, 256, 5)
    G.train()
    a = G(x, e, e)
    with torch.no_grad():
        b = gref.generator_forward(sd, x, e, e, 16, 16, training=True)
    for u, v in zip(a, b):
        assert (u.detach() - v).abs().max() < 2e-5
