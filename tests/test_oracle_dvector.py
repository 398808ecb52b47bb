"""The D_VECTOR restatement (oracle/model_bl_ref.py) against the outputs of the UNMODIFIED reference module
(tests/golden/dvector.npz, oracle/gen_golden_dvector.py): same seed -> same init -> outputs within fp32 rounding."""
import numpy as np
import torch

from oracle import model_bl_ref as bref
from tests.helpers import load_golden


def _seeded_state_dict():
    torch.manual_seed(0)      # registration order of model_bl.py:8-11: lstm (3 layers), then embedding
    lstm = torch.nn.LSTM(input_size=80, hidden_size=768, num_layers=3, batch_first=True)
    emb = torch.nn.Linear(768, 256)
    sd = {"lstm." + k: v.detach() for k, v in lstm.state_dict().items()}
    sd.update({"embedding." + k: v.detach() for k, v in emb.state_dict().items()})
    return sd


def test_restatement_matches_reference_module():
    g = load_golden("dvector")
    sd = _seeded_state_dict()
    assert list(sd.keys()) == g["names"].tolist()
    got = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in sd.values()])
    np.testing.assert_allclose(got, g["param_digest"], rtol=1e-12)          # the init is reproduced from the seed
    B, T, seed = [int(v) for v in g["meta"][:3]]
    with torch.no_grad():
        y = bref.dvector_forward(sd, bref.synth_mels(B, T, seed))
    assert np.abs(y.numpy() - g["y32"]).max() < 2e-6
    assert np.abs(y.numpy() - g["y64"]).max() < 2e-6
    np.testing.assert_allclose(np.linalg.norm(y.numpy(), axis=-1), 1.0, atol=1e-6)
