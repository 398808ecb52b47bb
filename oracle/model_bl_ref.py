"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of model_bl.D_VECTOR.forward (model_bl.py:13-20) over a state_dict, on the
explicit-time-loop LSTM of oracle/generator_ref.py.  Pinned by oracle/gen_golden_dvector.py, which runs the UNMODIFIED
reference module (same seed -> same init) and stores its outputs in tests/golden/dvector.npz."""
import torch

from oracle import generator_ref as gref


def dvector_forward(sd, x, num_layers=3):
    out = gref.lstm(sd, "lstm", x, num_layers, False)                    # :15
    emb = out[:, -1, :] @ sd["embedding.weight"].t() + sd["embedding.bias"]   # :16
    return emb / emb.norm(p=2, dim=-1, keepdim=True)                     # :17-19


def synth_mels(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, T, 80, generator=g)
