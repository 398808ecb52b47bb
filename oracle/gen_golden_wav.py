"""Golden fixtures for the waveform variant from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden_wav.py

Imports ``/root/reference/model_vc_wav.py`` as it is (it pulls Encoder / Decoder / ConvNorm /
LinearNorm from the reference's own model_vc_mel.py), seeds torch, runs the 'wav' branch of the
Solver step (solver_encoder.py:264-290: four loss terms incl. SI-SNR, then :293-300 zero_grad /
backward / Adam) on CPU in fp32 and stores outputs, losses, gradient digests, BN buffers and
post-Adam parameter digests under tests/golden/wav_*.npz.  Weights are reproduced from the seed.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle.gen_golden import digest  # noqa: E402
from oracle.generator_wav_ref import synth_wav_inputs  # noqa: E402


def ref_wav_step(G, x_real, emb_org, lambda_cd=1.0, lambda_sisnr=1.0):
    """solver_encoder.py:264-290, verbatim arithmetic."""
    x_convtas, x_identic, gen_outputs, code_real = G(x_real, emb_org, emb_org)
    g_loss_id = F.mse_loss(x_real.squeeze(), x_identic.squeeze())
    g_loss_gen = F.mse_loss(x_convtas.squeeze(), gen_outputs.squeeze())
    code_reconst = G(x_identic, emb_org, None)
    g_loss_cd = F.l1_loss(code_real, code_reconst)
    dot = torch.sum(x_identic * x_real, dim=1, keepdim=True)
    s_target_energy = torch.sum(x_real ** 2, dim=1, keepdim=True)
    scaled_target = dot * x_real / s_target_energy
    e_noise = x_identic - scaled_target
    losses = torch.sum(scaled_target ** 2, dim=1) / (torch.sum(e_noise ** 2, dim=1))
    losses = (10 * torch.log10(losses))
    g_loss_sisnr = -(losses.mean())
    g_loss = g_loss_id + lambda_sisnr * g_loss_sisnr + g_loss_gen + lambda_cd * g_loss_cd
    return g_loss, (g_loss_id, g_loss_gen, g_loss_cd, g_loss_sisnr), (x_convtas, x_identic, gen_outputs, code_real, code_reconst)


def make_wav_golden(name, dim_neck, freq, depth, B, L=33536, wseed=0, iseed=4321, steps=2):
    from model_vc_wav import GeneratorWav
    torch.manual_seed(wseed)
    G = GeneratorWav(dim_neck, 256, 512, freq, depth)
    G.train()
    x, e = synth_wav_inputs(B, L, 256, iseed)
    out = {"meta": np.array([dim_neck, freq, depth, B, L, wseed, iseed, steps], np.int64)}
    out["param_names"] = np.array([k for k, _ in G.named_parameters()])
    out["param_digest0"] = np.stack([digest(p) for p in G.parameters()])
    out["state_dict_keys"] = np.array(list(G.state_dict().keys()))
    opt = torch.optim.Adam(G.parameters(), 1e-4)
    for s in range(steps):
        g_loss, ls, outs = ref_wav_step(G, x, e)
        opt.zero_grad()
        g_loss.backward()
        out[f"s{s}_losses"] = np.array([g_loss.item()] + [l.item() for l in ls], np.float64)
        if s == 0:
            for k, v in zip(("x_convtas", "x_identic", "gen_outputs", "code_real", "code_reconst"), outs):
                out["s0_" + k] = v.detach().numpy()
            out["s0_grad_digest"] = np.stack([digest(p.grad) for p in G.parameters()])
            for k, v in G.state_dict().items():
                if "running" in k or "num_batches" in k:
                    out["s0_buf/" + k] = v.detach().numpy().copy()
        opt.step()
        out[f"s{s}_param_digest"] = np.stack([digest(p) for p in G.parameters()])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: out[k] for k in out if k.endswith("losses")})


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 8)
    make_wav_golden("wav_16_16_d1_b2", 16, 16, 1, 2)
    make_wav_golden("wav_32_32_d3_b3", 32, 32, 3, 3)
