"""CPU oracle for the waveform variant (GeneratorWav).  TEST INFRASTRUCTURE ONLY.

The checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it.

Functional PyTorch-CPU restatement over a ``state_dict`` of (paths relative to the upstream repo)

* ``ConvTasNetEncoder.forward``   model_vc_wav.py:29-33  (Conv1d(1->512, k=1024, s=256), then depth x [Conv1d k3 -> PReLU -> BN])
* ``ConvTasNetDecoder.forward``   model_vc_wav.py:54-58  (depth x [ConvTranspose1d k3 -> PReLU -> BN], ConvTranspose1d(512->1, k=1024, s=256))
* ``GeneratorWav.forward``        model_vc_wav.py:74-102
* the 'wav' branch of the step    solver_encoder.py:264-290 (four loss terms incl. SI-SNR), :293-300

on top of ``oracle.generator_ref`` (Encoder / Decoder restatements, explicit LSTM time loop).

Pinning: ``oracle/gen_golden_wav.py`` runs the *unmodified* ``/root/reference/model_vc_wav.py`` in the build container and
stores its outputs in ``tests/golden/wav_*.npz``; ``tests/test_oracle_wav.py`` checks this restatement against them.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import generator_ref as gref

Tensor = torch.Tensor
N_FILT, K_FILT, S_FILT = 512, 1024, 256          # model_vc_wav.py:14-16


def synth_wav_inputs(B, L, dim_emb, seed):
    """Waveform crops shaped like the loader's (B, L, 1): band-limited noise with a slow envelope, |x| < 1."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(B, L, generator=g)
    w = 0.1 * (w + torch.roll(w, 1, 1) + torch.roll(w, 2, 1)) / 3 ** 0.5
    env = 0.5 + 0.5 * torch.sin(torch.linspace(0, 6.0, L))[None, :] * torch.rand(B, 1, generator=g)
    x = (w * env).clamp(-1, 1).unsqueeze(-1)
    e = F.normalize(torch.randn(B, dim_emb, generator=g), dim=-1) * 0.8
    return x, e


def build_wav_module(dim_neck: int, dim_emb: int, dim_pre: int, freq: int, depth: int):
    """torch.nn parameter tree with the reference's registration order and initialisers (model_vc_wav.py:62-73:
    tasEncoder, Encoder, Decoder, replacement encoder conv0 (ConvNorm default gain 'linear'), replacement
    linear_projection, tasDecoder).  ``torch.manual_seed(s)`` before the call reproduces the reference init bit for bit."""
    import torch.nn as nn

    class _TasEnc(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1x1 = nn.Conv1d(1, N_FILT, kernel_size=K_FILT, stride=S_FILT, padding=0)
            self.convD = nn.ModuleList([nn.Sequential(nn.Conv1d(N_FILT, N_FILT, 3, 1, 1), nn.PReLU(), nn.BatchNorm1d(N_FILT))
                                        for _ in range(depth)])

    class _TasDec(nn.Module):
        def __init__(self):
            super().__init__()
            self.convTD = nn.ModuleList([nn.Sequential(nn.ConvTranspose1d(N_FILT, N_FILT, 3, 1, 1), nn.PReLU(),
                                                       nn.BatchNorm1d(N_FILT)) for _ in range(depth)])
            self.convT1x1 = nn.ConvTranspose1d(N_FILT, 1, kernel_size=K_FILT, stride=S_FILT, padding=0)

    class _Wav(nn.Module):
        def __init__(self):
            super().__init__()
            self.tasEncoder = _TasEnc()
            core = gref.build_reference_like_module(dim_neck, dim_emb, dim_pre, freq, with_postnet=False)
            self.encoder, self.decoder = core.encoder, core.decoder
            conv0 = type(self.encoder.convolutions[0][0])(N_FILT + dim_emb, 512, "linear")
            self.encoder.convolutions[0][0] = conv0
            self.decoder.linear_projection = type(self.decoder.linear_projection)(1024, N_FILT)
            self.tasDecoder = _TasDec()

    return _Wav()


def _conv_prelu_bn(sd, prefix: str, x: Tensor, transposed: bool, training: bool) -> Tensor:
    """nn.Sequential(Conv1d | ConvTranspose1d (k3, s1, p1), PReLU, BatchNorm1d) on channel-first x (model_vc_wav.py:22-25, :44-47)."""
    w, b = sd[prefix + ".0.weight"], sd[prefix + ".0.bias"]
    y = F.conv_transpose1d(x, w, b, stride=1, padding=1) if transposed else F.conv1d(x, w, b, stride=1, padding=1)
    y = F.prelu(y, sd[prefix + ".1.weight"])
    y = F.batch_norm(y, sd[prefix + ".2.running_mean"], sd[prefix + ".2.running_var"], sd[prefix + ".2.weight"],
                     sd[prefix + ".2.bias"], training=training, momentum=gref.BN_MOMENTUM, eps=gref.BN_EPS)
    if training:
        sd[prefix + ".2.num_batches_tracked"] += 1
    return y


def _depth(sd, stem: str) -> int:
    d = 0
    while f"{stem}.{d}.0.weight" in sd:
        d += 1
    return d


def tas_encoder_forward(sd, x: Tensor, training: bool) -> Tensor:
    """(B, 1, L) -> (B, 512, T)."""
    x = F.conv1d(x, sd["tasEncoder.conv1x1.weight"], sd["tasEncoder.conv1x1.bias"], stride=S_FILT)
    for i in range(_depth(sd, "tasEncoder.convD")):
        x = _conv_prelu_bn(sd, f"tasEncoder.convD.{i}", x, False, training)
    return x


def tas_decoder_forward(sd, x: Tensor, training: bool) -> Tensor:
    """(B, 512, T) -> (B, 1, L)."""
    for i in range(_depth(sd, "tasDecoder.convTD")):
        x = _conv_prelu_bn(sd, f"tasDecoder.convTD.{i}", x, True, training)
    return F.conv_transpose1d(x, sd["tasDecoder.convT1x1.weight"], sd["tasDecoder.convT1x1.bias"], stride=S_FILT)


def generator_wav_forward(sd, x: Tensor, c_org: Tensor, c_trg: Optional[Tensor], dim_neck: int, freq: int,
                          training: bool = True):
    """model_vc_wav.py:74-102; x (B, L, 1)."""
    x = tas_encoder_forward(sd, x.permute(0, 2, 1), training)                   # :80-81
    x_ct = x.clone()                                                            # :82
    x = x.permute(0, 2, 1)                                                      # :85
    codes = gref.encoder_forward(sd, x, c_org, dim_neck, freq, training)        # :86
    if c_trg is None:
        return torch.cat(codes, dim=-1)                                         # :88-89
    T = x.size(1)
    code_exp = torch.cat([c.unsqueeze(1).expand(-1, T // len(codes), -1) for c in codes], dim=1)   # :90-93
    enc_out = torch.cat((code_exp, c_trg.unsqueeze(1).expand(-1, T, -1)), dim=-1)                  # :95
    x_dec = gref.decoder_forward(sd, enc_out, training).permute(0, 2, 1)        # :96
    x_identic = tas_decoder_forward(sd, x_dec, training).permute(0, 2, 1)       # :99
    return x_ct, x_identic, x_dec, torch.cat(codes, dim=-1)                     # :100-101


def wav_losses(sd, x_real: Tensor, emb_org: Tensor, dim_neck: int, freq: int, lambda_cd: float = 1.0,
               lambda_sisnr: float = 1.0):
    """solver_encoder.py:264-290."""
    x_convtas, x_identic, gen_outputs, code_real = generator_wav_forward(sd, x_real, emb_org, emb_org, dim_neck, freq)
    l_id = F.mse_loss(x_real.squeeze(), x_identic.squeeze())
    l_gen = F.mse_loss(x_convtas.squeeze(), gen_outputs.squeeze())
    code_reconst = generator_wav_forward(sd, x_identic, emb_org, None, dim_neck, freq)
    l_cd = F.l1_loss(code_real, code_reconst)
    dot = torch.sum(x_identic * x_real, dim=1, keepdim=True)
    s_target_energy = torch.sum(x_real ** 2, dim=1, keepdim=True)
    scaled_target = dot * x_real / s_target_energy
    e_noise = x_identic - scaled_target
    losses = torch.sum(scaled_target ** 2, dim=1) / (torch.sum(e_noise ** 2, dim=1))
    l_sisnr = -((10 * torch.log10(losses)).mean())
    g_loss = l_id + lambda_sisnr * l_sisnr + l_gen + lambda_cd * l_cd
    outs = {"x_convtas": x_convtas, "x_identic": x_identic, "gen_outputs": gen_outputs, "code_real": code_real,
            "code_reconst": code_reconst}
    return g_loss, (l_id, l_gen, l_cd, l_sisnr), outs


def wav_train_step(sd: Dict[str, Tensor], x_real: Tensor, emb_org: Tensor, dim_neck: int, freq: int):
    """Losses + backward (solver_encoder.py:264-294).  ``sd`` is modified in place (BN buffers).  Returns
    (losses dict, outputs dict, grads OrderedDict keyed like the state_dict's parameters)."""
    params, _ = gref.split_state_dict(sd)
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    work = dict(sd)
    work.update(leaves)
    g_loss, (l_id, l_gen, l_cd, l_sisnr), outs = wav_losses(work, x_real, emb_org, dim_neck, freq)
    grads = OrderedDict(zip(leaves.keys(), torch.autograd.grad(g_loss, list(leaves.values()))))
    for k in sd:
        if k not in leaves:
            sd[k] = work[k]
    losses = {"g_loss": g_loss.detach(), "L_id": l_id.detach(), "L_gen": l_gen.detach(), "L_cd": l_cd.detach(),
              "L_SISNR": l_sisnr.detach()}
    return losses, {k: v.detach() for k, v in outs.items()}, grads
