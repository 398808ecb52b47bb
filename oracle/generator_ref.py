"""CPU oracle for the AutoVC mel Generator training step.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product path (``autovc_b200``) must never route through anything in ``oracle/``.

It restates, as plain functional PyTorch on CPU tensors over a ``state_dict``, what the
reference computes on this path (paths relative to the upstream repo):

* ``Encoder.forward``      model_vc_mel.py:63-81
* up-sampling + concat     model_vc_mel.py:186-192
* ``Decoder.forward``      model_vc_mel.py:108-122
* ``Postnet.forward``      model_vc_mel.py:163-169
* ``Generator.forward``    model_vc_mel.py:181-203
* the step maths           solver_encoder.py:227-243 (losses), :293-300 (backward + Adam)

The primitives underneath (``F.conv1d``, ``F.batch_norm``, ``torch.nn.LSTM`` semantics,
``F.mse_loss``/``F.l1_loss``) are the same third-party PyTorch ops the reference calls; the
LSTM is written out as an explicit time loop (gate order i,f,g,o, zero initial state) so
the recurrence the CUDA kernels implement is spelled out rather than hidden in a library.

Pinning: ``oracle/gen_golden.py`` runs the *unmodified* reference module from
``/root/reference`` in this container and stores its outputs in ``tests/golden/``;
``tests/test_oracle_generator.py`` checks this restatement against those vectors.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------
# parameter construction with the reference's RNG consumption order
# ----------------------------------------------------------------------------------------
def _xavier_uniform(shape, gain, gen=None):
    # torch.nn.init.xavier_uniform_ (model_vc_mel.py:12-14, :33-34)
    recept = 1
    for s in shape[2:]:
        recept *= s
    fan_in, fan_out = shape[1] * recept, shape[0] * recept
    bound = gain * math.sqrt(6.0 / (fan_in + fan_out))
    return bound


def layer_plan(dim_neck: int, dim_emb: int, dim_pre: int, n_bins: int = 80):
    """Static description of the 14 conv layers / 4 LSTMs / 1 linear of the Generator.

    Returns a dict used by the oracle and by tests to enumerate layers in registration
    order (model_vc_mel.py:48-61, :90-106, :130-161, :177-179)."""
    enc = [(n_bins + dim_emb, 512), (512, 512), (512, 512)]
    dec = [(dim_pre, dim_pre)] * 3
    post = [(n_bins, 512), (512, 512), (512, 512), (512, 512), (512, n_bins)]
    return {"enc_convs": enc, "dec_convs": dec, "post_convs": post,
            "enc_lstm": (512, dim_neck, 2, True),
            "lstm1": (2 * dim_neck + dim_emb, dim_pre, 1, False),
            "lstm2": (dim_pre, 1024, 2, False),
            "linear": (1024, n_bins)}


# ----------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------
def conv_bn_act(sd: Dict[str, Tensor], prefix: str, x: Tensor, act: str, training: bool,
                update_buffers: bool = True) -> Tensor:
    """``act(BatchNorm1d(Conv1d_k5_p2(x)))`` on channel-first ``x`` (B,C,T).

    prefix e.g. ``encoder.convolutions.0`` -> keys ``.0.conv.{weight,bias}`` (ConvNorm,
    model_vc_mel.py:20-38) and ``.1.{weight,bias,running_mean,running_var,
    num_batches_tracked}`` (nn.BatchNorm1d, model_vc_mel.py:57)."""
    w, b = sd[prefix + ".0.conv.weight"], sd[prefix + ".0.conv.bias"]
    y = F.conv1d(x, w, b, stride=1, padding=(w.shape[-1] - 1) // 2)
    rm, rv = sd[prefix + ".1.running_mean"], sd[prefix + ".1.running_var"]
    if training and not update_buffers:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(y, rm, rv, sd[prefix + ".1.weight"], sd[prefix + ".1.bias"],
                     training=training, momentum=BN_MOMENTUM, eps=BN_EPS)
    if training and update_buffers:
        sd[prefix + ".1.num_batches_tracked"] += 1
    if act == "relu":
        return F.relu(y)
    if act == "tanh":
        return torch.tanh(y)
    return y


def lstm_layer(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor,
               reverse: bool = False) -> Tensor:
    """One LSTM layer-direction, batch-first ``x`` (B,T,I) -> (B,T,H); h0 = c0 = 0.

    gates = x W_ih^T + b_ih + h W_hh^T + b_hh, row blocks i,f,g,o; c' = f c + i g;
    h' = o tanh(c')  (torch.nn.LSTM semantics used at model_vc_mel.py:61,:90,:104)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    pre = x @ w_ih.t() + (b_ih + b_hh)
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out: List[Optional[Tensor]] = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = pre[:, t] + h @ w_hh.t()
        i, f, gg, o = g.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return torch.stack(out, dim=1)


def lstm(sd: Dict[str, Tensor], prefix: str, x: Tensor, num_layers: int, bidirectional: bool) -> Tensor:
    for l in range(num_layers):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)) if bidirectional else (("", False),):
            outs.append(lstm_layer(x,
                                   sd[f"{prefix}.weight_ih_l{l}{suffix}"], sd[f"{prefix}.weight_hh_l{l}{suffix}"],
                                   sd[f"{prefix}.bias_ih_l{l}{suffix}"], sd[f"{prefix}.bias_hh_l{l}{suffix}"],
                                   reverse=rev))
        x = torch.cat(outs, dim=-1) if bidirectional else outs[0]
    return x


# ----------------------------------------------------------------------------------------
# modules
# ----------------------------------------------------------------------------------------
def encoder_forward(sd, x: Tensor, c_org: Tensor, dim_neck: int, freq: int, training: bool,
                    update_buffers: bool = True) -> List[Tensor]:
    """model_vc_mel.py:63-81.  ``x`` (B,T,n_bins) or (B,1,T,n_bins)."""
    x = x.squeeze(1).transpose(2, 1)                                   # :64
    c = c_org.unsqueeze(-1).expand(-1, -1, x.size(-1))                 # :65
    x = torch.cat((x, c), dim=1)                                       # :66
    for i in range(3):                                                 # :68-69
        x = conv_bn_act(sd, f"encoder.convolutions.{i}", x, "relu", training, update_buffers)
    x = x.transpose(1, 2)                                              # :70
    out = lstm(sd, "encoder.lstm", x, 2, True)                         # :73
    fwd, bwd = out[:, :, :dim_neck], out[:, :, dim_neck:]              # :74-75
    codes = []
    for i in range(0, out.size(1), freq):                              # :77-79
        codes.append(torch.cat((fwd[:, i + freq - 1, :], bwd[:, i, :]), dim=-1))
    return codes


def decoder_forward(sd, enc_out: Tensor, training: bool, update_buffers: bool = True) -> Tensor:
    """model_vc_mel.py:108-122."""
    x = lstm(sd, "decoder.lstm1", enc_out, 1, False)                   # :111
    x = x.transpose(1, 2)
    for i in range(3):                                                 # :114-115
        x = conv_bn_act(sd, f"decoder.convolutions.{i}", x, "relu", training, update_buffers)
    x = x.transpose(1, 2)
    x = lstm(sd, "decoder.lstm2", x, 2, False)                         # :118
    return F.linear(x, sd["decoder.linear_projection.linear_layer.weight"],
                    sd["decoder.linear_projection.linear_layer.bias"])  # :120


def postnet_forward(sd, x: Tensor, training: bool, update_buffers: bool = True) -> Tensor:
    """model_vc_mel.py:163-169; ``x`` channel-first (B,n_bins,T)."""
    for i in range(4):
        x = conv_bn_act(sd, f"postnet.convolutions.{i}", x, "tanh", training, update_buffers)
    return conv_bn_act(sd, "postnet.convolutions.4", x, "none", training, update_buffers)


def generator_forward(sd, x: Tensor, c_org: Tensor, c_trg: Optional[Tensor], dim_neck: int, freq: int,
                      training: bool = True, update_buffers: bool = True):
    """model_vc_mel.py:181-203."""
    codes = encoder_forward(sd, x, c_org, dim_neck, freq, training, update_buffers)
    if c_trg is None:                                                  # :183-184
        return torch.cat(codes, dim=-1)
    T = x.size(1)
    rep = T // len(codes)
    code_exp = torch.cat([c.unsqueeze(1).expand(-1, rep, -1) for c in codes], dim=1)   # :186-190
    enc_out = torch.cat((code_exp, c_trg.unsqueeze(1).expand(-1, T, -1)), dim=-1)      # :192
    x_identic = decoder_forward(sd, enc_out, training, update_buffers)                 # :194
    post = postnet_forward(sd, x_identic.transpose(2, 1), training, update_buffers)    # :196
    x_identic_psnt = x_identic + post.transpose(2, 1)                                  # :197
    return x_identic.unsqueeze(1), x_identic_psnt.unsqueeze(1), torch.cat(codes, dim=-1)


def train_losses(sd, x_real: Tensor, emb_org: Tensor, dim_neck: int, freq: int, lambda_cd: float = 1.0,
                 update_buffers: bool = True):
    """solver_encoder.py:227-243 — returns (g_loss, L_id, L_id_psnt, L_cd, outputs dict)."""
    x_identic, x_identic_psnt, code_real = generator_forward(sd, x_real, emb_org, emb_org, dim_neck, freq,
                                                             True, update_buffers)               # :228
    l_id = F.mse_loss(x_real.squeeze(), x_identic.squeeze())                                     # :230
    l_id_psnt = F.mse_loss(x_real, x_identic_psnt.squeeze(1))                                    # :233
    code_reconst = generator_forward(sd, x_identic_psnt, emb_org, None, dim_neck, freq,
                                     True, update_buffers)                                        # :235
    l_cd = F.l1_loss(code_real, code_reconst)                                                    # :236
    g_loss = l_id + l_id_psnt + lambda_cd * l_cd                                                 # :243
    outs = {"x_identic": x_identic, "x_identic_psnt": x_identic_psnt, "code_real": code_real,
            "code_reconst": code_reconst}
    return g_loss, l_id, l_id_psnt, l_cd, outs


PARAM_SUFFIXES = ("weight", "bias")


def split_state_dict(sd: Dict[str, Tensor]) -> Tuple["OrderedDict[str, Tensor]", "OrderedDict[str, Tensor]"]:
    """Split a Generator state_dict into (parameters, buffers) preserving order."""
    params, bufs = OrderedDict(), OrderedDict()
    for k, v in sd.items():
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            bufs[k] = v
        else:
            params[k] = v
    return params, bufs


def train_step(sd: Dict[str, Tensor], x_real: Tensor, emb_org: Tensor, dim_neck: int, freq: int,
               lambda_cd: float = 1.0, adam_state: Optional[dict] = None, lr: float = 1e-4):
    """One full step: losses (solver :227-243), backward (:293-294), optional Adam (:300,
    built at :130 with default betas (0.9, 0.999), eps 1e-8).  ``sd`` is modified in place
    (BN buffers; parameters too when ``adam_state`` is given).  Returns losses, outputs, grads."""
    params, _ = split_state_dict(sd)
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    work = dict(sd)
    work.update(leaves)
    g_loss, l_id, l_id_psnt, l_cd, outs = train_losses(work, x_real, emb_org, dim_neck, freq, lambda_cd)
    grads = torch.autograd.grad(g_loss, list(leaves.values()))
    grads = OrderedDict(zip(leaves.keys(), grads))
    for k in sd:                       # propagate buffer updates done on `work`
        if k not in leaves:
            sd[k] = work[k]
    if adam_state is not None:
        adam_state["t"] = adam_state.get("t", 0) + 1
        t = adam_state["t"]
        b1, b2, eps = 0.9, 0.999, 1e-8
        for k, g in grads.items():
            m = adam_state.setdefault("m", {}).setdefault(k, torch.zeros_like(g))
            v = adam_state.setdefault("v", {}).setdefault(k, torch.zeros_like(g))
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
            sd[k] = sd[k] - (lr / (1 - b1 ** t)) * (m / denom)
    losses = {"g_loss": g_loss.detach(), "L_id": l_id.detach(), "L_id_psnt": l_id_psnt.detach(),
              "L_cd": l_cd.detach()}
    return losses, {k: v.detach() for k, v in outs.items()}, grads


# ----------------------------------------------------------------------------------------
# fast CPU reference used only for the *timed* cpu_baseline / --impl reference legs:
# identical maths, but the recurrences go through torch.nn.LSTM (oneDNN) exactly like the
# reference's own module does (model_vc_mel.py:61,:90,:104), so the timing reflects the
# reference's CPU path rather than a Python time loop.
# ----------------------------------------------------------------------------------------
def build_reference_like_module(dim_neck: int, dim_emb: int, dim_pre: int, freq: int, n_bins: int = 80,
                                with_postnet: bool = True):
    """An nn.Module tree with the reference's registration order, parameter names and
    initialisers (ConvNorm xavier gains model_vc_mel.py:33-34; LinearNorm :12-14), built
    from torch.nn parts.  ``torch.manual_seed(s)`` before the call reproduces the reference
    init bit-for-bit (checked in tests against the golden state_dict)."""
    import torch.nn as nn

    class _ConvNorm(nn.Module):
        def __init__(self, cin, cout, gain):
            super().__init__()
            self.conv = nn.Conv1d(cin, cout, kernel_size=5, stride=1, padding=2)
            nn.init.xavier_uniform_(self.conv.weight, gain=nn.init.calculate_gain(gain))

        def forward(self, x):
            return self.conv(x)

    class _LinearNorm(nn.Module):
        def __init__(self, i, o):
            super().__init__()
            self.linear_layer = nn.Linear(i, o)
            nn.init.xavier_uniform_(self.linear_layer.weight, gain=nn.init.calculate_gain("linear"))

        def forward(self, x):
            return self.linear_layer(x)

    def stack(dims, gains):
        return nn.ModuleList([nn.Sequential(_ConvNorm(ci, co, g), nn.BatchNorm1d(co))
                              for (ci, co), g in zip(dims, gains)])

    plan = layer_plan(dim_neck, dim_emb, dim_pre, n_bins)

    class _Enc(nn.Module):
        def __init__(self):
            super().__init__()
            self.convolutions = stack(plan["enc_convs"], ["relu"] * 3)
            self.lstm = nn.LSTM(512, dim_neck, 2, batch_first=True, bidirectional=True)

    class _Dec(nn.Module):
        def __init__(self):
            super().__init__()
            self.lstm1 = nn.LSTM(2 * dim_neck + dim_emb, dim_pre, 1, batch_first=True)
            self.convolutions = stack(plan["dec_convs"], ["relu"] * 3)
            self.lstm2 = nn.LSTM(dim_pre, 1024, 2, batch_first=True)
            self.linear_projection = _LinearNorm(1024, n_bins)

    class _Post(nn.Module):
        def __init__(self):
            super().__init__()
            self.convolutions = stack(plan["post_convs"], ["tanh"] * 4 + ["linear"])

    class _Gen(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder, self.decoder = _Enc(), _Dec()
            self.postnet = _Post() if with_postnet else None      # GeneratorWav builds none (model_vc_wav.py:66-73)
            self.dim_neck, self.freq = dim_neck, freq

        def _enc(self, x, c):
            x = torch.cat((x.squeeze(1).transpose(2, 1), c.unsqueeze(-1).expand(-1, -1, x.squeeze(1).size(1))), 1)
            for conv in self.encoder.convolutions:
                x = F.relu(conv(x))
            o, _ = self.encoder.lstm(x.transpose(1, 2))
            n, f = self.dim_neck, self.freq
            return [torch.cat((o[:, i + f - 1, :n], o[:, i, n:]), -1) for i in range(0, o.size(1), f)]

        def forward(self, x, c_org, c_trg):
            codes = self._enc(x, c_org)
            if c_trg is None:
                return torch.cat(codes, -1)
            T = x.size(1)
            up = torch.cat([c.unsqueeze(1).expand(-1, T // len(codes), -1) for c in codes], 1)
            h, _ = self.decoder.lstm1(torch.cat((up, c_trg.unsqueeze(1).expand(-1, T, -1)), -1))
            h = h.transpose(1, 2)
            for conv in self.decoder.convolutions:
                h = F.relu(conv(h))
            h, _ = self.decoder.lstm2(h.transpose(1, 2))
            xi = self.decoder.linear_projection(h)
            p = xi.transpose(2, 1)
            for i, conv in enumerate(self.postnet.convolutions):
                p = conv(p) if i == 4 else torch.tanh(conv(p))
            xp = xi + p.transpose(2, 1)
            return xi.unsqueeze(1), xp.unsqueeze(1), torch.cat(codes, -1)

    return _Gen()


def module_train_step(G, opt, x_real, emb_org, lambda_cd: float = 1.0):
    """solver_encoder.py:228-243, :293-300 on a module + optimizer (timed CPU baseline)."""
    x_identic, x_identic_psnt, code_real = G(x_real, emb_org, emb_org)
    l_id = F.mse_loss(x_real.squeeze(), x_identic.squeeze())
    l_id_psnt = F.mse_loss(x_real, x_identic_psnt.squeeze(1))
    code_reconst = G(x_identic_psnt, emb_org, None)
    l_cd = F.l1_loss(code_real, code_reconst)
    g_loss = l_id + l_id_psnt + lambda_cd * l_cd
    opt.zero_grad()
    g_loss.backward()
    opt.step()
    return l_id.item(), l_id_psnt.item(), l_cd.item()
