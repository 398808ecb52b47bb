"""Generate the committed golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python oracle/gen_golden.py

Generator goldens: imports ``/root/reference/model_vc_mel.py`` (and ``model_vc_stft.py``)
as they are, seeds torch, runs the reference module / the step maths of
solver_encoder.py:227-243,:293-300 on CPU in fp32 (and an fp64 copy as "truth"), and stores
inputs' seeds, outputs, losses, gradient digests, updated BN buffers and post-Adam
parameter digests.  Weights are NOT stored (113 MB): they are reproduced from
``torch.manual_seed(seed)`` + the same module registration order, and the fixture carries
per-tensor digests so a test can prove the reproduction is exact.

Front-end goldens: a subset of the bundled ``wavs/<spk>/<utt>.wav`` ->
``spmel/<spk>/<utt>.npy`` pairs, each with the int16 samples, the speaker seed and the
offset of the utterance in the speaker's dither stream (make_spect.py:68,:76).
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def digest(t: torch.Tensor, n: int = 48):
    """Small fingerprint of a tensor: [sum, abs-sum, l2, first n, n strided samples]."""
    f = t.detach().double().flatten()
    step = max(1, f.numel() // n)
    head = f[:n]
    strided = f[::step][:n]
    pad = lambda v: torch.cat([v, v.new_zeros(n - v.numel())])
    return torch.cat([torch.stack([f.sum(), f.abs().sum(), f.norm()]), pad(head), pad(strided)]).numpy()


def synth_inputs(B, T, n_bins, dim_emb, seed):
    """mel uniform [0,1); embeddings = 0.8 * unit-normalised gaussian."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, T, n_bins, generator=g)
    e = F.normalize(torch.randn(B, dim_emb, generator=g), dim=-1) * 0.8
    e2 = F.normalize(torch.randn(B, dim_emb, generator=g), dim=-1) * 0.8
    return x, e, e2


def ref_step(G, x_real, emb_org, lambda_cd=1.0):
    x_identic, x_identic_psnt, code_real = G(x_real, emb_org, emb_org)
    l_id = F.mse_loss(x_real.squeeze(), x_identic.squeeze())
    l_id_psnt = F.mse_loss(x_real, x_identic_psnt.squeeze())
    code_reconst = G(x_identic_psnt, emb_org, None)
    l_cd = F.l1_loss(code_real, code_reconst)
    g_loss = l_id + l_id_psnt + lambda_cd * l_cd
    return g_loss, (l_id, l_id_psnt, l_cd), (x_identic, x_identic_psnt, code_real, code_reconst)


def make_train_golden(name, dim_neck, freq, B, T, n_bins=80, wseed=0, iseed=1234, stft=False, steps=2):
    if stft:
        from model_vc_stft import GeneratorSTFT
        torch.manual_seed(wseed)
        G = GeneratorSTFT(dim_neck, 256, 512, freq).model
    else:
        from model_vc_mel import Generator
        torch.manual_seed(wseed)
        G = Generator(dim_neck, 256, 512, freq)
    G.train()
    x, e, _ = synth_inputs(B, T, n_bins, 256, iseed)
    out = {"meta": np.array([dim_neck, freq, B, T, n_bins, wseed, iseed, steps], np.int64)}
    names = [k for k, _ in G.named_parameters()]
    out["param_names"] = np.array(names)
    out["param_digest0"] = np.stack([digest(p) for p in G.parameters()])
    sd_names = list(G.state_dict().keys())
    out["state_dict_keys"] = np.array(sd_names)

    # fp64 truth for the first step (error budgeting)
    G64 = copy.deepcopy(G).double()
    g64, l64, o64 = ref_step(G64, x.double(), e.double())
    out["f64_losses"] = np.array([g64.item()] + [l.item() for l in l64])
    out["f64_x_identic_psnt"] = o64[1].detach().numpy()
    out["f64_code_real"] = o64[2].detach().numpy()
    out["f64_code_reconst"] = o64[3].detach().numpy()

    opt = torch.optim.Adam(G.parameters(), 1e-4)                    # solver_encoder.py:130
    for s in range(steps):
        g_loss, ls, outs = ref_step(G, x, e)
        opt.zero_grad()
        g_loss.backward()
        out[f"s{s}_losses"] = np.array([g_loss.item()] + [l.item() for l in ls], np.float64)
        if s == 0:
            out["s0_x_identic"] = outs[0].detach().numpy()
            out["s0_x_identic_psnt"] = outs[1].detach().numpy()
            out["s0_code_real"] = outs[2].detach().numpy()
            out["s0_code_reconst"] = outs[3].detach().numpy()
            out["s0_grad_digest"] = np.stack([digest(p.grad) for p in G.parameters()])
            bufs = {k: v for k, v in G.state_dict().items() if "running" in k or "num_batches" in k}
            for k, v in bufs.items():
                out["s0_buf/" + k] = v.detach().numpy().copy()
        opt.step()
        out[f"s{s}_param_digest"] = np.stack([digest(p) for p in G.parameters()])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: out[k] for k in out if k.endswith("losses")})


def make_eval_golden(name, dim_neck, freq, B, T, wseed=0, iseed=77):
    """Conversion maths (conversion.py:47,:91-92): eval-mode forward with c_trg != c_org.
    BN running stats are made non-trivial first by two train-mode forwards."""
    from model_vc_mel import Generator
    torch.manual_seed(wseed)
    G = Generator(dim_neck, 256, 512, freq)
    x, e, e2 = synth_inputs(B, T, 80, 256, iseed)
    G.train()
    with torch.no_grad():
        G(x, e, e)
        G(x.flip(0), e2, e)
    G.eval()
    with torch.no_grad():
        xi, xp, codes = G(x, e, e2)
    out = {"meta": np.array([dim_neck, freq, B, T, 80, wseed, iseed], np.int64),
           "x_identic": xi.numpy(), "x_identic_psnt": xp.numpy(), "codes": codes.numpy()}
    for k, v in G.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["buf/" + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, xi.shape, float(xp.abs().mean()))


def make_frontend_goldens(max_samples=100000):
    from scipy.io import wavfile
    import warnings
    picks = []
    for spk in sorted(os.listdir(os.path.join(REF, "wavs"))):
        files = sorted(os.listdir(os.path.join(REF, "wavs", spk)))
        offset = 0
        for fn in files:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                fs, w = wavfile.read(os.path.join(REF, "wavs", spk, fn))
            assert fs == 16000 and w.dtype == np.int16
            npy = os.path.join(REF, "spmel", spk, fn[:-4] + ".npy")
            if os.path.exists(npy) and len(w) <= max_samples:
                picks.append((spk, fn, offset, w, np.load(npy)))
            offset += len(w)
    want = {"p003/p003_020.wav", "p001/p001_020.wav", "p002/p002_020.wav", "p225/p225_003.wav", "p225/p225_022.wav"}
    chosen = [p for p in picks if f"{p[0]}/{p[1]}" in want]
    out = {"names": np.array([f"{p[0]}/{p[1]}" for p in chosen]),
           "offsets": np.array([p[2] for p in chosen], np.int64)}
    for i, p in enumerate(chosen):
        out[f"wav{i}"] = p[3]
        out[f"spmel{i}"] = p[4]
    np.savez_compressed(os.path.join(OUT, "frontend_bundled.npz"), **out)
    print("frontend:", [(f"{p[0]}/{p[1]}", len(p[3]), p[2], p[4].shape) for p in chosen])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    make_train_golden("train_16_16_b2_t128", 16, 16, 2, 128)
    make_train_golden("train_32_32_b3_t64", 32, 32, 3, 64, steps=1)
    make_train_golden("train_16_16_b16_t128", 16, 16, 16, 128, steps=1)   # bf16-mode gate (rel-L2) at a less noisy batch
    make_train_golden("train_stft_16_16_b2_t32", 16, 16, 2, 32, n_bins=513, stft=True, steps=1)
    make_eval_golden("eval_32_32_b2_t96", 32, 32, 2, 96)
    make_frontend_goldens()
