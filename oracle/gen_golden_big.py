"""Golden fixtures at the BENCHED shapes, generated from the UNMODIFIED reference (test infrastructure).

Run in the build container only (needs /root/reference):

    python oracle/gen_golden_big.py            # ~2 min on 8 cores

BASELINE.json configs[1] (B=256, T=128, 16/16), configs[2] per GPU (B=128, T=256, 32/32) and the config-5 eval
forward at T=640 (10 s utterances: 626 frames padded to a multiple of 32, conversion.py:40-44).  One training
step of the reference module (solver_encoder.py:227-243,:293-300) in fp32 on CPU.  Full outputs at these
sizes are 10 MB each, so the fixture is COMPACT: the content codes in full, the mel outputs for SEL_UTTS
utterances in full plus a digest (3 reductions + 2048 strided samples) over the whole tensor, the losses,
gradient digests of all parameters, BN buffers and the post-Adam parameter digests.  Train-mode BatchNorm
couples every utterance of the batch, so the selected utterances test the whole batch's statistics.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.gen_golden import OUT, digest, ref_step, synth_inputs  # noqa: E402

N_STRIDED = 2048


def sel_utts(B):
    return sorted({0, 1, B // 2 - 1, B // 2, B - 2, B - 1})


def make_train_golden_compact(name, dim_neck, freq, B, T, wseed=0, iseed=1234):
    from model_vc_mel import Generator
    torch.manual_seed(wseed)
    G = Generator(dim_neck, 256, 512, freq).train()
    x, e, _ = synth_inputs(B, T, 80, 256, iseed)
    sel = sel_utts(B)
    out = {"meta": np.array([dim_neck, freq, B, T, 80, wseed, iseed, 1], np.int64), "sel": np.array(sel, np.int64),
           "param_names": np.array([k for k, _ in G.named_parameters()]),
           "param_digest0": np.stack([digest(p) for p in G.parameters()])}
    opt = torch.optim.Adam(G.parameters(), 1e-4)                    # solver_encoder.py:130
    t0 = time.time()
    g_loss, ls, outs = ref_step(G, x, e)
    opt.zero_grad()
    g_loss.backward()
    out["s0_losses"] = np.array([g_loss.item()] + [l.item() for l in ls], np.float64)
    for k, v in zip(("x_identic", "x_identic_psnt", "code_real", "code_reconst"), outs):
        v = v.detach()
        out[f"s0_{k}_digest"] = digest(v, N_STRIDED)
        if k.startswith("code"):
            out["s0_" + k] = v.numpy()
        else:
            out[f"s0_{k}_sel"] = v[sel].numpy()
    out["s0_grad_digest"] = np.stack([digest(p.grad) for p in G.parameters()])
    for k, v in G.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["s0_buf/" + k] = v.detach().numpy().copy()
    opt.step()
    out["s0_param_digest"] = np.stack([digest(p) for p in G.parameters()])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, out["s0_losses"], f"{time.time() - t0:.1f}s", flush=True)


def make_eval_golden_long(name, dim_neck, freq, B, T, wseed=0, iseed=77):
    """Conversion maths (conversion.py:47,:91-92) at the config-5 length.  The BN running statistics are made
    non-trivial by two train-mode forwards on T=128 crops of the same inputs."""
    from model_vc_mel import Generator
    torch.manual_seed(wseed)
    G = Generator(dim_neck, 256, 512, freq)
    x, e, e2 = synth_inputs(B, T, 80, 256, iseed)
    G.train()
    with torch.no_grad():
        G(x[:, :128].contiguous(), e, e)
        G(x[:, 128:256].flip(0).contiguous(), e2, e)
    G.eval()
    with torch.no_grad():
        xi, xp, codes = G(x, e, e2)
    out = {"meta": np.array([dim_neck, freq, B, T, 80, wseed, iseed], np.int64),
           "x_identic": xi.numpy(), "x_identic_psnt": xp.numpy(), "codes": codes.numpy()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, xi.shape, float(xp.abs().mean()), flush=True)


if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", "8")))
    make_eval_golden_long("eval_32_32_b2_t640", 32, 32, 2, 640)
    make_eval_golden_long("eval_16_16_b3_t640", 16, 16, 3, 640)
    make_train_golden_compact("train_c2_16_16_b256_t128", 16, 16, 256, 128)
    make_train_golden_compact("train_c3_32_32_b128_t256", 32, 32, 128, 256)
