#!/usr/bin/env python
"""Loss-curve gate (north_star: a 1k-step loss curve from identical init within 2%).

Trains the drop-in Generator from the same seeded init on the same data stream in fp32 mode (the
parity-validated <=1e-4 mode, standing in for the reference's fp32 path) and in a reduced-precision mode,
and compares the moving-average total loss (window 25; SURVEY 7.2: per-step values are chaotic at small
batch even between two fp32 reference runs).  Optionally also runs the CPU oracle for the first steps.

    python scripts/loss_curve.py --steps 1000 --batch 16 --modes tf32 bf16
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

import autovc_b200
from autovc_b200 import solver


def corpus(n_utt=64, frames=400, seed=7):
    """Synthetic 'speakers': smooth band-limited mel-like trajectories in [0,1] + per-speaker embeddings."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(n_utt, frames // 8 + 2, 80, generator=g)
    x = F.interpolate(base.permute(0, 2, 1), size=frames, mode="linear", align_corners=True).permute(0, 2, 1)
    x = (0.7 * x + 0.3 * torch.rand(n_utt, frames, 80, generator=g)).clamp(0, 1)
    e = F.normalize(torch.randn(n_utt, 256, generator=g), dim=-1) * 0.8
    return x, e


def run(mode, steps, B, T, seed=0):
    torch.manual_seed(seed)
    G = autovc_b200.Generator(16, 256, 512, 16, precision=mode).cuda().train()
    # the fp32 reference curve keeps torch.optim.Adam; the reduced-precision runs use the shipped configuration (FusedAdam)
    opt = torch.optim.Adam(G.parameters(), 1e-4) if mode == "fp32" else autovc_b200.FusedAdam(G.parameters(), 1e-4)
    X, E = corpus()
    X, E = X.cuda(), E.cuda()
    rs = np.random.RandomState(123)
    losses = []
    for i in range(steps):
        idx = rs.randint(0, X.shape[0], size=B)
        off = rs.randint(0, X.shape[1] - T, size=B)
        xb = torch.stack([X[j, o:o + T] for j, o in zip(idx, off)])
        out = solver.train_step(G, opt, xb.contiguous(), E[idx].contiguous())
        losses.append([out["g_loss"], out["L_id"], out["L_id_psnt"], out["L_cd"]])
    return np.array(losses)


def movavg(v, w=25):
    c = np.cumsum(np.insert(v, 0, 0.0))
    return (c[w:] - c[:-w]) / w


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--len-crop", type=int, default=128)
    ap.add_argument("--modes", nargs="+", default=["tf32"])
    ap.add_argument("--out", default="gpurun_out/loss_curve.json")
    a = ap.parse_args()
    ref = run("fp32", a.steps, a.batch, a.len_crop)
    res = {"steps": a.steps, "batch": a.batch, "fp32_first": ref[0].tolist(), "fp32_last100_mean": float(ref[-100:, 0].mean())}
    for m in a.modes:
        cur = run(m, a.steps, a.batch, a.len_crop)
        ma_r, ma_c = movavg(ref[:, 0]), movavg(cur[:, 0])
        rel = np.abs(ma_c - ma_r) / ma_r
        res[m] = {"max_rel_dev_movavg25": float(rel.max()), "final_rel_dev_movavg25": float(rel[-1]),
                  "max_rel_dev_per_step": float((np.abs(cur[:, 0] - ref[:, 0]) / ref[:, 0]).max()),
                  "last100_mean": float(cur[-100:, 0].mean()),
                  "curve_every_50": cur[::50, 0].round(5).tolist()}
        print(m, {k: v for k, v in res[m].items() if k != "curve_every_50"}, flush=True)
    res["fp32_curve_every_50"] = ref[::50, 0].round(5).tolist()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print("fp32 loss", ref[0, 0], "->", res["fp32_last100_mean"])
