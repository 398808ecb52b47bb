"""1000-step fp32 loss curve of the UNMODIFIED reference Generator on CPU (test infrastructure).

Run in the build container only (needs /root/reference; ~20 min on 4 cores):

    python oracle/gen_loss_curve_ref.py

north_star: "a 1k-step loss curve from identical init within 2%".  The reference module
(/root/reference/model_vc_mel.py) is trained with the step of solver_encoder.py:227-243,:293-300 and
torch.optim.Adam(lr 1e-4) (:130) from torch.manual_seed(0) on the data stream of tests.helpers
(loss_curve_corpus / loss_curve_batches: B = CURVE_B (16; a second curve at 64) crops of 128 frames).  The per-step losses go
to tests/golden/loss_curve_ref_b<B>.npz; tests/test_gpu_loss_curve.py trains the drop-in from the same init on the
same stream and compares 25-step moving averages.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.gen_golden import OUT, ref_step  # noqa: E402
from tests.helpers import loss_curve_batches, loss_curve_corpus  # noqa: E402

if __name__ == "__main__":
    steps, B, T = int(os.environ.get("CURVE_STEPS", "1000")), int(os.environ.get("CURVE_B", "16")), 128
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", "4")))
    from model_vc_mel import Generator
    torch.manual_seed(0)
    G = Generator(16, 256, 512, 16).train()
    opt = torch.optim.Adam(G.parameters(), 1e-4)
    X, E = loss_curve_corpus()
    losses = []
    t0 = time.time()
    for i, (idx, off) in enumerate(loss_curve_batches(steps, B, T)):
        xb = torch.stack([X[j, o:o + T] for j, o in zip(idx, off)]).contiguous()
        g_loss, ls, _ = ref_step(G, xb, E[idx].contiguous())
        opt.zero_grad()
        g_loss.backward()
        opt.step()
        losses.append([g_loss.item()] + [l.item() for l in ls])
        if i % 50 == 0:
            print(i, losses[-1], f"{time.time() - t0:.0f}s", flush=True)
    np.savez_compressed(os.path.join(OUT, f"loss_curve_ref_b{B}.npz"), losses=np.array(losses, np.float64),
                        meta=np.array([steps, B, T, 0, 123], np.int64))
    print("done", losses[0], "->", np.mean([l[0] for l in losses[-100:]]))
