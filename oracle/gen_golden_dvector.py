"""tests/golden/dvector.npz: outputs of the UNMODIFIED reference model_bl.D_VECTOR(dim_input=80, dim_cell=768, dim_emb=256)
(make_metadata.py:42) with torch.manual_seed(0) init on seeded synthetic mel crops, fp32 and an fp64 copy.
Run in the build container only:  python oracle/gen_golden_dvector.py"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle import model_bl_ref as bref  # noqa: E402


def main():
    from model_bl import D_VECTOR
    torch.manual_seed(0)
    C = D_VECTOR(dim_input=80, dim_cell=768, dim_emb=256).eval()
    x = bref.synth_mels(6, 128, 21)
    with torch.no_grad():
        y32 = C(x)
        y64 = copy.deepcopy(C).double()(x.double())
    w = dict(C.state_dict())
    digest = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in w.values()])
    out = os.path.join(ROOT, "tests", "golden", "dvector.npz")
    np.savez_compressed(out, y32=y32.numpy(), y64=y64.numpy(), names=np.array(list(w.keys())), param_digest=digest,
                        meta=np.array([6, 128, 21, 0]))
    print("wrote", out, os.path.getsize(out), "max |y32-y64|", float((y32.double() - y64).abs().max()))


if __name__ == "__main__":
    main()
