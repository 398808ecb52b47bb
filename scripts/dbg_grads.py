import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import autovc_b200
from autovc_b200 import solver
from tests.helpers import load_golden, synth_inputs, digest
g = load_golden("train_16_16_b16_t128")
for prec in ["tf32"]:
    torch.manual_seed(0)
    G = autovc_b200.Generator(16,256,512,16, precision=prec).cuda().train()
    x,e,_ = synth_inputs(16,128,80,256,1234)
    opt = torch.optim.Adam(G.parameters(), 1e-4)
    out = solver.train_step(G, opt, x.cuda(), e.cuda(), return_outputs=True)
    ref = g["s0_grad_digest"]
    for i,(n,p) in enumerate(G.named_parameters()):
        d = digest(out["grads"][n])
        print(f"{prec} {n:55s} ours {d[2]:.5e} ref {ref[i][2]:.5e} ratio {d[2]/max(ref[i][2],1e-30):.4f}")
