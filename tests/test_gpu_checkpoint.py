"""model_EMA (solver_encoder.py:168-177) and the checkpoint of :333-346 / :147-153 on the drop-in Generator."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import synth_inputs

pytestmark = pytest.mark.gpu


def _reference_model_EMA(G, ema):
    """The literal statements of solver_encoder.py:168-177."""
    flat_params = torch.cat([param.data.view(-1) for param in G.parameters()], 0)
    avg_params = ema * flat_params + (1 - ema) * flat_params
    offset = 0
    for param in G.parameters():
        param.data.copy_(avg_params[offset:offset + param.nelement()].view(param.size()))
        offset += param.nelement()


@pytest.mark.parametrize("ema", [0.999, 0.9999, 0.5])
def test_model_EMA_is_bit_exact(ema):
    import autovc_b200
    from autovc_b200 import solver
    torch.manual_seed(0)
    G = autovc_b200.Generator(16, 256, 512, 16).cuda()
    torch.manual_seed(0)
    R = autovc_b200.Generator(16, 256, 512, 16).cuda()
    before = [p.detach().clone() for p in G.parameters()]
    solver.model_EMA(G, ema)
    _reference_model_EMA(R, ema)
    changed = 0
    for a, b, p0 in zip(G.parameters(), R.parameters(), before):
        assert torch.equal(a, b)
        changed += int((a != p0).sum())
    assert changed > 0 or ema == 0.5          # fp32: the "identity" moves parameters by an ulp; exact for 0.5 + 0.5


def test_checkpoint_layout_and_resume(tmp_path):
    """Train two steps, checkpoint (EMA + async write), keep training; a fresh model resumed from the file (through the
    reference's own resume statements, solver_encoder.py:147-153) reproduces the following step exactly."""
    import autovc_b200
    from autovc_b200 import solver
    torch.manual_seed(0)
    G = autovc_b200.Generator(16, 256, 512, 16, precision="half").cuda().train()
    opt = autovc_b200.FusedAdam(G.parameters(), 1e-4)
    x, e, _ = synth_inputs(4, 32, 80, 256, 7)
    x, e = x.cuda(), e.cuda()
    for _ in range(2):
        out = solver.train_step(G, opt, x, e)
    path = os.path.join(tmp_path, "chkpnt_spmel_test.ckpt")
    h = solver.save_checkpoint(G, opt, epoch=2, loss={"G/loss_id": out["L_id"]}, path=path, ema=0.999)
    nxt = solver.train_step(G, opt, x, e)          # the loop goes on while the file is written
    h.wait()
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert sorted(ck.keys()) == ["epoch", "loss", "optimizer", "state_dict"] and ck["epoch"] == 2
    assert list(ck["state_dict"].keys()) == list(G.state_dict().keys())
    assert sorted(ck["optimizer"].keys()) == ["param_groups", "state"] and len(ck["optimizer"]["state"]) == len(list(G.parameters()))
    # resume with the reference's statements into a fresh module + torch.optim.Adam (the reference's optimizer class)
    torch.manual_seed(1)
    G2 = autovc_b200.Generator(16, 256, 512, 16, precision="half").cuda().train()
    opt2 = torch.optim.Adam(G2.parameters(), 1e-4)
    checkpoint = torch.load(path, map_location="cuda", weights_only=False)
    G2.load_state_dict(checkpoint["state_dict"])
    opt2.load_state_dict(checkpoint["optimizer"])
    nxt2 = solver.train_step(G2, opt2, x, e)
    for k in ("g_loss", "L_id", "L_id_psnt", "L_cd"):
        assert abs(nxt[k] - nxt2[k]) <= 1e-6 * abs(nxt[k]), (k, nxt[k], nxt2[k])
    # and through load_checkpoint into FusedAdam
    torch.manual_seed(2)
    G3 = autovc_b200.Generator(16, 256, 512, 16, precision="half").cuda().train()
    opt3 = autovc_b200.FusedAdam(G3.parameters(), 1e-4)
    epoch, loss = solver.load_checkpoint(path, G3, opt3)
    assert epoch == 2 and "G/loss_id" in loss
    assert all(int(st["step"]) == 2 for st in ck["optimizer"]["state"].values())      # not the live counter of the running loop
    nxt3 = solver.train_step(G3, opt3, x, e)
    for k in ("g_loss", "L_id", "L_id_psnt", "L_cd"):
        assert abs(nxt[k] - nxt3[k]) <= 1e-6 * abs(nxt[k]), (k, nxt[k], nxt3[k])
    for a, b in zip(G.parameters(), G3.parameters()):
        assert float((a - b).abs().max()) <= 1e-5 * float(a.abs().max()) + 1e-9      # a tenth of one Adam step (lr 1e-4)


def test_fused_adam_invalidates_packed_weights():
    """FusedAdam writes the parameters through raw pointers: an encoder-only call right after the step must see the NEW
    weights (advisor finding: the pack cache keyed on the autograd version counter)."""
    import autovc_b200
    from autovc_b200 import ops, solver
    torch.manual_seed(0)
    G = autovc_b200.Generator(16, 256, 512, 16, precision="half").cuda().train()
    opt = autovc_b200.FusedAdam(G.parameters(), 1e-2)          # a large step so that stale packs would be visible
    x, e, _ = synth_inputs(4, 32, 80, 256, 9)
    x, e = x.cuda(), e.cuda()
    solver.train_step(G, opt, x, e)
    G.eval()
    with torch.enable_grad():                      # the path that does NOT clear the cache on entry
        got = G(x, e, None).detach()
    ops._GLOBAL_CACHE.begin_step()
    with torch.no_grad():
        ref = G(x, e, None)
    assert torch.equal(got, ref)
