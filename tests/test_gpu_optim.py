"""FusedAdam (avc_adam_step) against torch.optim.Adam as solver_encoder.py:130 configures it: same updates, same
state_dict layout, checkpoints interchangeable in both directions."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import autovc_b200

DEV = "cuda"
SHAPES = [(512, 336, 5), (4096,), (64, 16), (7,), (1,), (4097,), (2048, 288), (80, 1024), (3, 5, 7)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in SHAPES]


def _grads(params, seed, unaligned=False):
    g = torch.Generator().manual_seed(seed)
    for p in params:
        gr = torch.randn(p.shape, generator=g).to(DEV) * 10 ** float(torch.randint(-6, 1, (1,), generator=g))
        if unaligned:      # a gradient that is a 4-byte-aligned view into a flat bucket (GradBucketReducer)
            flat = torch.empty(p.numel() + 1, device=DEV)
            flat[1:].copy_(gr.flatten())
            gr = flat[1:].view_as(p)
        p.grad = gr


@pytest.mark.parametrize("unaligned", [False, True])
def test_fused_adam_matches_torch_adam(unaligned):
    pa, pb = _params(0), _params(0)
    ref = torch.optim.Adam(pa, 1e-3)
    ours = autovc_b200.FusedAdam(pb, 1e-3)
    for it in range(5):
        _grads(pa, 100 + it)
        _grads(pb, 100 + it, unaligned)
        ref.step()
        ours.step()
        for a, b in zip(pa, pb):
            # same formula in the same order; torch's foreach kernels round a few intermediates differently
            assert torch.allclose(a, b, rtol=2e-6, atol=1e-9), (it, a.shape, float((a - b).abs().max()))
    sa, sb = ref.state_dict(), ours.state_dict()
    assert sa["param_groups"][0].keys() == sb["param_groups"][0].keys()
    assert list(sa["state"].keys()) == list(sb["state"].keys())
    for k in sa["state"]:
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 5
        for name in ("exp_avg", "exp_avg_sq"):
            assert torch.allclose(sa["state"][k][name], sb["state"][k][name], rtol=2e-6, atol=1e-30)


def test_fused_adam_resumes_from_a_torch_adam_checkpoint():
    pa, pb = _params(1), _params(1)
    ref = torch.optim.Adam(pa, 1e-4)
    for it in range(2):
        _grads(pa, 7 + it)
        ref.step()
    ours = autovc_b200.FusedAdam(pb, 1e-4)
    with torch.no_grad():
        for a, b in zip(pa, pb):
            b.copy_(a)
    # (deepcopy = what torch.load of a checkpoint gives; load_state_dict itself aliases same-device tensors)
    ours.load_state_dict(copy.deepcopy(ref.state_dict()))          # solver_encoder.py:150-152 style resume
    _grads(pa, 9)
    _grads(pb, 9)
    ref.step()
    ours.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-9)
    back = torch.optim.Adam(_params(1), 1e-4)
    back.load_state_dict(copy.deepcopy(ours.state_dict()))         # and the other way round
    assert float(back.state_dict()["state"][0]["step"]) == 3


def test_fused_adam_skips_parameters_without_gradient_and_rejects_unsupported():
    ps = _params(2)
    ours = autovc_b200.FusedAdam(ps, 1e-3)
    _grads(ps, 3)
    ps[2].grad = None
    before = ps[2].detach().clone()
    ours.step()
    assert torch.equal(ps[2], before) and len(ours.state[ps[2]]) == 0
    with pytest.raises(autovc_b200.AvcError):
        autovc_b200.FusedAdam(ps, 1e-3, weight_decay=0.1)
    cpu = [torch.nn.Parameter(torch.zeros(4))]
    opt = autovc_b200.FusedAdam(cpu, 1e-3)
    cpu[0].grad = torch.ones(4)
    with pytest.raises(autovc_b200.AvcError):
        opt.step()
