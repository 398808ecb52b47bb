"""CPU-side checks of the boundary: the shared library loads and exports every symbol that
include/autovc_b200.h declares (no compute calls without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "autovc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    from autovc_b200 import _lib
    names = _declared()
    assert len(names) >= 30
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.avc_version() == 100


def test_workspace_queries_run_without_gpu():
    from autovc_b200 import _lib
    assert _lib.query("avc_gemm_tn_workspace_bytes", 256, 128, 512, 512, 5, 0) > 0
    assert _lib.query("avc_lstm_bwd_workspace_bytes", 256, 128, 1024, 0) == 256 * 1024 * 4
    assert _lib.query("avc_logmel_workspace_bytes", 4, 16000) > 4 * 16000 * 8


def test_ops_refuse_cpu_tensors():
    import torch
    from autovc_b200 import AvcError, Generator
    torch.manual_seed(0)
    G = Generator(16, 256, 512, 16)
    with pytest.raises(AvcError):
        G(torch.rand(1, 16, 80), torch.rand(1, 256), torch.rand(1, 256))


def test_fused_adam_is_a_torch_adam_and_refuses_cpu():
    """Host logic of the optimizer drop-in without a GPU: class relationship, state_dict layout, loud failure on CPU."""
    import torch
    from autovc_b200 import AvcError, FusedAdam
    p = [torch.nn.Parameter(torch.zeros(8)), torch.nn.Parameter(torch.zeros(3, 5))]
    opt = FusedAdam(p, 1e-4)
    assert isinstance(opt, torch.optim.Adam)
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(8)), torch.nn.Parameter(torch.zeros(3, 5))], 1e-4)
    assert opt.state_dict()["param_groups"][0].keys() == ref.state_dict()["param_groups"][0].keys()
    for q in p:
        q.grad = torch.ones_like(q)
    with pytest.raises(AvcError):
        opt.step()
    with pytest.raises(AvcError):
        FusedAdam(p, 1e-4, amsgrad=True)


def test_nccl_env_defaults_do_not_override_the_user(monkeypatch):
    from autovc_b200 import solver
    monkeypatch.delenv("NCCL_MAX_CTAS", raising=False)
    solver.nccl_env_defaults()
    assert os.environ["NCCL_MAX_CTAS"] == "16"
    monkeypatch.setenv("NCCL_MAX_CTAS", "32")
    solver.nccl_env_defaults()
    assert os.environ["NCCL_MAX_CTAS"] == "32"
