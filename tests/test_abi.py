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
    assert _lib.query("avc_logmel_workspace_bytes", 4, 16000) > 4 * 16000 * 12


def test_ops_refuse_cpu_tensors():
    import torch
    from autovc_b200 import AvcError, Generator
    torch.manual_seed(0)
    G = Generator(16, 256, 512, 16)
    with pytest.raises(AvcError):
        G(torch.rand(1, 16, 80), torch.rand(1, 256), torch.rand(1, 256))
