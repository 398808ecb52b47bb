"""ctypes binding of libautovc_b200.so (the C-ABI declared in include/autovc_b200.h).

There is NO CPU fallback: if the shared library is missing this raises at import of any op,
and every op raises on non-CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libautovc_b200.so")

PREC_FP32, PREC_BF16, PREC_TF32, PREC_HALF, PREC_FP32X3 = 0, 1, 2, 3, 4
FMT_FP32, FMT_BF16, FMT_FP16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
ACT_CODES = {"none": ACT_NONE, "linear": ACT_NONE, "relu": ACT_RELU, "tanh": ACT_TANH}

P = c_void_p  # every device pointer / stream crosses the boundary as void*

# name -> (restype, argtypes); mirrors include/autovc_b200.h one to one
SIGNATURES = {
    "avc_version": (c_int, []),
    "avc_last_error": (ctypes.c_char_p, []),
    "avc_launch_count": (c_ulonglong, []),
    "avc_gemm_nt_taps": (c_int, [P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, c_int, P, c_size_t, P]),
    "avc_gemm_nt_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "avc_gemm_tn_taps": (c_int, [P, c_int, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_size_t, P]),
    "avc_gemm_tn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "avc_pack_conv_weight": (c_int, [P, P, P, c_int, c_int, c_int, P]),
    "avc_pack_lstm_weight": (c_int, [P, P, P, c_int, c_int, P]),
    "avc_pack_lstm_bias": (c_int, [P, P, P, c_int, P]),
    "avc_transpose": (c_int, [P, P, c_int, c_int, P]),
    "avc_channel_stats": (c_int, [P, c_int, c_int, c_int, P, P]),
    "avc_bn_finalize": (c_int, [P, c_int, c_int, c_float, c_float, P, P, P, P, P]),
    "avc_bn_eval_stats": (c_int, [P, P, c_int, c_float, P, P, P]),
    "avc_bn_act_fwd": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, P]),
    "avc_bn_act_bwd_reduce": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, P]),
    "avc_bn_act_bwd_apply": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "avc_bn_act_bwd_reduce_y": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, P]),
    "avc_bn_act_bwd_apply_y": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, P, P, c_int, c_int, c_int, c_int, P]),
    "avc_colsum": (c_int, [P, c_int, c_int, c_int, P, P, c_int, c_int, P, c_size_t, P]),
    "avc_colsum16": (c_int, [P, c_int, c_int, c_int, c_int, P, P, c_int, c_int, P, c_size_t, P]),
    "avc_lstm_seq_fwd": (c_int, [P, P, P, c_int, P, P, c_int, c_int, c_int, c_int, c_int, P, c_size_t, P]),
    "avc_lstm_fwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "avc_lstm_seq_bwd": (c_int, [P, c_int, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, c_size_t, P]),
    "avc_lstm_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "avc_debug_set_trace": (None, [P]),
    "avc_gemm_nt_taps_h": (c_int, [P, c_int, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, c_int, P, c_size_t, P]),
    "avc_gemm_nt_h_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "avc_gemm_tn_taps_h": (c_int, [P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_size_t, P]),
    "avc_gemm_tn_h_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "avc_cast16": (c_int, [P, c_int, P, c_int, c_size_t, c_int, c_int, P]),
    "avc_bn_act_fwd_h": (c_int, [P, P, P, P, P, P, P, P, c_int, P, c_int, c_int, c_int, c_int, P]),
    "avc_bn_act_bwd_apply_h": (c_int, [P, P, P, P, P, P, P, P, P, c_int, P, P, c_int, c_int, c_int, c_int, P]),
    "avc_lstm_seq_fwd_h": (c_int, [P, P, c_int, P, c_int, P, P, P, c_int, P, c_int, c_int, c_int, c_int, P, c_size_t, P]),
    "avc_lstm_seq_bwd_h": (c_int, [P, c_int, P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, P, c_size_t, P]),
    "avc_pack_conv_weight_h": (c_int, [P, P, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    "avc_pack_lstm_weight_h": (c_int, [P, P, c_int, c_int, P, c_int, c_int, c_int, c_int, P]),
    "avc_gemm_nt_taps_hw": (c_int, [P, c_int, c_int, P, c_int, c_int, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int,
                                    P, c_size_t, P]),
    "avc_concat_bcast": (c_int, [P, c_int, P, P, c_int, c_int, c_int, c_int, P]),
    "avc_codes_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "avc_codes_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "avc_upsample_concat_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "avc_upsample_concat_bwd": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    "avc_copy2d": (c_int, [P, c_int, P, c_int, c_int, c_int, P]),
    "avc_mse_loss_fwd": (c_int, [P, P, c_size_t, P, P, P]),
    "avc_l1_loss_fwd": (c_int, [P, P, c_size_t, P, P, P]),
    "avc_loss_bwd": (c_int, [P, P, c_size_t, P, c_int, P, P, c_int, P]),
    "avc_l2_normalize_rows": (c_int, [P, P, c_int, c_int, P]),
    "avc_crop_batch": (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "avc_adam_chunk_elems": (c_int, []),
    "avc_adam_step": (c_int, [P, P, c_int, c_double, c_double, c_double, c_double, c_int, c_float, P]),
    "avc_ema_blend": (c_int, [P, P, c_int, c_double, P]),
    "avc_logmel_frontend": (c_int, [P, P, P, c_int, c_int, P, P, P, P, c_int, P, c_size_t, P]),
    "avc_logmel_workspace_bytes": (c_size_t, [c_int, c_int]),
    "avc_logstft_frontend": (c_int, [P, P, P, c_int, c_int, P, P, P, P, c_int, P, c_size_t, P]),
    "avc_prelu_fwd": (c_int, [P, P, P, P, c_int, c_int, P]),
    "avc_prelu_bwd": (c_int, [P, P, P, P, P, c_int, P, c_size_t, P]),
    "avc_sum_all": (c_int, [P, c_size_t, P, P, c_int, P]),
    "avc_copy_rows3d": (c_int, [P, c_int, P, c_int, c_int, c_int, P]),
    "avc_permute021": (c_int, [P, P, c_int, c_int, c_int, P]),
    "avc_sisnr_fwd": (c_int, [P, P, c_int, c_int, P, P, P, P]),
    "avc_sisnr_bwd": (c_int, [P, P, P, P, c_int, c_int, P, c_int, P]),
}

_lib = None


class AvcError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and bind every signature.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AvcError(f"{LIB_PATH} not found: build it with `make -C autovc_b200/csrc` "
                       f"(or __graft_entry__.build()); autovc_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_timing = None   # None, or a list of (name, args, start_event, end_event) while bench.py profiles


def enable_timing(on: bool):
    """Bracket every C-ABI call with CUDA events on the launching stream (bench.py roofline leg)."""
    global _timing
    _timing = [] if on else None


def collect_timing():
    """[(name, args, milliseconds)] for the calls since enable_timing(True); caller synchronises first."""
    out = [(n, a, s.elapsed_time(e)) for n, a, s, e in (_timing or [])]
    if _timing is not None:
        _timing.clear()
    return out


def call(name: str, *args):
    """Call an int-returning entry point; raise AvcError(avc_last_error()) on failure."""
    lib = load()
    if _timing is not None:
        import torch
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = getattr(lib, name)(*args)
        e.record()
        _timing.append((name, [a if isinstance(a, int) else None for a in args], s, e))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise AvcError(f"{name} failed (code {rc}): {lib.avc_last_error().decode()}")


def query(name: str, *args):
    return getattr(load(), name)(*args)


def launch_count() -> int:
    return int(load().avc_launch_count())
