"""ctypes binding of libtt_b200.so (the C ABI declared in include/tt_b200.h).

There is NO CPU fallback: if the shared library is missing, or the current
device is not a B200-class (cc 10.x) GPU, every op raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtt_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "tt_b200.h")

P = c_void_p
_SIGNATURES = {
    "tt_abi_version": (c_int, []),
    "tt_last_error": (ctypes.c_char_p, []),
    "tt_device_info": (c_int, [P, P, P]),
    "tt_emb_gather_pool_fwd": (c_int, [P, c_int, c_int64, c_int, P, c_int64, c_int, c_int, c_int64, P, c_int64, P, P, P]),
    "tt_emb_segment_grad_workspace": (c_int, [c_int64, c_int, P]),
    "tt_emb_segment_grad": (c_int, [P, c_int64, c_int, c_int, c_int64, c_int64, P, c_int64, P, c_int, P, P, P, P, P,
                                    c_size_t, P]),
    "tt_emb_segment_grad_lists": (c_int, [P, c_int64, c_int64, c_int64, P, c_int64, P, c_int64, c_int64, c_int, P, P, P, P, P,
                                          c_size_t, P]),
    "tt_shard_route": (c_int, [P, c_int64, c_int, c_int64, c_int64, c_int, P, c_int64, c_int64, c_int64, c_int64, P, P, P,
                               c_size_t, P]),
    "tt_shard_owner_gather": (c_int, [P, c_int, c_int64, c_int, c_int, P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int,
                                      P, c_int64, c_int64, P, P]),
    "tt_shard_combine": (c_int, [P, c_int64, c_int64, c_int, P, c_int64, c_int, c_int64, c_int64, c_int, P, c_int64, c_int64,
                                 c_int64, P, P, c_int, P, c_int64, P]),
    "tt_shard_grad_pack": (c_int, [P, c_int64, c_int64, c_int, c_int, c_int, c_int, P, c_int64, c_int64, P, c_int64, c_int64,
                                   c_int64, P, c_int64, c_int64, P]),
    "tt_bn_workspace": (c_int, [c_int64, c_int, P]),
    "tt_bn_stats": (c_int, [P, c_int64, c_int, c_int64, P, P, c_size_t, P]),
    "tt_bn_apply": (c_int, [P, c_int64, c_int, c_int64, P, c_int, P, P, c_int, c_float, c_int, c_float, P, c_int64, P, c_int64,
                            P, P, P, P, P, c_float, P, P]),
    "tt_bn_bwd_stats": (c_int, [P, c_int64, P, c_int64, c_int, c_int64, P, P, P, P, c_int, c_int, c_float, P, c_int64, P, P, P,
                                c_int, P, c_size_t, P]),
    "tt_bn_bwd_apply": (c_int, [P, c_int64, P, c_int64, c_int, c_int64, P, P, P, P, c_int, c_int, c_float, P, c_int64, P,
                                c_int, c_double, P, c_int64, P]),
    "tt_p2p_allgather_small": (c_int, [P, c_int, c_int, c_int, P, c_int64, c_int64, c_int64, P, P]),
    "tt_emb_segment_adam_lists": (c_int, [c_int64, P, P, c_int64, c_int64, c_int, P, c_int64, P, c_size_t, P, P, P, P, c_double,
                                          c_double, c_double, c_double, P, P, P]),
    "tt_emb_rowwise_adam": (c_int, [P, c_int, P, P, c_int, P, P, P, c_int64, P, c_double, c_double, c_double, c_double, P, P, P]),
    "tt_emb_scatter_rows": (c_int, [P, c_int, P, P, P, c_int64, P]),
    "tt_sq_norm_accum": (c_int, [P, c_int64, P, P, c_size_t, P]),
    "tt_clip_coef": (c_int, [P, c_int, c_float, P, P, P]),
    "tt_adam_flat": (c_int, [P, P, P, P, c_int64, P, c_double, c_double, c_double, c_double, P, P, P]),
    "tt_ce_workspace": (c_int, [c_int64, c_int64, c_int, c_int, P]),
    "tt_ce_fwd_f32": (c_int, [P, P, P, P, c_int, P, c_int64, c_int64, c_int, c_float, P, P, P, P, P, c_size_t, P]),
    "tt_ce_bwd_f32": (c_int, [P, P, P, P, c_int, P, c_int64, c_int64, c_int, c_float, P, P, P, P, P, P, P, c_size_t, P]),
    "tt_ce_tc_workspace": (c_int, [c_int64, c_int64, c_int, c_int, P]),
    "tt_ce_fwd_tc": (c_int, [P, P, P, P, c_int, P, c_int64, c_int64, c_int, c_float, P, P, P, P, P, c_size_t, P]),
    "tt_ce_bwd_tc_workspace": (c_int, [c_int64, c_int64, c_int, c_int, P]),
    "tt_ce_bwd_tc": (c_int, [P, P, c_int, c_int64, c_int64, c_int, c_float, P, P, P, P, P, P, P, c_size_t, P, c_size_t, P]),
    "tt_ce_tc_workspace_rect": (c_int, [c_int64, c_int64, c_int64, c_int, c_int, P]),
    "tt_ce_fwd_tc_rect": (c_int, [P, P, P, c_int64, P, c_int, P, c_int64, c_int64, c_int64, c_int, c_float, P, P, P, P, P,
                                  c_size_t, P]),
    "tt_ce_fwd_tc_rect_bits": (c_int, [P, P, P, c_int64, P, c_int, P, c_int64, c_int64, c_int64, c_int, c_float, P, P, P, P, P,
                                       c_size_t, c_int, P]),
    "tt_ce_fwd_tc_fused": (c_int, [P, P, P, c_int64, P, c_int64, c_int64, c_int64, c_int, c_float, P, P, P, P, P, c_size_t, P,
                                   c_size_t, c_int, P]),
    "tt_ce_bwd_tc_fused": (c_int, [P, c_int64, c_int64, c_int64, c_int, c_float, P, P, P, P, P, P, c_size_t, P, c_size_t, P]),
    "tt_ce_bwd_tc_workspace_rect": (c_int, [c_int64, c_int64, c_int64, c_int, c_int, P]),
    "tt_ce_bwd_tc_rect": (c_int, [P, P, c_int, c_int64, c_int64, c_int64, c_int, c_float, P, P, P, P, P, P, P, c_size_t, P,
                                  c_size_t, P]),
    "tt_ce_tc_debug_trace": (c_int, [P]),
    "tt_score_topk_workspace": (c_int, [c_int64, c_int64, c_int, c_int, P]),
    "tt_score_topk_f32": (c_int, [P, c_int64, P, c_int64, c_int, c_int, c_int64, P, P, P, P, P, c_size_t, P]),
    "tt_topk_tc_prepare_corpus": (c_int, [P, c_int64, c_int, P, P, P]),
    "tt_score_topk_tc_workspace": (c_int, [c_int64, c_int64, c_int, c_int, c_int, P]),
    "tt_score_topk_tc": (c_int, [P, c_int64, P, P, P, c_int64, c_int, c_int, c_int64, P, P, P, P, P, c_int, P, c_size_t, P]),
    "tt_topk_merge": (c_int, [P, P, c_int, c_int64, c_int, P, P, P]),
    "tt_linear_fwd": (c_int, [P, P, P, c_int64, c_int, c_int, c_int, P, P]),
    "tt_linear_dgrad": (c_int, [P, P, c_int64, c_int, c_int, P, P]),
    "tt_linear_wgrad_workspace": (c_int, [c_int64, c_int, c_int, P]),
    "tt_linear_wgrad": (c_int, [P, P, c_int64, c_int, c_int, P, P, c_int, P, c_size_t, P]),
    "tt_linear_tc_supported": (c_int, [c_int64, c_int, c_int]),
    "tt_linear_fwd_tc": (c_int, [P, P, P, c_int64, c_int, c_int, c_int, P, P]),
    "tt_linear_dgrad_tc": (c_int, [P, P, c_int64, c_int, c_int, P, P]),
    "tt_linear_wgrad_tc_workspace": (c_int, [c_int64, c_int, c_int, P]),
    "tt_linear_wgrad_tc": (c_int, [P, P, c_int64, c_int, c_int, P, P, c_int, P, c_size_t, P]),
    "tt_l2_normalize_fwd": (c_int, [P, c_int64, c_int, c_float, P, P, P]),
    "tt_l2_normalize_bwd": (c_int, [P, P, P, c_int64, c_int, c_float, P, P]),
    "tt_attn_small_fwd": (c_int, [P, P, c_int64, c_int, c_int, c_int, c_float, P, c_int64, P, P]),
    "tt_attn_small_bwd": (c_int, [P, P, P, c_int64, c_int, c_int, c_int, c_float, P, c_int64, P, P]),
    "tt_add_dropout_ln_fwd": (c_int, [P, P, c_int64, c_int, P, P, c_float, c_float, P, c_int64, P, P, P, P]),
    "tt_add_dropout_ln_bwd_workspace": (c_int, [c_int64, c_int, P]),
    "tt_add_dropout_ln_bwd": (c_int, [P, P, P, P, c_int64, c_int, c_float, P, c_int64, P, P, P, P, c_int, P, c_size_t, P]),
}

_lib = None


class TTError(RuntimeError):
    pass


def declared_symbols(header_path: str = HEADER_PATH):
    """Names of every TT_API function declared in include/tt_b200.h."""
    with open(header_path) as f:
        text = f.read()
    return sorted(set(re.findall(r"TT_API\s+[\w\s\*]+?\b(tt_[a-z0-9_]+)\s*\(", text)))


def load():
    """Load the shared library (once) and attach signatures.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TTError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  recommendsystemproject_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.tt_abi_version() != 2:
        raise TTError("libtt_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().tt_last_error().decode(errors="replace")
        raise TTError(f"{what} failed (rc={rc}): {msg}")


def require_device():
    """Fail loudly unless the current CUDA device is cc 10.x."""
    import torch
    if not torch.cuda.is_available():
        raise TTError("recommendsystemproject_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    lib = load()
    sms, major, minor = c_int(0), c_int(0), c_int(0)
    check(lib.tt_device_info(ctypes.byref(sms), ctypes.byref(major), ctypes.byref(minor)), "tt_device_info")
    return sms.value, major.value, minor.value


__all__ = ["load", "check", "require_device", "declared_symbols", "TTError", "LIB_PATH",
           "c_double", "c_int32", "c_size_t"]
