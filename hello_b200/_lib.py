"""ctypes binding of include/hello_moe.h (libhello_moe.so, built in-tree by hello_b200/build.py).

There is no CPU fallback: if the library is missing or does not load, importing the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# HELLO_MOE_LIB: developer hook for A/B runs of two builds of the library on one GPU box (tools/ab/); never a CPU path.
LIB_PATH = os.environ.get("HELLO_MOE_LIB") or os.path.join(HERE, "libhello_moe.so")
ABI_VERSION = 3

LAYOUT_RCL, LAYOUT_RLC = 0, 1
META_NONE, META_SITE, META_REF = 0, 1, 2
COMBINE_NONE, COMBINE_CONV, COMBINE_SUM = 0, 1, 2
PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}

EXPORTS = (
    "hello_moe_abi_version", "hello_moe_create", "hello_moe_destroy", "hello_moe_last_error",
    "hello_moe_workspace_bytes", "hello_moe_forward", "hello_moe_forward_range", "hello_moe_launch_count",
    "hello_moe_run_net",
    "hello_moe_profile_enable", "hello_moe_profile_collect", "hello_moe_readconv_debug",
    "hello_moe_headconv_debug",
    # include/hello_encode.h
    "hello_encode_reads", "hello_encode_last_error",
)


class HelloCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("n_tech", C.c_int32), ("read_channels", C.c_int32 * 2),
        ("xattn_present", C.c_int32 * 3), ("has_combiners", C.c_int32), ("meta_kind", C.c_int32),
        ("feature_length", C.c_int32), ("precision", C.c_int32),
        ("max_chunk_sites", C.c_int32),
    ]


class HelloBatch(C.Structure):
    _fields_ = [
        ("n_sites", C.c_int64), ("n_alleles", C.c_int64), ("n_reads", C.c_int64 * 2),
        ("input_layout", C.c_int32), ("reserved", C.c_int32),
        ("d_reads", C.c_void_p * 2), ("d_allele_read_off", C.c_void_p * 2), ("h_allele_read_off", C.c_void_p * 2),
        ("d_site_allele_off", C.c_void_p), ("h_site_allele_off", C.c_void_p), ("d_ref_onehot", C.c_void_p),
        ("d_allele_rank", C.c_void_p), ("d_pair_off", C.c_void_p),
    ]


class HelloResult(C.Structure):
    _fields_ = [
        ("d_logits", C.c_void_p), ("d_meta", C.c_void_p), ("d_pair_prob", C.c_void_p),
        ("d_pair_mix64", C.c_void_p), ("d_best_pair", C.c_void_p), ("d_best_prob", C.c_void_p),
        ("d_call_pair", C.c_void_p), ("d_call_qual", C.c_void_p), ("d_best_expert", C.c_void_p),
    ]


class HelloMoEError(RuntimeError):
    pass


_lib = None


def load():
    """Load libhello_moe.so or raise; never falls back to another implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HelloMoEError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). hello_b200 has no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.hello_moe_abi_version.restype = C.c_int
    lib.hello_moe_create.restype = C.c_int
    lib.hello_moe_create.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(HelloCfg), C.c_int, C.POINTER(C.c_void_p)]
    lib.hello_moe_destroy.restype = None
    lib.hello_moe_destroy.argtypes = [C.c_void_p]
    lib.hello_moe_last_error.restype = C.c_char_p
    lib.hello_moe_last_error.argtypes = [C.c_void_p]
    lib.hello_moe_workspace_bytes.restype = C.c_size_t
    lib.hello_moe_workspace_bytes.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
    lib.hello_moe_forward.restype = C.c_int
    lib.hello_moe_forward.argtypes = [C.c_void_p, C.POINTER(HelloBatch), C.POINTER(HelloResult), C.c_void_p,
                                      C.c_size_t, C.c_void_p]
    lib.hello_moe_forward_range.restype = C.c_int
    lib.hello_moe_forward_range.argtypes = [C.c_void_p, C.POINTER(HelloBatch), C.POINTER(HelloResult), C.c_int64, C.c_int64,
                                            C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.hello_moe_launch_count.restype = C.c_int64
    lib.hello_moe_launch_count.argtypes = [C.c_void_p]
    lib.hello_moe_run_net.restype = C.c_int
    lib.hello_moe_run_net.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.hello_moe_profile_enable.restype = C.c_int
    lib.hello_moe_profile_enable.argtypes = [C.c_void_p, C.c_int]
    lib.hello_moe_profile_collect.restype = C.c_int
    lib.hello_moe_profile_collect.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.hello_moe_readconv_debug.restype = C.c_int
    lib.hello_moe_readconv_debug.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hello_moe_headconv_debug.restype = C.c_int
    lib.hello_moe_headconv_debug.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
    if lib.hello_moe_abi_version() != ABI_VERSION:
        raise HelloMoEError("libhello_moe.so ABI %d != binding ABI %d; rebuild" % (lib.hello_moe_abi_version(),
                                                                                 ABI_VERSION))
    _lib = lib
    return lib
