"""ctypes binding of libzutis_b200.so (the C ABI declared in include/zutis_b200.h).

The library is the product: there is no Python / PyTorch / CPU fallback behind these
functions.  A missing .so, a missing symbol or a non-sm_100 device raises.
PyTorch's part is limited to owning device memory and streams (``tensor.data_ptr()``,
``torch.cuda.current_stream().cuda_stream``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzutis_b200.so")

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE, ERR_NO_DEVICE = range(6)
GT_U8, GT_I16, GT_I32, GT_I64 = range(4)
DECODE_AUTO, DECODE_GENERIC, DECODE_TILED, DECODE_CELLS = range(4)
DECODE_WORKSPACE_ZEROED = 0x100
GEMM_FP32_SIMT, GEMM_TF32X3, GEMM_TF32 = 0, 1, 2
GEMM_SIGMOID = 16
GEMM_A_PREPARED = 32


class ZutisError(RuntimeError):
    """A libzutis_b200 call returned a non-zero status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"libzutis_b200 status {status}: {message}")
        self.status = status


class ZutisBadArgument(ZutisError, ValueError):
    pass


class ZutisUnsupported(ZutisError, NotImplementedError):
    pass


_vp, _i, _l, _f, _sz = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/zutis_b200.h declares
SIGNATURES = {
    "zutis_last_error_string": (C.c_char_p, []),
    "zutis_abi_version": (_i, []),
    "zutis_device_check": (_i, [_i]),
    "zutis_gemm_workspace_bytes": (_sz, [_i, _l, _i, _i, _i]),
    "zutis_gemm_logits": (_i, [_vp, _l, _l, _vp, _l, _l, _vp, _l, _l, _l, _i, _l, _i, _i, _i, _vp, _sz, _vp]),
    "zutis_decode_score": (_i, [_vp, _l, _l, _l, _l, _i, _i, _i, _i, _i, _i, _vp, _i, _l, _vp, _vp, _i, _i, _vp]),
    "zutis_decode_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "zutis_decode_score_ws": (_i, [_vp, _l, _l, _l, _l, _i, _i, _i, _i, _i, _i, _vp, _i, _l, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "zutis_score_labels": (_i, [_vp, _i, _vp, _i, _l, _vp, _i, _vp]),
    "zutis_hist_merge": (_i, [_vp, _i, _vp, _l, _i, _vp]),
    "zutis_allreduce_hist": (_i, [_vp, _l, _vp, _vp]),
    "zutis_p2p_create": (_i, [_i, _i, _l, _vp, _vp]),
    "zutis_p2p_connect": (_i, [_i, _vp]),
    "zutis_allreduce_hist_p2p": (_i, [_i, _vp, _l, _vp, _vp]),
    "zutis_merge_allreduce_hist_p2p": (_i, [_i, _vp, _vp, _l, _vp, _vp]),
    "zutis_p2p_destroy": (_i, [_i]),
    "zutis_upsample_bilinear": (_i, [_vp, _l, _l, _l, _l, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "zutis_decode_threshold": (_i, [_vp, _l, _l, _l, _l, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "zutis_unpack_mask_bits": (_i, [_vp, _l, _i, _i, _vp, _vp]),
    "zutis_pairwise_mask_intersections": (_i, [_vp, _i, _l, _vp, _vp]),
    "zutis_instance_nms_hard": (_i, [_vp, _vp, _vp, _i, _i, C.c_double, _f, _vp, _vp, _vp]),
    "zutis_mask_rle": (_i, [_vp, _l, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "zutis_rle_to_string": (_i, [_vp, _vp, _vp, _i, _vp, _l, _vp, _vp, _vp, _vp]),
    "zutis_instance_lowres_stats": (_i, [_vp, _l, _l, _l, _l, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "zutis_instance_stats_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "zutis_instance_lowres_stats_ws": (_i, [_vp, _l, _l, _l, _l, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "zutis_instance_categories": (_i, [_vp, _l, _vp, _i, _i, _f, _vp, _vp, _vp]),
    "zutis_image_norm_workspace_bytes": (_sz, [_i, _l, _i]),
    "zutis_image_layernorm_l2norm": (_i, [_vp, _i, _l, _i, _i, _f, _f, _vp, _sz, _vp]),
    "zutis_semantic_eval_host": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i]),
    "zutis_semantic_eval_host_h2d_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i, _i, _i]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libzutis_b200.so (no build attempt, no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m zutis_b200.build` "
                "(nvcc, sm_100a).  zutis_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().zutis_last_error_string().decode("utf-8", "replace")


def check(status: int) -> None:
    if status == OK:
        return
    msg = last_error()
    if status == ERR_BAD_ARG:
        raise ZutisBadArgument(status, msg)
    if status == ERR_UNSUPPORTED:
        raise ZutisUnsupported(status, msg)
    raise ZutisError(status, msg)


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))
