"""zutis_b200 -- B200 (sm_100a) implementation of ZUTIS's dense mask-decode + scoring path.

Host side: drop-ins for the reference's call surface on this path
(``predict`` / ``get_mask_proposals`` of networks/zutis.py, ``RunningScore`` of
utils/running_score.py, ``compute_iou`` of utils/iou.py).  Device side: hand-written CUDA kernels
in ``csrc/`` behind the C ABI of ``include/zutis_b200.h`` (``libzutis_b200.so``), reached through
ctypes.  PyTorch supplies device memory, streams and ``torch.distributed`` only.

Importing this package does not need a GPU; calling into it does (there is no CPU fallback).
"""
from . import _ffi
from ._ffi import ZutisBadArgument, ZutisError, ZutisUnsupported

__all__ = [
    "RunningScore", "compute_iou", "predict", "get_mask_proposals", "decode_and_score", "ZutisDecoder", "StreamingScorer",
    "install", "shard_range", "ZutisError", "ZutisBadArgument", "ZutisUnsupported",
]


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import zutis_b200` stays cheap
    import importlib
    lazy = {
        "RunningScore": "running_score", "compute_iou": "iou",
        "predict": "decode", "get_mask_proposals": "decode", "decode_and_score": "decode",
        "ZutisDecoder": "decode", "StreamingScorer": "decode", "install": "decode", "image_to_text_space": "decode",
        "shard_range": "distributed", "init_distributed": "distributed",
    }
    if name in lazy:
        return getattr(importlib.import_module(f"{__name__}.{lazy[name]}"), name)
    if name in ("ops", "distributed", "running_score", "iou"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
