"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/zutis_b200.h declares, the ctypes table covers them, and the host logic that needs no
GPU (sharding, size parsing, loud failure without a device) behaves."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import zutis_b200
from zutis_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zutis_b200.h")).read()
    return sorted(set(re.findall(r"ZUTIS_API\s+[\w\s\*]+?\b(zutis_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ["zutis_gemm_logits", "zutis_decode_score", "zutis_hist_merge", "zutis_decode_threshold",
                 "zutis_last_error_string", "zutis_semantic_eval_host"]:
        assert must in syms
    assert len(syms) >= 15


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_ffi.LIB_PATH), "libzutis_b200.so missing: run `python -m zutis_b200.build`"
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/zutis_b200.h but not exported"


def test_ctypes_table_matches_header():
    assert sorted(_ffi.SIGNATURES) == declared_symbols()
    _ffi.lib()                                   # binds every signature; raises on a missing symbol
    assert _ffi.lib().zutis_abi_version() == 1


def test_argument_counts_match_header():
    text = open(os.path.join(ROOT, "include", "zutis_b200.h")).read()
    for name, (_, args) in _ffi.SIGNATURES.items():
        m = re.search(r"ZUTIS_API\s+[\w\s\*]+?\b" + name + r"\s*\(([^;]*?)\)\s*;", text, re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, ctypes table has {len(args)}"


def test_host_entry_h2d_bytes_accounts_for_narrowed_labels():
    """int64 / int32 labels cross PCIe as uint8 (n <= 255) or int16; no histogram requested -> the caller's type."""
    f = _ffi.lib().zutis_semantic_eval_host_h2d_bytes
    B, Q, D, h, w, H, W = 4, 81, 512, 40, 40, 320, 320
    fixed = Q * D * 4 + B * h * w * D * 4
    assert f(_ffi.GT_I64, 1, B, Q, D, h, w, H, W) == fixed + B * H * W
    assert f(_ffi.GT_I32, 1, B, Q, D, h, w, H, W) == fixed + B * H * W
    assert f(_ffi.GT_U8, 1, B, Q, D, h, w, H, W) == fixed + B * H * W
    assert f(_ffi.GT_I64, 1, B, 920, D, h, w, H, W) == 920 * D * 4 + B * h * w * D * 4 + 2 * B * H * W
    assert f(_ffi.GT_I64, 0, B, Q, D, h, w, H, W) == fixed + 8 * B * H * W


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu():
    assert _ffi.lib().zutis_device_check(0) == _ffi.ERR_NO_DEVICE
    assert "no CPU fallback" in _ffi.last_error()
    with pytest.raises(RuntimeError):
        zutis_b200.RunningScore(3)
    with pytest.raises(RuntimeError):
        zutis_b200.compute_iou(np.zeros((2, 2), bool), np.zeros((2, 2), bool))
    with pytest.raises(TypeError):
        zutis_b200.ops.decode_score(torch.zeros(1, 2, 3, 3), (6, 6))
    with pytest.raises(TypeError):
        zutis_b200.ZutisDecoder(torch.zeros(3, 8))


def test_status_codes_map_to_exceptions():
    # argument validation happens before any device work, so it can be exercised without a GPU
    L = _ffi.lib()
    st = L.zutis_decode_score(None, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, None, 0, 0, None, None, 0, 0, None)
    assert st == _ffi.ERR_BAD_ARG and "logits is NULL" in _ffi.last_error()
    with pytest.raises(zutis_b200.ZutisBadArgument):
        _ffi.check(st)
    st = L.zutis_hist_merge(None, 1, None, 4, 1, None)
    assert st == _ffi.ERR_BAD_ARG
    buf = (ctypes.c_float * 4)()
    st = L.zutis_gemm_logits(buf, 2, 0, buf, 2, 0, buf, 1, 1, 1, 0, 1, 2, 1, 0, None, 0, None)
    assert st == _ffi.ERR_BAD_ARG and "non-positive" in _ffi.last_error()


def test_shard_range_partitions_exactly():
    from zutis_b200.distributed import shard_range
    for n in (0, 1, 7, 64, 65, 5000):
        for g in (1, 2, 3, 8):
            spans = [shard_range(n, r, g) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_size_pair_accepts_reference_forms():
    from zutis_b200.ops import size_pair
    assert size_pair((3, 4)) == (3, 4)
    assert size_pair(torch.zeros(1, 3, 5, 7).shape[-2:]) == (5, 7)
    assert size_pair([torch.tensor([97]), torch.tensor([131])]) == (97, 131)      # trainer.py:322-323
    assert size_pair(None) is None


def test_scores_from_counts_matches_oracle(golden):
    from oracle import oracle as O
    from zutis_b200.distributed import scores_from_counts
    g = golden("model_cfg1")
    s, c = scores_from_counts(g["confusion"])
    so, co = O.scores_from_hist(g["confusion"])
    assert s == so
    assert np.array_equal(np.array(list(c.values())), np.array(list(co.values())), equal_nan=True)
    assert np.array_equal(np.array([s["Pixel Acc"], s["Mean Acc"], s["FreqW Acc"], s["Mean IoU"]]), g["scores"])
