"""CPU: the C-ABI library loads, exports every symbol include/vats_attn.h declares, fails loudly without a GPU,
and its host-side tile-skipping arithmetic never drops an allowed (i, j) pair."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from oracle import mask_predicate
from vats_multimodal_lm_b200 import _ffi


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "vats_attn.h")).read()
    declared = set(re.findall(r"\b(vats_attn_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_ffi.EXPORTED_SYMBOLS), declared ^ set(_ffi.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} not exported"
    assert _ffi.version() == 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_compute_calls_fail_loudly_without_gpu():
    z3 = (0, 0, 0)
    with pytest.raises(_ffi.VatsAttnError) as e:
        _ffi.prefill(16, 16, 16, 16, None, None, 1, 4, 4, 2, 1, 8, z3, z3, z3, z3, 1.0, True, -1, -1, None)
    assert e.value.code in (2, 3)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        torch.ops.vats.gqa_swa_prefill(torch.zeros(1, 2, 2, 8, dtype=torch.bfloat16),
                                       torch.zeros(1, 2, 1, 8, dtype=torch.bfloat16),
                                       torch.zeros(1, 2, 1, 8, dtype=torch.bfloat16), None, None, 1.0, True, -1, -1, 0)


def test_argument_validation_precedes_device_check():
    z3 = (8, 8, 8)
    with pytest.raises(_ffi.VatsAttnError) as e:
        _ffi.prefill(16, 16, 16, 16, None, None, 1, 4, 4, 3, 2, 8, z3, z3, z3, z3, 1.0, True, -1, -1, None)
    assert e.value.code == 1 and "divisible" in str(e.value)
    with pytest.raises(_ffi.VatsAttnError) as e:
        _ffi.prefill(16, 16, 16, 16, None, None, 1, 4, 4, 2, 2, 8, z3, z3, z3, z3, -1.0, True, -1, -1, None)
    assert e.value.code == 1 and "scale" in str(e.value)


def test_prepare_entry_points_validate_before_touching_the_device():
    """vats_attn_decode_prepare / vats_attn_prefill_prepare: argument errors are reported with code 1 (or 2 for an
    unsupported geometry) whether or not a GPU is present; on a box without one a well-formed call fails loudly."""
    z2, z3 = (8, 8), (8, 8, 8)
    with pytest.raises(_ffi.VatsAttnError) as e:   # in_dtype must be 0 / 1 (raw call: the wrapper takes a bool)
        _ffi._check(_ffi.load().vats_attn_decode_prepare(16, 16, 16, 7, 16, 16, 16, 16, None, None, 1, 2, 1, 8, 4,
                                                          _ffi._i64x2(*z2), _ffi._i64x2(*z2), _ffi._i64x2(*z2),
                                                          _ffi._i64x2(*z2), _ffi._i64x3(*z3), _ffi._i64x3(*z3), 1, 1e-6,
                                                          None))
    assert e.value.code == 1 and "in_dtype" in str(e.value)
    with pytest.raises(_ffi.VatsAttnError) as e:   # cos without sin
        _ffi.decode_prepare(16, 16, 16, False, 16, 16, 16, 16, 16, None, 1, 2, 1, 8, 4, z2, z2, z2, z2, z3, z3, True,
                            1e-6, None)
    assert e.value.code == 1 and "cos_table" in str(e.value)
    with pytest.raises(_ffi.VatsAttnError) as e:   # RoPE needs an even head_dim
        _ffi.prefill_prepare(16, 16, 16, True, 16, 16, 16, 16, 16, 1, 4, 2, 1, 7, 0, z3, z3, z3, z3, z3, z3, True, 1e-6,
                             None)
    assert e.value.code == 2 and "even head_dim" in str(e.value)
    with pytest.raises(_ffi.VatsAttnError) as e:   # head_dim > 256
        _ffi.prefill_prepare(16, 16, 16, True, 16, 16, 16, None, None, 1, 4, 2, 1, 300, 0, z3, z3, z3, z3, z3, z3, True,
                             1e-6, None)
    assert e.value.code == 2
    if not torch.cuda.is_available():
        with pytest.raises(_ffi.VatsAttnError) as e:
            _ffi.prefill_prepare(16, 16, 16, True, 16, 16, 16, None, None, 1, 4, 2, 1, 8, 0, z3, z3, z3, z3, z3, z3,
                                 True, 1e-6, None)
        assert e.value.code in (2, 3)
        with pytest.raises(RuntimeError):
            torch.ops.vats.prefill_prepare(torch.zeros(1, 2, 2, 8), torch.zeros(1, 2, 1, 8), torch.zeros(1, 2, 1, 8),
                                           None, None, 0, True, 1e-6)


CASES = [
    # Tq, Tk, causal, left, right
    (300, 300, True, -1, -1), (300, 300, True, 64, 0), (300, 300, True, 0, 0), (300, 300, False, 40, 40),
    (300, 300, False, -1, -1), (1, 500, True, 128, 0), (100, 500, True, 200, 0), (500, 100, True, -1, 0),
    (257, 129, False, 5, -1), (129, 257, False, -1, 7), (64, 1000, True, 1000, 0), (1000, 64, False, 10, 10),
    (130, 130, True, 129, 0), (130, 130, True, 130, 0), (130, 130, True, 1, 0),
]


@pytest.mark.parametrize("Tq,Tk,causal,left,right", CASES)
@pytest.mark.parametrize("bm,bn", [(128, 128), (8, 32), (32, 32)])
def test_tile_range_covers_every_allowed_pair(Tq, Tk, causal, left, right, bm, bn):
    m = mask_predicate(1, Tq, Tk, causal, left, right)[0]
    n_kt = (Tk + bn - 1) // bn
    for q0 in range(0, Tq, bm):
        first, last = _ffi.debug_tile_range(q0, bm, bn, Tq, Tk, causal, left, right)
        rows = m[q0:q0 + bm]
        for t in range(n_kt):
            blk = rows[:, t * bn:(t + 1) * bn]
            visited = first <= t <= last
            if blk.any():
                assert visited, f"tile {t} of q-block {q0} holds allowed pairs but is skipped"
            if visited and _ffi.debug_tile_is_full(t, q0, bm, bn, Tq, Tk, causal, left, right):
                assert blk.all() and blk.shape[1] == bn, f"tile {t} of q-block {q0} flagged full but is masked"
        # tightness: the first and last visited tiles are not both needlessly empty for band masks
        if first <= last:
            assert rows[:, first * bn:(last + 1) * bn].any()


def test_decode_workspace_size_contract():
    # large enough for either decode kernel: the CUDA-core kernel's 256-key splits (17 of them for a 4097-key window)
    assert _ffi.decode_workspace_bytes(64, 32, 8, 128, 8192, 4096) >= 64 * 32 * 17 * 130 * 4
    assert 16 <= _ffi.decode_workspace_bytes(2, 4, 2, 64, 100, -1) < 1 << 20
    assert _ffi.decode_workspace_bytes(0, 4, 2, 64, 100, -1) == 0
    assert _ffi.decode_workspace_bytes(2, 4, 3, 64, 100, -1) == 0   # H % G != 0
