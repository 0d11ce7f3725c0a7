"""Device-wide primitives under the broadphase (csrc/prims.cuh) against numpy."""
import ctypes as C

import numpy as np
import pytest

from chipmunk2d_b200.engine import load_engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 255, 256, 257, 2047, 2048, 2049, 100003, 1 << 20])
@pytest.mark.parametrize("bits", [32, 45])
def test_radix_sort_pairs(n, bits):
    lib = load_engine()
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 1 << bits, size=n, dtype=np.uint64)
    if n > 10:
        keys[: n // 3] = keys[0]          # runs of duplicates: the sort must be stable
    vals = np.arange(n, dtype=np.int32)
    k = keys.copy(); v = vals.copy()
    rc = lib.cpb200_debug_sort_pairs(0, n, bits, k.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
    assert rc == 0, lib.cpb200_last_error()
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])


@pytest.mark.parametrize("n", [0, 1, 5, 1023, 1024, 1025, 65536, 1048577, 3000001])
def test_exclusive_scan(n):
    lib = load_engine()
    rng = np.random.default_rng(n)
    data = rng.integers(0, 1000, size=n, dtype=np.uint32)
    d = data.copy()
    rc = lib.cpb200_debug_exclusive_scan(0, n, d.ctypes.data_as(C.c_void_p))
    assert rc == 0, lib.cpb200_last_error()
    expect = np.concatenate([[0], np.cumsum(data, dtype=np.uint64)[:-1]]).astype(np.uint32) if n else data
    assert np.array_equal(d, expect)
