"""Host plan (csrc/plan.cpp) read back through the C ABI, against the oracle and
the reference's golden tables."""
import ctypes
import math

import numpy as np
import pytest

from conftest import CAbi, key
from oracle import sn_oracle as O


def _level(lib, plan, k):
    npart = ctypes.c_int()
    parts = ctypes.POINTER(ctypes.c_int32)()
    dims = ctypes.POINTER(ctypes.c_int32)()
    assert lib.snfft_plan_level_info(plan, k, ctypes.byref(npart), ctypes.byref(parts), ctypes.byref(dims)) == 0
    ps = [tuple(v for v in (parts[r * k + i] for i in range(k)) if v > 0) for r in range(npart.value)]
    return ps, [dims[r] for r in range(npart.value)]


@pytest.mark.parametrize("n", [2, 5, 8, 10, 11])
def test_partitions_dims_offsets(emu, n, golden_tables):
    plan = CAbi(emu).plan(n, np.float32)
    try:
        for k in range(1, n + 1):
            ps, ds = _level(emu, plan, k)
            assert ps == list(O.generate_partitions(k))
            assert ds == [O.dim(p) for p in ps]
            flat = golden_tables[f"partitions_n{k}"]
            assert ps == [tuple(int(v) for v in row if v > 0) for row in flat]
        npart = ctypes.c_int()
        offs = ctypes.POINTER(ctypes.c_int64)()
        nn, dt = ctypes.c_int(), ctypes.c_int()
        assert emu.snfft_plan_info(plan, ctypes.byref(nn), ctypes.byref(dt), ctypes.byref(npart), None, None,
                                   ctypes.byref(offs)) == 0
        assert nn.value == n and dt.value == 0
        ds = [O.dim(p) for p in O.generate_partitions(n)]
        assert [offs[i] for i in range(npart.value + 1)] == [0] + list(np.cumsum([d * d for d in ds]))
        assert offs[npart.value] == math.factorial(n)
    finally:
        emu.snfft_plan_destroy(plan)


@pytest.mark.parametrize("n", [4, 6, 8])
def test_yor_tables(emu, n, golden_tables):
    plan = CAbi(emu).plan(n, np.float64)
    try:
        for k in range(2, n + 1):
            for s, lam in enumerate(O.generate_partitions(k)):
                for j in range(k - 1):
                    d = ctypes.c_int()
                    partner = ctypes.POINTER(ctypes.c_int32)()
                    diag = ctypes.POINTER(ctypes.c_double)()
                    off = ctypes.POINTER(ctypes.c_double)()
                    assert emu.snfft_plan_yor_table(plan, k, s, j, ctypes.byref(d), ctypes.byref(partner),
                                                    ctypes.byref(diag), ctypes.byref(off)) == 0
                    dd = d.value
                    p = np.ctypeslib.as_array(partner, (dd,))
                    a = np.ctypeslib.as_array(diag, (dd,))
                    b = np.ctypeslib.as_array(off, (dd,))
                    op, oa, ob = O.transposition_table(lam, j)
                    assert np.array_equal(p, op) and np.array_equal(a, oa) and np.array_equal(b, ob), (lam, j)
                    if k <= 6:          # dense matrix of the reference itself (irreps.py:91-109), bit exact
                        m = np.diag(a.copy())
                        rows = np.arange(dd)
                        m[rows, p] += np.where(p != rows, b, 0.0)
                        assert np.array_equal(m, golden_tables[f"n{k}_yor_{key(lam)}_{j}"])
    finally:
        emu.snfft_plan_destroy(plan)


def test_bad_arguments(emu):
    h = ctypes.c_void_p()
    assert emu.snfft_plan_create(1, 0, 0, ctypes.byref(h)) == -1
    assert emu.snfft_plan_create(13, 0, 0, ctypes.byref(h)) == -1
    assert emu.snfft_plan_create(5, 7, 0, ctypes.byref(h)) == -1
    assert b"dtype" in emu.snfft_last_error()
    plan = CAbi(emu).plan(4, np.float32)
    f = np.zeros((1, 24), dtype=np.float32)
    out = np.zeros(24, dtype=np.float32)
    ws = np.zeros(8, dtype=np.uint8)
    assert emu.snfft_forward(plan, f.ctypes.data, 1, out.ctypes.data, ws.ctypes.data, ws.nbytes, None) == -4
    assert emu.snfft_forward(plan, f.ctypes.data, -1, out.ctypes.data, ws.ctypes.data, ws.nbytes, None) == -1
    assert emu.snfft_forward(plan, None, 1, out.ctypes.data, ws.ctypes.data, ws.nbytes, None) == -1
    assert emu.snfft_forward(plan, f.ctypes.data, 0, out.ctypes.data, None, 0, None) == 0     # empty batch
    emu.snfft_plan_destroy(plan)
