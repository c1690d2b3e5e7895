"""The oracle's C inner loops (oracle/sn_oracle_c.c) against the numpy oracle and the reference goldens."""
import math

import numpy as np
import pytest

from conftest import key, load_transform_golden, rel_err
from oracle import sn_oracle as O


@pytest.mark.parametrize("n", [4, 6, 8])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_c_matches_numpy_bitwise(n, dt):
    f = np.random.default_rng(n).standard_normal((2, math.factorial(n))).astype(dt)
    a, b = O.sn_fft_packed(f, n), O.sn_fft_packed_c(f, n)
    if dt == np.float64:
        for lam in a:
            assert np.array_equal(a[lam], b[lam]), lam           # same operations in the same order
        assert np.array_equal(O.sn_ifft_packed(a, n), O.sn_ifft_packed_c(b, n))
    else:
        # the numpy path promotes float32 data to float64 (its tables are float64); the C path stays in float32
        for lam in a:
            assert b[lam].dtype == np.float32 and rel_err(b[lam], a[lam]) < 1e-5, lam
        assert rel_err(O.sn_ifft_packed_c(b, n), f) < 1e-5


@pytest.mark.parametrize("n", [5, 7])
def test_c_matches_reference_golden(n):
    g = load_transform_golden(n)
    f = g[f"n{n}_f64_f"]
    ft = O.sn_fft_packed_c(f, n)
    for lam in O.generate_partitions(n):
        ref = g[f"n{n}_f64_ft_{key(lam)}"]
        assert rel_err(ft[lam].reshape(ref.shape), ref) < 1e-12
    assert rel_err(O.sn_ifft_packed_c(ft, n), g[f"n{n}_f64_ift"]) < 1e-12
