"""The CPU oracle against the reference's golden outputs (tests/golden, made by
oracle/make_golden.py from the unmodified reference) and the properties the
reference's own tests check (tests/test_fourier.py:55-159 there)."""
import math

import numpy as np
import pytest

from conftest import key, load_transform_golden, rel_err
from oracle import sn_oracle as O


@pytest.mark.parametrize("n", range(1, 13))
def test_partition_order(n, golden_tables):
    flat = golden_tables[f"partitions_n{n}"]
    ref = [tuple(int(v) for v in row if v > 0) for row in flat]
    assert list(O.generate_partitions(n)) == ref


@pytest.mark.parametrize("n", range(2, 8))
def test_dims(n, golden_tables):
    assert [O.dim(p) for p in O.generate_partitions(n)] == list(golden_tables[f"n{n}_dims"])


def test_known_dims():
    # reference tests/test_tableau.py:67-84
    assert O.dim((4, 3, 2, 1)) == 768 and O.dim((5, 4, 1)) == 288 and O.dim((3, 2)) == 5
    for n in range(1, 11):
        assert sum(O.dim(p) ** 2 for p in O.generate_partitions(n)) == math.factorial(n)


@pytest.mark.parametrize("n", range(2, 7))
def test_yor_matrices_bit_exact(n, golden_tables):
    for lam in O.generate_partitions(n):
        for j in range(n - 1):
            assert np.array_equal(O.transposition_matrix(lam, j), golden_tables[f"n{n}_yor_{key(lam)}_{j}"])


@pytest.mark.parametrize("n", range(2, 8))
def test_coset_index_and_rank(n, golden_tables):
    perms = O.all_permutations(n)
    assert np.array_equal(O.perm_rank(perms), np.arange(math.factorial(n)))   # test_permutations.py:84-87
    for i in range(n):
        assert np.array_equal(O.coset_index(n, i), golden_tables[f"n{n}_coset_{i}"].reshape(-1))
    if n <= 6:
        assert np.array_equal(O.perm_unrank(np.arange(math.factorial(n)), n), perms)


@pytest.mark.parametrize("n", range(2, 9))
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_forward_matches_reference(n, dt):
    g = load_transform_golden(n)
    f = g[f"n{n}_{dt}_f"]
    mine = O.sn_fft_packed(f, n)
    tol = 1e-12 if dt == "f64" else 1e-5
    for lam in O.generate_partitions(n):
        ref = g[f"n{n}_{dt}_ft_{key(lam)}"]
        assert rel_err(mine[lam].reshape(ref.shape), ref) < tol, lam


@pytest.mark.parametrize("n", range(2, 8))
def test_inverse_matches_reference(n):
    g = load_transform_golden(n)
    ft = {lam: g[f"n{n}_f64_ft_{key(lam)}"] for lam in O.generate_partitions(n)}
    B = g[f"n{n}_f64_f"].shape[0]
    ft = {lam: v.reshape(B, O.dim(lam), O.dim(lam)) for lam, v in ft.items()}
    inv = O.sn_ifft_packed(ft, n)
    assert rel_err(inv, g[f"n{n}_f64_ift"]) < 1e-12
    assert rel_err(inv, g[f"n{n}_f64_f"]) < 1e-12


@pytest.mark.parametrize("n", [3, 4, 5])
def test_fast_equals_dense(n):
    rng = np.random.default_rng(n)
    f = rng.standard_normal((2, math.factorial(n)))
    fast, slow = O.sn_fft_packed(f, n), O.slow_sn_ft_packed(f, n)
    for lam in fast:
        assert np.allclose(fast[lam], slow[lam], atol=1e-11)
    assert np.allclose(O.slow_sn_ift_packed(fast, n), f, atol=1e-11)


@pytest.mark.parametrize("n", [4, 5, 6])
def test_plancherel(n):
    # reference tests/test_fourier.py:75-84
    rng = np.random.default_rng(10 + n)
    f = rng.standard_normal((3, math.factorial(n)))
    ft = O.sn_fft_packed(f, n)
    power = sum(O.dim(lam) / math.factorial(n) * (v ** 2).sum(axis=(1, 2)) for lam, v in ft.items())
    assert np.allclose(power, (f ** 2).sum(1), rtol=1e-12)


@pytest.mark.parametrize("n", [4, 5, 6])
def test_left_translation_equivariance(n):
    # reference tests/test_fourier.py:112-138: f_perm[k] = f[idx(pi^-1 * p_k)]  =>  FT(f_perm) = rho(pi) FT(f)
    rng = np.random.default_rng(20 + n)
    perms = O.all_permutations(n)
    f = rng.standard_normal((1, math.factorial(n)))
    pi = perms[rng.integers(len(perms))]
    pinv = np.argsort(pi)
    # (pinv * p)[i] = p[pinv[i]]  (permutations.py:85-88)
    action = O.perm_rank(perms[:, pinv])
    ft, ftp = O.sn_fft_packed(f, n), O.sn_fft_packed(f[:, action], n)
    for lam in ft:
        assert np.allclose(ftp[lam][0], O.rho(lam, pi) @ ft[lam][0], atol=1e-10), lam


@pytest.mark.parametrize("n", [3, 4])
def test_convolution_theorem(n):
    # reference tests/test_fourier.py:22-34, 87-109: (g*f)(pi) = sum_sigma g(sigma) f(pi sigma^-1)
    rng = np.random.default_rng(30 + n)
    perms = O.all_permutations(n)
    N = len(perms)
    f, g = rng.standard_normal(N), rng.standard_normal(N)
    conv = np.zeros(N)
    for i, pi in enumerate(perms):
        for j, sg in enumerate(perms):
            sinv = np.argsort(sg)
            prod = sinv[pi]                     # (pi * sinv)[x] = sinv[pi[x]]
            conv[i] += g[j] * f[O.perm_rank(prod[None])[0]]
    fc = O.sn_fft_packed(conv[None], n)
    ff, fg = O.sn_fft_packed(f[None], n), O.sn_fft_packed(g[None], n)
    for lam in fc:
        assert np.allclose(fc[lam][0], ff[lam][0] @ fg[lam][0], atol=1e-10), lam


def test_n9_roundtrip_and_plancherel():
    n = 9
    rng = np.random.default_rng(9)
    f = rng.standard_normal((1, math.factorial(n)))
    ft = O.sn_fft_packed(f, n)
    power = sum(O.dim(lam) / math.factorial(n) * (v ** 2).sum() for lam, v in ft.items())
    assert abs(power - (f ** 2).sum()) < 1e-9 * power
    assert rel_err(O.sn_ifft_packed(ft, n), f) < 1e-12
