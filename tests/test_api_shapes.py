"""Python wrappers (algebraist_b200.fourier): dict layout, key order and the
shape quirks of the reference, checked on CPU tensors by routing the
package through the emulation build.  Values are compared with the reference's
goldens; the shapes below were measured on the unmodified reference."""
import math

import numpy as np
import pytest
import torch

from conftest import key, load_transform_golden, package_on_emu, rel_err
from oracle import sn_oracle as O


@pytest.mark.parametrize("n", [3, 5, 6])
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_sn_fft_values_and_keys(emu, n, dt):
    import algebraist_b200 as A
    g = load_transform_golden(n)
    f = torch.from_numpy(g[f"n{n}_{dt}_f"])
    with package_on_emu(emu):
        ft = A.sn_fft(f, n)
        assert list(ft.keys()) == list(O.generate_partitions(n))
        for lam, v in ft.items():
            ref = g[f"n{n}_{dt}_ft_{key(lam)}"]
            assert tuple(v.shape) == ref.shape and v.dtype == f.dtype
            assert rel_err(v.numpy(), ref) < (1e-12 if dt == "f64" else 1e-5)
            assert v.is_contiguous()
        inv = A.sn_ifft(ft, n)
        assert tuple(inv.shape) == tuple(f.shape)
        assert rel_err(inv.numpy(), f.numpy()) < (1e-12 if dt == "f64" else 1e-5)


@pytest.mark.parametrize("n", [4, 5])
def test_shapes_small_n(emu, n):
    """n <= 5: the reference squeezes d > 1 entries (fourier.py:129); 1-d irreps keep the leading shape."""
    import algebraist_b200 as A
    nf = math.factorial(n)
    with package_on_emu(emu):
        ft = A.sn_fft(torch.randn(nf), n)
        for lam, v in ft.items():
            d = O.dim(lam)
            assert tuple(v.shape) == (() if d == 1 else (d, d))
        assert tuple(A.sn_ifft(ft, n).shape) == (nf,)
        ft = A.sn_fft(torch.randn(1, nf), n)
        for lam, v in ft.items():
            d = O.dim(lam)
            assert tuple(v.shape) == ((1,) if d == 1 else (d, d))
        assert tuple(A.sn_ifft(ft, n).shape) == (nf,)             # reference tests/test_fourier.py:61-62
        ft = A.sn_fft(torch.randn(5, nf), n)
        for lam, v in ft.items():
            d = O.dim(lam)
            assert tuple(v.shape) == ((5,) if d == 1 else (5, d, d))
        assert tuple(A.sn_ifft(ft, n).shape) == (5, nf)
        ft = A.sn_fft(torch.randn(2, 3, nf), n)
        assert tuple(ft[(n - 1, 1)].shape) == (2, 3, n - 1, n - 1)


def test_shapes_large_n(emu):
    """n >= 6: batch of one is preserved."""
    import algebraist_b200 as A
    n, nf = 6, 720
    with package_on_emu(emu):
        ft = A.sn_fft(torch.randn(nf), n)
        assert tuple(ft[(6,)].shape) == () and tuple(ft[(3, 2, 1)].shape) == (16, 16)
        assert tuple(A.sn_ifft(ft, n).shape) == (nf,)
        ft = A.sn_fft(torch.randn(1, nf), n)
        assert tuple(ft[(6,)].shape) == (1,) and tuple(ft[(3, 2, 1)].shape) == (1, 16, 16)
        assert tuple(A.sn_ifft(ft, n).shape) == (1, nf)
        ft = A.sn_fft(torch.randn(2, 2, nf, dtype=torch.float64), n)       # general leading dims (superset)
        assert tuple(ft[(5, 1)].shape) == (2, 2, 5, 5)
        assert tuple(A.sn_ifft(ft, n).shape) == (2, 2, nf)


def test_ifft_accepts_foreign_dict(emu):
    """A dict that is not a view of one packed buffer (e.g. user-built) is packed by copy."""
    import algebraist_b200 as A
    n = 5
    rng = np.random.default_rng(0)
    ft = {lam: torch.from_numpy(rng.standard_normal((2, O.dim(lam), O.dim(lam)))) for lam in O.generate_partitions(n)}
    ft[(5,)] = ft[(5,)].reshape(2)
    ft[(1, 1, 1, 1, 1)] = ft[(1, 1, 1, 1, 1)].reshape(2)
    ref = O.sn_ifft_packed({lam: v.numpy().reshape(2, O.dim(lam), O.dim(lam)) for lam, v in ft.items()}, n)
    with package_on_emu(emu):
        assert rel_err(A.sn_ifft(ft, n).numpy(), ref) < 1e-12
        assert rel_err(A.slow_sn_ift(ft, n).numpy(), ref) < 1e-12


@pytest.mark.parametrize("n", [3, 4, 5])
def test_slow_sn_ft(emu, n):
    """reference tests/test_fourier.py:141-150 (fast == dense) and slow_sn_ft's own shapes (fourier.py:96-106)."""
    import algebraist_b200 as A
    nf = math.factorial(n)
    f = torch.randn(nf, dtype=torch.float64)
    with package_on_emu(emu):
        slow, fast = A.slow_sn_ft(f, n), A.sn_fft(f, n)
        for lam in fast:
            d = O.dim(lam)
            assert tuple(slow[lam].shape) == ((1,) if d == 1 else (d, d))
            assert torch.allclose(slow[lam].reshape(fast[lam].shape), fast[lam], atol=1e-10)


def test_errors(emu):
    import algebraist_b200 as A
    with package_on_emu(emu):
        with pytest.raises(ValueError):
            A.sn_fft(torch.randn(100), 5)
        with pytest.raises(ValueError):
            A.sn_fft(torch.randn(120).half(), 5)
        with pytest.raises(ValueError):
            A.sn_fft(torch.randn(1), 1)
        with pytest.raises(ValueError):
            A.sn_ifft({(5,): torch.randn(1)}, 5)


@pytest.mark.parametrize("n", [3, 4, 5])
def test_decomposition_and_power_vs_reference(emu, n):
    """sn_fourier_decomposition (fourier.py:291-295) and calc_power (fourier.py:349-393) against the
    unmodified reference (tests/golden/next_rows.npz, made by oracle/make_golden.py)."""
    import os
    import algebraist_b200 as A
    from conftest import GOLDEN
    g = load_transform_golden(n)
    nx = np.load(os.path.join(GOLDEN, "next_rows.npz"))
    f = torch.from_numpy(g[f"n{n}_f64_f"])
    with package_on_emu(emu):
        ft = A.sn_fft(f, n)
        dec = A.sn_fourier_decomposition(ft, n)
        assert list(dec.keys()) == list(O.generate_partitions(n))
        total = 0
        for lam, v in dec.items():
            ref = nx[f"n{n}_decomp_{key(lam)}"]
            assert tuple(v.shape) == ref.shape
            assert np.allclose(v.numpy(), ref, atol=1e-12), lam
            total = total + v
        assert np.allclose(total.numpy(), f.numpy(), atol=1e-12)           # reference tests/test_fourier.py:66-72
        pw = A.calc_power(ft, n)
        for lam, v in pw.items():
            ref = nx[f"n{n}_power_{key(lam)}"]
            assert tuple(v.shape) == ref.shape and np.allclose(v.numpy(), ref, rtol=1e-12), lam
        assert np.allclose(sum(pw.values()).numpy(), (f ** 2).sum(1).numpy())   # tests/test_fourier.py:75-84
        one = A.sn_fft(f[0], n)
        pw1 = A.calc_power(one, n)
        for lam, v in pw1.items():
            ref = nx[f"n{n}_power1_{key(lam)}"]
            assert tuple(v.shape) == ref.shape and np.allclose(v.numpy(), ref, rtol=1e-12), lam


def test_decomposition_large_n(emu):
    """n = 6 (the reference raises there): the components still add up to the inverse transform and
    Plancherel holds for calc_power."""
    import algebraist_b200 as A
    n = 6
    torch.manual_seed(6)
    f = torch.randn(2, math.factorial(n), dtype=torch.float64)
    with package_on_emu(emu):
        ft = A.sn_fft(f, n)
        dec = A.sn_fourier_decomposition(ft, n)
        assert np.allclose(sum(dec.values()).numpy(), f.numpy(), atol=1e-12)
        pw = A.calc_power(ft, n)
        assert all(tuple(v.shape) == (2,) for v in pw.values())
        assert np.allclose(sum(pw.values()).numpy(), (f ** 2).sum(1).numpy())
