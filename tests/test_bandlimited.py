"""Band-limited / subset transforms against the oracle restricted to the same partitions: on the
emulation build (CPU) and on the GPU.  The subset pipeline prunes the level steps k >= 8, so n = 8, 9 exercise it."""
import math

import numpy as np
import pytest
import torch

from conftest import package_on_emu, rel_err
from oracle import sn_oracle as O


def _oracle_inverse_of_subset(ref, keep, n):
    ft = {lam: (v if lam in keep else np.zeros_like(v)) for lam, v in ref.items()}
    return O.sn_ifft_packed(ft, n)


def _check(A, n, B, dtype, device, k=None, partitions=None):
    tol = 1e-11 if dtype == torch.float64 else 2e-5
    rng = np.random.default_rng(n * 7 + B)
    f = rng.standard_normal((B, math.factorial(n)))
    x = torch.from_numpy(f).to(dtype).to(device)
    ref = O.sn_fft_packed_c(f, n) if n >= 8 else O.sn_fft_packed(f, n)
    got = A.sn_fft_bandlimited(x, n, k=k, partitions=partitions)
    want_keys = [lam for lam in O.generate_partitions(n) if (lam[0] >= n - k if k is not None else lam in set(partitions))]
    assert list(got) == want_keys
    for lam, v in got.items():
        assert rel_err(v.cpu().numpy().reshape(ref[lam].shape), ref[lam]) < tol, lam
    back = A.sn_ifft_bandlimited(got, n)
    want = _oracle_inverse_of_subset(ref, set(want_keys), n)
    assert rel_err(back.cpu().numpy().reshape(want.shape), want) < tol
    return got


@pytest.mark.parametrize("n,B,k", [(8, 2, 1), (8, 1, 2), (6, 3, 1)])
def test_bandlimited_on_emulation(emu, n, B, k):
    with package_on_emu(emu):
        import algebraist_b200 as A
        _check(A, n, B, torch.float64, "cpu", k=k)
        assert A.bandlimited_partitions(n, k) == [lam for lam in O.generate_partitions(n) if lam[0] >= n - k]
        with pytest.raises(ValueError):
            A.sn_fft_bandlimited(torch.zeros(math.factorial(n), dtype=torch.float64), n)
        with pytest.raises(ValueError):
            A.sn_fft_bandlimited(torch.zeros(math.factorial(n), dtype=torch.float64), n, partitions=[(n + 1,)])


def test_decomposition_through_subset_plans_on_emulation(emu):
    """sn_fourier_decomposition now runs one pruned inverse per partition: the components are the oracle's and add up."""
    n, B = 8, 1
    rng = np.random.default_rng(3)
    f = rng.standard_normal((B, math.factorial(n)))
    ref = O.sn_fft_packed_c(f, n)
    with package_on_emu(emu):
        import algebraist_b200 as A
        ft = A.sn_fft(torch.from_numpy(f), n)
        comps = A.sn_fourier_decomposition(ft, n)
        for lam in [(8,), (7, 1), (4, 2, 1, 1), (1,) * 8]:
            want = _oracle_inverse_of_subset(ref, {lam}, n)
            assert rel_err(comps[lam].numpy().reshape(want.shape), want) < 1e-11, lam
        assert rel_err(sum(comps.values()).numpy(), f) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("n,B,dtype,k", [(8, 5, torch.float32, 2), (9, 4, torch.float64, 1), (9, 3, torch.float32, 3),
                                         (10, 2, torch.float64, 2), (7, 3, torch.float64, 1)])
def test_bandlimited_on_device(n, B, dtype, k):
    import algebraist_b200 as A
    _check(A, n, B, dtype, "cuda", k=k)


@pytest.mark.gpu
def test_subset_of_arbitrary_partitions_and_decomposition_on_device():
    import algebraist_b200 as A
    n = 9
    _check(A, n, 2, torch.float64, "cuda", partitions=[(4, 3, 1, 1), (9,), (3, 3, 3), (2, 1, 1, 1, 1, 1, 1, 1)])
    f = torch.randn(2, math.factorial(n), dtype=torch.float64, device="cuda")
    ft = A.sn_fft(f, n)
    comps = A.sn_fourier_decomposition(ft, n)
    assert torch.allclose(sum(comps.values()), f, atol=1e-10)
    one = A.sn_ifft_bandlimited({(5, 2, 2): ft[(5, 2, 2)]}, n)
    assert torch.allclose(one, comps[(5, 2, 2)], atol=1e-12)
    # projection property: transforming a component back gives the same block and nothing else
    back = A.sn_fft(comps[(5, 2, 2)], n)
    assert torch.allclose(back[(5, 2, 2)], ft[(5, 2, 2)], atol=1e-9)
    assert back[(4, 3, 2)].abs().max() < 1e-9
