"""Group-action utilities: sn_translate and rho against the oracle on the CPU (emulation build) and,
on the GPU, the reference's equivariance property (tests/test_fourier.py:112-138) entirely on the device at
n = 8, 9, 10 with no dimension cap."""
import ctypes
import math

import numpy as np
import pytest
import torch

from conftest import package_on_emu
from oracle import sn_oracle as O


@pytest.mark.parametrize("n", [3, 4, 5, 6])
def test_translate_and_rho_on_emulation(emu, n):
    rng = np.random.default_rng(n)
    nf = math.factorial(n)
    with package_on_emu(emu):
        import algebraist_b200 as A
        f = rng.standard_normal((2, nf))
        pi = O.perm_unrank(np.array([rng.integers(nf)]), n)[0]
        pinv = np.argsort(pi)
        perms = O.perm_unrank(np.arange(nf), n)
        left = O.perm_rank(perms[:, pinv])                 # (pi^-1 * p_k)[i] = p_k[pi^-1[i]]
        right = O.perm_rank(pinv[perms])                   # (p_k * pi^-1)[i] = pi^-1[p_k[i]]
        got_l = A.sn_translate(torch.from_numpy(f), pi, n).numpy()
        got_r = A.sn_translate(torch.from_numpy(f), torch.from_numpy(pi), n, side="right").numpy()
        assert np.array_equal(got_l, f[:, left]) and np.array_equal(got_r, f[:, right])
        for lam in O.generate_partitions(n):
            want = O.rho(lam, pi)
            got = A.rho(lam, pi, n, dtype=torch.float64, device="cpu").numpy()
            assert np.allclose(got, want.reshape(got.shape), atol=1e-13), lam
        with pytest.raises(ValueError):
            A.sn_translate(torch.from_numpy(f), [0] * n, n)
        with pytest.raises(ValueError):
            A.rho((n + 1,), pi, n, dtype=torch.float64, device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("n,dtype,tol", [(5, torch.float64, 1e-11), (8, torch.float64, 1e-11), (9, torch.float64, 1e-11),
                                         (10, torch.float32, 1e-5), (10, torch.float64, 1e-11)])
def test_equivariance_on_device(n, dtype, tol):
    """FT(left translate of f by pi)[lam] == rho_lam(pi) @ FT(f)[lam] and FT(right translate)[lam] == FT(f)[lam] @ rho_lam(pi)
    for every partition (d up to 768 at n = 10): translation, rho and both transforms run on the GPU."""
    import algebraist_b200 as A
    gen = torch.Generator(device="cuda")
    gen.manual_seed(n)
    nf = math.factorial(n)
    f = torch.randn(nf, dtype=dtype, device="cuda", generator=gen)
    pi = A.all_permutations(n)[int(torch.randint(nf, (1,)).item())] if n <= 9 else torch.randperm(n)
    ft = A.sn_fft(f, n)
    ft_l = A.sn_fft(A.sn_translate(f, pi, n), n)
    ft_r = A.sn_fft(A.sn_translate(f, pi, n, side="right"), n)
    for lam, F in ft.items():
        R = A.rho(lam, pi, n, dtype=dtype)
        if F.dim() == 0:
            want_l = want_r = R.reshape(()) * F
        else:
            want_l, want_r = R @ F, F @ R
        scale = F.norm().clamp_min(1e-30)
        assert (ft_l[lam] - want_l).norm() / scale < tol, ("left", lam)
        assert (ft_r[lam] - want_r).norm() / scale < tol, ("right", lam)


@pytest.mark.gpu
def test_rho_homomorphism_and_index_kernels_on_device():
    """rho(p * q) == rho(p) @ rho(q) (tests/test_irreps.py:108-127) and permutation_index(all_permutations) == arange
    (tests/test_permutations.py:84-87) on the GPU."""
    import algebraist_b200 as A
    n = 7
    perms = A.all_permutations(n)
    assert torch.equal(A.permutation_index(perms), torch.arange(math.factorial(n), device="cuda"))
    p, q = perms[1234], perms[4321]
    pq = q[p]                                               # (p * q)[i] = q[p[i]]
    for lam in [(4, 2, 1), (3, 2, 2), (7,), (1,) * 7, (5, 1, 1)]:
        lhs = A.rho(lam, pq, n, dtype=torch.float64)
        rhs = A.rho(lam, p, n, dtype=torch.float64) @ A.rho(lam, q, n, dtype=torch.float64)
        assert torch.allclose(lhs, rhs, atol=1e-12), lam
