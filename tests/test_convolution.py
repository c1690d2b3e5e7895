"""sn_convolve against the reference's direct group convolution (tests/golden/convolution.npz, made by
oracle/make_golden.py::convolution with the reference's Permutation class, tests/test_fourier.py:22-34) and the
convolution theorem the reference tests (tests/test_fourier.py:87-109)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, package_on_emu, rel_err


def _check(A, n, device, dt, tol):
    g = np.load(os.path.join(GOLDEN, "convolution.npz"))
    f = torch.from_numpy(g[f"n{n}_f"]).to(dt).to(device)
    h = torch.from_numpy(g[f"n{n}_g"]).to(dt).to(device)
    ref = g[f"n{n}_conv_f_g"]
    got = A.sn_convolve(f, h, n)
    assert tuple(got.shape) == tuple(f.shape)
    assert rel_err(got.cpu().numpy(), ref) < tol
    # theorem in the reference's form: FT(convolve(g, f)) = FT(f) @ FT(g)
    swapped = A.sn_fft(A.sn_convolve(h, f, n), n)
    ff, fh = A.sn_fft(f, n), A.sn_fft(h, n)
    for lam, v in swapped.items():
        want = ff[lam] * fh[lam] if ff[lam].dim() == 0 else ff[lam] @ fh[lam]
        assert rel_err(v.cpu().numpy(), want.cpu().numpy()) < 10 * tol, lam
    # batched call = per-row calls
    fb = torch.stack([f, h, f + h])
    hb = torch.stack([h, h, f])
    gb = A.sn_convolve(fb, hb, n)
    assert rel_err(gb[0].cpu().numpy(), ref) < tol
    assert rel_err(gb[2].cpu().numpy(), A.sn_convolve(f + h, f, n).cpu().numpy()) < 10 * tol


@pytest.mark.parametrize("n", [3, 4, 5])
def test_convolution_vs_reference_emulated(emu, n):
    import algebraist_b200 as A
    with package_on_emu(emu):
        _check(A, n, "cpu", torch.float64, 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("n,dt,tol", [(3, torch.float64, 1e-12), (5, torch.float64, 1e-12), (5, torch.float32, 1e-5)])
def test_convolution_vs_reference_gpu(n, dt, tol):
    import algebraist_b200 as A
    _check(A, n, "cuda", dt, tol)


@pytest.mark.gpu
def test_convolution_with_delta_is_translation_gpu():
    """n = 8: convolving with the indicator of one permutation permutes the function (norm preserved), and the
    indicator of the identity is the unit."""
    import math
    import algebraist_b200 as A
    n = 8
    nf = math.factorial(n)
    torch.manual_seed(8)
    f = torch.randn(2, nf, device="cuda")
    e = torch.zeros(2, nf, device="cuda")
    e[:, 0] = 1.0                                  # lexicographic index 0 = identity
    assert rel_err(A.sn_convolve(f, e, n).cpu().numpy(), f.cpu().numpy()) < 1e-5
    assert rel_err(A.sn_convolve(e, f, n).cpu().numpy(), f.cpu().numpy()) < 1e-5
    d = torch.zeros(2, nf, device="cuda")
    d[:, 12345] = 1.0
    t = A.sn_convolve(f, d, n)
    assert abs(t.norm().item() / f.norm().item() - 1) < 1e-5
    assert rel_err(torch.sort(t, dim=1).values.cpu().numpy(), torch.sort(f, dim=1).values.cpu().numpy()) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("n,B,dt,tol", [(10, 2, torch.float32, 2e-6), (9, 3, torch.float64, 1e-14), (7, 5, torch.float32, 2e-6)])
def test_block_matmul_vs_torch_gpu(n, B, dt, tol):
    """snfft_block_matmul (C ABI) on random packed coefficients, every partition (d up to 768: many tiles and K
    steps, ragged edges) against torch.matmul in fp64."""
    import math
    from algebraist_b200 import _lib
    from algebraist_b200.plan import get_plan
    dev = torch.device("cuda", 0)
    plan = get_plan(n, dt, dev)
    nf = math.factorial(n)
    torch.manual_seed(n)
    a = torch.randn(B * nf, dtype=dt, device=dev)
    b = torch.randn(B * nf, dtype=dt, device=dev)
    c = torch.empty_like(a)
    _lib.check(_lib.load().snfft_block_matmul(plan.handle, a.data_ptr(), b.data_ptr(), c.data_ptr(), B,
                                              torch.cuda.current_stream(dev).cuda_stream), "snfft_block_matmul")
    for d, off in zip(plan.dims, plan.offsets):
        blk = slice(off * B, (off + d * d) * B)
        want = torch.matmul(a[blk].view(B, d, d).double(), b[blk].view(B, d, d).double())
        got = c[blk].view(B, d, d).double()
        assert ((got - want).norm() / want.norm()).item() < tol, d
