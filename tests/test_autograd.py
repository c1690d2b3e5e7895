"""Backpropagation through sn_fft / sn_ifft.  The reference transforms are compositions of differentiable torch
ops (fourier.py:111-131, 223-278); here the backward passes are the adjoint pipelines (fourier.py `_ForwardFn`,
`_InverseFn`).  Checked against gradients taken through the dense path (`slow_sn_ft` / `slow_sn_ift`: einsums over
rho(sigma), differentiable by torch itself).  CPU: the package routed through the emulation build; GPU: the product."""
import math

import pytest
import torch

from conftest import package_on_emu


def _grad_checks(A, n, B, dt, device, tol):
    nf = math.factorial(n)
    gen = torch.Generator().manual_seed(n * 10 + B)
    f = torch.randn(B, nf, dtype=dt, generator=gen).to(device)
    # forward: loss = sum_lambda <W_lambda, F_lambda>
    f1 = f.clone().requires_grad_(True)
    ft = A.sn_fft(f1, n)
    weights = {lam: torch.randn(v.shape, dtype=dt, generator=gen).to(device) for lam, v in ft.items()}
    sum((weights[lam] * v).sum() for lam, v in ft.items()).backward()
    f2 = f.clone().requires_grad_(True)
    slow = A.slow_sn_ft(f2, n)
    sum((weights[lam] * slow[lam].reshape(weights[lam].shape)).sum() for lam in slow).backward()
    assert (f1.grad - f2.grad).norm() / f2.grad.norm() < tol
    # inverse: loss = <w, sn_ifft(ft)>, gradients with respect to every block
    base = {lam: v.detach() for lam, v in A.sn_fft(f, n).items()}
    w = torch.randn(B, nf, dtype=dt, generator=gen).to(device)
    a = {lam: v.clone().requires_grad_(True) for lam, v in base.items()}
    (w * A.sn_ifft(a, n).reshape(B, nf)).sum().backward()
    b = {lam: v.clone().requires_grad_(True) for lam, v in base.items()}
    (w * A.slow_sn_ift(b, n).reshape(B, nf)).sum().backward()
    for lam in a:
        assert (a[lam].grad - b[lam].grad).norm() / b[lam].grad.norm() < tol, lam
    # no graph is recorded when nothing requires grad
    assert not A.sn_fft(f, n)[next(iter(base))].requires_grad


@pytest.mark.parametrize("n,B", [(4, 3), (6, 2)])
def test_autograd_matches_dense_path_emulated(emu, n, B):
    import algebraist_b200 as A
    with package_on_emu(emu):
        _grad_checks(A, n, B, torch.float64, "cpu", 1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("n,B,dt,tol", [(5, 4, torch.float64, 1e-11), (7, 2, torch.float64, 1e-11),
                                        (7, 3, torch.float32, 2e-5)])
def test_autograd_matches_dense_path_gpu(n, B, dt, tol):
    import algebraist_b200 as A
    _grad_checks(A, n, B, dt, "cuda", tol)
