"""GPU parity tests (run on the B200 box: pytest -m gpu).

Every test calls the CUDA kernels through the C ABI (ctypes, the same binding
the package uses) or through the Python drop-in API, and compares with
  * the reference's golden outputs (tests/golden, n = 2..8), and
  * the CPU oracle (oracle/sn_oracle.py) on the same seeded inputs (n = 9, 10),
plus size-independent properties at BASELINE.json's full sizes (round trip,
Plancherel, linearity, equivariance).
Tolerances (BASELINE.json north_star): fp32 relative 1e-5, fp64 relative 1e-12,
relative = ||delta||_F / ||ref||_F per partition."""
import ctypes
import math

import numpy as np
import pytest
import torch

from conftest import key, load_transform_golden, rel_err
from oracle import sn_oracle as O

pytestmark = pytest.mark.gpu

TDT = {"f32": torch.float32, "f64": torch.float64}
TOL = {"f32": 1e-5, "f64": 1e-12}


@pytest.fixture(scope="module")
def A():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import algebraist_b200
    from algebraist_b200 import _lib
    lib = _lib.load()                       # raises if libsnfft_b200.so is missing: no fallback
    assert b"sm_100a" in lib.snfft_build_info()
    return algebraist_b200


def _check_dict(ft, ref, tol, B=None):
    for lam, r in ref.items():
        v = ft[lam].detach().cpu().numpy()
        assert rel_err(v.reshape(np.shape(r)), r) < tol, (lam, rel_err(v.reshape(np.shape(r)), r))


@pytest.mark.parametrize("n", range(2, 9))
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_forward_inverse_vs_reference_golden(A, n, dt):
    g = load_transform_golden(n)
    f = torch.from_numpy(g[f"n{n}_{dt}_f"]).cuda()
    before = A._lib.load().snfft_launch_count()
    ft = A.sn_fft(f, n)
    assert A._lib.load().snfft_launch_count() > before
    assert list(ft.keys()) == list(O.generate_partitions(n))
    ref = {lam: g[f"n{n}_{dt}_ft_{key(lam)}"] for lam in ft}
    for lam, v in ft.items():
        assert v.is_cuda and v.dtype == TDT[dt] and tuple(v.shape) == ref[lam].shape
    _check_dict(ft, ref, TOL[dt])
    inv = A.sn_ifft(ft, n)
    assert rel_err(inv.cpu().numpy(), f.cpu().numpy()) < TOL[dt]
    if n <= 7:
        assert rel_err(inv.cpu().numpy(), g[f"n{n}_{dt}_ift"]) < TOL[dt]


@pytest.mark.parametrize("n,B", [(9, 3), (10, 2), (9, 4), (10, 4)])
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_vs_oracle_large_n(A, n, B, dt):
    """B = 4 (fp64: also B = 2): full instance vectors, so the top level's T kernel reads / writes the packed
    coefficients itself; B = 3 (fp32: also 2): the separate pack / unpack kernels."""
    rng = np.random.default_rng(100 * n + B)
    f = rng.standard_normal((B, math.factorial(n)))
    ref = O.sn_fft_packed(f, n)
    x = torch.from_numpy(f).to(TDT[dt]).cuda()
    ft = A.sn_fft(x, n)
    _check_dict(ft, ref, TOL[dt])
    inv = A.sn_ifft(ft, n)
    assert rel_err(inv.cpu().numpy(), f) < TOL[dt]
    # inverse of arbitrary coefficients against the oracle's adjoint pipeline
    coef = {lam: rng.standard_normal(v.shape) for lam, v in ref.items()}
    ref_inv = O.sn_ifft_packed(coef, n)
    as_api = {lam: torch.from_numpy(v.reshape(B) if O.dim(lam) == 1 else v).to(TDT[dt]).cuda()
              for lam, v in coef.items()}
    got = A.sn_ifft(as_api, n)
    assert rel_err(got.cpu().numpy(), ref_inv) < TOL[dt]


@pytest.mark.parametrize("n,B,rows", [(10, 128, (0, 77, 127)), (9, 1024, (0, 511, 1023))])
def test_full_size_batches_rows_vs_oracle(A, n, B, rows):
    """BASELINE configs 3 and 4 at their FULL batch sizes (S_9 x 1024, S_10 x 128, fp32): the rows `rows` of the
    transform of the whole batch against the oracle's transform of those functions (every partition, 1e-5), and
    the same rows of the inverse of the whole coefficient set against the inputs.  (The first / a middle / the last
    row sit in different instance vectors, tiles and CTAs of every kernel.)"""
    g = torch.Generator(device="cuda")
    g.manual_seed(31 * n + B)
    x = torch.randn(B, math.factorial(n), device="cuda", generator=g)
    ft = A.sn_fft(x, n)
    sel = list(rows)
    ref = O.sn_fft_packed_c(x[sel].double().cpu().numpy(), n)
    _check_dict({lam: v[sel] for lam, v in ft.items()}, ref, 1e-5)
    back = A.sn_ifft(ft, n)
    assert rel_err(back[sel].cpu().numpy(), x[sel].cpu().numpy()) < 1e-5
    # the oracle's inverse of the device coefficients of those rows agrees with the device inverse
    coef = {lam: v[sel].double().cpu().numpy().reshape(len(sel), O.dim(lam), O.dim(lam)) for lam, v in ft.items()}
    assert rel_err(back[sel].cpu().numpy(), O.sn_ifft_packed_c(coef, n)) < 1e-5


@pytest.mark.parametrize("n,B,dt", [(8, 512, "f32"), (9, 128, "f32"), (10, 16, "f32"), (10, 8, "f64")])
def test_properties_at_size(A, n, B, dt):
    """Round trip, Plancherel and linearity on a full-width batch (reference tests/test_fourier.py:55-84)."""
    torch.manual_seed(n)
    nf = math.factorial(n)
    f = torch.randn(B, nf, dtype=TDT[dt], device="cuda")
    ft = A.sn_fft(f, n)
    inv = A.sn_ifft(ft, n)
    err = ((inv - f).double().norm(dim=1) / f.double().norm(dim=1)).max().item()
    assert err < TOL[dt], err
    power = sum(v.double().reshape(B, -1).pow(2).sum(1) * (1 if v.dim() == 1 else v.shape[-1]) / nf
                for v in ft.values())
    ref = f.double().pow(2).sum(1)
    assert ((power - ref).abs() / ref).max().item() < 10 * TOL[dt]
    # linearity: FT(a f0 + f1) = a FT(f0) + FT(f1)
    comb = (0.5 * f[0] + f[1]).unsqueeze(0)
    fc = A.sn_fft(comb, n)
    for lam, v in fc.items():
        want = 0.5 * ft[lam][0] + ft[lam][1]
        assert (v[0] - want).double().norm().item() <= 4 * TOL[dt] * max(want.double().norm().item(), 1e-30)


@pytest.mark.parametrize("n", [6, 8])
def test_left_translation_equivariance(A, n):
    """reference tests/test_fourier.py:112-138 on the device path."""
    rng = np.random.default_rng(n)
    nf = math.factorial(n)
    f = rng.standard_normal(nf)
    pi = O.perm_unrank(np.array([rng.integers(nf)]), n)[0]
    pinv = np.argsort(pi)
    perms = O.perm_unrank(np.arange(nf), n) if n <= 6 else O.all_permutations(n)
    action = O.perm_rank(perms[:, pinv])
    ft = A.sn_fft(torch.from_numpy(f).cuda(), n)
    ftp = A.sn_fft(torch.from_numpy(f[action]).cuda(), n)
    for lam in ft:
        if O.dim(lam) == 1 or O.dim(lam) > 40:
            continue
        want = O.rho(lam, pi) @ ft[lam].cpu().numpy()
        assert np.allclose(ftp[lam].cpu().numpy(), want, atol=1e-9), lam


@pytest.mark.parametrize("n", [3, 5, 6])
def test_slow_sn_ft_matches_fast(A, n):
    f = torch.randn(2, math.factorial(n), dtype=torch.float64, device="cuda")
    slow, fast = A.slow_sn_ft(f, n), A.sn_fft(f, n)
    for lam in fast:
        assert torch.allclose(slow[lam].reshape(fast[lam].shape), fast[lam], atol=1e-9), lam
    assert torch.allclose(A.slow_sn_ift(fast, n).reshape(f.shape), f, atol=1e-10)


@pytest.mark.parametrize("cls", [0, 1, 2, 3, 4, 5])
def test_every_kernel_class_on_device(A, cls):
    """Each kernel class (tile width / row groups) on the GPU, S_6 forced through it."""
    from algebraist_b200 import plan as P
    lib = A._lib.load()
    n, B = 6, 5
    rng = np.random.default_rng(cls)
    f = rng.standard_normal((B, 720))
    ref = O.sn_fft_packed(f, n)
    for dt in ("f32", "f64"):
        lib.snfft_debug_force_class(cls)
        P._cache.pop((n, TDT[dt], "cuda", torch.cuda.current_device()), None)
        try:
            x = torch.from_numpy(f).to(TDT[dt]).cuda()
            ft = A.sn_fft(x, n)
            _check_dict(ft, ref, TOL[dt])
            assert rel_err(A.sn_ifft(ft, n).cpu().numpy(), f) < TOL[dt]
        finally:
            lib.snfft_debug_force_class(-1)
            P._cache.pop((n, TDT[dt], "cuda", torch.cuda.current_device()), None)


@pytest.mark.parametrize("n", [4, 7, 9])
def test_index_kernels_bit_exact(A, n, golden_tables):
    lib = A._lib.load()
    nf = math.factorial(n)
    m = min(nf, 50000)
    rng = np.random.default_rng(n)
    ranks = np.sort(rng.choice(nf, size=m, replace=False)).astype(np.int64)
    d_ranks = torch.from_numpy(ranks).cuda()
    d_perms = torch.empty((m, n), dtype=torch.int8, device="cuda")
    assert lib.snfft_perm_unrank(n, d_ranks.data_ptr(), m, d_perms.data_ptr(), None) == 0
    perms = d_perms.cpu().numpy().astype(np.int64)
    assert np.array_equal(O.perm_rank(perms), ranks)
    if n <= 7:
        assert np.array_equal(perms, O.all_permutations(n)[ranks])
    d_back = torch.empty(m, dtype=torch.int64, device="cuda")
    assert lib.snfft_perm_rank(n, d_perms.data_ptr(), m, d_back.data_ptr(), None) == 0
    assert torch.equal(d_back, d_ranks)
    for i in range(n):
        idx = torch.empty(math.factorial(n - 1), dtype=torch.int64, device="cuda")
        assert lib.snfft_coset_index(n, i, idx.data_ptr(), None) == 0
        assert np.array_equal(idx.cpu().numpy(), O.coset_index(n, i))
        if n <= 7:
            assert np.array_equal(idx.cpu().numpy(), golden_tables[f"n{n}_coset_{i}"].reshape(-1))
    p1 = torch.empty(nf, dtype=torch.int64, device="cuda")
    assert lib.snfft_chain_index(n, p1.data_ptr(), None) == 0
    assert np.array_equal(O.gather_index(n)[p1.cpu().numpy()], np.arange(nf))


def test_host_buffer_entry_points(A):
    """snfft_forward_host / snfft_inverse_host: host in, host out (the e2e path of bench.py)."""
    from algebraist_b200.plan import get_plan
    lib = A._lib.load()
    n, B = 7, 9
    rng = np.random.default_rng(1)
    f = rng.standard_normal((B, 5040)).astype(np.float32)
    plan = get_plan(n, torch.float32, torch.device("cuda", torch.cuda.current_device()))
    F = np.zeros(B * 5040, dtype=np.float32)
    assert lib.snfft_forward_host(plan.handle, f.ctypes.data, B, F.ctypes.data, None) == 0
    ref = O.sn_fft_packed(f.astype(np.float64), n)
    off = 0
    for lam in O.generate_partitions(n):
        d = O.dim(lam)
        assert rel_err(F[off * B:(off + d * d) * B].reshape(B, d, d), ref[lam]) < 1e-5
        off += d * d
    back = np.zeros_like(f)
    assert lib.snfft_inverse_host(plan.handle, F.ctypes.data, B, back.ctypes.data, None) == 0
    assert rel_err(back, f) < 1e-5


def test_cpu_tensor_round_trip_and_streams(A):
    """CPU tensors are staged through the GPU (no CPU compute); non-default streams are honoured."""
    f = torch.randn(3, 720)
    ft = A.sn_fft(f, 6)
    assert not ft[(6,)].is_cuda
    assert torch.allclose(A.sn_ifft(ft, 6), f, atol=1e-5)
    s = torch.cuda.Stream()
    x = torch.randn(4, 5040, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        y = A.sn_ifft(A.sn_fft(x, 7), 7)
    s.synchronize()
    assert torch.allclose(x, y, atol=1e-5)


def test_empty_batch(A):
    ft = A.sn_fft(torch.empty(0, 120, device="cuda"), 5)
    assert tuple(ft[(5,)].shape) == (0,) and tuple(ft[(3, 2)].shape) == (0, 5, 5)


@pytest.mark.parametrize("mode", [0, -1])
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_kernel_families_vs_oracle(A, mode, dt):
    """Level steps 6..10 through the table-driven sweep kernels (mode 0) and through the automatic choice
    (column kernels 6..8, top / bottom kernels 9, 10), forward and inverse."""
    from algebraist_b200 import plan as P
    n, B = 10, 2
    lib = A._lib.load()
    rng = np.random.default_rng(1000 + mode)
    f = rng.standard_normal((B, math.factorial(n)))
    ref = O.sn_fft_packed(f, n)
    lib.snfft_debug_tb_mode(mode)
    saved = dict(P._cache)
    P._cache.clear()
    try:
        x = torch.from_numpy(f).to(TDT[dt]).cuda()
        ft = A.sn_fft(x, n)
        _check_dict(ft, ref, TOL[dt])
        inv = A.sn_ifft(ft, n)
        assert rel_err(inv.cpu().numpy(), f) < TOL[dt]
    finally:
        lib.snfft_debug_tb_mode(-1)
        P._cache.clear()
        P._cache.update(saved)


@pytest.mark.parametrize("n,dt", [(9, "f64"), (10, "f32"), (11, "f32")])
def test_coset_partials_sum_to_full_transform(A, n, dt):
    """snfft_forward_cosets (the per-rank work of the coset-sharded transform, BASELINE config 5): the partial
    transforms of a partition of the top-level cosets add up to the full transform; n = 11 also checks
    Plancherel on the 39.9M-entry function."""
    from algebraist_b200.distributed import coset_shards, sn_fft_coset_partial
    torch.manual_seed(n)
    nf = math.factorial(n)
    f = torch.randn(nf, dtype=TDT[dt], device="cuda")
    full = A.sn_fft(f, n)
    total = None
    for lo, hi in coset_shards(n, 4):
        plan, packed, batch = sn_fft_coset_partial(f, n, range(lo, hi))
        total = packed.clone() if total is None else total + packed
    for lam, d, off in zip(plan.partitions, plan.dims, plan.offsets):
        got = total[off:off + d * d].reshape(full[lam].shape)
        ref = full[lam]
        assert ((got - ref).double().norm() / ref.double().norm()).item() < 10 * TOL[dt], lam
    power = sum(v.double().pow(2).sum() * (1 if v.dim() == 0 else v.shape[-1]) / nf for v in full.values())
    assert abs(power.item() / f.double().pow(2).sum().item() - 1) < 10 * TOL[dt]


def test_vs_oracle_n11(A):
    """One S_11 function (39.9M entries, BASELINE config 5's size) against the oracle, fp32."""
    n = 11
    rng = np.random.default_rng(11)
    f = rng.standard_normal((1, math.factorial(n)))
    ref = O.sn_fft_packed_c(f, n)
    ft = A.sn_fft(torch.from_numpy(f[0]).float().cuda(), n)
    _check_dict(ft, ref, TOL["f32"])
    # the same function as row 2 of a batch of 4: full instance vectors, so level 11 runs the T kernel on the packed
    # layout (MB = 8 bottom operators, three top sweeps) and the inverse reads the packed coefficients; rows 0, 1, 3
    # are other functions
    g = torch.Generator(device="cuda")
    g.manual_seed(1111)
    x = torch.randn(4, math.factorial(n), device="cuda", generator=g)
    x[2] = torch.from_numpy(f[0]).float().cuda()
    ft4 = A.sn_fft(x, n)
    _check_dict({lam: v[2:3] for lam, v in ft4.items()}, ref, TOL["f32"])
    for lam, v in ft.items():                # scalar threads (B = 1) and vector threads run the same arithmetic in the
        w = ft4[lam][2]                      # same order: identical values
        assert torch.equal(w.reshape(v.shape), v), lam
    back = A.sn_ifft(ft4, n)
    assert rel_err(back.cpu().numpy(), x.cpu().numpy()) < TOL["f32"]


@pytest.mark.parametrize("n,dt", [(5, "f64"), (7, "f32"), (8, "f64")])
def test_decomposition_and_power_gpu(A, n, dt):
    """sn_fourier_decomposition / calc_power on the device: reference goldens at n = 5, and at every n the
    components add up to f and the power spectrum adds up to ||f||^2 (reference tests/test_fourier.py:66-84)."""
    import os
    from conftest import GOLDEN
    torch.manual_seed(1000 + n)
    f = torch.randn(3, math.factorial(n), dtype=torch.float64).to(TDT[dt]).cuda()
    ft = A.sn_fft(f, n)
    dec = A.sn_fourier_decomposition(ft, n)
    assert rel_err(sum(dec.values()).cpu().numpy(), f.cpu().numpy()) < TOL[dt]
    pw = A.calc_power(ft, n)
    tot = sum(v.double() for v in pw.values())
    assert ((tot - f.double().pow(2).sum(1)).abs() / tot).max().item() < 10 * TOL[dt]
    if n == 5:
        nx = np.load(os.path.join(GOLDEN, "next_rows.npz"))
        for lam in dec:
            assert np.allclose(dec[lam].cpu().numpy(), nx[f"n5_decomp_{key(lam)}"], atol=1e-12)
            assert np.allclose(pw[lam].cpu().numpy(), nx[f"n5_power_{key(lam)}"], rtol=1e-12)


def test_stream_batches_host_pipeline(A):
    """Host-resident batch streamed through the GPU in chunks (ragged last chunk): same result as one call."""
    n, B = 8, 37
    torch.manual_seed(3)
    f_host = torch.randn(B, math.factorial(n)).pin_memory()
    out = A.stream_batches(lambda x: A.sn_ifft(A.sn_fft(x, n), n), f_host, chunk=8)
    assert out.shape == f_host.shape and rel_err(out.numpy(), f_host.numpy()) < TOL["f32"]
    ft_host = A.stream_batches(lambda x: A.sn_fft(x, n)[(n - 1, 1)], f_host, chunk=16)
    ref = A.sn_fft(f_host.cuda(), n)[(n - 1, 1)].cpu()
    assert torch.equal(ft_host, ref)


@pytest.mark.parametrize("n,B", [(8, 1), (8, 33), (9, 1), (9, 5), (9, 37), (10, 3)])
def test_ragged_batches_full_pipeline(A, n, B):
    """Batches that are not a multiple of 4 / 128 (scalar transfers, ragged last tiles of the column kernels):
    row 0 against the oracle, every row against the transform of that row alone, round trip."""
    rng = np.random.default_rng(1000 * n + B)
    f = rng.standard_normal((B, math.factorial(n)))
    x = torch.from_numpy(f).float().cuda()
    ft = A.sn_fft(x, n)
    _check_dict({lam: v[:1] for lam, v in ft.items()}, O.sn_fft_packed(f[:1], n), 1e-5)
    last = A.sn_fft(x[B - 1:], n)
    for lam, v in ft.items():
        assert (v[B - 1:] - last[lam]).double().norm().item() <= 1e-6 * max(last[lam].double().norm().item(), 1e-30)
    assert rel_err(A.sn_ifft(ft, n).cpu().numpy(), f) < 1e-5


def test_config2_s8_batch_4096(A):
    """BASELINE config 2: forward FFT on S_8, batch 4096 fp32 (661 MB).  Rows 0..3 against the reference's golden
    outputs (same inputs), Plancherel and round trip on every row."""
    n, B = 8, 4096
    g = load_transform_golden(n)
    f4 = torch.from_numpy(g[f"n{n}_f32_f"])
    nf = math.factorial(n)
    torch.manual_seed(8)
    f = torch.randn(B, nf, device="cuda")
    f[:f4.shape[0]] = f4.cuda()
    ft = A.sn_fft(f, n)
    for lam, v in ft.items():
        ref = g[f"n{n}_f32_ft_{key(lam)}"]
        assert rel_err(v[:f4.shape[0]].cpu().numpy().reshape(ref.shape), ref) < 1e-5, lam
    power = sum(v.double().reshape(B, -1).pow(2).sum(1) * (1 if v.dim() == 1 else v.shape[-1]) / nf for v in ft.values())
    ref = f.double().pow(2).sum(1)
    assert ((power - ref).abs() / ref).max().item() < 1e-4
    inv = A.sn_ifft(ft, n)
    assert ((inv - f).double().norm(dim=1) / f.double().norm(dim=1)).max().item() < 1e-5


def test_n12_single_function_round_trip(A):
    """One function on S_12 (479 M values, 1.9 GB): the largest n of the C ABI; levels 6..9 run the column kernels
    with NINST = 12!/k! instances.  Round trip and Plancherel."""
    n = 12
    nf = math.factorial(n)
    torch.manual_seed(12)
    f = torch.randn(1, nf, device="cuda")
    ft = A.sn_fft(f, n)
    power = sum(v.double().reshape(1, -1).pow(2).sum(1) * (1 if v.dim() == 1 else v.shape[-1]) / nf for v in ft.values())
    ref = f.double().pow(2).sum(1)
    assert ((power - ref).abs() / ref).max().item() < 1e-4
    inv = A.sn_ifft(ft, n)
    del ft
    assert ((inv - f).double().norm() / f.double().norm()).item() < 1e-5
