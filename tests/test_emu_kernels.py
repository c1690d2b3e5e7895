"""Kernel logic on the CPU: algebraist_b200/csrc/snfft.cu compiled by g++ against
tests/emu/cuda_emu.h (blocks and __syncthreads emulated with fibers), called
through the C ABI and compared with the oracle and the reference goldens.
The GPU parity tests proper are in test_gpu_parity.py."""
import ctypes
import math

import numpy as np
import pytest

from conftest import CAbi, key, load_transform_golden, pack, rel_err, unpack
from oracle import sn_oracle as O

DT = {"f32": np.float32, "f64": np.float64}
TOL = {"f32": 1e-5, "f64": 1e-12}


@pytest.mark.parametrize("n", range(2, 8))
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_forward_inverse_vs_reference_golden(emu, n, dt):
    g = load_transform_golden(n)
    f = g[f"n{n}_{dt}_f"]
    B = f.shape[0]
    abi = CAbi(emu)
    plan = abi.plan(n, DT[dt])
    try:
        packed = abi.forward(plan, f)
        ft = unpack(packed, n, B)
        for lam in O.generate_partitions(n):
            ref = g[f"n{n}_{dt}_ft_{key(lam)}"]
            assert rel_err(ft[lam].reshape(ref.shape), ref) < TOL[dt], lam
        inv = abi.inverse(plan, packed, B)
        assert rel_err(inv, g[f"n{n}_{dt}_ift"]) < TOL[dt]
        assert rel_err(inv, f) < TOL[dt]
    finally:
        emu.snfft_plan_destroy(plan)


@pytest.mark.parametrize("cls", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_every_kernel_class(emu, cls, dt):
    """Force all shapes of S_5 / S_6 through each kernel class (tile widths 32..2, row groups 1..256)."""
    n, B = (6, 3) if cls <= 2 else (5, 5)
    rng = np.random.default_rng(cls)
    f = rng.standard_normal((B, math.factorial(n))).astype(DT[dt])
    emu.snfft_debug_force_class(cls)
    abi = CAbi(emu)
    plan = abi.plan(n, DT[dt])
    emu.snfft_debug_force_class(-1)
    try:
        packed = abi.forward(plan, f)
        ref = O.sn_fft_packed(f.astype(np.float64), n)
        ft = unpack(packed, n, B)
        for lam in ref:
            assert rel_err(ft[lam], ref[lam]) < TOL[dt], lam
        assert rel_err(abi.inverse(plan, packed, B), f) < TOL[dt]
    finally:
        emu.snfft_plan_destroy(plan)


@pytest.mark.parametrize("dt", ["f64", "f32"])
@pytest.mark.parametrize("B", [1, 5, 37])
def test_generated_kernels_equal_sweep_kernels(emu, dt, B):
    """n = 7: levels 2..5 in the generated register kernel and 6, 7 in the column kernels; the table-driven
    per-level sweep kernels (debug hook 1) must agree."""
    n = 7
    rng = np.random.default_rng(B)
    f = rng.standard_normal((B, math.factorial(n))).astype(DT[dt])
    abi = CAbi(emu)
    plan = abi.plan(n, DT[dt])
    try:
        fused = abi.forward(plan, f)
        emu.snfft_debug_disable_base(1)
        plain = abi.forward(plan, f)
        # the column kernels carry compile-time scales through the rotations (one FMA per rotation output) and
        # multiply them back at the stores: same numbers up to the rounding order
        assert rel_err(fused, plain) < (1e-14 if dt == "f64" else 2e-6)
        inv_plain = abi.inverse(plan, fused, B)
        emu.snfft_debug_disable_base(0)
        inv_fused = abi.inverse(plan, fused, B)
        # the inverse column kernels add the two terms of the top sweep to the running sums one at a time
        # (one child block of F_lambda live at a time): same numbers up to the rounding order
        assert rel_err(inv_fused, inv_plain) < (1e-14 if dt == "f64" else 2e-6)
        assert rel_err(inv_fused, f) < TOL[dt]
    finally:
        emu.snfft_debug_disable_base(0)
        emu.snfft_plan_destroy(plan)


@pytest.mark.parametrize("B", [1, 2, 7, 33, 70])
def test_ragged_batches(emu, B):
    n = 5
    rng = np.random.default_rng(B)
    f = rng.standard_normal((B, 120))
    abi = CAbi(emu)
    plan = abi.plan(n, np.float64)
    try:
        packed = abi.forward(plan, f)
        ref = O.sn_fft_packed(f, n)
        assert rel_err(packed, pack(ref, n, B, np.float64)) < 1e-12
        assert rel_err(abi.inverse(plan, packed, B), f) < 1e-12
    finally:
        emu.snfft_plan_destroy(plan)


def test_inverse_of_random_coefficients(emu):
    """sn_ifft on arbitrary Fourier data (reference tests/test_fourier.py:153-159)."""
    n, B = 6, 2
    rng = np.random.default_rng(5)
    ft = {lam: rng.standard_normal((B, O.dim(lam), O.dim(lam))) for lam in O.generate_partitions(n)}
    abi = CAbi(emu)
    plan = abi.plan(n, np.float64)
    try:
        inv = abi.inverse(plan, pack(ft, n, B, np.float64), B)
        assert rel_err(inv, O.sn_ifft_packed(ft, n)) < 1e-12
        assert rel_err(inv, O.slow_sn_ift_packed(ft, n)) < 1e-12
    finally:
        emu.snfft_plan_destroy(plan)


@pytest.mark.parametrize("n", [3, 5, 7])
def test_index_kernels_bit_exact(emu, n, golden_tables):
    nf = math.factorial(n)
    perms = O.all_permutations(n).astype(np.int8)
    ranks = np.zeros(nf, dtype=np.int64)
    assert emu.snfft_perm_rank(n, perms.ctypes.data, nf, ranks.ctypes.data, None) == 0
    assert np.array_equal(ranks, np.arange(nf))
    back = np.zeros((nf, n), dtype=np.int8)
    assert emu.snfft_perm_unrank(n, ranks.ctypes.data, nf, back.ctypes.data, None) == 0
    assert np.array_equal(back, perms)
    for i in range(n):
        idx = np.zeros(math.factorial(n - 1), dtype=np.int64)
        assert emu.snfft_coset_index(n, i, idx.ctypes.data, None) == 0
        assert np.array_equal(idx, golden_tables[f"n{n}_coset_{i}"].reshape(-1))
    p1 = np.zeros(nf, dtype=np.int64)
    assert emu.snfft_chain_index(n, p1.ctypes.data, None) == 0
    g = O.gather_index(n)                       # instance -> rank, from restrict_to_coset recursion
    assert np.array_equal(g[p1], np.arange(nf))


@pytest.mark.parametrize("n", [3, 4, 5])
def test_dense_rep(emu, n):
    abi = CAbi(emu)
    plan = abi.plan(n, np.float64)
    perms = O.all_permutations(n)
    try:
        for part, lam in enumerate(O.generate_partitions(n)):
            d = O.dim(lam)
            out = np.zeros((len(perms), d, d))
            assert emu.snfft_dense_rep(plan, part, out.ctypes.data, None) == 0
            ref = np.stack([O.rho(lam, s) for s in perms])
            assert np.allclose(out, ref, atol=1e-13), lam
    finally:
        emu.snfft_plan_destroy(plan)


def test_launch_counter(emu):
    before = emu.snfft_launch_count()
    abi = CAbi(emu)
    plan = abi.plan(4, np.float32)
    abi.forward(plan, np.zeros((1, 24), dtype=np.float32))
    emu.snfft_plan_destroy(plan)
    assert emu.snfft_launch_count() - before >= 4      # gather + levels 2..4


# (n, kernel mode, base mode): mode -1 is the product's automatic choice (column kernels at levels 6..8), mode 0 the
# table-driven sweep kernels; base 1 sends the levels <= 5 through the sweep kernels too.
FAMILY_CASES = [(7, -1, 0), (7, 0, 0), (6, 0, 1), (5, 0, 1), (8, -1, 0), (8, 0, 0), (8, 0, 1)]


@pytest.mark.parametrize("n,mode,base", FAMILY_CASES)
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_kernel_families_vs_oracle(emu, n, mode, base, dt):
    """Both implementations of the level steps against the oracle, forward, inverse of arbitrary coefficients, round trip."""
    B = 1 if n == 8 else 3
    rng = np.random.default_rng(10 * n + mode)
    f = rng.standard_normal((B, math.factorial(n))).astype(DT[dt])
    emu.snfft_debug_tb_mode(mode)
    emu.snfft_debug_disable_base(base)
    abi = CAbi(emu)
    plan = abi.plan(n, DT[dt])
    emu.snfft_debug_tb_mode(-1)
    try:
        packed = abi.forward(plan, f)
        ref = O.sn_fft_packed(f.astype(np.float64), n)
        ft = unpack(packed, n, B)
        for lam in ref:
            assert rel_err(ft[lam], ref[lam]) < TOL[dt], lam
        coef = {lam: rng.standard_normal(v.shape) for lam, v in ref.items()}
        got = abi.inverse(plan, pack(coef, n, B, DT[dt]), B)
        assert rel_err(got, O.sn_ifft_packed(coef, n)) < TOL[dt]
        assert rel_err(abi.inverse(plan, packed, B), f) < TOL[dt]
    finally:
        emu.snfft_debug_disable_base(0)
        emu.snfft_plan_destroy(plan)


@pytest.mark.parametrize("B", [1, 2])
def test_level9_top_bottom_kernels_vs_oracle(emu, B):
    """n = 9 in automatic mode: the level step k = 9 runs the top / bottom kernels (tz_kernels.cuh: level-7 column
    operators on pieces of the inputs, then the two top sweeps) in both directions; against the oracle.  B = 2: full
    instance vectors, so the T kernel of the top level reads / writes the packed coefficients itself (no pack / unpack pass)."""
    n = 9
    rng = np.random.default_rng(9)
    f = rng.standard_normal((B, math.factorial(n)))
    abi = CAbi(emu)
    plan = abi.plan(n, np.float64)
    try:
        packed = abi.forward(plan, f)
        ref = O.sn_fft_packed(f, n)
        ft = unpack(packed, n, B)
        for lam in ref:
            assert rel_err(ft[lam], ref[lam]) < 1e-12, lam
        assert rel_err(abi.inverse(plan, packed, B), f) < 1e-12
    finally:
        emu.snfft_plan_destroy(plan)


def test_column_and_sweep_kernels_agree_ragged_batch(emu):
    """Odd batch (scalar transfer path, ragged last tile): both kernel families give the same coefficients."""
    n, B = 7, 5
    rng = np.random.default_rng(7)
    f = rng.standard_normal((B, math.factorial(n)))
    outs = []
    for mode in (0, -1):
        emu.snfft_debug_tb_mode(mode)
        abi = CAbi(emu)
        plan = abi.plan(n, np.float64)
        emu.snfft_debug_tb_mode(-1)
        try:
            packed = abi.forward(plan, f)
            outs.append(packed)
            assert rel_err(abi.inverse(plan, packed, B), f) < 1e-12
        finally:
            emu.snfft_plan_destroy(plan)
    assert rel_err(outs[1], outs[0]) < 1e-13


@pytest.mark.parametrize("n,B,dt", [(4, 4, "f32"), (6, 8, "f32"), (7, 36, "f32"), (5, 2, "f64"), (6, 18, "f64")])
def test_vector_gather_scatter(emu, n, B, dt):
    """Batches that are a multiple of 16 bytes take the vectorised gather / scatter (register-block
    transposes, cached chain-index table); it must agree bit for bit with the scalar tile kernel."""
    rng = np.random.default_rng(n * 100 + B)
    f = rng.standard_normal((B, math.factorial(n))).astype(DT[dt])
    abi = CAbi(emu)
    plan = abi.plan(n, DT[dt])
    try:
        packed = abi.forward(plan, f)
        back = abi.inverse(plan, packed, B)
        emu.snfft_debug_scalar_gather(1)
        packed_s = abi.forward(plan, f)
        back_s = abi.inverse(plan, packed, B)
        assert np.array_equal(packed, packed_s) and np.array_equal(back, back_s)
        ref = O.sn_fft_packed(f.astype(np.float64), n)
        ft = unpack(packed, n, B)
        for lam in ref:
            assert rel_err(ft[lam], ref[lam]) < TOL[dt], lam
        assert rel_err(back, f) < TOL[dt]
    finally:
        emu.snfft_debug_scalar_gather(0)
        emu.snfft_plan_destroy(plan)
