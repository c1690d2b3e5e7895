"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python oracle/make_golden.py            # n = 2..7  (seconds)
    python oracle/make_golden.py --n8       # adds n = 8 (about a minute, ~8 GB)

What is recorded (every array comes from the reference's own code):
  partitions_n{n}            generate_partitions(n)            tableau.py:103-123
  n{n}_dims                  SnIrrep.dim per partition         irreps.py:40-44
  n{n}_yor_{lam}_{j}         adj_transposition_matrix(j, j+1)  irreps.py:91-109   (n <= 6)
  n{n}_coset_{i}             argwhere(sn_perms[:, i] == n-1)   fourier.py:77-80   (n <= 7)
  n{n}_{dt}_f                seeded random input (B, n!)
  n{n}_{dt}_ft_{lam}         sn_fft(f, n)[lam]                 fourier.py:281-288
  n{n}_{dt}_ift              inverse of that ft: sn_ifft for n <= 5 (fourier.py:298-299);
                             for n = 6, 7 the reference's dense inverse reached by
                             BASE_CASE = n (fourier.py:134-165) because the recursive
                             inverse raises there (fourier.py:202-214)
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
import algebraist.fourier as RF                                     # noqa: E402
from algebraist.irreps import SnIrrep                               # noqa: E402
from algebraist.tableau import generate_partitions                  # noqa: E402
from algebraist.utils import generate_all_permutations              # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def key(lam):
    return "_".join(str(x) for x in lam)


def tables():
    out = {}
    for n in range(1, 13):
        parts = generate_partitions(n)
        flat = np.zeros((len(parts), n), dtype=np.int64)
        for r, p in enumerate(parts):
            flat[r, :len(p)] = p
        out[f"partitions_n{n}"] = flat
    for n in range(2, 8):
        perms = generate_all_permutations(n)
        for i in range(n):
            out[f"n{n}_coset_{i}"] = torch.argwhere(perms[:, i] == n - 1).squeeze().numpy()
        dims = []
        for lam in generate_partitions(n):
            ir = SnIrrep(n, lam)
            dims.append(ir.dim)
            if n <= 6:
                for j in range(n - 1):
                    out[f"n{n}_yor_{key(lam)}_{j}"] = ir.adj_transposition_matrix(j, j + 1)
        out[f"n{n}_dims"] = np.array(dims, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **out)


def transforms(n, batch, dtypes, with_inverse=True):
    out = {}
    for dt_name, dt in dtypes:
        torch.manual_seed(1000 + n)
        f = torch.randn(batch, math.factorial(n), dtype=torch.float64).to(dt)
        ft = RF.sn_fft(f, n)
        out[f"n{n}_{dt_name}_f"] = f.numpy()
        for lam, v in ft.items():
            out[f"n{n}_{dt_name}_ft_{key(lam)}"] = v.numpy()
        if with_inverse:
            old = RF.BASE_CASE
            if n > old:
                RF.BASE_CASE = n          # dense inverse path, fourier.py:164-165
            try:
                out[f"n{n}_{dt_name}_ift"] = RF.sn_ifft(ft, n).numpy()
            finally:
                RF.BASE_CASE = old
    np.savez_compressed(os.path.join(OUT, f"transform_n{n}.npz"), **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n8", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    both = [("f64", torch.float64), ("f32", torch.float32)]
    if args.n8:
        transforms(8, 1, both, with_inverse=False)
    else:
        tables()
        for n in range(2, 8):
            transforms(n, 3, both)


def next_rows():
    """tests/golden/next_rows.npz: sn_fourier_decomposition (fourier.py:291-295) and calc_power
    (fourier.py:362-393) of the UNMODIFIED reference on the ft of transform_n{n}.npz, n = 3..5
    (the reference's recursive inverse only works for n <= 5)."""
    out = {}
    for n in range(3, 6):
        torch.manual_seed(1000 + n)
        f = torch.randn(3, math.factorial(n), dtype=torch.float64)
        ft = RF.sn_fft(f, n)
        for lam, v in RF.sn_fourier_decomposition(ft, n).items():
            out[f"n{n}_decomp_{key(lam)}"] = v.numpy()
        for lam, v in RF.calc_power(ft, n).items():
            out[f"n{n}_power_{key(lam)}"] = v.numpy()
        one = {lam: v[0] for lam, v in ft.items()}            # unbatched call
        for lam, v in RF.calc_power(one, n).items():
            out[f"n{n}_power1_{key(lam)}"] = np.asarray(v.numpy())
    np.savez_compressed(os.path.join(OUT, "next_rows.npz"), **out)


def convolution():
    """tests/golden/convolution.npz: group convolution exactly as the reference's own test helper computes it
    (tests/test_fourier.py:22-34: result[i] = sum_j f[j] * g[index(perms[i] * perms[j].inverse)], with the
    reference's `Permutation` product and `full_group` order), n = 3..5, fp64, seeded inputs.  The convolution
    theorem the reference tests (tests/test_fourier.py:87-109) is FT(convolve(g, f)) = FT(f) @ FT(g).

        python -c "import sys; sys.path.insert(0, 'oracle'); import make_golden as M; M.convolution()"
    """
    from algebraist.permutations import Permutation
    out = {}
    for n in range(3, 6):
        perms = list(Permutation.full_group(n))
        index = {p.sigma: i for i, p in enumerate(perms)}
        torch.manual_seed(2000 + n)
        f = torch.randn(math.factorial(n), dtype=torch.float64).numpy()
        g = torch.randn(math.factorial(n), dtype=torch.float64).numpy()
        res = np.zeros_like(f)
        for i, pi in enumerate(perms):
            for j, sigma in enumerate(perms):
                res[i] += f[j] * g[index[(pi * sigma.inverse).sigma]]
        out[f"n{n}_f"], out[f"n{n}_g"], out[f"n{n}_conv_f_g"] = f, g, res        # = convolve(f, g, n) of the test helper
    np.savez_compressed(os.path.join(OUT, "convolution.npz"), **out)
