"""Small transforms through every kernel family, for compute-sanitizer (one tool per run):

    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py

Each case checks its result against the oracle as well, so a sanitizer-clean run that computes garbage still fails.
Families: automatic choice (generated levels 2..5, column kernels 6..8, top / bottom kernels 9: vector and scalar T
kernel), table-driven sweep kernels at every level, scalar / small-batch / vector gather and
pack, ragged batches, the two coset-sharded partial transforms, translate / rho / dense rep / power / block matmul.
"""
import math
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import algebraist_b200 as A  # noqa: E402
from algebraist_b200 import _lib, plan as P  # noqa: E402
from algebraist_b200.distributed import coset_pair_shards, sn_fft_coset2_partial, sn_fft_coset_partial  # noqa: E402
from oracle import sn_oracle as O  # noqa: E402

lib = _lib.load()
TOL = {torch.float32: 2e-5, torch.float64: 1e-11}
ran = []


def check(n, B, dtype, tag):
    rng = np.random.default_rng(n * 100 + B)
    f = rng.standard_normal((B, math.factorial(n)))
    x = torch.from_numpy(f).to(dtype).cuda()
    ft = A.sn_fft(x, n)
    ref = O.sn_fft_packed_c(f, n) if n >= 8 else O.sn_fft_packed(f, n)
    for lam, v in ft.items():
        r = ref[lam].reshape(tuple(v.shape))
        err = np.linalg.norm(v.double().cpu().numpy() - r) / max(np.linalg.norm(r), 1e-300)
        assert err < TOL[dtype], (tag, n, B, lam, err)
    back = A.sn_ifft(ft, n)
    err = ((back - x).double().norm() / x.double().norm()).item()
    assert err < TOL[dtype], (tag, "round trip", err)
    ran.append(f"{tag}: n={n} B={B} {str(dtype)[6:]}")


def fresh():
    P._cache.clear()


# automatic family choice
for n, B, dt in ((7, 4, torch.float32), (7, 3, torch.float64), (8, 5, torch.float32), (9, 4, torch.float32),
                 (9, 3, torch.float32), (9, 2, torch.float64), (9, 1, torch.float64)):
    check(n, B, dt, "auto")
# sweep kernels only / TB kernels / base kernel variants
for mode in (0,):
    lib.snfft_debug_tb_mode(mode)
    fresh()
    check(7, 5, torch.float32, f"tb_mode={mode}")
    check(8, 2, torch.float64, f"tb_mode={mode}")
lib.snfft_debug_tb_mode(-1)
for base in (1,):
    lib.snfft_debug_disable_base(base)
    fresh()
    check(7, 3, torch.float32, f"base={base}")
lib.snfft_debug_disable_base(0)
lib.snfft_debug_scalar_gather(1)
fresh()
check(7, 4, torch.float32, "scalar gather/pack")
lib.snfft_debug_scalar_gather(0)
fresh()
os.environ["SNFFT_TZ_VEC"] = "1"            # read once per process by the T launchers: only effective if set before the first call
# coset-sharded partial transforms add up to the full transform
n = 8
f = torch.randn(math.factorial(n), dtype=torch.float64, device="cuda")
full = A.sn_fft(f, n)
for name, parts in (("cosets", [sn_fft_coset_partial(f, n, c)[1] for c in ([0, 3, 4], [1, 2, 5, 6, 7])]),
                    ("coset pairs", [sn_fft_coset2_partial(f, n, pr)[1] for pr in coset_pair_shards(n, 3)])):
    tot = sum(parts)
    plan = P.get_plan(n, torch.float64, f.device)
    for lam, d, off in zip(plan.partitions, plan.dims, plan.offsets):
        got = tot[off:off + d * d].view(full[lam].shape)
        assert torch.allclose(got, full[lam], atol=1e-10), (name, lam)
    ran.append(f"{name}: n={n}")
# group action, dense rep, power, convolution product
n = 7
g = torch.randn(2, math.factorial(n), dtype=torch.float64, device="cuda")
pi = torch.randperm(n)
ftg, ftl = A.sn_fft(g, n), A.sn_fft(A.sn_translate(g, pi, n), n)
for lam in ftg:
    R = A.rho(lam, pi, n, dtype=torch.float64)
    want = R.reshape(()) * ftg[lam] if ftg[lam].dim() == 1 else R @ ftg[lam]
    assert torch.allclose(ftl[lam], want, atol=1e-10), lam
pw = A.calc_power(ftg, n)
assert torch.allclose(sum(pw.values()), g.pow(2).sum(-1), rtol=1e-10)
slow = A.slow_sn_ft(g[:1, :math.factorial(5)].contiguous(), 5)
fast = A.sn_fft(g[:1, :math.factorial(5)].contiguous(), 5)
for lam in fast:
    assert torch.allclose(slow[lam].reshape(fast[lam].shape), fast[lam], atol=1e-10)
h = torch.randn_like(g)
conv = A.sn_convolve(g, h, n)
assert conv.shape == g.shape and torch.isfinite(conv).all()
ran.append("translate / rho / power / dense rep / convolve: n=7")
torch.cuda.synchronize()
print("SANITIZE_CASES_OK", len(ran))
for r in ran:
    print("  ", r)
