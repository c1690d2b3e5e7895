"""Group-action utilities on the device: translation of functions on S_n by a permutation and
rho_lambda(pi) in Young's orthogonal form, as the reference's equivariance test builds them on the CPU
(tests/test_fourier.py:112-138 with Permutation.inverse / __mul__ / permutation_index, permutations.py:82-100,187-192,
and SnIrrep.matrix_representations, irreps.py:122-141).  Both run in this package's kernels: the translation
unranks, composes and ranks every index on the GPU (bit exact), rho applies the word of adjacent transpositions
of pi to the identity with the plan's YOR tables -- no n! x d x d table, no dimension cap."""
import ctypes
import math

import torch

from . import _lib
from . import fourier as _f            # looked up at call time (the CPU test-suite swaps its device helpers)
from .fourier import _check_input
from .plan import _DTYPES, get_plan


def _as_perm(perm, n: int):
    sig = [int(x) for x in (perm.tolist() if isinstance(perm, torch.Tensor) else perm)]
    if sorted(sig) != list(range(n)):
        raise ValueError(f"perm must be a one-line permutation of range({n}), got {sig}")
    return (ctypes.c_int * n)(*sig)


def sn_translate(fn_vals: torch.Tensor, perm, n: int, side: str = "left") -> torch.Tensor:
    """Translate f by the permutation `perm` (one-line notation, sequence or tensor of ints).

    side="left":  out[..., k] = f[..., index(perm^-1 * p_k)]  -- the action of tests/test_fourier.py:112-121, for which
                  sn_fft(out)[lam] == rho(lam, perm) @ sn_fft(f)[lam];
    side="right": out[..., k] = f[..., index(p_k * perm^-1)], for which sn_fft(out)[lam] == sn_fft(f)[lam] @ rho(lam, perm)
    with p_k the k-th permutation in lexicographic order and (a * b)[i] = b[a[i]] (permutations.py:85-88)."""
    _check_input(fn_vals, n)
    if side not in ("left", "right"):
        raise ValueError("side must be 'left' or 'right'")
    sig = _as_perm(perm, n)
    x, home = _f._to_device(fn_vals)
    f2d = x.reshape(-1, x.shape[-1]).contiguous()
    out = torch.empty_like(f2d)
    if f2d.shape[0]:
        with _f._device_ctx(f2d.device):
            _lib.check(_lib.load().snfft_translate(n, _DTYPES[f2d.dtype], f2d.data_ptr(), f2d.shape[0], sig,
                                                   1 if side == "right" else 0, out.data_ptr(), _f._stream_ptr(f2d.device)),
                       "snfft_translate")
    out = out.view(tuple(fn_vals.shape))
    return out if home == out.device else out.to(home)


def rho(partition, perm, n: int, dtype: torch.dtype = torch.float32, device=None) -> torch.Tensor:
    """rho_partition(perm): the (d, d) matrix of `perm` in the irrep `partition` of S_n, Young's orthogonal form in the
    reference's basis order (SnIrrep(n, partition).matrix_representations[perm], irreps.py:122-141), on `device`
    (default: the current CUDA device).  rho(p * q) == rho(p) @ rho(q)."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("algebraist_b200 needs a CUDA device (there is no CPU fallback)")
        device = torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device)
    plan = get_plan(n, dtype, dev)
    lam = tuple(int(v) for v in partition)
    if lam not in plan.index:
        raise ValueError(f"{lam} is not a partition of {n}")
    part = plan.index[lam]
    d = plan.dims[part]
    sig = _as_perm(perm, n)
    out = torch.empty((d, d), dtype=dtype, device=plan.device)
    with _f._device_ctx(plan.device):
        _lib.check(_lib.load().snfft_rho(plan.handle, part, sig, out.data_ptr(), _f._stream_ptr(plan.device)), "snfft_rho")
    return out


def permutation_index(perms: torch.Tensor) -> torch.Tensor:
    """Lexicographic ranks of one-line permutations, (m, n) -> (m,) int64 (Permutation.permutation_index,
    permutations.py:187-192), on the device of `perms` (CUDA)."""
    if perms.dim() != 2:
        raise ValueError("perms must be (m, n)")
    x, home = _f._to_device(perms)
    p8 = x.to(torch.int8).contiguous()
    m, n = p8.shape
    out = torch.empty(m, dtype=torch.int64, device=p8.device)
    if m:
        with _f._device_ctx(p8.device):
            _lib.check(_lib.load().snfft_perm_rank(n, p8.data_ptr(), m, out.data_ptr(), _f._stream_ptr(p8.device)),
                       "snfft_perm_rank")
    return out if home == out.device else out.to(home)


def all_permutations(n: int, device=None) -> torch.Tensor:
    """generate_all_permutations(n) (utils.py:61-63): the (n!, n) int64 table in lexicographic order, built on the
    device by the unrank kernel."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    m = math.factorial(n)
    ranks = torch.arange(m, dtype=torch.int64, device=dev)
    out = torch.empty((m, n), dtype=torch.int8, device=dev)
    with _f._device_ctx(dev):
        _lib.check(_lib.load().snfft_perm_unrank(n, ranks.data_ptr(), m, out.data_ptr(), _f._stream_ptr(dev)), "snfft_perm_unrank")
    return out.to(torch.int64)
