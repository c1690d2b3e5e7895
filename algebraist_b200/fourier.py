"""Drop-in for `algebraist.fourier`: sn_fft, sn_ifft, slow_sn_ft, slow_sn_ift.

Same signatures, dict[partition -> tensor] layout, key order and shape quirks
as the reference (/root/reference/src/algebraist/fourier.py); the arithmetic
runs in hand-written sm_100a CUDA kernels behind the C ABI of include/snfft.h.
There is no CPU path: CPU tensors are copied to the current CUDA device,
transformed there and copied back."""
import math

import torch

from . import _lib
from .plan import get_plan

# The reference switches to its dense path at n <= BASE_CASE (fourier.py:27);
# the only observable effect kept here is the `.squeeze()` it applies there
# (fourier.py:129), which collapses size-1 leading dimensions for d > 1.
BASE_CASE = 5


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _device_ctx(device):
    return torch.cuda.device(device)


def _to_device(t: torch.Tensor):
    """Returns (cuda tensor, original device)."""
    if t.is_cuda:
        return t, t.device
    if not torch.cuda.is_available():
        raise RuntimeError("algebraist_b200 needs a CUDA device (there is no CPU fallback)")
    return t.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=False), t.device


def _forward_packed(f2d: torch.Tensor, n: int):
    """f2d: (B, n!) contiguous CUDA tensor -> (plan, packed coefficients (B * n!,))."""
    plan = get_plan(n, f2d.dtype, f2d.device)
    batch = f2d.shape[0]
    packed = torch.empty(batch * plan.nfact, dtype=f2d.dtype, device=f2d.device)
    if batch:
        with _device_ctx(f2d.device):
            ws = plan.workspace(batch)
            _lib.check(_lib.load().snfft_forward(plan.handle, f2d.data_ptr(), batch, packed.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), _stream_ptr(f2d.device)),
                       "snfft_forward")
    return plan, packed


def _inverse_packed(plan, packed: torch.Tensor, batch: int) -> torch.Tensor:
    out = torch.empty((batch, plan.nfact), dtype=packed.dtype, device=packed.device)
    if batch:
        with _device_ctx(packed.device):
            ws = plan.workspace(batch)
            _lib.check(_lib.load().snfft_inverse(plan.handle, packed.data_ptr(), batch, out.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), _stream_ptr(packed.device)),
                       "snfft_inverse")
    return out


def _scale_blocks(plan, packed: torch.Tensor, batch: int, inverse_weights: bool) -> torch.Tensor:
    """packed * diag(w_lambda), w = n!/d_lambda (inverse_weights) or d_lambda/n!, in place."""
    views = [packed[off * batch:(off + d * d) * batch] for d, off in zip(plan.dims, plan.offsets)]
    w = [plan.nfact / d if inverse_weights else d / plan.nfact for d in plan.dims]
    torch._foreach_mul_(views, w)
    return packed


class _ForwardFn(torch.autograd.Function):
    """Autograd for sn_fft.  The reference transform is a composition of differentiable torch ops
    (einsum / matmul, fourier.py:111-131, 223-278), so callers may backpropagate through it.  The transform
    T is linear and T^-1 = T^T diag(d_lambda / n!) (Plancherel), hence T^T g = T^-1 (n!/d_lambda g): the
    backward pass is the inverse pipeline on rescaled gradients."""

    @staticmethod
    def forward(ctx, f2d, n):
        plan, packed = _forward_packed(f2d, n)
        ctx.plan, ctx.batch = plan, f2d.shape[0]
        batch = ctx.batch
        outs = tuple(packed[off * batch:(off + d * d) * batch] for d, off in zip(plan.dims, plan.offsets))
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grads):
        plan, batch = ctx.plan, ctx.batch
        like = next(g for g in grads if g is not None)
        packed = torch.zeros(batch * plan.nfact, dtype=like.dtype, device=like.device)
        for g, d, off in zip(grads, plan.dims, plan.offsets):
            if g is not None:
                packed[off * batch:(off + d * d) * batch].copy_(g.reshape(-1))
        _scale_blocks(plan, packed, batch, inverse_weights=True)
        return _inverse_packed(plan, packed, batch), None


class _InverseFn(torch.autograd.Function):
    """Autograd for sn_ifft: (T^-1)^T h = diag(d_lambda / n!) T h, the forward pipeline on the incoming
    gradient with every block rescaled."""

    @staticmethod
    def forward(ctx, n, batch, *pieces):
        some = pieces[0]
        plan = get_plan(n, some.dtype, some.device)
        packed = torch.cat([p.reshape(-1) for p in pieces])
        ctx.plan, ctx.batch, ctx.shapes, ctx.n = plan, batch, [tuple(p.shape) for p in pieces], n
        return _inverse_packed(plan, packed, batch)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        plan, batch = ctx.plan, ctx.batch
        _, packed = _forward_packed(grad_out.reshape(batch, plan.nfact).contiguous(), ctx.n)
        _scale_blocks(plan, packed, batch, inverse_weights=False)
        grads = tuple(packed[off * batch:(off + d * d) * batch].view(shape)
                      for d, off, shape in zip(plan.dims, plan.offsets, ctx.shapes))
        return (None, None) + grads


def _check_input(fn_vals, n):
    if not isinstance(fn_vals, torch.Tensor):
        raise TypeError("fn_vals must be a torch.Tensor")
    if not isinstance(n, int) or n < 2:
        raise ValueError(f"n must be an integer >= 2, got {n!r}")
    if fn_vals.dim() < 1 or fn_vals.shape[-1] != math.factorial(n):
        raise ValueError(f"last dimension of fn_vals must be n! = {math.factorial(n)}, got shape {tuple(fn_vals.shape)}")
    if fn_vals.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"only float32 and float64 are supported, got {fn_vals.dtype}")


def sn_fft(fn_vals: torch.Tensor, n: int, verbose: bool = False) -> dict:
    """Fast Fourier transform on S_n (reference fourier.py:281-288).

    fn_vals: (n!,) or (..., n!) values of f in lexicographic order of one-line
    permutations.  Returns {partition: F(partition)} in generate_partitions(n)
    order; 1-dimensional irreps have the leading shape, the others
    leading + (d, d).  For n <= 5 the reference squeezes size-1 leading
    dimensions of the d > 1 entries (fourier.py:129); that is reproduced."""
    _check_input(fn_vals, n)
    x, home = _to_device(fn_vals)
    lead = tuple(x.shape[:-1])
    f2d = x.reshape(-1, x.shape[-1]).contiguous()
    batch = f2d.shape[0]
    if torch.is_grad_enabled() and f2d.requires_grad and batch:
        plan = get_plan(n, f2d.dtype, f2d.device)
        blocks = _ForwardFn.apply(f2d, n)
    else:
        plan, packed = _forward_packed(f2d, n)
        blocks = [packed[off * batch:(off + d * d) * batch] for d, off in zip(plan.dims, plan.offsets)]
    result = {}
    for lam, d, blk in zip(plan.partitions, plan.dims, blocks):
        if d == 1:
            val = blk.view(lead)
        else:
            val = blk.view(lead + (d, d))
            if n <= BASE_CASE:
                val = val.squeeze()
        result[lam] = val if home == x.device else val.to(home)
    return result


def _pack_ft(ft: dict, n: int):
    some = next(iter(ft.values()))
    if not isinstance(some, torch.Tensor):
        raise TypeError("ft must map partitions to tensors")
    x, home = _to_device(some)
    if x.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"only float32 and float64 are supported, got {x.dtype}")
    plan = get_plan(n, x.dtype, x.device)
    missing = [p for p in plan.partitions if p not in ft]
    if missing:
        raise ValueError(f"ft is missing partitions {missing[:3]}...")
    triv = ft[plan.partitions[0]]
    lead = tuple(triv.shape)
    batch = triv.numel()
    # zero-copy when ft is the dict sn_fft produced (views into one packed buffer)
    pieces, contiguous_run, base_ptr = [], True, None
    for lam, d, off in zip(plan.partitions, plan.dims, plan.offsets):
        t = ft[lam]
        if t.numel() != batch * d * d:
            raise ValueError(f"ft[{lam}] has {t.numel()} elements, expected {batch * d * d}")
        if not (t.device == x.device and t.dtype == x.dtype and t.is_contiguous()):
            contiguous_run = False
        elif base_ptr is None:
            base_ptr = t.data_ptr()
            if off != 0:
                contiguous_run = False
        elif t.data_ptr() != base_ptr + off * batch * t.element_size():
            contiguous_run = False
        pieces.append(t)
    if contiguous_run and base_ptr is not None:
        storage_ok = all(p.untyped_storage().data_ptr() == pieces[0].untyped_storage().data_ptr() for p in pieces)
        if storage_ok:
            packed = torch.as_strided(pieces[0], (batch * plan.nfact,), (1,), pieces[0].storage_offset())
            return plan, packed, batch, lead, home
    packed = torch.cat([p.to(device=x.device, dtype=x.dtype).reshape(-1) for p in pieces])
    return plan, packed, batch, lead, home


def sn_ifft(ft: dict, n: int) -> torch.Tensor:
    """Inverse fast Fourier transform (reference fourier.py:298-299).

    The reference's recursive inverse raises for n > 5 (fourier.py:202-214);
    this returns the mathematical inverse
        f(sigma) = 1/n! sum_lambda d_lambda <F(lambda), rho_lambda(sigma)>
    at every n, shape leading + (n!,) (squeezed for n <= 5 like the reference)."""
    if torch.is_grad_enabled() and any(isinstance(v, torch.Tensor) and v.requires_grad for v in ft.values()):
        some = next(iter(ft.values()))
        x, home = _to_device(some)
        plan = get_plan(n, x.dtype, x.device)
        lead = tuple(ft[plan.partitions[0]].shape)
        batch = ft[plan.partitions[0]].numel()
        pieces = []
        for lam, d in zip(plan.partitions, plan.dims):
            if lam not in ft or ft[lam].numel() != batch * d * d:
                raise ValueError(f"ft[{lam}] is missing or has the wrong number of elements")
            pieces.append(ft[lam].to(device=x.device, dtype=x.dtype))
        out = _InverseFn.apply(n, batch, *pieces).view(lead + (plan.nfact,))
    else:
        plan, packed, batch, lead, home = _pack_ft(ft, n)
        out = _inverse_packed(plan, packed, batch).view(lead + (plan.nfact,))
    if n <= BASE_CASE:
        out = out.squeeze()
    return out if home == out.device else out.to(home)


# The pruned level steps run the table-driven sweep kernels, which are several times slower per coefficient than
# the generated kernels of the full pipeline: the subset path wins while the kept partitions hold less than
# about 3 % of the coefficients (S_10 fp32, batch 128: k = 1, 2, 3 -> 3.7 / 3.9 / 5.0 ms against 7.6 ms for the
# full forward transform, k = 4 with 9 % of the coefficients 8.7 ms; profiles r02j).
_SUBSET_MAX_FRACTION = 0.03


def _inverse_subset_packed(plan, sub, packed: torch.Tensor, batch: int) -> torch.Tensor:
    """Inverse of a packed coefficient array that is zero outside the subset plan's partitions."""
    if plan.subplan_info(sub)[1] > _SUBSET_MAX_FRACTION:
        return _inverse_packed(plan, packed, batch)
    out = torch.empty((batch, plan.nfact), dtype=packed.dtype, device=packed.device)
    if batch:
        with _device_ctx(packed.device):
            ws = plan.workspace(batch)
            _lib.check(_lib.load().snfft_inverse_subset(plan.handle, sub, packed.data_ptr(), batch, out.data_ptr(),
                                                        ws.data_ptr(), ws.numel(), _stream_ptr(packed.device)),
                       "snfft_inverse_subset")
    return out


def sn_fourier_decomposition(ft: dict, n: int) -> dict:
    """Per-partition components of the inverse transform (reference fourier.py:291-295):
    {lambda: d_lambda/n! <F(lambda), rho_lambda(.)>}, whose sum over lambda is sn_ifft(ft, n).
    Each component is the inverse pipeline restricted to one partition (`snfft_inverse_subset`: the level steps
    k >= 8 only touch the shapes contained in lambda); the reference's own recursion raises for n > 5
    (fourier.py:202-214)."""
    plan, packed, batch, lead, home = _pack_ft(ft, n)
    one = torch.zeros_like(packed)
    out = {}
    for p, (lam, d, off) in enumerate(zip(plan.partitions, plan.dims, plan.offsets)):
        blk = slice(off * batch, (off + d * d) * batch)
        one[blk].copy_(packed[blk])
        comp = _inverse_subset_packed(plan, plan.subplan([p]), one, batch).view(lead + (plan.nfact,))
        one[blk].zero_()
        if n <= BASE_CASE:
            comp = comp.squeeze()
        out[lam] = comp if home == comp.device else comp.to(home)
    return out


def bandlimited_partitions(n: int, k: int):
    """The partitions of n with lambda_1 >= n - k, in generate_partitions order: the "low-frequency" irreps of
    ranking / partial-ranking analyses (k = 1: the trivial and the (n-1, 1) irrep).  Pure integer code
    (tableau.py:103-123 order), no device needed."""
    return [lam for lam in _generate_partitions(n) if lam[0] >= n - k]


_SMALL_PARTITIONS = {
    1: [(1,)], 2: [(2,), (1, 1)], 3: [(3,), (2, 1), (1, 1, 1)], 4: [(4,), (3, 1), (2, 2), (2, 1, 1), (1, 1, 1, 1)],
    5: [(5,), (4, 1), (3, 2), (3, 1, 1), (2, 2, 1), (2, 1, 1, 1), (1, 1, 1, 1, 1)],
}


def _generate_partitions(n: int):
    """generate_partitions(n) of the reference (tableau.py:103-123): hard-coded lists for n <= 5, else (n,) followed by
    sorted((k, *p)) for k = 1..n-1 and p a partition of n-k, first occurrence kept."""
    if n in _SMALL_PARTITIONS:
        return list(_SMALL_PARTITIONS[n])
    out = [(n,)]
    seen = {(n,)}
    for k in range(1, n):
        for p in _generate_partitions(n - k):
            q = tuple(sorted((k, *p), reverse=True))
            if q not in seen:
                seen.add(q)
                out.append(q)
    return out


def sn_fft_bandlimited(fn_vals: torch.Tensor, n: int, k: int = None, partitions=None) -> dict:
    """Forward transform restricted to a subset of the partitions: the partitions with
    lambda_1 >= n - k, or an explicit list `partitions`.  Returns {partition: F(partition)} for those partitions
    only, equal to the same entries of sn_fft(fn_vals, n), in generate_partitions order.  The level steps k >= 8
    compute only the shapes contained in a selected partition, so a small band costs a fraction of the full
    transform (about 40 % at n = 10, k <= 3)."""
    _check_input(fn_vals, n)
    if (k is None) == (partitions is None):
        raise ValueError("give either k or partitions")
    x, home = _to_device(fn_vals)
    lead = tuple(x.shape[:-1])
    f2d = x.reshape(-1, x.shape[-1]).contiguous()
    batch = f2d.shape[0]
    plan = get_plan(n, f2d.dtype, f2d.device)
    if partitions is None:
        if not isinstance(k, int) or k < 0:
            raise ValueError("k must be a non-negative integer")
        wanted = [lam for lam in plan.partitions if lam[0] >= n - k]
    else:
        wanted = [tuple(int(v) for v in lam) for lam in partitions]
        bad = [lam for lam in wanted if lam not in plan.index]
        if bad:
            raise ValueError(f"{bad[0]} is not a partition of {n}")
    idx = sorted({plan.index[lam] for lam in wanted})
    sub = plan.subplan(idx)
    if plan.subplan_info(sub)[1] > _SUBSET_MAX_FRACTION:
        _, packed = _forward_packed(f2d, n)            # a wide band: the full pipeline is faster, select afterwards
    else:
        packed = torch.empty(batch * plan.nfact, dtype=f2d.dtype, device=f2d.device)
    if batch and plan.subplan_info(sub)[1] <= _SUBSET_MAX_FRACTION:
        with _device_ctx(f2d.device):
            ws = plan.workspace(batch)
            _lib.check(_lib.load().snfft_forward_subset(plan.handle, sub, f2d.data_ptr(), batch, packed.data_ptr(),
                                                        ws.data_ptr(), ws.numel(), _stream_ptr(f2d.device)),
                       "snfft_forward_subset")
    result = {}
    for p in idx:
        lam, d, off = plan.partitions[p], plan.dims[p], plan.offsets[p]
        blk = packed[off * batch:(off + d * d) * batch]
        if d == 1:
            val = blk.view(lead)
        else:
            val = blk.view(lead + (d, d))
            if n <= BASE_CASE:
                val = val.squeeze()
        # a small band is cloned out of the n!-sized packed buffer so that the buffer can be freed
        val = val.clone() if plan.offsets[-1] > 8 * sum(plan.dims[q] ** 2 for q in idx) else val
        result[lam] = val if home == x.device else val.to(home)
    return result


def sn_ifft_bandlimited(ft: dict, n: int) -> torch.Tensor:
    """Inverse transform of a band-limited coefficient dict: the partitions missing from `ft` count as zero,
        f(sigma) = 1/n! sum_{lambda in ft} d_lambda <F(lambda), rho_lambda(sigma)>
    -- the projection of a function onto the isotypic components present in `ft`, e.g.
    sn_ifft_bandlimited(sn_fft_bandlimited(f, n, k=1), n) is the first-order (one-point marginal) approximation of f."""
    if not ft:
        raise ValueError("ft is empty")
    some = next(iter(ft.values()))
    if not isinstance(some, torch.Tensor):
        raise TypeError("ft must map partitions to tensors")
    x, home = _to_device(some)
    if x.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"only float32 and float64 are supported, got {x.dtype}")
    plan = get_plan(n, x.dtype, x.device)
    keys = [tuple(int(v) for v in lam) for lam in ft]
    bad = [lam for lam in keys if lam not in plan.index]
    if bad:
        raise ValueError(f"{bad[0]} is not a partition of {n}")
    first = plan.index[keys[0]]
    d0 = plan.dims[first]
    t0 = ft[list(ft)[0]]
    batch = t0.numel() // (d0 * d0)
    lead = tuple(t0.shape) if d0 == 1 else tuple(t0.shape[:-2])
    if n <= BASE_CASE and math.prod(lead or (1,)) != batch:
        lead = (batch,) if batch != 1 else ()
    packed = torch.zeros(batch * plan.nfact, dtype=x.dtype, device=x.device)
    for lam, v in zip(keys, ft.values()):
        p = plan.index[lam]
        d, off = plan.dims[p], plan.offsets[p]
        if v.numel() != batch * d * d:
            raise ValueError(f"ft[{lam}] has {v.numel()} elements, expected {batch * d * d}")
        packed[off * batch:(off + d * d) * batch].copy_(v.to(device=x.device, dtype=x.dtype).reshape(-1))
    out = _inverse_subset_packed(plan, plan.subplan([plan.index[lam] for lam in keys]), packed, batch).view(lead + (plan.nfact,))
    if n <= BASE_CASE:
        out = out.squeeze()
    return out if home == out.device else out.to(home)


def calc_power(ft: dict, n: int) -> dict:
    """Power spectrum {lambda: d_lambda/n! ||F(lambda)||_F^2} (reference fourier.py:349-393): shape () per
    partition for an unbatched ft, (batch,) for a batched one.  The values add up to ||f||^2."""
    plan, packed, batch, lead, home = _pack_ft(ft, n)
    if len(lead) > 1:
        raise ValueError("calc_power takes an unbatched or a singly batched ft like the reference")
    out = torch.empty((len(plan.partitions), batch), dtype=packed.dtype, device=packed.device)
    if batch:
        with _device_ctx(packed.device):
            _lib.check(_lib.load().snfft_power(plan.handle, packed.data_ptr(), batch, out.data_ptr(),
                                               _stream_ptr(packed.device)), "snfft_power")
    result = {}
    for p, lam in enumerate(plan.partitions):
        v = out[p].view(lead)
        result[lam] = v if home == v.device else v.to(home)
    return result


def sn_convolve(f: torch.Tensor, g: torch.Tensor, n: int) -> torch.Tensor:
    """Group convolution through the transform (the reference tests the theorem at
    tests/test_fourier.py:87-109 with the direct sum of tests/test_fourier.py:22-34):

        (f * g)(pi) = sum_sigma f(sigma) g(pi sigma^-1)        <=>        FT(f * g) = FT(g) @ FT(f)

    f, g: (n!,) or (..., n!) with equal shapes; returns the same shape.  Both transforms and the inverse run in
    this package's kernels, and so do the per-partition d x d products (`snfft_block_matmul`: an FFMA / DFMA tile
    kernel, because the fp32 1e-5 / fp64 1e-12 contract rules out the TF32 tensor path)."""
    _check_input(f, n)
    _check_input(g, n)
    if f.shape != g.shape or f.dtype != g.dtype:
        raise ValueError("f and g must have the same shape and dtype")
    xf, home = _to_device(f)
    xg, _ = _to_device(g)
    xg = xg.to(xf.device)
    plan, pf = _forward_packed(xf.reshape(-1, xf.shape[-1]).contiguous(), n)
    _, pg = _forward_packed(xg.reshape(-1, xg.shape[-1]).contiguous(), n)
    batch = pf.numel() // plan.nfact
    prod = torch.empty_like(pf)
    if batch:
        with _device_ctx(pf.device):
            _lib.check(_lib.load().snfft_block_matmul(plan.handle, pg.data_ptr(), pf.data_ptr(), prod.data_ptr(), batch,
                                                      _stream_ptr(pf.device)), "snfft_block_matmul")
    out = _inverse_packed(plan, prod, batch).view(tuple(f.shape))
    return out if home == out.device else out.to(home)


def _dense_rep(plan, part: int) -> torch.Tensor:
    d = plan.dims[part]
    out = torch.empty((plan.nfact, d, d), dtype=plan.dtype, device=plan.device)
    with _device_ctx(plan.device):
        _lib.check(_lib.load().snfft_dense_rep(plan.handle, part, out.data_ptr(), _stream_ptr(plan.device)),
                   "snfft_dense_rep")
    return out


def slow_sn_ft(fn_vals: torch.Tensor, n: int) -> dict:
    """Dense transform sum_g f(g) rho(g) (reference fourier.py:82-108): the
    cross-check for sn_fft.  rho_lambda(g) for all g is built on the device
    from the Young's-orthogonal-form tables (irreps.py:122-148)."""
    _check_input(fn_vals, n)
    if fn_vals.dim() > 2:
        raise ValueError("slow_sn_ft takes (n!,) or (batch, n!) input like the reference")
    x, home = _to_device(fn_vals)
    if x.dim() == 1:
        x = x.unsqueeze(0)                                   # fourier.py:96-97
    plan = get_plan(n, x.dtype, x.device)
    results = {}
    for part, (lam, d) in enumerate(zip(plan.partitions, plan.dims)):
        mats = _dense_rep(plan, part)
        if d == 1:
            res = torch.einsum('bi,i->b', x, mats.view(-1))            # fourier.py:103
        else:
            res = torch.einsum('bi,ijk->bjk', x, mats).squeeze()       # fourier.py:105
        results[lam] = res if home == x.device else res.to(home)
    return results


def slow_sn_ift(ft: dict, n: int) -> torch.Tensor:
    """Dense inverse (reference fourier.py:302-327), dtype preserving."""
    some = next(iter(ft.values()))
    x, home = _to_device(some)
    plan = get_plan(n, x.dtype, x.device)
    triv = ft[plan.partitions[0]]
    lead = tuple(triv.shape)
    batch = triv.numel()
    tot = torch.zeros((batch, plan.nfact), dtype=x.dtype, device=x.device)
    for part, (lam, d) in enumerate(zip(plan.partitions, plan.dims)):
        mats = _dense_rep(plan, part)
        blk = ft[lam].to(device=x.device, dtype=x.dtype).reshape(batch, d, d)
        tot += d * torch.einsum('bij,gij->bg', blk, mats)              # fourier.py:157
    out = (tot / plan.nfact).view(lead + (plan.nfact,)).squeeze()
    return out if home == out.device else out.to(home)
