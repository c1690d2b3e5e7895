"""Multi-GPU use of the transform (one process per GPU, torch.distributed).

Batched transforms shard over the batch and need no collective: the rows of
`fn_vals` are independent functions (the reference vmaps over them,
fourier.py:263), so every rank transforms its own rows with its own plan.
`sn_fft_batch_sharded` / `sn_ifft_batch_sharded` do exactly that; an
all_gather is issued only if the caller asks for the full result."""
import math

import torch
import torch.distributed as dist

from .fourier import sn_fft, sn_ifft


def shard_bounds(batch: int, world_size: int, rank: int):
    """Rows [lo, hi) of a batch of `batch` functions owned by `rank`: contiguous,
    sizes differ by at most one, the first `batch % world_size` ranks get the extra row."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _irrep_dim(lam) -> int:
    """Hook-length formula (tableau.py:236-265)."""
    n = sum(lam)
    prod = 1
    for i, row in enumerate(lam):
        for j in range(row):
            prod *= row - j + sum(1 for r in lam[i + 1:] if r > j)
    return math.factorial(n) // prod


def sn_fft_batch_sharded(fn_vals: torch.Tensor, n: int, group=None, gather: bool = False) -> dict:
    """Forward transform of the rows this rank owns of a (B, n!) tensor that every
    rank holds (or of a local shard if `fn_vals` is already the local rows: pass
    the shard and use sn_fft directly).  Returns the local dict; with gather=True
    the per-rank dicts are all_gathered into full (B, ...) tensors."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if fn_vals.dim() != 2:
        raise ValueError("sn_fft_batch_sharded takes a (batch, n!) tensor")
    lo, hi = shard_bounds(fn_vals.shape[0], world, rank)
    local = sn_fft(fn_vals[lo:hi], n)
    if not gather or world == 1:
        return local
    out = {}
    full = fn_vals.shape[0]
    spans = [shard_bounds(full, world, r) for r in range(world)]
    longest = max(b - a for a, b in spans)
    for lam, v in local.items():
        d = _irrep_dim(lam)         # from the partition (a rank that owns no rows cannot infer it from its tensor)
        # n <= 5 squeezes a single local row (fourier.py:129): restore the batch axis before gathering
        v = v.reshape((hi - lo,) + ((d, d) if d > 1 else ()))
        # all_gather wants equal shapes: pad the short shards with zero rows, trim afterwards
        padded = torch.zeros((longest,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        padded[:hi - lo] = v
        pieces = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(pieces, padded, group=group)
        out[lam] = torch.cat([pc[:b - a] for pc, (a, b) in zip(pieces, spans)], dim=0)
    return out


def sn_ifft_batch_sharded(ft: dict, n: int) -> torch.Tensor:
    """Inverse of a local (per-rank) coefficient dict: no communication."""
    return sn_ifft(ft, n)


# ---------------------------------------------------------------------------
# One large function sharded over the top-level cosets (BASELINE config 5)
# ---------------------------------------------------------------------------
def coset_shards(n: int, world_size: int):
    """Top-level cosets {sigma : sigma[i] = n-1}, i in range(n), dealt to the ranks as contiguous
    ranges whose sizes differ by at most one (rank r owns cosets [lo_r, hi_r))."""
    return [shard_bounds(n, world_size, r) for r in range(world_size)]


def coset_pair_shards(n: int, world_size: int):
    """The n (n-1) two-level cosets (i_n, i_{n-1}) -- sigma[i_n] = n-1 and, inside that coset, value n-2 at
    position i_{n-1} -- dealt to the ranks as contiguous runs of the (i_n major) list whose sizes differ by at
    most one: 110 units at n = 11, i.e. 14 / 13 per rank on 8 GPUs instead of 2 / 1 top-level cosets."""
    pairs = [(i, j) for i in range(n) for j in range(n - 1)]
    return [pairs[lo:hi] for lo, hi in (shard_bounds(len(pairs), world_size, r) for r in range(world_size))]


def partition_owners(dims, world_size: int):
    """Which rank ends up with which partition after the reduce-scatter.

    Greedy bin packing of the partitions by d^2 (largest first into the emptiest bin); returns
    (owner[p] for every partition, bin_size) with bin_size = the largest bin: ncclReduceScatter wants
    equal receive counts, so every rank's slot of the send buffer is padded to that size."""
    order = sorted(range(len(dims)), key=lambda p: -dims[p] * dims[p])
    load = [0] * world_size
    owner = [0] * len(dims)
    for p in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        owner[p] = r
        load[r] += dims[p] * dims[p]
    return owner, max(load) if load else 0


def _reduce_scatter_sum(send: torch.Tensor, recv: torch.Tensor, group=None):
    """recv <- sum over ranks of this rank's slot of `send` ([world][slot])."""
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(recv, send, op=dist.ReduceOp.SUM, group=group)
    else:
        # gloo (the CPU test-suite's transport) has no reduce_scatter: all_reduce, keep the own slot
        dist.all_reduce(send, op=dist.ReduceOp.SUM, group=group)
        recv.copy_(send.view(dist.get_world_size(group), -1)[dist.get_rank(group)])


def sn_fft_coset_partial(fn_vals: torch.Tensor, n: int, cosets):
    """The terms of the top-level Clausen sum (fourier.py:253-273) that `cosets` own, as a packed
    coefficient tensor ((B * n!,), same layout as sn_fft's internal buffer) plus the plan."""
    from . import _lib
    from .fourier import _check_input, _device_ctx, _stream_ptr, _to_device
    from .plan import get_plan
    import ctypes
    _check_input(fn_vals, n)
    if n < 3:
        raise ValueError("coset sharding needs n >= 3")
    x, _ = _to_device(fn_vals)
    f2d = x.reshape(-1, x.shape[-1]).contiguous()
    batch = f2d.shape[0]
    cosets = sorted(int(c) for c in cosets)
    plan = get_plan(n, f2d.dtype, f2d.device)
    sub = get_plan(n - 1, f2d.dtype, f2d.device)
    packed = torch.empty(batch * plan.nfact, dtype=f2d.dtype, device=f2d.device)
    if batch:
        lib = _lib.load()
        arr = (ctypes.c_int * max(len(cosets), 1))(*cosets)
        with _device_ctx(f2d.device):
            nbytes = lib.snfft_forward_cosets_workspace_bytes(plan.handle, sub.handle, batch, len(cosets))
            ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=f2d.device)
            _lib.check(lib.snfft_forward_cosets(plan.handle, sub.handle, f2d.data_ptr(), batch, arr, len(cosets),
                                                packed.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(f2d.device)),
                       "snfft_forward_cosets")
    return plan, packed, batch


def sn_fft_coset2_partial(fn_vals: torch.Tensor, n: int, pairs):
    """The terms of the two top Clausen sums (fourier.py:253-273 applied at levels n and n-1) that the two-level
    cosets `pairs` = [(i_n, i_{n-1}), ...] own, as a packed coefficient tensor plus the plan: the pairs'
    restrictions run through the S_{n-2} pipeline as one batch and the last two level steps run on their
    (otherwise zero) level array."""
    from . import _lib
    from .fourier import _check_input, _device_ctx, _stream_ptr, _to_device
    from .plan import get_plan
    import ctypes
    _check_input(fn_vals, n)
    if n < 4:
        raise ValueError("two-level coset sharding needs n >= 4")
    x, _ = _to_device(fn_vals)
    f2d = x.reshape(-1, x.shape[-1]).contiguous()
    batch = f2d.shape[0]
    pairs = sorted((int(i), int(j)) for i, j in pairs)
    plan = get_plan(n, f2d.dtype, f2d.device)
    sub = get_plan(n - 2, f2d.dtype, f2d.device)
    packed = torch.empty(batch * plan.nfact, dtype=f2d.dtype, device=f2d.device)
    if batch:
        lib = _lib.load()
        top = (ctypes.c_int * max(len(pairs), 1))(*[i for i, _ in pairs])
        low = (ctypes.c_int * max(len(pairs), 1))(*[j for _, j in pairs])
        with _device_ctx(f2d.device):
            nbytes = lib.snfft_forward_cosets2_workspace_bytes(plan.handle, sub.handle, batch, len(pairs))
            ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=f2d.device)
            _lib.check(lib.snfft_forward_cosets2(plan.handle, sub.handle, f2d.data_ptr(), batch, top, low, len(pairs),
                                                 packed.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(f2d.device)),
                       "snfft_forward_cosets2")
    return plan, packed, batch


def sn_fft_coset_sharded(fn_vals: torch.Tensor, n: int, group=None, levels: int = None) -> dict:
    """Forward transform of one large function (or a small batch) that every rank holds, sharded over cosets:
    rank r transforms the restrictions of f to its cosets, applies its terms of the top level steps, and one sum
    reduce-scatter over the packed coefficients leaves every rank with a subset of the partitions.  Returns
    {partition: F(partition)} for the partitions THIS rank owns (shapes as sn_fft); `partition_owners` tells
    which rank has which.

    levels = 2 (default for n >= 4): the units are the n (n-1) two-level cosets (`coset_pair_shards`, S_{n-2}
    sub-transforms); levels = 1: the n top-level cosets (`coset_shards`, S_{n-1} sub-transforms)."""
    if not dist.is_initialized():
        raise RuntimeError("sn_fft_coset_sharded needs an initialised torch.distributed process group")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if levels is None:
        levels = 2 if n >= 4 else 1
    if levels == 2:
        plan, packed, batch = sn_fft_coset2_partial(fn_vals, n, coset_pair_shards(n, world)[rank])
    elif levels == 1:
        lo, hi = coset_shards(n, world)[rank]
        plan, packed, batch = sn_fft_coset_partial(fn_vals, n, range(lo, hi))
    else:
        raise ValueError("levels must be 1 or 2")
    owner, bin_size = partition_owners(plan.dims, world)
    # send buffer [world][bin_size * batch]: the partitions of bin r back to back, zero padded (one batched copy)
    pieces, slots, fill = [[] for _ in range(world)], {}, [0] * world
    for p, (d, off) in enumerate(zip(plan.dims, plan.offsets)):
        r = owner[p]
        pieces[r].append(packed[off * batch:(off + d * d) * batch])
        slots[p] = fill[r]
        fill[r] += d * d
    pad = torch.zeros(max(bin_size - min(fill), 0) * batch, dtype=packed.dtype, device=packed.device)
    flat = []
    for r in range(world):
        flat += pieces[r]
        if fill[r] < bin_size:
            flat.append(pad[:(bin_size - fill[r]) * batch])
    send = torch.cat(flat)
    recv = torch.empty(bin_size * batch, dtype=packed.dtype, device=packed.device)
    _reduce_scatter_sum(send, recv, group)
    lead = tuple(fn_vals.shape[:-1])
    out = {}
    for p, (lam, d) in enumerate(zip(plan.partitions, plan.dims)):
        if owner[p] != rank:
            continue
        blk = recv[slots[p] * batch:(slots[p] + d * d) * batch]
        out[lam] = blk.view(lead) if d == 1 else blk.view(lead + (d, d))
    return out
