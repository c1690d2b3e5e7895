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
        d = 1 if v.numel() == hi - lo else int(round(math.sqrt(v.numel() // max(hi - lo, 1))))
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
