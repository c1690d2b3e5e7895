"""Multi-process path on CPU: world_size 2 over gloo.  The batch is sharded with no
collective on the data path; the optional gather uses all_gather.
The ranks run the package on the emulation build (no GPU in this container)."""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, build_emu


def test_shard_bounds():
    from algebraist_b200.distributed import shard_bounds
    for batch in (0, 1, 7, 128, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, n, batch, emu_path, out_dir):
    import contextlib
    import ctypes
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from algebraist_b200 import _lib
    from algebraist_b200.distributed import shard_bounds, sn_fft_batch_sharded, sn_ifft_batch_sharded
    from conftest import package_on_emu
    emu = _lib.bind(ctypes.CDLL(emu_path))
    torch.manual_seed(7)                                   # every rank builds the same full batch
    f = torch.randn(batch, math.factorial(n), dtype=torch.float64)
    with package_on_emu(emu):
        local = sn_fft_batch_sharded(f, n)
        lo, hi = shard_bounds(batch, world, rank)
        back = sn_ifft_batch_sharded(local, n)
        assert torch.allclose(back.reshape(hi - lo, f.shape[1]), f[lo:hi], atol=1e-12)
        full = sn_fft_batch_sharded(f, n, gather=True)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{"_".join(map(str, k)): v.numpy() for k, v in full.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,batch", [(6, 5), (4, 4), (6, 1)])
def test_batch_sharding_world2_gloo(tmp_path, n, batch):
    from oracle import sn_oracle as O
    emu_path = build_emu()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n, batch, emu_path, str(tmp_path)), nprocs=2, join=True)
    torch.manual_seed(7)
    f = torch.randn(batch, math.factorial(n), dtype=torch.float64).numpy()
    ref = O.sn_fft_packed(f, n)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        for lam, r in ref.items():
            v = got["_".join(map(str, lam))]
            assert np.allclose(v.reshape(r.shape), r, atol=1e-11), (rank, lam)


def test_coset_shards_and_partition_owners():
    from algebraist_b200.distributed import coset_shards, partition_owners
    from oracle import sn_oracle as O
    for n, world in ((11, 8), (11, 2), (12, 4), (5, 8)):
        spans = coset_shards(n, world)
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    dims = [O.dim(lam) for lam in O.generate_partitions(9)]
    for world in (1, 2, 4, 8):
        owner, bin_size = partition_owners(dims, world)
        loads = [sum(d * d for d, o in zip(dims, owner) if o == r) for r in range(world)]
        assert sum(loads) == math.factorial(9) and max(loads) == bin_size
        assert bin_size <= math.factorial(9) / world + max(dims) ** 2     # greedy: within one block of even


def _coset_worker(rank, world, port, n, batch, emu_path, out_dir):
    import ctypes
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from algebraist_b200 import _lib
    from algebraist_b200.distributed import sn_fft_coset_sharded
    from conftest import package_on_emu
    emu = _lib.bind(ctypes.CDLL(emu_path))
    torch.manual_seed(11)                                  # every rank holds the same function(s)
    f = torch.randn(batch, math.factorial(n), dtype=torch.float64) if batch else torch.randn(math.factorial(n), dtype=torch.float64)
    with package_on_emu(emu):
        mine = sn_fft_coset_sharded(f, n)
    np.savez(os.path.join(out_dir, f"coset_rank{rank}.npz"), **{"_".join(map(str, k)): v.numpy() for k, v in mine.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,batch", [(6, 0), (5, 2)])
def test_coset_sharding_world2_gloo(tmp_path, n, batch):
    """One function sharded over the top-level cosets: the partitions the two ranks end up with are
    disjoint, cover everything, and equal the oracle's transform."""
    from oracle import sn_oracle as O
    from algebraist_b200.distributed import partition_owners
    emu_path = build_emu()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_coset_worker, args=(2, port, n, batch, emu_path, str(tmp_path)), nprocs=2, join=True)
    torch.manual_seed(11)
    f = (torch.randn(batch, math.factorial(n), dtype=torch.float64) if batch
         else torch.randn(math.factorial(n), dtype=torch.float64)).numpy()
    ref = O.sn_fft_packed(f.reshape(-1, f.shape[-1]), n)
    parts = list(O.generate_partitions(n))
    owner, _ = partition_owners([O.dim(lam) for lam in parts], 2)
    seen = set()
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"coset_rank{rank}.npz"))
        for k in got.files:
            lam = tuple(int(x) for x in k.split("_"))
            assert owner[parts.index(lam)] == rank and lam not in seen
            seen.add(lam)
            assert np.allclose(got[k].reshape(ref[lam].shape), ref[lam], atol=1e-11), (rank, lam)
    assert seen == set(parts)
