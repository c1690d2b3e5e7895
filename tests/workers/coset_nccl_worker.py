"""torchrun worker of tests/test_gpu_multirank.py: ONE function sharded over cosets on N NCCL ranks
(sn_fft_coset_sharded, both granularities) against the single-GPU transform of the same function."""
import math
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    from algebraist_b200 import sn_fft
    from algebraist_b200.distributed import partition_owners, sn_fft_coset_sharded
    from algebraist_b200.plan import get_plan
    gen = torch.Generator(device=dev)
    gen.manual_seed(123)
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-12)):
        f = torch.randn(math.factorial(n), dtype=dtype, device=dev, generator=gen)
        full = sn_fft(f, n)
        plan = get_plan(n, dtype, dev)
        owner, _ = partition_owners(plan.dims, world)
        for levels in (2, 1):
            mine = sn_fft_coset_sharded(f, n, levels=levels)
            assert set(mine) == {lam for lam, o in zip(plan.partitions, owner) if o == rank}, "ownership"
            worst = 0.0
            for lam, v in mine.items():
                assert v.shape == full[lam].shape
                worst = max(worst, ((v - full[lam]).double().norm() / full[lam].double().norm().clamp_min(1e-300)).item())
            assert worst < tol, (dtype, levels, worst)
            count = torch.tensor([len(mine)], device=dev)
            dist.all_reduce(count)
            assert count.item() == len(plan.partitions)
    dist.barrier()
    if rank == 0:
        print(f"COSET_NCCL_OK n={n} world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
