"""NCCL paths on real GPUs: needs >= 2 visible devices (gpurun --gpus 2); skipped on a one-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("n", [8, 10])
def test_coset_sharded_nccl(n):
    """sn_fft_coset_sharded on every visible GPU (two-level and top-level cosets, fp32 and fp64): every rank's
    partitions equal the single-GPU transform and together they cover all partitions."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    world = min(ngpu, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "workers", "coset_nccl_worker.py"), str(n)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert f"COSET_NCCL_OK n={n} world={world}" in res.stdout
