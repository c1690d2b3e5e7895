"""One forward + inverse transform for profiler captures (ncu): no warm-up loops, no timing.

    ncu --set full --clock-control none --import-source on -k regex:level_kernel -o gpurun_out/x \
        python tools/prof_step.py --n 10 --batch 32
"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from algebraist_b200 import sn_fft, sn_ifft

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
dt = torch.float32 if a.dtype == "f32" else torch.float64
torch.manual_seed(0)
f = torch.randn(a.batch, math.factorial(a.n), dtype=dt, device="cuda")
for _ in range(a.reps):
    ft = sn_fft(f, a.n)
    g = sn_ifft(ft, a.n)
torch.cuda.synchronize()
print("rel err", float((g - f).norm() / f.norm()))
