"""Per-launch CUDA-event times of ONE rank's share of the coset-sharded transform (snfft_forward_cosets2 on a run of
two-level cosets: BASELINE config 5), both plans profiled.

    python tools/coset_times.py --n 11 --pairs 14 [--reps 5]
"""
import argparse
import ctypes
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from algebraist_b200 import _lib  # noqa: E402
from algebraist_b200.distributed import coset_pair_shards, sn_fft_coset2_partial  # noqa: E402
from algebraist_b200.plan import get_plan  # noqa: E402

KINDS = {0: "fwd_level", 1: "inv_level", 2: "gather", 3: "scatter", 4: "fwd_base", 5: "inv_base", 6: "pack", 7: "unpack"}
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=11)
ap.add_argument("--pairs", type=int, default=14)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
f = torch.randn(math.factorial(a.n), device=dev)
pairs = [(i, j) for i in range(a.n) for j in range(a.n - 1)][:a.pairs]
plans = [get_plan(a.n, torch.float32, dev), get_plan(a.n - 2, torch.float32, dev)]
for _ in range(3):
    sn_fft_coset2_partial(f, a.n, pairs)
torch.cuda.synchronize()
for p in plans:
    _lib.check(lib.snfft_profile_begin(p.handle), "begin")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    sn_fft_coset2_partial(f, a.n, pairs)
e1.record()
torch.cuda.synchronize()
print(f"n={a.n} pairs={a.pairs}: {e0.elapsed_time(e1) / a.reps:.3f} ms per partial transform")
for name, p in zip(("top plan", "sub plan"), plans):
    recs = (_lib.ProfileRecord * 256)()
    nrec = ctypes.c_int()
    _lib.check(lib.snfft_profile_end(p.handle, recs, 256, ctypes.byref(nrec)), "end")
    tot = 0.0
    for i in range(nrec.value):
        r = recs[i]
        tot += r.total_ms / a.reps
        print(f"  {name}: {KINDS[r.kind]}_k{r.level}_c{r.kernel_class:<3d} {r.total_ms / r.launches:7.3f} ms x {r.launches // a.reps}")
    print(f"  {name}: profiled launches {tot:.3f} ms")
