"""Per-launch CUDA-event times of one forward + inverse (snfft_profile_begin / end).

    python tools/kernel_times.py --n 11 --batch 1 [--dtype f64] [--reps 5] [--forward-only]
"""
import argparse
import ctypes
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from algebraist_b200 import _lib  # noqa: E402
from algebraist_b200.plan import get_plan  # noqa: E402

KINDS = {0: "fwd_level", 1: "inv_level", 2: "gather", 3: "scatter", 4: "fwd_base", 5: "inv_base", 6: "pack", 7: "unpack"}

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=11)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--forward-only", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
tdt = torch.float32 if a.dtype == "f32" else torch.float64
lib = _lib.load()
plan = get_plan(a.n, tdt, dev)
nf = math.factorial(a.n)
f = torch.randn(a.batch, nf, dtype=tdt, device=dev)
F = torch.empty(a.batch * nf, dtype=tdt, device=dev)
ws = plan.workspace(a.batch)
st = torch.cuda.current_stream(dev).cuda_stream


def step():
    _lib.check(lib.snfft_forward(plan.handle, f.data_ptr(), a.batch, F.data_ptr(), ws.data_ptr(), ws.numel(), st), "fwd")
    if not a.forward_only:
        _lib.check(lib.snfft_inverse(plan.handle, F.data_ptr(), a.batch, f.data_ptr(), ws.data_ptr(), ws.numel(), st), "inv")


for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.check(lib.snfft_profile_begin(plan.handle), "begin")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    step()
e1.record()
torch.cuda.synchronize()
recs = (_lib.ProfileRecord * 256)()
nrec = ctypes.c_int()
_lib.check(lib.snfft_profile_end(plan.handle, recs, 256, ctypes.byref(nrec)), "end")
print(f"n={a.n} batch={a.batch} {a.dtype}: {e0.elapsed_time(e1) / a.reps:.3f} ms per step")
esize = 4 if a.dtype == "f32" else 8
for i in range(nrec.value):
    r = recs[i]
    ninst = a.batch * (nf // math.factorial(r.level)) if r.kind in (0, 1, 4, 5, 6, 7) else a.batch
    b = 2 * r.elems_per_instance * ninst * esize
    ms = r.total_ms / r.launches
    print(f"   {KINDS[r.kind]}_k{r.level}_c{r.kernel_class:<3d} x{r.launches:<3d} {ms:8.3f} ms  {b / ms / 1e6:8.0f} GB/s")
