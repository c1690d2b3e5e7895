"""Experiment: forward+inverse of 128 S_10 functions as one call vs two half-batches on two CUDA streams."""
import ctypes, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from algebraist_b200 import _lib
from algebraist_b200.plan import get_plan
n, B = 10, 128
dev = torch.device("cuda", 0)
lib = _lib.load()
plan = get_plan(n, torch.float32, dev)
nf = math.factorial(n)
f = torch.randn(B, nf, device=dev)
def make(bs):
    return dict(F=torch.empty(bs * nf, device=dev), f2=torch.empty(bs, nf, device=dev), ws=plan.workspace(bs))
def run(fpart, bufs, stream):
    bs = fpart.shape[0]
    _lib.check(lib.snfft_forward(plan.handle, fpart.data_ptr(), bs, bufs["F"].data_ptr(), bufs["ws"].data_ptr(), bufs["ws"].numel(), stream.cuda_stream), "f")
    _lib.check(lib.snfft_inverse(plan.handle, bufs["F"].data_ptr(), bs, bufs["f2"].data_ptr(), bufs["ws"].data_ptr(), bufs["ws"].numel(), stream.cuda_stream), "i")
def timeit(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
full = make(B)
s0 = torch.cuda.current_stream()
print("one call      :", timeit(lambda: run(f, full, s0)), "ms")
for parts in (2, 4):
    bs = B // parts
    bufs = [make(bs) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    def go():
        ev = torch.cuda.Event(); ev.record(s0)
        for p in range(parts):
            streams[p].wait_event(ev)
            run(f[p * bs:(p + 1) * bs], bufs[p], streams[p])
        for p in range(parts):
            e = torch.cuda.Event(); e.record(streams[p]); s0.wait_event(e)
    print(parts, "streams     :", timeit(go), "ms")
    def seq():
        for p in range(parts): run(f[p * bs:(p + 1) * bs], bufs[p], s0)
    print(parts, "sequential  :", timeit(seq), "ms")
