"""End-to-end pipeline study on one GPU: PCIe floor, device time per chunk size, stream_batches settings.

    python tools/e2e_sweep.py [--n 10] [--batch 128]  > gpurun_out/e2e_sweep.json
"""
import argparse
import json
import math
import sys
import os
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from algebraist_b200 import sn_fft, sn_ifft, stream_batches  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10)
    ap.add_argument("--batch", type=int, default=128)
    a = ap.parse_args()
    n, B = a.n, a.batch
    nf = math.factorial(n)
    dev = torch.device("cuda", 0)
    out = {"n": n, "batch": B}
    f_host = torch.randn(B, nf).pin_memory()
    o_host = torch.empty(B, nf).pin_memory()
    d_in = torch.empty(B, nf, device=dev)
    d_out = torch.randn(B, nf, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(fn, reps=5):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        return ts

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(f_host, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            o_host.copy_(d_out, non_blocking=True)

    def both(chunk=None):
        if chunk is None:
            h2d()
            d2h()
            return
        for lo in range(0, B, chunk):
            with torch.cuda.stream(s1):
                d_in[lo:lo + chunk].copy_(f_host[lo:lo + chunk], non_blocking=True)
            with torch.cuda.stream(s2):
                o_host[lo:lo + chunk].copy_(d_out[lo:lo + chunk], non_blocking=True)

    gb = B * nf * 4 / 1e9
    out["h2d_ms"] = timed(h2d)
    out["d2h_ms"] = timed(d2h)
    out["both_ms"] = timed(both)
    out["both_chunk16_ms"] = timed(lambda: both(16))
    out["gb_per_direction"] = gb
    out["device_ms_by_batch"] = {}
    for b in (4, 8, 16, 32, 64, 128):
        if b > B:
            continue
        x = torch.randn(b, nf, device=dev)
        for _ in range(2):
            sn_ifft(sn_fft(x, n), n)
        out["device_ms_by_batch"][b] = min(timed(lambda: sn_ifft(sn_fft(x, n), n), 4))
        del x
    out["e2e"] = {}
    cfgs = [(16, 3, True), (16, 4, True), (32, 3, True), (8, 3, True), (16, 3, False), (64, 2, True), (8, 4, True), (8, 3, False),
            (4, 3, True), (12, 3, True)]
    for rep, (chunk, nslots, ramp) in enumerate(cfgs + cfgs[::-1]):
        def step():
            stream_batches(lambda x: sn_ifft(sn_fft(x, n), n), f_host, o_host, chunk=chunk, device=dev, nslots=nslots, ramp=ramp)
        for _ in range(2):
            step()
        out["e2e"][f"{rep}_chunk{chunk}_slots{nslots}_ramp{int(ramp)}"] = [round(t, 2) for t in timed(step, 8)]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
