"""Device times of the band-limited transforms against the full ones (S_10 fp32, batch 128).

    python tools/bandlimited_time.py > gpurun_out/bandlimited.json
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import algebraist_b200 as A  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n, B = 10, 128
f = torch.randn(B, math.factorial(n), device="cuda")
out = {"n": n, "batch": B, "dtype": "f32", "sn_fft_ms": timed(lambda: A.sn_fft(f, n))}
ft = A.sn_fft(f, n)
out["sn_ifft_ms"] = timed(lambda: A.sn_ifft(ft, n))
for k in (1, 2, 3, 4):
    parts = A.bandlimited_partitions(n, k)
    frac = sum(ft[lam][0].numel() for lam in parts) / math.factorial(n)
    band = A.sn_fft_bandlimited(f, n, k=k)
    out[f"k={k}"] = {"partitions": len(parts), "fraction_of_coefficients": frac,
                     "sn_fft_bandlimited_ms": timed(lambda: A.sn_fft_bandlimited(f, n, k=k)),
                     "sn_ifft_bandlimited_ms": timed(lambda: A.sn_ifft_bandlimited(band, n))}
small = f[:8].contiguous()
fts = A.sn_fft(small, n)
out["sn_fourier_decomposition_batch8_ms"] = timed(lambda: A.sn_fourier_decomposition(fts, n), reps=2)
out["full_inverse_batch8_x42_ms"] = 42 * timed(lambda: A.sn_ifft(fts, n))
print(json.dumps(out, indent=1))
