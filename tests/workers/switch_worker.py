"""Worker of tests/test_gpu_switches.py: forward + inverse of seeded inputs with whatever kernel switches the
environment sets (SNFFT_COL_V2, SNFFT_TZ_ZV2, SNFFT_TZ_PACK, SNFFT_S6: read once per process by the library);
prints one SHA-256 per case over the packed coefficients and the reconstruction."""
import hashlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from algebraist_b200 import sn_fft, sn_ifft
    for n, B in ((7, 8), (8, 12), (9, 8), (10, 4)):
        g = torch.Generator(device="cuda")
        g.manual_seed(1000 * n + B)
        f = torch.randn(B, math.factorial(n), device="cuda", generator=g)
        ft = sn_fft(f, n)
        back = sn_ifft(ft, n)
        h = hashlib.sha256()
        for lam, v in ft.items():
            h.update(v.contiguous().cpu().numpy().tobytes())
        h.update(back.cpu().numpy().tobytes())
        err = ((back - f).norm() / f.norm()).item()
        print(f"CASE n={n} B={B} sha={h.hexdigest()} err={err:.3e}")


if __name__ == "__main__":
    main()
