"""Time the UNMODIFIED reference in the build container, where /root/reference exists.

    python tools/time_reference_here.py > profiles/r02_reference_cpu_build_container.json

The reference cannot travel to the GPU box, so bench.py reports these numbers as context next to the CPU port it
times on the box itself.  Config 1: S_5, one fp32 function, sn_fft + sn_ifft round trip (README.md:34-53), median of
20.  Config 2: S_8 fp32 forward, cold call on a 4-row batch (includes the table construction the reference performs on
every call), scaled linearly to 4096 rows.
"""
import json
import math
import os
import statistics
import sys
import time

sys.path.insert(0, "/root/reference/src")
import torch  # noqa: E402
from algebraist.fourier import sn_fft, sn_ifft  # noqa: E402


def cpu_model():
    for ln in open("/proc/cpuinfo"):
        if ln.startswith("model name"):
            return ln.split(":", 1)[1].strip()
    return "unknown"


torch.manual_seed(0)
f5 = torch.randn(math.factorial(5))
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    sn_ifft(sn_fft(f5, 5), 5)
    ts.append(time.perf_counter() - t0)
out = {"where": "build container (no GPU)", "cpu_model": cpu_model(), "cpus": os.cpu_count(),
       "torch_threads": torch.get_num_threads(), "torch": torch.__version__,
       "config1_S5_round_trip_ms_median_of_20": 1e3 * statistics.median(ts)}
if "--skip-s8" not in sys.argv:
    f8 = torch.randn(4, math.factorial(8))
    t0 = time.perf_counter()
    sn_fft(f8, 8)
    dt = time.perf_counter() - t0
    out["config2_S8_forward_batch4_cold_s"] = dt
    out["config2_S8_forward_batch4096_scaled_s"] = dt * 1024
    out["config2_S8_forward_per_s"] = 4 / dt
print(json.dumps(out, indent=1))
