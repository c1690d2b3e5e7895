"""Device time of sn_convolve (two forward transforms, the per-partition products, one inverse) at the bench workload."""
import json
import math
import sys

import torch

sys.path.insert(0, ".")
from algebraist_b200 import sn_convolve

n, B = 10, 128
f = torch.randn(B, math.factorial(n), device="cuda")
g = torch.randn(B, math.factorial(n), device="cuda")
for _ in range(2):
    h = sn_convolve(f, g, n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    h = sn_convolve(f, g, n)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"workload": "sn_convolve S10 f32 batch 128", "ms": ms, "convolutions_per_s": B / ms * 1e3}))
