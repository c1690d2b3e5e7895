"""Pinned host <-> device copy bandwidth of this box (context for bench.py's e2e number)."""
import json
import torch

n = 1857945600 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
o = torch.empty(n, dtype=torch.float32).pin_memory()
res = {}
s2 = torch.cuda.Stream()
for name in ("h2d", "d2h", "both"):
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if name in ("h2d", "both"):
            d.copy_(h, non_blocking=True)
        if name == "d2h":
            o.copy_(d, non_blocking=True)
        if name == "both":
            with torch.cuda.stream(s2):
                o.copy_(d, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[name] = {"ms": round(best, 2), "GB/s_per_direction": round(n * 4 / best / 1e6, 1)}
print(json.dumps(res))
