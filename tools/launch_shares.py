"""Kernel shares of ONE device-timed step from an ncu launch list of bench.py (gpu__time_duration.sum per launch).

    python tools/launch_shares.py profiles/r02z_launches.csv [profiles/r02z_bench.json]

The bench command also runs warm-up steps and the end-to-end leg (chunks of 8 functions): only launches with the full
batch's grid (the largest grid of each kernel) are counted, averaged per launch, and set beside the CUDA-event shares of
the bench line when one is given.  ncu's times are cold-cache and serialised: the SHARES must agree, not the absolutes."""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, gi, vi = h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Value")
by = collections.defaultdict(list)
for r in rows[1:]:
    name = r[ki]
    if "snfft::" not in name:
        continue
    m = re.search(r"(gather_vec_kernel<float, [01]>|s6_kernel<float, [01]>|col_kernel<[^>]*>|z_kernel<[^>]*>|t_kernel<snfft::TzVec<[^>]*>[^>]*>|s5_kernel<[^>]*>|pack_vec_kernel<[^>]*>)", name)
    if not m:
        continue
    grid = int(r[gi].strip("()").split(",")[0])
    by[m.group(1)].append((grid, float(r[vi].replace(",", ""))))
full = {}
for k, v in by.items():
    grids = sorted({g for g, _ in v}, reverse=True)
    # the Z kernel runs for k = 9 and k = 10 (two full-size grids)
    keep = grids[:2] if k.startswith("z_kernel") else grids[:1]
    tmax = max(t for _, t in v)
    for g in keep:
        # kernels whose grid does not depend on the batch (gather / scatter walk over the batch inside the CTA): the
        # full-batch launches are the long ones
        ts = [t for gg, t in v if gg == g and t >= 0.5 * tmax]
        if ts and sum(ts) / len(ts) >= 1e5:          # variants that only run in the chunked end-to-end leg are tiny
            full[(k, g)] = (sum(ts) / len(ts) / 1e6, len(ts))
tot = sum(t for t, _ in full.values())
print(f"one step = {tot:.2f} ms of kernel time under ncu ({len(full)} launches)")
ev = {}
if len(sys.argv) > 2:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    ev = {k["kernel"]: k for k in d["kernels"]}
    print(f"bench line: {d['ms_per_step']:.2f} ms/step by CUDA events")
for (k, g), (t, n) in sorted(full.items(), key=lambda x: -x[1][0]):
    print(f"  {k:58s} grid {g:7d}  {t:6.3f} ms  {100 * t / tot:5.1f} %   ({n} launches)")
if ev:
    print("CUDA-event shares of the same step (bench line):")
    for k in sorted(ev.values(), key=lambda k: -k["share"]):
        print(f"  {k['kernel']:24s} {k['avg_ms']:6.3f} ms  {100 * k['share']:5.1f} %")
