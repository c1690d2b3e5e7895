"""Print the headline numbers and the per-kernel table of bench.py JSON lines (gpurun_out/bench_*.json)."""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(f"{path}: {d['value']:.0f} {d['unit']}  {d['ms_per_step']:.2f} ms/step  e2e {d['e2e']['value']:.0f}  "
          f"err {d['config'].get('round_trip_rel_err')}")
    for k in d.get("kernels", [])[:24]:
        print(f"   {k['kernel']:24s} {k['avg_ms']:8.3f} ms  {k['share'] * 100:5.1f}%  {k['alg_gbs']:7.0f} GB/s")
