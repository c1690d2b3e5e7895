"""Print the headline numbers and the per-kernel table of bench.py JSON lines (gpurun_out/bench_*.json)."""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(f"{path}: {d['value']:.0f} {d['unit']}  {d['ms_per_step']:.2f} ms/step  e2e {d['e2e']['value']:.0f}  "
          f"(floor {d['e2e'].get('pcie_floor_ms')} ms, e2e {d['e2e'].get('ms_per_step')} ms)  "
          f"err {d.get('checks', d['config']).get('round_trip_rel_err')}")
    for sr in d.get("sub_records", []):
        print(f"   sub {sr['config']['workload']}: {sr['value']:.0f}  {sr['ms_per_step']:.2f} ms/step  top {sr['roofline']['kernel']} "
              f"frac {sr['roofline']['frac']:.3f}")
    if d.get("coset"):
        c = d["coset"]
        print(f"   coset {c['config']['workload']}: {c['ms_per_step']:.3f} ms  {c.get('collective')}  compute {c.get('compute_ms_max_over_ranks')} "
              f"err {c.get('rel_err_vs_single_gpu')}")
    for k in d.get("kernels", [])[:24]:
        print(f"   {k['kernel']:24s} {k['avg_ms']:8.3f} ms  {k['share'] * 100:5.1f}%  {k['alg_gbs']:7.0f} GB/s")
