"""Write profiles/<name>_traffic.json from the summary CSV of an `ncu --set full` capture of
tools/prof_step.py (tools/ncu_summary.py): DRAM read + write bytes per launch of the kernels the bench
line names.  The level kernels of one forward + inverse run in a fixed order, so the k-th launch of a
kernel family is matched to the bench name by its template arguments and grid.

    python tools/ncu_traffic.py summary.csv batch dtype out.json [n]
"""
import csv
import re
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def to_bytes(s):
    v, u = s.split()
    return float(v) * UNIT[u]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    batch, dtype, out = int(sys.argv[2]), sys.argv[3], sys.argv[4]
    h = rows[0]
    ki, ti, ri, wi = h.index("kernel"), h.index("time"), h.index("dram_rd"), h.index("dram_wr")
    recs = []
    # dominant kernels of the bench line: the largest launch of each sweep-kernel class, by family
    fam = {"inv_level_kernel<float, 4, 4, 64": "inv_level_k10_c3", "fwd_level_kernel<float, 4, 4, 64": "fwd_level_k10_c3",
           "inv_level_kernel<double, 2, 4, 64": "inv_level_k10_c3", "fwd_level_kernel<double, 2, 4, 64": "fwd_level_k10_c3"}
    # column kernels (col_kernels.cuh): one launch per level and direction
    # (names as ncu prints them: col_kernel<float, 8, 0, 640>, col_kernel<F2, 8, 0, 384> -- the F2 instantiation runs two
    # fp32 instances per thread)
    cpat = re.compile(r"col_kernel<(?:snfft::)?(float|double|F2), (?:\(int\))?(\d+), (?:\(bool\))?([01])")
    # top / bottom kernels: the T kernel names its level and direction; the Z kernel (level-7 operators) runs for
    # k = 9 and 10 in a fixed order inside one forward + inverse: forward 9, forward 10, inverse 10, inverse 9
    # t_kernel<TzVec<float, 4>, 10, 0, 1>: level, direction, packed (the top level's kernel on the public layout: class -7)
    tpat = re.compile(r"t_kernel<.*?(float|double)(?:, (?:\(int\))?\d)?>?, (?:\(int\))?(\d+), (?:\(bool\))?([01])(?:, (?:\(bool\))?([01]))?")
    zorder = {"0": ["fwd_level_k9_c-4", "fwd_level_k10_c-4"], "1": ["inv_level_k10_c-4", "inv_level_k9_c-4"]}
    zseen = {"0": 0, "1": 0}
    best = {}
    for r in rows[1:]:
        m = tpat.search(r[ki])
        if m:
            name = f"{'inv' if m.group(3) == '1' else 'fwd'}_level_k{m.group(2)}_c{'-7' if m.group(4) == '1' else '-5'}"
            best[name] = (float(r[ti].split()[0]), to_bytes(r[ri]) + to_bytes(r[wi]), r[ti])
            continue
        m = re.search(r"s6_kernel<float, (?:\(bool\))?([01])>", r[ki])
        if m:
            name = f"{'inv' if m.group(1) == '1' else 'fwd'}_base_k6_c-8"
            best[name] = (float(r[ti].split()[0]), to_bytes(r[ri]) + to_bytes(r[wi]), r[ti])
            continue
        m = cpat.search(r[ki])
        if m:
            name = f"{'inv' if m.group(3) == '1' else 'fwd'}_level_k{m.group(2)}_c-3"
            best[name] = (float(r[ti].split()[0]), to_bytes(r[ri]) + to_bytes(r[wi]), r[ti])
            continue
        if "z_kernel<" in r[ki]:
            inv = "1" if (", 1>" in r[ki] or "true>" in r[ki]) else "0"
            if zseen[inv] < len(zorder[inv]):
                name = zorder[inv][zseen[inv]]
                zseen[inv] += 1
                best[name] = (float(r[ti].split()[0]), to_bytes(r[ri]) + to_bytes(r[wi]), r[ti])
            continue
        for pat, name in fam.items():
            if r[ki].startswith(pat) or (("col_kernel" in pat or "t_kernel" in pat) and pat in r[ki]):
                t = float(r[ti].split()[0])
                if name not in best or t > best[name][0]:
                    best[name] = (t, to_bytes(r[ri]) + to_bytes(r[wi]), r[ti])
    import os
    import subprocess
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                                cwd=os.path.dirname(os.path.abspath(__file__))).stdout.strip() or None
    except OSError:
        commit = None
    commit = os.environ.get("SNFFT_COMMIT", commit)
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    for name, (t, b, ts) in best.items():
        recs.append({"kernel": name, "n": n, "batch": batch, "dtype": dtype, "dram_bytes_per_launch": b, "ncu_time": ts,
                     "capture": os.path.basename(sys.argv[1]), "commit": commit})
    json.dump(recs, open(out, "w"), indent=1)
    print(recs)


main()
