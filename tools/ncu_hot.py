"""Top SASS instructions by stall samples from an .ncu-rep captured with --import-source on.

    python tools/ncu_hot.py rep.ncu-rep [N] > hot.txt
"""
import csv, subprocess, sys, io
rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 80
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": line, "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(line)
for b in blocks:
    rows = list(csv.reader(b["rows"]))
    if not rows:
        continue
    h = rows[0]
    si, ii, src, ex = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
    data = []
    tot = 0
    toti = 0
    for k, r in enumerate(rows[1:]):
        try:
            s = int(r[si]); n = int(r[ii])
        except Exception:
            continue
        tot += s; toti += n
        data.append((s, n, k, r[src]))
    print("=====", b["name"][:160], "instructions:", len(data), "samples:", tot, "warp-inst:", toti)
    # opcode histogram by executed count and samples
    hist = {}
    for s, n, k, t in data:
        op = t.strip().split()[0] if t.strip() else "?"
        if op.startswith("@"):
            op = t.strip().split()[1]
        op = op.split(".")[0]
        e = hist.setdefault(op, [0, 0]); e[0] += n; e[1] += s
    print("opcode: executed%  samples%")
    for op, (n, s) in sorted(hist.items(), key=lambda x: -x[1][1])[:18]:
        print(f"  {op:10s} {100*n/max(toti,1):6.1f} {100*s/max(tot,1):6.1f}")
    for s, n, k, t in sorted(data, reverse=True)[:N]:
        print(f"{100*s/max(tot,1):5.2f}% samples  exec {n:>10d}  #{k:5d}  {t.strip()[:110]}")
