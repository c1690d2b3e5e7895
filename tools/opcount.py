"""Work model of one forward level step: rotations per sweep index j with support restriction."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "algebraist_b200", "csrc"))
from gen_base import partitions, children, dim, boffs, table
import math

def level(k):
    tot_elems = math.factorial(k)
    rot = [0.0] * (k - 1)     # full rotations (both rows in support)
    exp = [0.0] * (k - 1)     # expanding rotations (one row zero)
    neg = [0.0] * (k - 1)
    supp_final = 0.0
    rowterm = 0.0
    for lam in partitions(k):
        dl = dim(lam)
        for b, mu in enumerate(children(lam)):
            bo, dm = boffs(lam)[b], dim(mu)
            ncols = dm
            for i in range(k):
                s = [False] * dl
                for r in range(dm):
                    s[bo + r] = True
                for j in range(k - 2, i - 1, -1):
                    tb = table(lam, j)
                    for t, (q, dist) in enumerate(tb):
                        if q > t:
                            if s[t] and s[q]:
                                rot[j] += ncols
                            elif s[t] or s[q]:
                                exp[j] += ncols
                                s[t] = s[q] = True
                        elif q == t and dist == -1 and s[t]:
                            neg[j] += ncols
                supp_final += ncols * sum(s)
                rowterm += ncols * dl
    print(f"k={k}: per output element: final-support rows/term-rows = {supp_final/rowterm:.3f}")
    tr = te = 0
    for j in range(k - 2, -1, -1):
        print(f"  s_{j}: full rot {rot[j]/tot_elems:6.3f}  expanding {exp[j]/tot_elems:6.3f}  neg {neg[j]/tot_elems:6.3f}   (per output element)")
        tr += rot[j]; te += exp[j]
    print(f"  total rot {tr/tot_elems:.2f} exp {te/tot_elems:.2f}; smem words now ~ {4*(tr+te)/tot_elems:.1f}; slots {(4*tr+2*te)/tot_elems:.1f}")
    print(f"  accumulate rows/elem: full {rowterm/tot_elems:.1f}  support-only {supp_final/tot_elems:.1f}")

for k in [int(a) for a in sys.argv[1:]] or [8, 9, 10]:
    level(k)
