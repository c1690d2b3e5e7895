"""Statistics for the T_s (fused top sweeps, gather form) + B_mb (register bottom) level design."""
import sys, os, math
from functools import lru_cache
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "algebraist_b200", "csrc"))
from gen_base import partitions, children, dim, boffs

@lru_cache(maxsize=None)
def paths(top, depth):
    """[(row offset in top, bottom shape)] for all chains of `depth` removals, in row order."""
    if depth == 0:
        return ((0, top),)
    out = []
    for c, o in zip(children(top), boffs(top)):
        for (o2, z) in paths(c, depth - 1):
            out.append((o + o2, z))
    return tuple(out)

def stats(k, s, mb):
    print(f"--- k={k} s={s} mb={mb}")
    max_nsb = 0; max_R = 0
    rows = []
    for lam in partitions(k):
        dl = dim(lam)
        sb = paths(lam, k - mb)            # sub-blocks: paths lam -> alpha |- mb
        nsb = len(sb)
        types = len(set(z for _, z in sb))
        for b, mu in enumerate(children(lam)):
            inp = paths(mu, s)              # paths b -> zeta
            zetas = sorted(set(z for _, z in inp))
            R = sum(dim(z) for z in zetas)
            outp = paths(lam, s + 1)
            nin = max(sum(1 for _, z in inp if z == ze) for ze in zetas)
            nout = max(sum(1 for _, z in outp if z == ze) for ze in zetas)
            rows.append((dl, dim(mu), nsb, types, R, len(zetas), nin, nout))
    rows.sort(reverse=True)
    print(" d_lam d_mu nsb types R(T items/slotvec) norbits maxnin maxnout")
    seen = set()
    for r in rows:
        if r[0] in seen: continue
        seen.add(r[0])
        if len(seen) <= 12: print("  ", r)
    print("  max nsb", max(r[2] for r in rows), "max R", max(r[4] for r in rows), "max nin", max(r[6] for r in rows), "max nout", max(r[7] for r in rows))

for a in sys.argv[1:]:
    k, s, mb = map(int, a.split(","))
    stats(k, s, mb)
