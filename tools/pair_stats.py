"""How much shared-memory traffic would fusing two consecutive sweeps (orbits of <s_a, s_{a-1}>) save?"""
import sys, os, math
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "algebraist_b200", "csrc"))
from gen_base import partitions, children, dim, boffs, table

def level(k):
    tot_elems = math.factorial(k)
    single = 0.0      # words per column with one sweep per pass (rot 4 words, neg 2 words), all chains
    fused = 0.0
    bysize = {1: 0.0, 2: 0.0, 3: 0.0, 6: 0.0}
    for lam in partitions(k):
        dl = dim(lam)
        for b, mu in enumerate(children(lam)):
            bo, dm = boffs(lam)[b], dim(mu)
            sup = [False] * dl
            for r in range(dm): sup[bo + r] = True
            # per sweep j: rotations/negs on support, support growth
            per_j = {}
            sups = {k - 1: list(sup)}
            for j in range(k - 2, -1, -1):
                tb = table(lam, j)
                rot = neg = 0
                for t, (q, dist) in enumerate(tb):
                    if q > t and (sup[t] or sup[q]):
                        rot += 1
                    elif q == t and dist == -1 and sup[t]:
                        neg += 1
                for t, (q, dist) in enumerate(tb):
                    if q != t and (sup[t] or sup[q]): sup[t] = sup[q] = True
                per_j[j] = (rot, neg)
                sups[j] = list(sup)
            # pairs from the top
            pair_cost = {}
            for a in range(k - 2, 0, -2):
                bsw = a - 1
                ta, tb2 = table(lam, a), table(lam, bsw)
                supb = sups[a + 1]       # support before sweep a
                seen = [False] * dl
                cost = 0
                for r in range(dl):
                    if seen[r]: continue
                    orb, stack = [], [r]
                    seen[r] = True
                    while stack:
                        x = stack.pop(); orb.append(x)
                        for tbl in (ta, tb2):
                            y = tbl[x][0]
                            if not seen[y]:
                                seen[y] = True; stack.append(y)
                    if not any(supb[x] for x in orb): continue
                    n = len(orb)
                    if n == 1:
                        s = (ta[orb[0]][1] == -1) != (tb2[orb[0]][1] == -1)
                        c = 2 if s else 0
                    else:
                        c = 2 * n
                    cost += c
                    bysize[n] += c * dm
                pair_cost[a] = cost
            # chains
            for i in range(k):
                # single-sweep formulation
                for j in range(k - 2, i - 1, -1):
                    rot, neg = per_j[j]
                    single += dm * (4 * rot + 2 * neg)
                # fused: pairs (a, a-1) with a-1 >= i, leftover single sweep
                j = k - 2
                while j - 1 >= i:
                    fused += dm * pair_cost[j]
                    j -= 2
                if j >= i:
                    rot, neg = per_j[j]
                    fused += dm * (4 * rot + 2 * neg)
    print(f"k={k}: sweep words/elem single {single/tot_elems:.1f}  fused pairs {fused/tot_elems:.1f}")

for k in [int(a) for a in sys.argv[1:]] or [9, 10]:
    level(k)
