"""Numerical model of the "top / bottom" (T / Z) factorisation of one level step, checked against the plain
sum over cosets, plus its operation and traffic counts.

    python tools/tz_model.py 9 7        # level k = 9, bottom level MB = 7

Forward step k (reference fourier.py:253-273), column c of block mu of F_lambda:
    F = sum_i rho(s_i) .. rho(s_{k-2}) embed_mu(x_i)
Row = path lambda -> mu_{k-1} -> .. -> mu_1; rho(s_j) changes mu_{j+1} only.  With MB <= k - 1, s = k - MB:
  * bottom sweeps s_i .. s_{MB-2} act inside the blocks alpha = mu_MB (partitions of MB),
  * top sweeps s_{MB-1} .. s_{k-2} act on (mu_{k-1} .. mu_MB) and leave (zeta = mu_{MB-1}, tableau) alone.
Hence  F[(P', alpha), r] = sum_{zeta in alpha, P: mu -> zeta} C[(P', alpha); (mu, P) | zeta] Z[(P, zeta) -> alpha][r]  + cosets i >= MB,
       Z[(P, zeta) -> alpha] = sum_{i < MB} rho_alpha(s_i .. s_{MB-2}) embed_zeta(x_i[(P, zeta), :])
i.e. a level-MB forward column operator on pieces of the inputs (independent of lambda and P'), then the top
sweeps once, after the sum over the cosets.
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "algebraist_b200", "csrc"))
from gen_base import boffs, children, dim, partitions, table  # noqa: E402


def parents(part):
    return [lam for lam in partitions(sum(part) + 1) if part in children(lam)]


def subblocks(part, depth):
    """[(path (part, c1, .., c_depth), row offset)] in row order."""
    out = [((part,), 0)]
    for _ in range(depth):
        nxt = []
        for path, off in out:
            for c, o in zip(children(path[-1]), boffs(path[-1])):
                nxt.append((path + (c,), off + o))
        out = nxt
    return out


def sweep(part, j, x):
    y = x.copy()
    for t, (q, dist) in enumerate(table(part, j)):
        if q > t:
            a, b = 1.0 / dist, math.sqrt(1.0 - 1.0 / (dist * dist))
            y[t] = a * x[t] + b * x[q]
            y[q] = b * x[t] - a * x[q]
        elif q == t and dist == -1:
            y[t] = -x[t]
    return y


def direct(k, lam, b, xs):
    """sum_i rho(s_i)..rho(s_{k-2}) embed_b(xs[i])."""
    mu = children(lam)[b]
    bo, dm, dl = boffs(lam)[b], dim(mu), dim(lam)
    out = np.zeros(dl)
    for i in range(k):
        v = np.zeros(dl)
        v[bo:bo + dm] = xs[i]
        for j in range(k - 2, i - 1, -1):
            v = sweep(lam, j, v)
        out += v
    return out


def tz(k, MB, lam, b, xs):
    s = k - MB
    mu = children(lam)[b]
    bo, dm, dl = boffs(lam)[b], dim(mu), dim(lam)
    # Z stage
    Z = {}
    for P, off in subblocks(mu, s):
        zeta = P[-1]
        dz = dim(zeta)
        for alpha in parents(zeta):
            bz = boffs(alpha)[children(alpha).index(zeta)]
            acc = np.zeros(dim(alpha))
            for i in range(MB):
                v = np.zeros(dim(alpha))
                v[bz:bz + dz] = xs[i][off:off + dz]
                for j in range(MB - 2, i - 1, -1):
                    v = sweep(alpha, j, v)
                acc += v
            Z[(P, alpha)] = acc
    # T stage
    out = np.zeros(dl)
    nterm = 0
    for Pp, offp in subblocks(lam, s):
        alpha = Pp[-1]
        da = dim(alpha)
        for P, off in subblocks(mu, s):
            zeta = P[-1]
            if zeta not in children(alpha):
                continue
            e = np.zeros(dl)
            e[bo + off] = 1.0
            for j in range(k - 2, MB - 2, -1):
                e = sweep(lam, j, e)
            cf = e[offp + boffs(alpha)[children(alpha).index(zeta)]]
            if cf != 0.0:
                out[offp:offp + da] += cf * Z[(P, alpha)]
                nterm += da
        # cosets i >= MB: sweeps s_{k-2} .. s_i keep alpha and the rows below
        for i in range(MB, k):
            for Q, offq in subblocks(mu, s - 1):
                if Q[-1] != alpha:
                    continue
                e = np.zeros(dl)
                e[bo + offq] = 1.0
                for j in range(k - 2, i - 1, -1):
                    e = sweep(lam, j, e)
                cf = e[offp]
                if cf != 0.0:
                    out[offp:offp + da] += cf * xs[i][offq:offq + da]
                    nterm += da
    return out, nterm


def check(k, MB):
    rng = np.random.default_rng(0)
    worst = 0.0
    terms = 0
    for lam in partitions(k):
        for b, mu in enumerate(children(lam)):
            xs = [rng.standard_normal(dim(mu)) for _ in range(k)]
            ref = direct(k, lam, b, xs)
            got, nt = tz(k, MB, lam, b, xs)
            terms += nt * dim(mu)
            worst = max(worst, np.abs(ref - got).max())
    print(f"k={k} MB={MB}: max abs diff {worst:.2e}; T terms per output element {terms / math.factorial(k):.2f}")


if __name__ == "__main__":
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    MB = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    check(k, MB)
