#!/usr/bin/env python
"""Generate base_gen.cuh: straight-line CUDA for the lowest levels of the S_n FFT.

    python algebraist_b200/csrc/gen_base.py        # rewrites base_gen.cuh next to this file

For k <= 6 the irreps are tiny (d <= 16), so a table-driven sweep spends its time
on loop and addressing overhead.  Here every thread owns one instance and runs
fully unrolled code in registers: the Young's-orthogonal-form coefficients
(reference irreps.py:91-109) become FFMA immediates, rows that are still zero
and the +-1 rows cost nothing, and nothing touches shared memory.

Generated device functions (T = float / double):
  s5_forward / s5_inverse   levels 2..5 (5..2) of one S_5 instance, 120 values in registers
(level 6 and above: gen_col.py, gen_tz.py, which import the combinatorics below)

The combinatorics below restate the reference's conventions (partition order
tableau.py:103-123, children irreps.py:64-68, basis order tableau.py:171-198,
axial distance tableau.py:214-220); the generated code is checked against the
CPU restatement and the table-driven kernels by the test-suite.
"""
import math
import os
from functools import lru_cache

SMALL = {
    1: [(1,)],
    2: [(2,), (1, 1)],
    3: [(3,), (2, 1), (1, 1, 1)],
    4: [(4,), (3, 1), (2, 2), (2, 1, 1), (1, 1, 1, 1)],
    5: [(5,), (4, 1), (3, 2), (3, 1, 1), (2, 2, 1), (2, 1, 1, 1), (1, 1, 1, 1, 1)],
}


@lru_cache(maxsize=None)
def partitions(n):
    if n <= 5:
        return tuple(SMALL[n])
    out = [(n,)]
    for k in range(1, n):
        for p in partitions(n - k):
            q = tuple(sorted((k, *p), reverse=True))
            if q not in out:
                out.append(q)
    return tuple(out)


def children(part):
    out = []
    for i in range(len(part)):
        nxt = part[i + 1] if i + 1 < len(part) else 0
        if part[i] > nxt:
            c = list(part)
            c[i] -= 1
            if c[-1] == 0:
                c.pop()
            out.append(tuple(c))
    return sorted(out)


@lru_cache(maxsize=None)
def dim(part):
    n = sum(part)
    if n <= 1:
        return 1
    prod = 1
    for i, row in enumerate(part):
        for j in range(row):
            prod *= row - j + sum(1 for r in part[i + 1:] if r > j)
    return math.factorial(n) // prod


def boffs(part):
    off, out = 0, []
    for c in children(part):
        out.append(off)
        off += dim(c)
    return out


@lru_cache(maxsize=None)
def coef_off(part):
    off = 0
    for p in partitions(sum(part)):
        if p == part:
            return off
        off += dim(p) ** 2
    raise KeyError(part)


def removed_box(big, small):
    for r in range(len(big)):
        s = small[r] if r < len(small) else 0
        if big[r] != s:
            return r, big[r] - 1
    raise ValueError


@lru_cache(maxsize=None)
def table(part, j):
    """Per row t of rho_part(s_j): (partner, signed axial distance)."""
    k = sum(part)
    if j < k - 2:
        out = []
        for mu, o in zip(children(part), boffs(part)):
            out += [(p + o, d) for p, d in table(mu, j)]
        return tuple(out)
    d_l = dim(part)
    if k == 2:
        return ((0, 1),) if part == (2,) else ((0, -1),)
    out = [None] * d_l
    ch, bo = children(part), boffs(part)
    for b, mu in enumerate(ch):
        rb, cb = removed_box(part, mu)
        for c, nu in enumerate(children(mu)):
            ra, ca = removed_box(mu, nu)
            dist = (ra - rb) + (cb - ca)
            lo = bo[b] + boffs(mu)[c]
            if abs(dist) > 1:
                mu2 = list(part)
                mu2[ra] -= 1
                if mu2[-1] == 0:
                    mu2.pop()
                mu2 = tuple(mu2)
                lo2 = bo[ch.index(mu2)] + boffs(mu2)[children(mu2).index(nu)]
            for t in range(dim(nu)):
                out[lo + t] = (lo2 + t if abs(dist) > 1 else lo + t, dist)
    return tuple(out)


# ----------------------------------------------------------------------------
# symbolic emission: a value is None (zero) or (sign, name)
# ----------------------------------------------------------------------------
class Emitter:
    def __init__(self):
        self.lines = []
        self.n = 0

    def tmp(self, expr):
        self.n += 1
        name = f"t{self.n}"
        self.lines.append(f"    const T {name} = {expr};")
        return name

    @staticmethod
    def lit(x):
        return f"T({x!r})"

    def lin(self, terms):
        """terms: list of (coef, value) -> value of sum coef*value (coef python float)."""
        terms = [(c * v[0], v[1]) for c, v in terms if v is not None and c != 0.0]
        if not terms:
            return None
        if len(terms) == 1 and abs(terms[0][0]) == 1.0:
            return (terms[0][0], terms[0][1])
        parts = []
        for c, name in terms:
            if c == 1.0:
                parts.append(f"+ {name}")
            elif c == -1.0:
                parts.append(f"- {name}")
            else:
                parts.append(f"+ {self.lit(c)} * {name}")
        expr = " ".join(parts)
        expr = expr[2:] if expr.startswith("+ ") else "-" + expr[2:]
        return (1.0, self.tmp(expr))

    def sweep(self, part, j, x):
        tb = table(part, j)
        for t, (q, dist) in enumerate(tb):
            if q > t:
                a = 1.0 / dist
                b = math.sqrt(1.0 - 1.0 / (dist * dist))
                x1, x2 = x[t], x[q]
                if x1 is None and x2 is None:
                    continue
                x[t] = self.lin([(a, x1), (b, x2)])
                x[q] = self.lin([(b, x1), (-a, x2)])
            elif q == t and dist == -1 and x[t] is not None:
                x[t] = (-x[t][0], x[t][1])

    def total(self, terms):
        return self.lin([(1.0, v) for v in terms])


class ScaledEmitter(Emitter):
    """Values are (scale, name) with an arbitrary compile-time scale instead of a sign: a linear combination
    c0 s0 N0 + c1 s1 N1 is emitted as ONE fused multiply-add  N = (c1 s1 / (c0 s0)) * N1 + N0  carrying the scale
    c0 s0 ("fast Givens": a 2 x 2 rotation costs 2 FFMA instead of 2 FMUL + 2 FFMA), a single term costs nothing,
    and the scales are multiplied back in where a value is stored.  Scaling by constants leaves the relative
    rounding errors unchanged; the scales are products of at most k - 1 factors 1/d, sqrt(1 - 1/d^2), |d| <= k."""

    def lin(self, terms):
        t = [(c * v[0], v[1]) for c, v in terms if v is not None and c != 0.0]
        if not t:
            return None
        s0, expr = t[0]
        if len(t) == 1:
            return (s0, expr)
        for c, name in t[1:]:
            r = c / s0
            if r == 1.0:
                expr = f"{expr} + {name}"
            elif r == -1.0:
                expr = f"{expr} - {name}"
            else:
                expr = f"sn_fma({self.lit(r)}, {name}, {expr})"     # vec2.cuh: FFMA / DFMA, FFMA2 for pairs of instances
        return (s0, self.tmp(expr))


def as_value(v):
    """A loader may return a plain name or an already scaled value."""
    return v if isinstance(v, tuple) else (1.0, v)


def forward_column(em, k, lam, b, c, load, store):
    """F_lam[:, boff + c] = sum_i rho(s_i)..rho(s_{k-2}) embed_b(F^(i)_mu[:, c])  (fourier.py:253-273)."""
    mu = children(lam)[b]
    bo, dm, dl = boffs(lam)[b], dim(mu), dim(lam)
    acc = [None] * dl                       # running sums keep few values live
    for i in range(k):
        x = [None] * dl
        for r in range(dm):
            x[bo + r] = as_value(load(mu, r, c, i))
        for j in range(k - 2, i - 1, -1):
            em.sweep(lam, j, x)
        for t in range(dl):
            if x[t] is not None:
                acc[t] = em.lin([(1.0, acc[t]), (1.0, x[t])])
    for t in range(dl):
        store(lam, t, bo + c, acc[t])


def inverse_column(em, k, mu, c, i, load, store):
    """G^(i)_mu[:, c] = sum_lam d_lam/(k d_mu) [rho(s_{k-2})..rho(s_i) F_lam[:, boff + c]]_mu  (fourier.py:191-199)."""
    dm = dim(mu)
    acc = [None] * dm
    for lam in partitions(k):
        ch = children(lam)
        if mu not in ch:
            continue
        b = ch.index(mu)
        bo, dl = boffs(lam)[b], dim(lam)
        # rows needed before each sweep: walk the sweeps backwards from the block rows
        need = [set(range(bo, bo + dm))]
        for j in range(k - 2, i - 1, -1):
            prev = set(need[-1])
            for t in need[-1]:
                prev.add(table(lam, j)[t][0])
            need.append(prev)
        need.reverse()                      # need[0] = rows to load, need[m] = rows needed after m sweeps
        x = [None] * dl
        for t in sorted(need[0]):
            x[t] = as_value(load(lam, t, bo + c))
        for m, j in enumerate(range(i, k - 1)):
            em.sweep(lam, j, x)
            for t in range(dl):
                if t not in need[m + 1]:
                    x[t] = None
        scale = dl / (k * dm)
        for r in range(dm):
            if x[bo + r] is not None:
                acc[r] = em.lin([(1.0, acc[r]), (scale, x[bo + r])])
    for r in range(dm):
        store(mu, r, c, acc[r])


def value_expr(v):
    if v is None:
        return "T(0)"
    if v[0] == 1.0:
        return v[1]
    return f"-{v[1]}" if v[0] == -1.0 else f"{Emitter.lit(v[0])} * {v[1]}"


def gen_s5(forward):
    """Levels 2..5 of one S_5 instance on 120 register values.
    x[p], p = chain index (p_1 = sum_k i_k 120/k!), y[e], e = packed S_5 coefficient index."""
    K0 = 5
    em = ScaledEmitter()
    if forward:
        cur = {((1,), 0, 0, p): f"x[{p}]" for p in range(120)}
        for k in range(2, K0 + 1):
            pk = 120 // math.factorial(k)
            nxt = {}
            for p in range(pk):
                for lam in partitions(k):
                    for b in range(len(children(lam))):
                        for c in range(dim(children(lam)[b])):
                            def load(mu, r, cc, i, p=p, pk=pk, cur=cur):
                                return cur[(mu, r, cc, i * pk + p)]

                            def store(l, t, col, v, p=p, nxt=nxt):
                                nxt[(l, t, col, p)] = v

                            forward_column(em, k, lam, b, c, load, store)
            # materialise signs so that the next level sees plain names
            cur = {}
            for key, v in nxt.items():          # the scales travel to the next level with the values
                cur[key] = (1.0, em.tmp("T(0)")) if v is None else v
        for lam in partitions(K0):
            d = dim(lam)
            for t in range(d):
                for col in range(d):
                    em.lines.append(f"    y[{coef_off(lam) + t * d + col}] = {value_expr(cur[(lam, t, col, 0)])};")
    else:
        cur = {}
        for lam in partitions(K0):
            d = dim(lam)
            for t in range(d):
                for col in range(d):
                    cur[(lam, t, col, 0)] = f"x[{coef_off(lam) + t * d + col}]"
        for k in range(K0, 1, -1):
            pk = 120 // math.factorial(k)
            nxt = {}
            for p in range(pk):
                for mu in partitions(k - 1):
                    for c in range(dim(mu)):
                        for i in range(k):
                            def load(l, t, col, p=p, cur=cur):
                                return cur[(l, t, col, p)]

                            def store(m, r, cc, v, p=p, pk=pk, i=i, nxt=nxt):
                                nxt[(m, r, cc, i * pk + p)] = v

                            inverse_column(em, k, mu, c, i, load, store)
            cur = {}
            for key, v in nxt.items():          # the scales travel to the next level with the values
                cur[key] = (1.0, em.tmp("T(0)")) if v is None else v
        for p in range(120):
            em.lines.append(f"    y[{p}] = {value_expr(cur[((1,), 0, 0, p)])};")
    name = "s5_forward" if forward else "s5_inverse"
    head = [f"template <typename T>",
            f"__device__ __forceinline__ void {name}(const T (&x)[120], T (&y)[120]) {{"]
    return "\n".join(head + em.lines + ["}"]) + "\n", em.n


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    parts = ["// GENERATED by gen_base.py -- do not edit.  Straight-line Young's-orthogonal-form code for the",
             "// lowest levels of the S_n FFT (see gen_base.py for the derivation and reference citations).",
             "#pragma once", "#include <stddef.h>", "#include \"vec2.cuh\"", "",
             "// keeps the loads of one column from being hoisted above the stores of the previous one",
             "// (otherwise the register allocator sees every column of the level at once and spills)",
             "#define SNFFT_CODE_FENCE() asm volatile(\"\" ::: \"memory\")", "", "namespace snfft {", ""]
    stats = {}
    for fn, fwd in ((gen_s5, True), (gen_s5, False)):
        code, n = fn(fwd)
        stats[(fn.__name__, fwd)] = n
        parts.append(code)
    parts.append("}  // namespace snfft")
    with open(os.path.join(here, "base_gen.cuh"), "w") as fh:
        fh.write("\n".join(parts) + "\n")
    print(stats)


if __name__ == "__main__":
    main()
