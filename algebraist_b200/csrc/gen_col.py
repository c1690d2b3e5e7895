#!/usr/bin/env python
"""Generate col_gen_<k>.cuh: register "column" code for the level steps k = 6, 7 and 8.

    python algebraist_b200/csrc/gen_col.py        # rewrites col_gen_<k>.cuh next to this file

A level step (reference fourier.py:253-273, inverse formula fourier.py:191-199) treats every column of
a block independently, and for k <= 8 the pieces of a column that the bottom sweeps keep together are
small (d <= 35), so one thread can own them in registers and the Young's-orthogonal-form coefficients
(irreps.py:91-109) become FFMA immediates -- no shared memory, no barriers, no tables:

  forward task (lambda, child block b, row set R)  ->  rows R of  F_lambda[:, boff_b + c]
       = sum_i rho(s_i) .. rho(s_{k-2}) embed_b(F^(i)_mu[:, c])           restricted to R
  inverse task (mu, row set R)                     ->  rows R of  G^(i)_mu[:, c],  i = 0 .. k-1
       = sum_lambda d_lambda/(k d_mu) [rho(s_{k-2}) .. rho(s_i) F_lambda[:, boff + c]]_{block mu}

with c a run-time column and R all rows (k <= 7) or one child block of the output (k = 8: the bottom sweeps
s_0..s_{k-3} never leave such a block, only the top sweep s_{k-2} couples two of them).  Rows that are still
zero (forward) or no longer needed (inverse) are dropped sweep by sweep, so the code has exactly the
support-restricted operation count of the table-driven kernels.

A thread is (task, column c, instance); the tasks that read the same child block mu form a group and run
side by side so that their re-reads hit in L1 / L2.

Generated per (K, direction):  ColInfo<K, INV> (groups for the host), col_run<T, K, INV>(task, c, in, out, s).
Combinatorics come from gen_base.py (same conventions, same citations).
"""
import math
import os

from gen_base import Emitter, ScaledEmitter, boffs, children, coef_off, dim, partitions, table

LEVELS = (6, 7, 8)
CHILD_MAJOR_MAX = 6       # levels whose tasks keep a whole child column in registers
# adjacent cosets per inverse task by level (the Z kernels of gen_tz.py use the entry of their bottom level MB):
# measured on a B200 (profiles r02l/r02m), bounded by the registers the running sums need (no spills at level 7 up
# to 4 cosets = 64 sums; level 8 spills beyond 2)
PREFETCH = os.environ.get("SNFFT_GEN_PREFETCH", "1") != "0"      # issue a chain's / block's loads one step ahead of the arithmetic
INV_COSETS_PER_TASK = {7: int(os.environ.get("SNFFT_GEN_INV_COSETS_7", "4")), 8: int(os.environ.get("SNFFT_GEN_INV_COSETS_8", "2"))}
# ... and at most this many running sums per inverse task (cosets x rows of mu): the two d = 35 shapes of level 7 with two
# cosets (70 sums) spilled 0.5-1.3 KB per thread in the six tasks that do most of level 8's work (profiles r02u: 2.2 GB
# written for 1.86 GB of output); they run one coset per task
INV_SUMS_CAP = {7: int(os.environ.get("SNFFT_GEN_INV_SUMS_7", "48")), 8: int(os.environ.get("SNFFT_GEN_INV_SUMS_8", "48"))}


def inv_cosets(k, d):
    """Adjacent cosets per inverse task of level k for a child shape of dimension d."""
    return max(1, min(INV_COSETS_PER_TASK[k], INV_SUMS_CAP[k] // d))


SPLIT_FWD = {}            # forward-only levels generated in parts (level 9, until the top / bottom kernels of gen_tz.py took it over)
# forward tasks: rows of output a task owns at most (level 8: adjacent child blocks are merged up to this many rows,
# so that small blocks share their loads; level 7: several parents of one child shape per task, as many as fit this
# many running sums).  Level 7 / Z kernels with F2 threads, measured (S_10, B = 128): 35 / 42 / 48 / 56 / 70 sums ->
# level 7 forward 0.69 / 0.67 / 0.67 / 0.68 / 0.70 ms, Z forward of level 9 0.56 / 0.56 / 0.55 / 0.58 / 0.61 ms; inverse
# sums cap 48 against 64 at level 7: 0.65 / 0.66 ms; 35 against 48 at level 8: 1.11 / 1.04 ms
FWD_ROWS_CAP = {7: int(os.environ.get("SNFFT_GEN_FWD_ROWS_7", "48")), 8: int(os.environ.get("SNFFT_GEN_FWD_ROWS_8", "35"))}


def partner_checked(part, j, t):
    q, dist = table(part, j)[t]
    if q != t:
        q2, dist2 = table(part, j)[q]
        assert q2 == t and dist2 == -dist
    return q, dist


def sweep_need(em, part, j, x, need):
    """x (dict row -> value) after rho_part(s_j), only on the rows in `need`."""
    new = {}
    for t in sorted(need):
        q, dist = partner_checked(part, j, t)
        if q == t:
            v = x.get(t)
            if v is not None:
                new[t] = v if dist == 1 else (-v[0], v[1])
            continue
        a = 1.0 / dist
        b = math.sqrt(1.0 - 1.0 / (dist * dist))
        v = em.lin([(a, x.get(t)), (b, x.get(q))])
        if v is not None:
            new[t] = v
    return new


def need_chain(part, seq, rows):
    """need[m] = rows that must be known after the first m sweeps of `seq` to get `rows` at the end."""
    need = [None] * (len(seq) + 1)
    need[-1] = set(rows)
    for m in range(len(seq) - 1, -1, -1):
        need[m] = set(need[m + 1])
        for t in need[m + 1]:
            need[m].add(table(part, seq[m])[t][0])
    return need


def forward_task(em, k, lam, b, rows, load, store):
    mu = children(lam)[b]
    bo, dm = boffs(lam)[b], dim(mu)
    acc = {t: None for t in rows}
    needs = [need_chain(lam, list(range(k - 2, i - 1, -1)), rows) for i in range(k)]

    def fetch(i):                       # the loads of coset i, issued one chain ahead (PREFETCH): with ~20 warps per SM the
        return {bo + r: (1.0, load(r, i)) for r in range(dm) if bo + r in needs[i][0]}     # latency of a chain's loads is not hidden otherwise

    nxt = fetch(0)
    for i in range(k):
        seq = list(range(k - 2, i - 1, -1))
        need = needs[i]
        x = nxt
        nxt = fetch(i + 1) if PREFETCH and i + 1 < k else None
        for m, j in enumerate(seq):
            x = sweep_need(em, lam, j, x, need[m + 1])
        for t in rows:
            if x.get(t) is not None:
                acc[t] = em.lin([(1.0, acc[t]), (1.0, x[t])])
        em.lines.append("    SNFFT_CODE_FENCE();")     # keeps the loads of the chains after the next one below this one (register pressure)
        if nxt is None and i + 1 < k:
            nxt = fetch(i + 1)
    for t in rows:
        store(t, acc[t])


def forward_task_multi(em, k, targets, load, store):
    """forward_task for several parents of ONE child shape at once.  targets = [(lam, b, rows)], all with the same
    children(lam)[b]; store(ti, t, value).  The chains of the parents share nothing but their inputs: a coset's piece
    of the child column is loaded once and every parent runs its own sweeps on it (the sibling tasks of forward_task
    re-read it from L2, one load per parent)."""
    mu = children(targets[0][0])[targets[0][1]]
    dm = dim(mu)
    acc = [{t: None for t in rows} for _, _, rows in targets]
    needs = [[need_chain(lam, list(range(k - 2, i - 1, -1)), rows) for lam, b, rows in targets] for i in range(k)]

    def fetch(i):                       # rows of the child column any parent needs from coset i, loaded once
        rows_in = sorted({r for ti, (lam, b, rows) in enumerate(targets) for r in range(dm) if boffs(lam)[b] + r in needs[i][ti][0]})
        return {r: load(r, i) for r in rows_in}

    nxt = fetch(0)
    for i in range(k):
        seq = list(range(k - 2, i - 1, -1))
        cache = nxt
        nxt = fetch(i + 1) if PREFETCH and i + 1 < k else None     # one coset ahead of the arithmetic
        for ti, (lam, b, rows) in enumerate(targets):
            assert children(lam)[b] == mu
            bo = boffs(lam)[b]
            need = needs[i][ti]
            x = {bo + r: (1.0, cache[r]) for r in range(dm) if bo + r in need[0]}
            for m, j in enumerate(seq):
                x = sweep_need(em, lam, j, x, need[m + 1])
            for t in rows:
                if x.get(t) is not None:
                    acc[ti][t] = em.lin([(1.0, acc[ti][t]), (1.0, x[t])])
        em.lines.append("    SNFFT_CODE_FENCE();")     # keeps the loads of later chains below this one (register pressure)
        if nxt is None and i + 1 < k:
            nxt = fetch(i + 1)
    for ti, (lam, b, rows) in enumerate(targets):
        for t in rows:
            store(ti, t, acc[ti][t])


def parent_groups(mu, cap):
    """The parents of `mu` in groups whose dimensions add up to at most `cap` running sums (largest with smallest)."""
    ps = sorted([lam for lam in partitions(sum(mu) + 1) if mu in children(lam)], key=dim, reverse=True)
    groups = []
    while ps:
        g = [ps.pop(0)]
        while ps and sum(dim(x) for x in g) + dim(ps[-1]) <= cap:
            g.append(ps.pop())
        groups.append(sorted(g, key=lambda lam: partitions(sum(lam)).index(lam)))
    return groups


def block_of(lam, t):
    """Index of the child block of `lam` that row t lies in."""
    for b, (c, o) in enumerate(zip(children(lam), boffs(lam))):
        if o <= t < o + dim(c):
            return b
    raise ValueError(t)


def inverse_task(em, k, mu, i, rows, load, store, unit_scale=False):
    """Rows `rows` of G^(i)_mu[:, c].  The bottom sweeps s_i..s_{k-3} never leave a child block of lambda, so the
    column of F_lambda is processed one child block at a time (loads, bottom sweeps, its term of the top sweep
    s_{k-2} added to the running sums): only the sums and one block are live at any point."""
    dm = dim(mu)
    acc = {r: None for r in rows}
    for lam in partitions(k):
        ch = children(lam)
        if mu not in ch:
            continue
        b = ch.index(mu)
        bo, dl = boffs(lam)[b], dim(lam)
        scale = 1.0 if unit_scale else dl / (k * dm)      # unit_scale: gen_tz.py applies the weights in its T kernel
        want = [bo + r for r in rows]
        if i > k - 2:                       # coset k-1: identity
            for r in rows:
                acc[r] = em.lin([(1.0, acc[r]), (scale, (1.0, load(lam, bo + r, bo)))])
            em.lines.append("    SNFFT_CODE_FENCE();")
            continue
        # top sweep s_{k-2} (applied last): out[t] = a_t y[t] + b_t y[q_t], y = rows after the bottom sweeps
        top = {}
        for t in want:
            q, dist = partner_checked(lam, k - 2, t)
            if q == t:
                top[t] = [(float(dist), t)]
            else:
                top[t] = [(1.0 / dist, t), (math.sqrt(1.0 - 1.0 / (dist * dist)), q)]
        by_block = {}
        for t in want:
            for coef, src in top[t]:
                by_block.setdefault(block_of(lam, src), []).append((t, coef, src))
        seq = list(range(i, k - 2))         # bottom sweeps, ascending
        for blk in sorted(by_block, key=lambda x: (x != b, x)):       # the block of mu first
            terms = by_block[blk]
            need = need_chain(lam, seq, sorted({src for _, _, src in terms}))
            x = {t: (1.0, load(lam, t, bo)) for t in sorted(need[0])}
            for m, j in enumerate(seq):
                x = sweep_need(em, lam, j, x, need[m + 1])
            for t, coef, src in terms:
                if x.get(src) is not None:
                    acc[t - bo] = em.lin([(1.0, acc[t - bo]), (scale * coef, x[src])])
            em.lines.append("    SNFFT_CODE_FENCE();")     # one block's loads at a time
    for r in rows:
        store(r, acc[r])


def inverse_task_multi(em, k, mu, cosets, rows, load, store, unit_scale=False):
    """inverse_task for several cosets at once: rows `rows` of G^(i)_mu[:, c] for every i in `cosets`.  The chains of
    the cosets share nothing but their INPUTS -- and those are what the inverse kernels are short of: one task per
    coset re-reads every parent column k times (58-65 % L2 hit rate, profiles r02h).  A block of a parent column is
    loaded once (the union of the rows the cosets need), then every coset runs its own bottom sweeps on it and adds
    its term of the top sweep to its own running sums.  The loads of the next block are issued before the arithmetic
    of the current one (PREFETCH).  store(i, r, value)."""
    dm = dim(mu)
    acc = {(i, r): None for i in cosets for r in rows}
    work = []                               # (lam, bo, scale, terms, plans, rows to load) per block of a parent column
    for lam in partitions(k):
        ch = children(lam)
        if mu not in ch:
            continue
        b = ch.index(mu)
        bo, dl = boffs(lam)[b], dim(lam)
        scale = 1.0 if unit_scale else dl / (k * dm)
        want = [bo + r for r in rows]
        top = {}
        for t in want:
            q, dist = partner_checked(lam, k - 2, t)
            if q == t:
                top[t] = [(float(dist), t)]
            else:
                top[t] = [(1.0 / dist, t), (math.sqrt(1.0 - 1.0 / (dist * dist)), q)]
        by_block = {}
        for t in want:
            for coef, src in top[t]:
                by_block.setdefault(block_of(lam, src), []).append((t, coef, src))
        for blk in sorted(by_block, key=lambda x: (x != b, x)):       # the block of mu first
            terms = by_block[blk]
            plans = []                      # per coset: (i, bottom sweeps, need chain) for this block
            for i in cosets:
                if i > k - 2:               # coset k-1: identity, only the rows of the block of mu itself
                    if blk == b:
                        plans.append((i, None, None))
                    continue
                seq = list(range(i, k - 2))
                plans.append((i, seq, need_chain(lam, seq, sorted({src for _, _, src in terms}))))
            rows_needed = set()
            for i, seq, need in plans:
                rows_needed |= set(want) if seq is None else need[0]
            if plans:
                work.append((lam, bo, scale, terms, plans, sorted(rows_needed)))

    def fetch(w):
        lam, bo, _, _, _, rows_needed = work[w]
        return {t: (1.0, load(lam, t, bo)) for t in rows_needed}

    nxt = fetch(0) if work else None
    for w, (lam, bo, scale, terms, plans, _) in enumerate(work):
        loaded = nxt
        nxt = fetch(w + 1) if PREFETCH and w + 1 < len(work) else None
        for i, seq, need in plans:
            if seq is None:
                for r in rows:
                    acc[(i, r)] = em.lin([(1.0, acc[(i, r)]), (scale, loaded[bo + r])])
                continue
            x = {t: loaded[t] for t in need[0]}
            for m, j in enumerate(seq):
                x = sweep_need(em, lam, j, x, need[m + 1])
            for t, coef, src in terms:
                if x.get(src) is not None:
                    acc[(i, t - bo)] = em.lin([(1.0, acc[(i, t - bo)]), (scale * coef, x[src])])
        em.lines.append("    SNFFT_CODE_FENCE();")     # at most two blocks' loads live at a time
        if nxt is None and w + 1 < len(work):
            nxt = fetch(w + 1)
    for i in cosets:
        for r in rows:
            store(i, r, acc[(i, r)])


def coset_groups(k, size):
    """Cosets 0..k-1 in runs of `size` adjacent ones: adjacent cosets need nearly the same rows (the rows coset i + 1
    needs are a subset of those coset i needs), so a run loads little more than its first member."""
    return [list(range(a, min(a + size, k))) for a in range(0, k, size)]


def value_expr(v):
    if v is None:
        return "T(0)"
    if v[0] == 1.0:
        return v[1]
    return f"-{v[1]}" if v[0] == -1.0 else f"{Emitter.lit(v[0])} * {v[1]}"


def row_sets(k, part):
    """Row sets a task of level k owns inside a column of `part`: everything (k = 7) or one child block."""
    d = dim(part)
    if k <= 7:
        return [list(range(d))]
    if k == 8:
        out, cap = [], FWD_ROWS_CAP[8]
        for c, o in zip(children(part), boffs(part)):       # adjacent child blocks, merged while they fit the cap
            blk = list(range(o, o + dim(c)))
            if out and len(out[-1]) + len(blk) <= cap:
                out[-1] += blk
            else:
                out.append(blk)
        return out
    out = []                    # k = 9: the blocks two levels down (partitions of 7, <= 35 rows)
    for c, o in zip(children(part), boffs(part)):
        out += [list(range(o + o2, o + o2 + dim(c2))) for c2, o2 in zip(children(c), boffs(c))]
    return out


def gen_level(k, inverse, only=None):
    """Returns (code, groups, number of tasks, statements); `only`: the child shapes mu (groups) to generate."""
    tag = f"col{k}{'i' if inverse else 'f'}"
    funcs, cases, groups = [], [], []
    total = 0
    task = 0
    for mu in partitions(k - 1):
        if only is not None and mu not in only:
            continue
        dm = dim(mu)
        task0 = task
        if not inverse and k <= CHILD_MAJOR_MAX:
            # child-major: one task per mu; the k * d_mu inputs of the column are loaded once and stay in registers
            # for every parent lambda (k * d_mu <= 36 values at k = 6)
            em = ScaledEmitter()
            cache = {}

            def load(r, i, em=em, cache=cache, dm=dm):
                if (r, i) not in cache:
                    cache[(r, i)] = em.tmp(f"in[(size_t){r * dm * k + i} * s]")
                return cache[(r, i)]

            for r in range(dm):                 # every load first (memory-level parallelism)
                for i in range(k):
                    load(r, i)
            for lam in partitions(k):
                if mu not in children(lam):
                    continue
                b = children(lam).index(mu)
                bo, dl = boffs(lam)[b], dim(lam)

                def store(t, v, em=em, lam=lam, dl=dl, bo=bo):
                    em.lines.append(f"    out[(size_t){coef_off(lam) + t * dl + bo} * s] = {value_expr(v)};")

                forward_task(em, k, lam, b, list(range(dl)), load, store)
            total += em.n
            funcs.append("\n".join([
                "template <typename T>",
                f"__device__ __noinline__ void {tag}_{task}(const T* __restrict__ in, T* __restrict__ out, unsigned s) {{"
                f"   // all parents of {mu}"] + em.lines + ["}"]) + "\n")
            cases.append(f"        case {task}: {tag}_{task}<T>(in + (size_t)({coef_off(mu)} + c) * {k} * s, out + (size_t)c * s, s); break;")
            task += 1
        elif inverse and k <= CHILD_MAJOR_MAX:
            # one task per mu: the columns of the parents (sum of d_lambda = k * d_mu values) are loaded once and
            # shared by the k cosets
            em = ScaledEmitter()
            cache = {}

            def load(lam, t, bo, em=em, cache=cache):
                if (lam, t) not in cache:
                    cache[(lam, t)] = em.tmp(f"in[(size_t){coef_off(lam) + t * dim(lam) + bo} * s]")
                return cache[(lam, t)]

            for lam in partitions(k):
                if mu in children(lam):
                    bo = boffs(lam)[children(lam).index(mu)]
                    for t in range(dim(lam)):
                        load(lam, t, bo)
            for i in range(k):
                def store(r, v, em=em, i=i, dm=dm):
                    em.lines.append(f"    out[(size_t){r * dm * k + i} * s] = {value_expr(v)};")

                inverse_task(em, k, mu, i, list(range(dm)), load, store)
            total += em.n
            funcs.append("\n".join([
                "template <typename T>",
                f"__device__ __noinline__ void {tag}_{task}(const T* __restrict__ in, T* __restrict__ out, unsigned s) {{"
                f"   // {mu}, every coset"] + em.lines + ["}"]) + "\n")
            cases.append(f"        case {task}: {tag}_{task}<T>(in + (size_t)c * s, out + (size_t)({coef_off(mu)} + c) * {k} * s, s); break;")
            task += 1
        elif not inverse and k == 7:
            # one task per group of parents: every coset's piece of the child column is loaded once per group
            for grp in parent_groups(mu, FWD_ROWS_CAP[7]):
                em = ScaledEmitter()
                targets = [(lam, children(lam).index(mu), list(range(dim(lam)))) for lam in grp]

                def load(r, i, em=em, dm=dm):
                    return em.tmp(f"in[(size_t){r * dm * k + i} * s]")

                def store(ti, t, v, em=em, targets=targets):
                    lam, b, _ = targets[ti]
                    em.lines.append(f"    out[(size_t){coef_off(lam) + t * dim(lam) + boffs(lam)[b]} * s] = {value_expr(v)};")

                forward_task_multi(em, k, targets, load, store)
                total += em.n
                funcs.append("\n".join([
                    "template <typename T>",
                    f"__device__ __noinline__ void {tag}_{task}(const T* __restrict__ in, T* __restrict__ out, unsigned s) {{"
                    f"   // {', '.join(str(l) for l in grp)} <- {mu}"] + em.lines + ["}"]) + "\n")
                cases.append(f"        case {task}: {tag}_{task}<T>(in + (size_t)({coef_off(mu)} + c) * {k} * s, out + (size_t)c * s, s); break;")
                task += 1
        elif not inverse:
            for lam in partitions(k):
                if mu not in children(lam):
                    continue
                b = children(lam).index(mu)
                bo, dl = boffs(lam)[b], dim(lam)
                for rows in row_sets(k, lam):
                    em = ScaledEmitter()
                    cache = {}

                    def load(r, i, em=em, cache=cache, dm=dm):
                        if (r, i) not in cache:
                            cache[(r, i)] = em.tmp(f"in[(size_t){r * dm * k + i} * s]")
                        return cache[(r, i)]

                    def store(t, v, em=em, lam=lam, dl=dl, bo=bo):
                        em.lines.append(f"    out[(size_t){coef_off(lam) + t * dl + bo} * s] = {value_expr(v)};")

                    forward_task(em, k, lam, b, rows, load, store)
                    total += em.n
                    funcs.append("\n".join([
                        "template <typename T>",
                        f"__device__ __noinline__ void {tag}_{task}(const T* __restrict__ in, T* __restrict__ out, unsigned s) {{"
                        f"   // {lam} <- {mu}, rows {rows[0]}..{rows[-1]}"] + em.lines + ["}"]) + "\n")
                    cases.append(f"        case {task}: {tag}_{task}<T>(in + (size_t)({coef_off(mu)} + c) * {k} * s, out + (size_t)c * s, s); break;")
                    task += 1
        else:
            # one task per (row set, coset i): a function that ran all k chains would keep the addresses of the rows of
            # F_lambda (the same in every chain) live throughout and spill
            for rows in row_sets(7, mu):                 # all rows of mu: the block of mu in F_lambda is then read once
                for grp in coset_groups(k, inv_cosets(k, dm)):
                    em = ScaledEmitter()

                    def load(lam, t, bo, em=em):
                        return em.tmp(f"in[(size_t){coef_off(lam) + t * dim(lam) + bo} * s]")

                    def store(i, r, v, em=em, dm=dm):
                        em.lines.append(f"    out[(size_t){r * dm * k + i} * s] = {value_expr(v)};")

                    inverse_task_multi(em, k, mu, grp, rows, load, store)
                    total += em.n
                    funcs.append("\n".join([
                        "template <typename T>",
                        f"__device__ __noinline__ void {tag}_{task}(const T* __restrict__ in, T* __restrict__ out, unsigned s) {{"
                        f"   // {mu}, rows {rows[0]}..{rows[-1]}, cosets {grp[0]}..{grp[-1]}"] + em.lines + ["}"]) + "\n")
                    cases.append(f"        case {task}: {tag}_{task}<T>(in + (size_t)c * s, out + (size_t)({coef_off(mu)} + c) * {k} * s, s); break;")
                    task += 1
        groups.append((dm, task0, task - task0))
    inv = "true" if inverse else "false"
    code = funcs
    code.append("\n".join([
        f"template <> struct ColInfo<{k}, {inv}> {{",
        f"    static constexpr int ngroups = {len(groups)};",
        f"    static constexpr int ntasks = {task};",
        "    static const int* dmu() { static const int v[] = {" + ", ".join(str(g[0]) for g in groups) + "}; return v; }",
        "    static const int* task0() { static const int v[] = {" + ", ".join(str(g[1]) for g in groups) + "}; return v; }",
        "    static const int* ntask() { static const int v[] = {" + ", ".join(str(g[2]) for g in groups) + "}; return v; }",
        "};"]) + "\n")
    code.append("\n".join([
        "template <typename T>",
        f"__device__ __forceinline__ void {tag}_run(int task, int c, const T* __restrict__ in, T* __restrict__ out, unsigned s) {{",
        "    switch (task) {"] + cases + ["        default: break;", "    }", "}"]) + "\n")
    return "\n".join(code), groups, task, total


def gen_s6(inverse):
    """Level-6 child-major tasks for the fused levels-2..6 kernel (s6_kernels.cuh): the level-5 side is a SHARED-MEMORY tile
    [coefficient of S_5][coset][S6_LANES instances] that the threads of the same CTA fill (forward) or consume (inverse)
    with the generated levels 2..5 code, so its stride is the compile-time S6_LANES; the level-6 side is the level array
    X_6 in global memory with the run-time stride s.  Same operations as the level-6 column tasks of gen_level()."""
    k = 6
    tag = f"s6{'i' if inverse else 'f'}"
    funcs, cases, items = [], [], []
    for task, mu in enumerate(partitions(k - 1)):
        dm = dim(mu)
        em = ScaledEmitter()
        cache = {}
        if not inverse:
            def load(r, i, em=em, cache=cache, dm=dm):
                if (r, i) not in cache:
                    cache[(r, i)] = em.tmp(f"sm[{r * dm * k + i} * S6_LANES]")
                return cache[(r, i)]

            for r in range(dm):
                for i in range(k):
                    load(r, i)
            for lam in partitions(k):
                if mu not in children(lam):
                    continue
                b = children(lam).index(mu)
                bo, dl = boffs(lam)[b], dim(lam)

                def store(t, v, em=em, lam=lam, dl=dl, bo=bo):
                    em.lines.append(f"    x6[(size_t){coef_off(lam) + t * dl + bo} * s] = {value_expr(v)};")

                forward_task(em, k, lam, b, list(range(dl)), load, store)
            funcs.append("\n".join([
                "template <typename T>",
                f"__device__ __forceinline__ void {tag}_{task}(const T* __restrict__ sm, T* __restrict__ x6, unsigned s) {{"
                f"   // all parents of {mu}"] + em.lines + ["}"]) + "\n")
        else:
            def load(lam, t, bo, em=em, cache=cache):
                if (lam, t) not in cache:
                    cache[(lam, t)] = em.tmp(f"x6[(size_t){coef_off(lam) + t * dim(lam) + bo} * s]")
                return cache[(lam, t)]

            for lam in partitions(k):
                if mu in children(lam):
                    bo = boffs(lam)[children(lam).index(mu)]
                    for t in range(dim(lam)):
                        load(lam, t, bo)
            for i in range(k):
                def store(r, v, em=em, i=i, dm=dm):
                    em.lines.append(f"    sm[{r * dm * k + i} * S6_LANES] = {value_expr(v)};")

                inverse_task(em, k, mu, i, list(range(dm)), load, store)
            funcs.append("\n".join([
                "template <typename T>",
                f"__device__ __forceinline__ void {tag}_{task}(const T* __restrict__ x6, T* __restrict__ sm, unsigned s) {{"
                f"   // {mu}, every coset"] + em.lines + ["}"]) + "\n")
        cases.append(f"        case {task}: {tag}_{task}<T>(a, b, s); break;")
        for c in range(dm):
            items.append((em.n, task, c, coef_off(mu) + c))
    code = funcs
    code.append("\n".join([
        "template <typename T, typename PA, typename PB>",
        f"__device__ __forceinline__ void {tag}_run(int task, PA a, PB b, unsigned s) {{",
        "    switch (task) {"] + cases + ["        default: break;", "    }", "}"]) + "\n")
    items.sort(key=lambda it: -it[0])         # most expensive first: warp w takes the items w, w + NW, w + 2 NW, ..
    return "\n".join(code), items


def write_s6(here):
    parts = ["// GENERATED by gen_col.py -- do not edit.  Level-6 tasks of the fused levels-2..6 kernel (s6_kernels.cuh): the",
             "// level-5 side is a shared-memory tile with the compile-time instance stride S6_LANES.",
             "#pragma once", "#include <stddef.h>", "#include \"vec2.cuh\"", "",
             "#ifndef SNFFT_CODE_FENCE",
             "#define SNFFT_CODE_FENCE() asm volatile(\"\" ::: \"memory\")",
             "#endif", "",
             "namespace snfft {", "namespace {", "",
             "constexpr int S6_LANES = 32;      // instances of level 6 per CTA", ""]
    for inverse in (False, True):
        code, items = gen_s6(inverse)
        parts.append(code)
        d = "i" if inverse else "f"
        parts.append(f"// (mu, c) columns of level 5 by decreasing cost: task = index of mu, c, coef_off(mu) + c")
        parts.append(f"constexpr int S6_NITEMS = {len(items)};" if not inverse else "")
        parts.append(f"static __device__ const unsigned char s6{d}_item_task[] = {{" + ", ".join(str(it[1]) for it in items) + "};")
        parts.append(f"static __device__ const unsigned char s6{d}_item_c[] = {{" + ", ".join(str(it[2]) for it in items) + "};")
        parts.append(f"static __device__ const unsigned char s6{d}_item_off[] = {{" + ", ".join(str(it[3]) for it in items) + "};")
        parts.append("")
    parts += ["}  // namespace", "}  // namespace snfft"]
    write_if_changed(os.path.join(here, "s6_gen.cuh"), "\n".join(parts) + "\n")


def header(k, codes, run_lines):
    parts = [f"// GENERATED by gen_col.py -- do not edit.  Register column code of the level step k = {k}",
             "// (see gen_col.py for the derivation and the reference citations).",
             "#pragma once", "#include <stddef.h>", "#include \"vec2.cuh\"", "",
             "#ifndef SNFFT_CODE_FENCE",
             "#define SNFFT_CODE_FENCE() asm volatile(\"\" ::: \"memory\")",
             "#endif", "",
             "// anonymous namespace: the parts of a split level define the same names with different contents in",
             "// different translation units (they must not be merged at link time)",
             "namespace snfft {", "namespace {", "",
             "template <int K, bool INV> struct ColInfo;", ""] + codes
    parts.append("template <typename T, int K, bool INV>")
    parts.append("__device__ __forceinline__ void col_run(int task, int c, const T* __restrict__ in, T* __restrict__ out, unsigned s) {")
    parts.append(f"    static_assert(K == {k}, \"this translation unit holds the code of one level\");")
    parts += run_lines
    parts.append("}")
    parts.append("")
    parts.append("}  // namespace")
    parts.append("}  // namespace snfft")
    return "\n".join(parts) + "\n"


def write_if_changed(path, text):
    if not os.path.exists(path) or open(path).read() != text:      # untouched levels keep their objects
        with open(path, "w") as fh:
            fh.write(text)


def split_groups(k, inverse, nparts):
    """Child shapes of level k in `nparts` sets of about equal generated code (largest first, greedy)."""
    cost = {mu: gen_level(k, inverse, only={mu})[3] for mu in partitions(k - 1)}
    sets, load = [set() for _ in range(nparts)], [0] * nparts
    for mu in sorted(cost, key=lambda m: -cost[m]):
        p = load.index(min(load))
        sets[p].add(mu)
        load[p] += cost[mu]
    return sets


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    stats = {}
    for k in LEVELS:            # one header per level: a translation unit only sees (and depends on) its own
        codes = []
        for inverse in (False, True):
            code, groups, ntask, n = gen_level(k, inverse)
            stats[(k, inverse)] = (ntask, n)
            codes.append(code)
        write_if_changed(os.path.join(here, f"col_gen_{k}.cuh"),
                         header(k, codes, [f"    if constexpr (!INV) col{k}f_run<T>(task, c, in, out, s);",
                                           f"    else col{k}i_run<T>(task, c, in, out, s);"]))
    for k, nparts in SPLIT_FWD.items():      # forward only, one header (= one kernel launch) per set of child shapes
        tot = [0, 0]
        for p, only in enumerate(split_groups(k, False, nparts)):
            code, groups, ntask, n = gen_level(k, False, only=only)
            tot[0] += ntask
            tot[1] += n
            write_if_changed(os.path.join(here, f"col_gen_{k}p{p}.cuh"),
                             header(k, [code], ["    static_assert(!INV, \"forward only\");",
                                                f"    col{k}f_run<T>(task, c, in, out, s);"]))
        stats[(k, False)] = tuple(tot)
    write_s6(here)
    print("(level, inverse): (tasks, statements)", stats)


if __name__ == "__main__":
    main()
