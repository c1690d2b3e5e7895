#!/usr/bin/env python
"""Generate tz_gen_z7.cuh and tz_gen_t<k>.cuh: the "top / bottom" factorisation of the level steps k = 9, 10.

    python algebraist_b200/csrc/gen_tz.py         # rewrites the headers next to this file

A forward level step (reference fourier.py:253-273) computes, for column c of block mu of F_lambda,
    F = sum_i rho(s_i) .. rho(s_{k-2}) embed_mu(x_i),          x_i = F^(i)_mu[:, c].
A row of lambda is a path lambda -> mu_{k-1} -> .. -> mu_1 in Young's lattice (restriction-adapted basis,
tableau.py:171-198) and rho(s_j) only changes mu_{j+1} (irreps.py:91-109).  With MB = 7 and s = k - MB:
  * the bottom sweeps s_i .. s_{MB-2} act inside the blocks alpha = mu_MB (partitions of 7, d <= 35) as the irrep
    alpha of S_7 -- the same for every lambda and every path above alpha;
  * the top sweeps s_{MB-1} .. s_{k-2} act on (mu_{k-1} .. mu_MB) and leave (zeta = mu_{MB-1}, tableau) alone.
So the sum over the cosets can be taken BEFORE the top sweeps:
    Z[(P, zeta) -> alpha] = sum_{i < MB} rho_alpha(s_i .. s_{MB-2}) embed_zeta(x_i[(P, zeta), :])        ("Z kernel")
    F[(P', alpha), r]     = sum_{zeta in alpha, P: mu -> zeta} C[(P', alpha); (mu, P) | zeta] Z[(P, zeta) -> alpha][r]
                            + the cosets i >= MB (top sweeps only)                                       ("T kernel")
with P the paths mu -> zeta, P' the paths lambda -> alpha (s steps each) and C the matrix elements of the top
sweeps, which do not depend on the row r of alpha.  Z is a level-7 forward column operator on pieces of the inputs
(the generated level-7 code of gen_col.py with other strides); it does not depend on lambda or P', so it is
shared by all parents of mu.  The T kernel is a small dense map per (mu, alpha), the same for every (r, c, instance):
a thread owns (r, c, instance), loads <= 30 values and writes <= 30.  Both kernels stream: every element of Z, of
the inputs and of the outputs moves once (plus the L2 re-reads of the Z kernel's sibling tasks), no shared memory.
Operation count at k = 10: 24.8 * 0.7 (Z) + 6.6 (T) FMA per output element against 45.6 for the plain sweeps
(tools/tz_model.py checks the factorisation numerically and counts).

The inverse (adjoint, fourier.py:191-199) runs the transposed maps in the opposite order:
    Z'[(P, zeta) -> alpha][r] = sum_{lambda, P'} d_lambda/(k d_mu) C[(P', alpha); (mu, P) | zeta] F_lambda[(P', alpha, r), c]
    G^(i)[(P, zeta), :]       = sum_alpha [rho_alpha(s_{MB-2}) .. rho_alpha(s_i) Z'[(P, zeta) -> alpha]]_{rows of zeta},   i < MB
and the cosets i >= MB straight from F.

Z layout: for mu of k-1 a block of MB d_mu rows x d_mu columns x NINST_k instances at MB * coef_off(mu) * NINST_k;
piece (P, zeta) at row offset `off` of mu owns the rows [MB off, MB (off + d_zeta)): its parents alpha in
partition order, d_alpha rows each (sum = MB d_zeta, the branching rule).
"""
import math
import os

from gen_base import ScaledEmitter, boffs, children, coef_off, dim, partitions
from gen_col import (FWD_ROWS_CAP, coset_groups, inv_cosets, forward_task, forward_task_multi, inverse_task_multi,
                     need_chain, parent_groups, row_sets, sweep_need, value_expr, write_if_changed)

LEVEL_MB = {9: 7, 10: 7, 11: 8}        # level k -> bottom level MB (the Z kernel runs level-MB column operators)
LEVELS = tuple(LEVEL_MB)


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


def aoff(zeta, alpha):
    """Row offset of the parent alpha inside the MB d_zeta rows of a piece (parents in partition order)."""
    return _aoff(zeta, alpha)


def _aoff(zeta, alpha):
    off = 0
    for a in parents(zeta):
        if a == alpha:
            return off
        off += dim(a)
    raise KeyError((zeta, alpha))


# ----------------------------------------------------------------------------
# Z kernel code: level-MB column operators with run-time strides
# ----------------------------------------------------------------------------
def gen_z(MB, inverse):
    """Level-MB column operators.  Forward: one task per (zeta, parent alpha, row set of alpha): all rows for
    MB = 7, one child block of alpha for MB = 8 (as in gen_col.py).  Inverse: one task per (zeta, run of adjacent cosets)."""
    tag = f"zb{MB}{'i' if inverse else 'f'}"
    funcs, cases, groups = [], [], []
    task = 0
    total = 0
    for zeta in partitions(MB - 1):
        dz = dim(zeta)
        task0 = task
        if not inverse and MB == 7:
            for grp in parent_groups(zeta, FWD_ROWS_CAP[7]):       # several parents per task: inputs loaded once per group
                em = ScaledEmitter()
                targets = [(alpha, children(alpha).index(zeta), list(range(dim(alpha)))) for alpha in grp]

                def load(r, i, em=em):
                    return em.tmp(f"in[(size_t){r} * rs + (size_t){i} * s]")

                def store(ti, t, v, em=em, targets=targets, zeta=zeta):
                    em.lines.append(f"    out[(size_t){aoff(zeta, targets[ti][0]) + t} * zs] = {value_expr(v)};")

                forward_task_multi(em, MB, targets, load, store)
                total += em.n
                funcs.append("\n".join([
                    "template <typename T>",
                    f"__device__ __noinline__ void {tag}_{task}(const T* __restrict__ in, T* __restrict__ out, unsigned rs, unsigned s, unsigned zs) {{"
                    f"   // {', '.join(str(a) for a in grp)} <- {zeta}"] + em.lines + ["}"]) + "\n")
                cases.append(f"        case {task}: {tag}_{task}<T>(in, out, rs, s, zs); break;")
                task += 1
        elif not inverse:
            for alpha in parents(zeta):
              for rows in row_sets(MB, alpha):
                em = ScaledEmitter()
                cache = {}

                def load(r, i, em=em, cache=cache):
                    if (r, i) not in cache:
                        cache[(r, i)] = em.tmp(f"in[(size_t){r} * rs + (size_t){i} * s]")
                    return cache[(r, i)]

                def store(t, v, em=em, ao=aoff(zeta, alpha)):
                    em.lines.append(f"    out[(size_t){ao + t} * zs] = {value_expr(v)};")

                forward_task(em, MB, alpha, children(alpha).index(zeta), rows, load, store)
                total += em.n
                funcs.append("\n".join([
                    "template <typename T>",
                    f"__device__ __noinline__ void {tag}_{task}(const T* __restrict__ in, T* __restrict__ out, unsigned rs, unsigned s, unsigned zs) {{"
                    f"   // {alpha} <- {zeta}, rows {rows[0]}..{rows[-1]}"] + em.lines + ["}"]) + "\n")
                cases.append(f"        case {task}: {tag}_{task}<T>(in, out, rs, s, zs); break;")
                task += 1
        else:
            for grp in coset_groups(MB, inv_cosets(MB, dz)):
                em = ScaledEmitter()

                def load(alpha, t, bo, em=em, zeta=zeta):
                    return em.tmp(f"in[(size_t){aoff(zeta, alpha) + t} * zs]")

                def store(i, r, v, em=em):
                    em.lines.append(f"    out[(size_t){r} * rs + (size_t){i} * s] = {value_expr(v)};")

                inverse_task_multi(em, MB, zeta, grp, list(range(dz)), load, store, unit_scale=True)
                total += em.n
                funcs.append("\n".join([
                    "template <typename T>",
                    f"__device__ __noinline__ void {tag}_{task}(const T* __restrict__ in, T* __restrict__ out, unsigned rs, unsigned s, unsigned zs) {{"
                    f"   // {zeta}, cosets {grp[0]}..{grp[-1]}"] + em.lines + ["}"]) + "\n")
                cases.append(f"        case {task}: {tag}_{task}<T>(in, out, rs, s, zs); break;")
                task += 1
        groups.append((dz, task0, task - task0))
    inv = "true" if inverse else "false"
    code = funcs
    code.append("\n".join([
        f"template <> struct ZInfo<{MB}, {inv}> {{",
        f"    static constexpr int ngroups = {len(groups)};",
        f"    static constexpr int ntasks = {task};",
        "    static const int* task0() { static const int v[] = {" + ", ".join(str(g[1]) for g in groups) + "}; return v; }",
        "    static const int* ntask() { static const int v[] = {" + ", ".join(str(g[2]) for g in groups) + "}; return v; }",
        "};"]) + "\n")
    code.append("\n".join([
        "template <typename T>",
        f"__device__ __forceinline__ void {tag}_run(int task, const T* __restrict__ in, T* __restrict__ out, unsigned rs, unsigned s, unsigned zs) {{",
        "    switch (task) {"] + cases + ["        default: break;", "    }", "}"]) + "\n")
    return "\n".join(code), task, total


# ----------------------------------------------------------------------------
# T kernel code: the top sweeps of level k, one function per (mu, alpha)
# ----------------------------------------------------------------------------
def t_tasks(k):
    MB = LEVEL_MB[k]
    s = k - MB
    out = []
    for mu in partitions(k - 1):
        seen = []
        for lam in parents(mu):
            for Pp, _ in subblocks(lam, s):
                if Pp[-1] not in seen:
                    seen.append(Pp[-1])
        for alpha in partitions(MB):
            if alpha in seen:
                out.append((mu, alpha))
    return out


def t_forward(em, k, mu, alpha):
    MB = LEVEL_MB[k]
    s = k - MB
    dmu, da = dim(mu), dim(alpha)
    zcache, xcache = {}, {}

    def load_z(off, zeta):
        key = (off, zeta)
        if key not in zcache:
            row = MB * off + aoff(zeta, alpha)
            zcache[key] = em.tmp(f"zb[(size_t){MB * coef_off(mu) + row * dmu} * s]")
        return zcache[key]

    def load_x(offq, i):
        key = (offq, i)
        if key not in xcache:
            xcache[key] = em.tmp(f"xb[(size_t){(coef_off(mu) + offq * dmu) * k + i} * s]")
        return xcache[key]

    pieces = [(P, off) for P, off in subblocks(mu, s) if P[-1] in children(alpha)]
    direct = [(Q, offq) for Q, offq in subblocks(mu, s - 1) if Q[-1] == alpha]
    for P, off in pieces:               # every load first (memory-level parallelism)
        load_z(off, P[-1])
    for Q, offq in direct:
        for i in range(MB, k):
            load_x(offq, i)
    stores = []
    for li, lam in enumerate(parents(mu)):
        outs = [offp for Pp, offp in subblocks(lam, s) if Pp[-1] == alpha]
        if not outs:
            continue
        bo, dl = boffs(lam)[children(lam).index(mu)], dim(lam)
        acc = {offp: None for offp in outs}
        bz = {zeta: boffs(alpha)[b] for b, zeta in enumerate(children(alpha))}
        # Z part: representative rows (.., zeta, t = 0); the top sweeps keep (zeta, t)
        x = {bo + off: (1.0, load_z(off, P[-1])) for P, off in pieces}
        want = sorted({offp + o for offp in outs for o in bz.values()})
        seq = list(range(k - 2, MB - 2, -1))
        need = need_chain(lam, seq, want)
        x = {t: v for t, v in x.items() if t in need[0]}
        for m, j in enumerate(seq):
            x = sweep_need(em, lam, j, x, need[m + 1])
        for offp in outs:
            acc[offp] = em.lin([(1.0, x.get(offp + o)) for o in bz.values()])
        # cosets i >= MB: sweeps s_{k-2} .. s_i keep alpha and the rows below (representative row r = 0)
        for i in range(MB, k):
            seq = list(range(k - 2, i - 1, -1))
            need = need_chain(lam, seq, outs)
            x = {bo + offq: (1.0, load_x(offq, i)) for Q, offq in direct if bo + offq in need[0]}
            for m, j in enumerate(seq):
                x = sweep_need(em, lam, j, x, need[m + 1])
            for offp in outs:
                acc[offp] = em.lin([(1.0, acc[offp]), (1.0, x.get(offp))])
        for offp in outs:
            stores.append(f"    ob{li}.st({offp * dl + bo}u, {value_expr(acc[offp])});")
    em.lines += stores
    return [(li, dim(lam), coef_off(lam)) for li, lam in enumerate(parents(mu))
            if any(Pp[-1] == alpha for Pp, _ in subblocks(lam, s))]


def t_inverse(em, k, mu, alpha):
    MB = LEVEL_MB[k]
    s = k - MB
    dmu, da = dim(mu), dim(alpha)
    pieces = [(P, off) for P, off in subblocks(mu, s) if P[-1] in children(alpha)]
    direct = [(Q, offq) for Q, offq in subblocks(mu, s - 1) if Q[-1] == alpha]
    bz = {zeta: boffs(alpha)[b] for b, zeta in enumerate(children(alpha))}
    zacc = {(off, P[-1]): None for P, off in pieces}
    xacc = {(offq, i): None for Q, offq in direct for i in range(MB, k)}
    used = []
    loads = {}
    for li, lam in enumerate(parents(mu)):      # every load first
        outs = [offp for Pp, offp in subblocks(lam, s) if Pp[-1] == alpha]
        if not outs:
            continue
        used.append((li, dim(lam), coef_off(lam)))
        bo, dl = boffs(lam)[children(lam).index(mu)], dim(lam)
        for offp in outs:
            loads[(li, offp)] = em.tmp(f"ib{li}.ld({offp * dl + bo}u)")
    for li, lam in enumerate(parents(mu)):
        outs = [offp for Pp, offp in subblocks(lam, s) if Pp[-1] == alpha]
        if not outs:
            continue
        bo, dl = boffs(lam)[children(lam).index(mu)], dim(lam)
        w = dl / (k * dmu)                       # fourier.py:191-199
        # Z' part, every zeta channel at once: F[(P', alpha), r] sits at the representative rows (P', alpha, zeta, 0)
        x = {}
        for offp in outs:
            for o in bz.values():
                x[offp + o] = (1.0, loads[(li, offp)])
        seq = list(range(MB - 1, k - 1))        # transposed: ascending
        want = sorted({bo + off for P, off in pieces})
        need = need_chain(lam, seq, want)
        x = {t: v for t, v in x.items() if t in need[0]}
        for m, j in enumerate(seq):
            x = sweep_need(em, lam, j, x, need[m + 1])
        for P, off in pieces:
            key = (off, P[-1])
            zacc[key] = em.lin([(1.0, zacc[key]), (w, x.get(bo + off))])
        for i in range(MB, k):
            seq = list(range(i, k - 1))
            want = sorted({bo + offq for Q, offq in direct})
            need = need_chain(lam, seq, want)
            x = {offp: (1.0, loads[(li, offp)]) for offp in outs if offp in need[0]}
            for m, j in enumerate(seq):
                x = sweep_need(em, lam, j, x, need[m + 1])
            for Q, offq in direct:
                xacc[(offq, i)] = em.lin([(1.0, xacc[(offq, i)]), (w, x.get(bo + offq))])
    for (off, zeta), v in zacc.items():
        row = MB * off + aoff(zeta, alpha)
        em.lines.append(f"    zb[(size_t){MB * coef_off(mu) + row * dmu} * s] = {value_expr(v)};")
    for (offq, i), v in xacc.items():
        em.lines.append(f"    xb[(size_t){(coef_off(mu) + offq * dmu) * k + i} * s] = {value_expr(v)};")
    return used


def gen_t(k, inverse):
    tag = f"tz{k}{'i' if inverse else 'f'}"
    funcs, cases = [], []
    tasks = t_tasks(k)
    total = 0
    for task, (mu, alpha) in enumerate(tasks):
        em = ScaledEmitter()
        used = (t_inverse if inverse else t_forward)(em, k, mu, alpha)
        total += em.n
        dmu = dim(mu)
        cq = "" if inverse else "const "
        cl = "const " if inverse else ""
        # the level-k side goes through TzAcc (tz_kernels.cuh): PK = false the level array X_k, PK = true the public
        # packed layout [lambda][B][d][d] (top level only: the pack / unpack pass is fused into this kernel)
        head = [
            "template <typename T, bool PK>",
            f"__device__ __noinline__ void {tag}_{task}({cq}T* __restrict__ zb, {cq}T* __restrict__ xb, {cl}T* __restrict__ lb, unsigned r, unsigned c, unsigned s, unsigned inst) {{"
            f"   // mu = {mu}, alpha = {alpha}",
            f"    zb += (size_t)(r * {dmu}u + c) * s;",
            f"    xb += (size_t)(r * {dmu}u + c) * {k}u * s;"]
        for li, dl, co in used:
            head.append(f"    const TzAcc<T, PK, {'true' if inverse else 'false'}> {'ib' if inverse else 'ob'}{li}(lb, {co}u, {dl}u, r, c, s, inst);")
        funcs.append("\n".join(head + em.lines + ["}"]) + "\n")
        cases.append(f"        case {task}: {tag}_{task}<T, PK>(zb, xb, lb, r, c, s, inst); break;")
    code = funcs
    code.append("\n".join([
        "template <typename T, bool PK>",
        f"__device__ __forceinline__ void {tag}_run(int task, {'' if inverse else 'const '}T* __restrict__ zb, {'' if inverse else 'const '}T* __restrict__ xb, "
        f"{'const ' if inverse else ''}T* __restrict__ lb, unsigned r, unsigned c, unsigned s, unsigned inst) {{",
        "    switch (task) {"] + cases + ["        default: break;", "    }", "}"]) + "\n")
    return "\n".join(code), total


def level_tables(k):
    """Static tables of level k: T tasks (d_alpha, d_mu) and the segments of the Z kernel by group (zeta)."""
    MB = LEVEL_MB[k]
    s = k - MB
    tasks = t_tasks(k)
    zetas = list(partitions(MB - 1))
    segs = []           # (group, in_off, z_off, d_mu)
    for g, zeta in enumerate(zetas):
        for mu in partitions(k - 1):
            dmu = dim(mu)
            for P, off in subblocks(mu, s):
                if P[-1] == zeta:
                    segs.append((g, (coef_off(mu) + off * dmu) * k, MB * (coef_off(mu) + off * dmu), dmu))
    lines = [
        f"template <> struct TzLevel<{k}> {{",
        f"    static constexpr int MB = {MB};",
        f"    static constexpr int ntasks = {len(tasks)};",
        f"    static constexpr int nsegs = {len(segs)};",
        "    static const int* t_dal() { static const int v[] = {" + ", ".join(str(dim(a)) for _, a in tasks) + "}; return v; }",
        "    static const int* t_dmu() { static const int v[] = {" + ", ".join(str(dim(m)) for m, _ in tasks) + "}; return v; }",
        "    static const int* seg_group() { static const int v[] = {" + ", ".join(str(x[0]) for x in segs) + "}; return v; }",
        "    static const int* seg_in() { static const int v[] = {" + ", ".join(str(x[1]) for x in segs) + "}; return v; }",
        "    static const int* seg_z() { static const int v[] = {" + ", ".join(str(x[2]) for x in segs) + "}; return v; }",
        "    static const int* seg_dmu() { static const int v[] = {" + ", ".join(str(x[3]) for x in segs) + "}; return v; }",
        "};"]
    return "\n".join(lines) + "\n"


HEAD = ["#pragma once", "#include <stddef.h>", "#include \"vec2.cuh\"", "",
        "#ifndef SNFFT_CODE_FENCE",
        "#define SNFFT_CODE_FENCE() asm volatile(\"\" ::: \"memory\")",
        "#endif", "",
        "// anonymous namespace: every translation unit sees one header of this family",
        "namespace snfft {", "namespace {", ""]
TAIL = ["}  // namespace", "}  // namespace snfft"]


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    stats = {}
    for MB in sorted(set(LEVEL_MB.values())):
        parts = [f"// GENERATED by gen_tz.py -- do not edit.  Level-{MB} column operators with run-time strides (Z kernel)."] + HEAD
        parts.append("template <int MB, bool INV> struct ZInfo;\n")
        for inverse in (False, True):
            code, ntask, n = gen_z(MB, inverse)
            stats[(f"z{MB}", inverse)] = (ntask, n)
            parts.append(code)
        write_if_changed(os.path.join(here, f"tz_gen_z{MB}.cuh"), "\n".join(parts + TAIL) + "\n")
    for k in LEVELS:
        parts = [f"// GENERATED by gen_tz.py -- do not edit.  Top sweeps of the level step k = {k} (T kernel) and the tables of",
                 "// its Z kernel launches."] + HEAD
        for inverse in (False, True):
            code, n = gen_t(k, inverse)
            stats[(k, inverse)] = (len(t_tasks(k)), n)
            parts.append(code)
        tasks = t_tasks(k)
        parts.append("\n".join([
            "// per task: d_alpha, d_mu (device copies of TzLevel<K>::t_dal / t_dmu)",
            f"static __device__ const unsigned short tz{k}_dal[] = {{" + ", ".join(str(dim(a)) for _, a in tasks) + "};",
            f"static __device__ const unsigned short tz{k}_dmu[] = {{" + ", ".join(str(dim(m)) for m, _ in tasks) + "};",
            "template <int K> struct TzDev;",
            f"template <> struct TzDev<{k}> {{",
            f"    static __device__ __forceinline__ unsigned dal(int t) {{ return tz{k}_dal[t]; }}",
            f"    static __device__ __forceinline__ unsigned dmu(int t) {{ return tz{k}_dmu[t]; }}",
            "    template <typename T, bool INV, bool PK, typename PZ, typename PL>",
            "    static __device__ __forceinline__ void run(int task, PZ zb, PZ xb, PL lb, unsigned r, unsigned c, unsigned s, unsigned inst) {",
            f"        if constexpr (!INV) tz{k}f_run<T, PK>(task, zb, xb, lb, r, c, s, inst);",
            f"        else tz{k}i_run<T, PK>(task, zb, xb, lb, r, c, s, inst);",
            "    }",
            "};"]) + "\n")
        write_if_changed(os.path.join(here, f"tz_gen_t{k}.cuh"), "\n".join(parts + TAIL) + "\n")
    parts = ["// GENERATED by gen_tz.py -- do not edit.  Task and segment tables of the T / Z kernels per level."] + HEAD
    parts.append("template <int K> struct TzLevel;\n")
    for k in LEVELS:
        parts.append(level_tables(k))
    write_if_changed(os.path.join(here, "tz_gen_tab.cuh"), "\n".join(parts + TAIL) + "\n")
    print("(kernel, inverse): (tasks, statements)", stats)


if __name__ == "__main__":
    main()
