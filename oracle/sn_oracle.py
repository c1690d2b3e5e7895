"""CPU oracle for the S_n fast Fourier transform  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the algorithm of dashstander/algebraist's
`sn_fft` / `sn_ifft` / `slow_sn_ft` (reference `src/algebraist/fourier.py`,
`irreps.py`, `tableau.py`, `permutations.py`, `utils.py`).  It exists so that
the CUDA path can be checked; nothing in `algebraist_b200/` imports it.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
arm may import or execute it.

Parity status: PINNED.  The reference ships no golden vectors for the
recursive path, so this oracle is pinned by executing the unmodified reference
in the build container (`oracle/make_golden.py`, which imports
`/root/reference/src`) and committing the outputs under `tests/golden/`:
  * forward  n = 3..8 : reference `sn_fft` (fourier.py:281-288), fp64 + fp32
  * inverse  n = 3..5 : reference `sn_ifft` (fourier.py:298-299)
  * inverse  n = 6..7 : reference dense inverse (`BASE_CASE = n`,
                        fourier.py:134-165) because the reference's recursive
                        inverse raises for n > 5 (fourier.py:202-214)
  * tables   n <= 7   : YOR adjacent-transposition matrices
                        (irreps.py:91-109), basis order (tableau.py:171-198),
                        coset index sets (fourier.py:77-80)
and by the reference's own known-answer tests (tests/test_tableau.py:67-100,
tests/test_permutations.py:84-87, tests/test_irreps.py:32-46).

Algorithm (all citations into /root/reference/src/algebraist):
  * partition order           tableau.py:103-123
  * children / block offsets  irreps.py:64-83, tableau.py:38-56
  * dimension (hook length)   tableau.py:236-265
  * basis (tableau id) order  tableau.py:171-198   (sort by position of the
                              largest entry, recurse)  ==> the basis of lambda
                              is the concatenation of its children's bases in
                              ascending child order
  * YOR coefficients          irreps.py:91-109, tableau.py:214-220
  * coset i = {sigma: sigma[i] = n-1}, lexicographic   fourier.py:63-80
  * coset representative matrices rho(c_i) = rho(s_i) rho(s_{i+1}) ... rho(s_{n-2})
                              irreps.py:28-35, 150-155
  * forward level step  F_lambda = sum_i rho_lambda(c_i) blockdiag_mu F_mu(f_i)
                              fourier.py:223-278
  * inverse level step (the formula the reference comments on at
    fourier.py:191-199; its implementation at :202-214 is defective)
        G^(i)_mu = sum_{lambda covers mu} d_lambda/(k d_mu) [rho_lambda(c_i)^T F_lambda]_{mu block}
  * dense transforms          fourier.py:82-108 (slow_sn_ft), 134-159, 302-327
"""
from __future__ import annotations

import math
from functools import lru_cache
from itertools import permutations as _iter_permutations

import numpy as np

# ----------------------------------------------------------------------------
# combinatorics
# ----------------------------------------------------------------------------

_SMALL_PARTITIONS = {  # tableau.py:105-112 (hard coded lists for n <= 5)
    0: [],
    1: [(1,)],
    2: [(2,), (1, 1)],
    3: [(3,), (2, 1), (1, 1, 1)],
    4: [(4,), (3, 1), (2, 2), (2, 1, 1), (1, 1, 1, 1)],
    5: [(5,), (4, 1), (3, 2), (3, 1, 1), (2, 2, 1), (2, 1, 1, 1), (1, 1, 1, 1, 1)],
}


@lru_cache(maxsize=None)
def generate_partitions(n: int) -> tuple:
    """Partitions of n in the reference's dict-key order (tableau.py:103-123)."""
    if n <= 5:
        return tuple(_SMALL_PARTITIONS[n])
    out = [(n,)]
    for k in range(1, n):
        for p in generate_partitions(n - k):
            q = tuple(sorted((k, *p), reverse=True))
            if q not in out:
                out.append(q)
    return tuple(out)


def children(part: tuple) -> list:
    """Partitions covered by `part`, ascending tuple order (irreps.py:64-68)."""
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
def dim(part: tuple) -> int:
    """Hook length formula (tableau.py:236-265)."""
    n = sum(part)
    if n <= 1:
        return 1
    prod = 1
    for i, row in enumerate(part):
        for j in range(row):
            prod *= row - j + sum(1 for r in part[i + 1:] if r > j)
    return math.factorial(n) // prod


def block_offsets(part: tuple) -> list:
    """Row offset of every child block (irreps.py:70-83)."""
    off, out = 0, []
    for c in children(part):
        out.append(off)
        off += dim(c)
    return out


def _removed_box(part: tuple, child: tuple):
    """(row, col) of the box in part but not in child."""
    for r in range(len(part)):
        cr = child[r] if r < len(child) else 0
        if part[r] != cr:
            return r, part[r] - 1
    raise ValueError("not a child")


@lru_cache(maxsize=None)
def last_transposition_table(part: tuple):
    """Young's orthogonal form of the LAST adjacent transposition s_{k-2} of S_k
    on the irrep `part` (k = |part| >= 2), as per-row arrays.

    Returns (partner, diag, off): rho[t,t] = diag[t]; rho[t, partner[t]] = off[t]
    (partner[t] == t and off[t] == 0 when the swapped tableau is not standard).
    Follows irreps.py:91-109 with the axial distance of tableau.py:214-220:
        d = (row_a - row_b) + (col_b - col_a),  a = k-2, b = k-1.
    Row t = (mu, nu, rest): entry k-1 sits in part/mu, entry k-2 in mu/nu.
    """
    k = sum(part)
    d_l = dim(part)
    partner = np.arange(d_l, dtype=np.int64)
    diag = np.zeros(d_l, dtype=np.float64)
    off = np.zeros(d_l, dtype=np.float64)
    ch = children(part)
    boff = block_offsets(part)
    for b, mu in enumerate(ch):
        rb, cb = _removed_box(part, mu)          # position of entry k-1
        sub = children(mu)
        sboff = block_offsets(mu)
        for c, nu in enumerate(sub):
            ra, ca = _removed_box(mu, nu)        # position of entry k-2
            dist = (ra - rb) + (cb - ca)
            lo = boff[b] + sboff[c]
            hi = lo + dim(nu)
            diag[lo:hi] = 1.0 / dist
            if abs(dist) > 1:
                # the swapped tableau: k-1 goes to (ra,ca), k-2 to (rb,cb)
                mu2 = list(part)
                mu2[ra] -= 1
                if mu2[-1] == 0:
                    mu2.pop()
                mu2 = tuple(mu2)
                b2 = ch.index(mu2)
                c2 = children(mu2).index(nu)
                lo2 = boff[b2] + block_offsets(mu2)[c2]
                partner[lo:hi] = np.arange(lo2, lo2 + dim(nu))
                off[lo:hi] = np.sqrt(1.0 - 1.0 / (dist * dist))
    return partner, diag, off


@lru_cache(maxsize=None)
def transposition_table(part: tuple, j: int):
    """Per-row YOR table of s_j = (j j+1) on `part`, any 0 <= j <= |part|-2.
    For j < k-2 it is the block-diagonal stack of the children's tables
    (the restriction property the basis order guarantees, tableau.py:171-198)."""
    k = sum(part)
    assert 0 <= j <= k - 2
    if j == k - 2:
        return last_transposition_table(part)
    ps, ds, os_ = [], [], []
    for mu, o in zip(children(part), block_offsets(part)):
        p, d, f = transposition_table(mu, j)
        ps.append(p + o)
        ds.append(d)
        os_.append(f)
    return np.concatenate(ps), np.concatenate(ds), np.concatenate(os_)


def transposition_matrix(part: tuple, j: int) -> np.ndarray:
    """Dense rho_part(s_j) (irreps.py:91-109)."""
    p, d, f = transposition_table(part, j)
    m = np.diag(d)
    rows = np.arange(len(p))
    m[rows, p] += np.where(p != rows, f, 0.0)
    return m


def apply_transposition(part: tuple, j: int, x: np.ndarray) -> np.ndarray:
    """rho_part(s_j) @ x for x of shape (d, ...)."""
    p, d, f = transposition_table(part, j)
    shp = (-1,) + (1,) * (x.ndim - 1)
    return d.reshape(shp) * x + f.reshape(shp) * x[p]


def coset_rep_matrix(part: tuple, i: int) -> np.ndarray:
    """rho(c_i) = rho(s_i) rho(s_{i+1}) ... rho(s_{k-2}) (irreps.py:28-35,150-155)."""
    k = sum(part)
    m = np.eye(dim(part))
    for j in range(k - 2, i - 1, -1):
        m = transposition_matrix(part, j) @ m
    return m


# ----------------------------------------------------------------------------
# permutation <-> index (Lehmer code)  and coset indexing
# ----------------------------------------------------------------------------

def perm_rank(perms: np.ndarray) -> np.ndarray:
    """Lexicographic rank of one-line permutations (permutations.py:187-192)."""
    perms = np.asarray(perms)
    n = perms.shape[-1]
    r = np.zeros(perms.shape[:-1], dtype=np.int64)
    for i in range(n):
        smaller = (perms[..., i + 1:] < perms[..., i:i + 1]).sum(-1)
        r += smaller * math.factorial(n - 1 - i)
    return r


def perm_unrank(ranks: np.ndarray, n: int) -> np.ndarray:
    """Inverse of perm_rank == rows of generate_all_permutations (utils.py:61-63)."""
    ranks = np.asarray(ranks, dtype=np.int64)
    out = np.zeros(ranks.shape + (n,), dtype=np.int64)
    flat = ranks.reshape(-1)
    res = out.reshape(-1, n)
    for q, r in enumerate(flat):
        avail = list(range(n))
        for i in range(n):
            f = math.factorial(n - 1 - i)
            dgt, r = divmod(int(r), f)
            res[q, i] = avail.pop(dgt)
    return out


def coset_index(n: int, i: int) -> np.ndarray:
    """Ranks (in S_n) of coset i = {sigma : sigma[i] = n-1} in lexicographic
    order == argwhere(sn_perms[:, i] == n-1) (fourier.py:77-80).  Closed form of
    SURVEY 3.4, vectorised."""
    m = math.factorial(n - 1)
    r = np.arange(m, dtype=np.int64)
    idx = np.full(m, (n - 1 - i) * math.factorial(n - 1 - i), dtype=np.int64)
    for j in range(n - 1):
        lj = (r // math.factorial(n - 2 - j)) % (n - 1 - j)
        pos = j if j < i else j + 1
        idx += lj * math.factorial(n - 1 - pos)
    return idx


@lru_cache(maxsize=8)
def gather_index(n: int) -> np.ndarray:
    """Input rank of every level-1 instance.

    Instances: at level k an instance is a chain prefix (i_n, ..., i_{k+1});
    p_{k-1} = i_k * P_k + p_k with P_n = 1, P_{k-1} = k * P_k.  Returns an
    int64 array g of n! entries with g[p_1] = rank in S_n of the permutation
    c^(n)_{i_n} * c^(n-1)_{i_{n-1}} * ... (fourier.py:253 applied recursively)."""
    idx = np.arange(math.factorial(n), dtype=np.int64)[None, :]     # (P_n, n!)
    for k in range(n, 1, -1):
        cos = np.stack([coset_index(k, i) for i in range(k)])        # (k, (k-1)!)
        # new[i * P_k + p, r] = idx[p, cos[i, r]]
        idx = idx[:, cos]                                            # (P_k, k, (k-1)!)
        idx = np.ascontiguousarray(idx.transpose(1, 0, 2)).reshape(-1, cos.shape[1])
    return idx[:, 0].copy()


# ----------------------------------------------------------------------------
# fast transforms (Clausen recursion with sparse YOR sweeps)
# ----------------------------------------------------------------------------

def _level_forward(k: int, prev: dict, ninst: int) -> dict:
    """One Clausen step S_{k-1} -> S_k (fourier.py:253-273).
    prev[mu] has shape (d_mu, d_mu, k * ninst), instance index i * ninst + p."""
    out = {}
    for lam in generate_partitions(k):
        d = dim(lam)
        ch, bo = children(lam), block_offsets(lam)
        acc = np.zeros((d, d, ninst), dtype=next(iter(prev.values())).dtype)
        for i in range(k):
            w = np.zeros_like(acc)
            for mu, o in zip(ch, bo):
                dm = dim(mu)
                w[o:o + dm, o:o + dm, :] = prev[mu][:, :, i * ninst:(i + 1) * ninst]
            for j in range(k - 2, i - 1, -1):          # rho(s_i)...rho(s_{k-2}) @ w
                w = apply_transposition(lam, j, w)
            acc += w
        out[lam] = acc
    return out


def _level_inverse(k: int, cur: dict, ninst: int) -> dict:
    """Adjoint Clausen step S_k -> S_{k-1} with the d_lambda/(k d_mu) weights
    (formula at fourier.py:191-199)."""
    dt = next(iter(cur.values())).dtype
    out = {mu: np.zeros((dim(mu), dim(mu), k * ninst), dtype=dt)
           for mu in generate_partitions(k - 1)}
    for lam in generate_partitions(k):
        d = dim(lam)
        ch, bo = children(lam), block_offsets(lam)
        for i in range(k):
            w = cur[lam]
            for j in range(i, k - 1):                  # rho(c_i)^T = s_{k-2}...s_i
                w = apply_transposition(lam, j, w)
            for mu, o in zip(ch, bo):
                dm = dim(mu)
                out[mu][:, :, i * ninst:(i + 1) * ninst] += (d / (k * dm)) * w[o:o + dm, o:o + dm, :]
    return out


def sn_fft_packed(f: np.ndarray, n: int) -> dict:
    """Forward transform.  f: (B, n!) -> dict lambda -> (B, d, d)."""
    f = np.asarray(f)
    assert f.ndim == 2 and f.shape[1] == math.factorial(n)
    bsz = f.shape[0]
    g = gather_index(n)
    # level-1 data: (1, 1, P_1 * B) with instance index p * B + b
    cur = {(1,): np.ascontiguousarray(f[:, g].T).reshape(1, 1, -1)}
    ninst = f.shape[1] * bsz
    for k in range(2, n + 1):
        ninst //= k
        cur = _level_forward(k, cur, ninst)
    return {lam: np.ascontiguousarray(np.moveaxis(v, 2, 0)) for lam, v in cur.items()}


def sn_ifft_packed(ft: dict, n: int) -> np.ndarray:
    """Inverse transform.  dict lambda -> (B, d, d) -> (B, n!)."""
    parts = generate_partitions(n)
    bsz = np.asarray(ft[parts[0]]).reshape(-1, 1, 1).shape[0] if dim(parts[0]) == 1 else None
    cur = {lam: np.moveaxis(np.asarray(ft[lam]).reshape(bsz, dim(lam), dim(lam)), 0, 2)
           for lam in parts}
    ninst = bsz
    for k in range(n, 1, -1):
        cur = _level_inverse(k, cur, ninst)
        ninst *= k
    x1 = cur[(1,)].reshape(-1, bsz)                    # (P_1, B)
    out = np.zeros((bsz, math.factorial(n)), dtype=x1.dtype)
    out[:, gather_index(n)] = x1.T
    return out


# ----------------------------------------------------------------------------
# dense ("slow") transforms, for cross-checks at small n
# ----------------------------------------------------------------------------

@lru_cache(maxsize=None)
def all_permutations(n: int) -> np.ndarray:
    """(n!, n) lexicographic table (utils.py:61-63)."""
    return np.array(list(_iter_permutations(range(n))), dtype=np.int64)


def chain_digits(sigma) -> list:
    """(i_n, ..., i_2): position of the largest remaining value, repeatedly."""
    s = list(sigma)
    out = []
    for k in range(len(s), 1, -1):
        i = s.index(k - 1)
        out.append(i)
        s.pop(i)
    return out


def rho(part: tuple, sigma) -> np.ndarray:
    """Dense rho_part(sigma), built as rho(c^(n)_{i_n}) rho(c^(n-1)_{i_{n-1}}) ...
    (sigma = c_{i_n} * tau, reference product permutations.py:85-88)."""
    n = sum(part)
    m = np.eye(dim(part))
    for k, i in zip(range(n, 1, -1), chain_digits(sigma)):
        # c^(k)_i = s_i ... s_{k-2} acting inside S_k <= S_n
        for j in range(i, k - 1):
            m = m @ transposition_matrix(part, j)
    return m


def slow_sn_ft_packed(f: np.ndarray, n: int) -> dict:
    """Dense forward transform sum_g f(g) rho(g) (fourier.py:82-108)."""
    f = np.asarray(f)
    perms = all_permutations(n)
    out = {}
    for lam in generate_partitions(n):
        mats = np.stack([rho(lam, s) for s in perms])        # (n!, d, d)
        out[lam] = np.einsum('bi,ijk->bjk', f, mats)
    return out


def slow_sn_ift_packed(ft: dict, n: int) -> np.ndarray:
    """Dense inverse (1/n!) sum_lambda d_lambda <F, rho(g)> (fourier.py:134-159,302-327)."""
    perms = all_permutations(n)
    tot = None
    for lam in generate_partitions(n):
        mats = np.stack([rho(lam, s) for s in perms])
        d = dim(lam)
        term = d * np.einsum('bij,gij->bg', np.asarray(ft[lam]).reshape(-1, d, d), mats)
        tot = term if tot is None else tot + term
    return tot / math.factorial(n)


# ----------------------------------------------------------------------------
# the same level steps with the inner loops in C (oracle/sn_oracle_c.c); the
# tables still come from this file.  Used for the CPU baseline of bench.py and
# for oracle checks at large n; pinned to the numpy path by tests/test_oracle_c.py.
# ----------------------------------------------------------------------------
_clib = None


def c_library():
    global _clib
    if _clib is None:
        import ctypes
        from oracle import build_oracle
        _clib = ctypes.CDLL(build_oracle.build())
    return _clib


@lru_cache(maxsize=None)
def _c_tables(lam: tuple):
    k = sum(lam)
    p = np.ascontiguousarray(np.stack([transposition_table(lam, j)[0] for j in range(k - 1)]).astype(np.int32))
    a = np.ascontiguousarray(np.stack([transposition_table(lam, j)[1] for j in range(k - 1)]))
    b = np.ascontiguousarray(np.stack([transposition_table(lam, j)[2] for j in range(k - 1)]))
    bo = np.array(block_offsets(lam), dtype=np.int32)
    dm = np.array([dim(c) for c in children(lam)], dtype=np.int32)
    return p, a, b, bo, dm


def _c_level(k: int, data: dict, ninst: int, forward: bool) -> dict:
    import ctypes
    lib = c_library()
    dt = next(iter(data.values())).dtype
    suf = "f32" if dt == np.float32 else "f64"
    fn = getattr(lib, ("snft_level_forward_" if forward else "snft_level_inverse_") + suf)
    vp = ctypes.c_void_p
    if forward:
        out = {}
    else:
        out = {mu: np.zeros((dim(mu), dim(mu), k * ninst), dtype=dt) for mu in generate_partitions(k - 1)}
    for lam in generate_partitions(k):
        p, a, b, bo, dm = _c_tables(lam)
        d = dim(lam)
        ch = children(lam)
        if forward:
            ptrs = (vp * len(ch))(*[data[mu].ctypes.data for mu in ch])
            res = np.empty((d, d, ninst), dtype=dt)
            rc = fn(k, d, len(ch), vp(bo.ctypes.data), vp(dm.ctypes.data), ptrs, ctypes.c_int64(ninst),
                    vp(p.ctypes.data), vp(a.ctypes.data), vp(b.ctypes.data), vp(res.ctypes.data))
            out[lam] = res
        else:
            ptrs = (vp * len(ch))(*[out[mu].ctypes.data for mu in ch])
            cur = np.ascontiguousarray(data[lam])
            rc = fn(k, d, len(ch), vp(bo.ctypes.data), vp(dm.ctypes.data), ptrs, ctypes.c_int64(ninst),
                    vp(p.ctypes.data), vp(a.ctypes.data), vp(b.ctypes.data), vp(cur.ctypes.data))
        assert rc == 0
    return out


def sn_fft_packed_c(f: np.ndarray, n: int) -> dict:
    """sn_fft_packed with C inner loops (same arithmetic order)."""
    f = np.asarray(f)
    assert f.ndim == 2 and f.shape[1] == math.factorial(n) and f.dtype in (np.float32, np.float64)
    bsz = f.shape[0]
    cur = {(1,): np.ascontiguousarray(f[:, gather_index(n)].T).reshape(1, 1, -1)}
    ninst = f.shape[1] * bsz
    for k in range(2, n + 1):
        ninst //= k
        cur = _c_level(k, cur, ninst, True)
    return {lam: np.ascontiguousarray(np.moveaxis(v, 2, 0)) for lam, v in cur.items()}


def sn_ifft_packed_c(ft: dict, n: int) -> np.ndarray:
    parts = generate_partitions(n)
    bsz = np.asarray(ft[parts[0]]).size
    cur = {lam: np.ascontiguousarray(np.moveaxis(np.asarray(ft[lam]).reshape(bsz, dim(lam), dim(lam)), 0, 2))
           for lam in parts}
    ninst = bsz
    for k in range(n, 1, -1):
        cur = _c_level(k, cur, ninst, False)
        ninst *= k
    x1 = cur[(1,)].reshape(-1, bsz)
    out = np.zeros((bsz, math.factorial(n)), dtype=x1.dtype)
    out[:, gather_index(n)] = x1.T
    return out
