// Host combinatorics of the S_n FFT plan.  Pure C++ (no CUDA).
//
// Everything here reproduces a convention of the reference
// (/root/reference/src/algebraist) that the transform's output depends on:
//   partition order          tableau.py:103-123  (dict key order of sn_fft)
//   children, ascending      irreps.py:64-68, tableau.py:38-56
//   dimension                tableau.py:236-265
//   basis (tableau id) order tableau.py:171-198  => basis(lambda) is the
//                            concatenation of the children's bases, so
//                            rho_lambda(s_j), j <= k-3, is block diagonal
//   YOR coefficients         irreps.py:91-109, tableau.py:214-220
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace snfft {

using Part = std::vector<int>;

int hook_dim(const Part& rows) {
    int n = 0;
    for (int r : rows) n += r;
    if (n <= 1) return 1;
    // n! / prod(hooks) computed with exact integer arithmetic: multiply the
    // factorial terms in and divide hooks out greedily (n <= 20 fits in u64
    // when interleaved; use long double fallback-free rational reduction).
    std::vector<int64_t> num, den;
    for (int i = 2; i <= n; ++i) num.push_back(i);
    for (size_t i = 0; i < rows.size(); ++i)
        for (int j = 0; j < rows[i]; ++j) {
            int below = 0;
            for (size_t r = i + 1; r < rows.size(); ++r)
                if (rows[r] > j) ++below;
            den.push_back(rows[i] - j + below);
        }
    // cancel common factors
    for (auto& d : den) {
        for (auto& m : num) {
            if (d == 1) break;
            int64_t g = std::__gcd(d, m);
            if (g > 1) { d /= g; m /= g; }
        }
        if (d != 1) throw std::logic_error("hook length did not divide n!");
    }
    int64_t r = 1;
    for (auto m : num) r *= m;
    return (int)r;
}

std::vector<Part> children_of(const Part& rows) {
    std::vector<Part> out;
    for (size_t i = 0; i < rows.size(); ++i) {
        int nxt = (i + 1 < rows.size()) ? rows[i + 1] : 0;
        if (rows[i] > nxt) {
            Part c = rows;
            c[i] -= 1;
            if (c.back() == 0) c.pop_back();
            out.push_back(c);
        }
    }
    std::sort(out.begin(), out.end());   // ascending tuple order == sorted() in irreps.py:68
    return out;
}

// generate_partitions(n), tableau.py:103-123 (hard coded lists for n <= 5,
// then [(n,)] + sorted((k,*p)) for k = 1..n-1, p in generate_partitions(n-k),
// first occurrence kept).
static const std::vector<Part>& partitions_of(int n, std::vector<std::vector<Part>>& memo) {
    if (!memo[n].empty() || n == 0) return memo[n];
    static const std::vector<std::vector<Part>> small = {
        {},
        {{1}},
        {{2}, {1, 1}},
        {{3}, {2, 1}, {1, 1, 1}},
        {{4}, {3, 1}, {2, 2}, {2, 1, 1}, {1, 1, 1, 1}},
        {{5}, {4, 1}, {3, 2}, {3, 1, 1}, {2, 2, 1}, {2, 1, 1, 1}, {1, 1, 1, 1, 1}},
    };
    if (n <= 5) {
        memo[n] = small[n];
        return memo[n];
    }
    std::vector<Part> out;
    out.push_back(Part{n});
    for (int k = 1; k < n; ++k) {
        for (const Part& p : partitions_of(n - k, memo)) {
            Part q = p;
            q.push_back(k);
            std::sort(q.begin(), q.end(), std::greater<int>());
            if (std::find(out.begin(), out.end(), q) == out.end()) out.push_back(q);
        }
    }
    memo[n] = out;
    return memo[n];
}

// (row, col) of the single box in `big` that is not in `small`.
static void removed_box(const Part& big, const Part& small, int& r, int& c) {
    for (size_t i = 0; i < big.size(); ++i) {
        int s = i < small.size() ? small[i] : 0;
        if (big[i] != s) { r = (int)i; c = big[i] - 1; return; }
    }
    throw std::logic_error("removed_box: not a child");
}

void build_host_plan(int n, HostPlan& plan) {
    if (n < 1 || n > 20) throw std::invalid_argument("n out of range");
    plan.n = n;
    plan.factorial[0] = 1;
    for (int i = 1; i <= 20; ++i) plan.factorial[i] = plan.factorial[i - 1] * i;
    plan.levels.assign(n + 1, HostLevel());
    std::vector<std::vector<Part>> memo(n + 1);

    for (int k = 1; k <= n; ++k) {
        HostLevel& L = plan.levels[k];
        L.k = k;
        const auto& parts = partitions_of(k, memo);
        L.shapes.resize(parts.size());
        int64_t off = 0;
        for (size_t s = 0; s < parts.size(); ++s) {
            HostShape& S = L.shapes[s];
            S.rows = parts[s];
            S.k = k;
            S.dim = hook_dim(parts[s]);
            S.coef_off = off;
            off += (int64_t)S.dim * S.dim;
            L.index[parts[s]] = (int)s;
            L.max_dim = std::max(L.max_dim, S.dim);
        }
        if (off != plan.factorial[k]) throw std::logic_error("sum d^2 != k!");
        L.partitions_flat.assign(parts.size() * k, 0);
        L.dims.resize(parts.size());
        L.coef_offsets.resize(parts.size() + 1);
        for (size_t s = 0; s < parts.size(); ++s) {
            for (size_t i = 0; i < parts[s].size(); ++i) L.partitions_flat[s * k + i] = parts[s][i];
            L.dims[s] = L.shapes[s].dim;
            L.coef_offsets[s] = L.shapes[s].coef_off;
        }
        L.coef_offsets[parts.size()] = off;
        if (k == 1) continue;

        HostLevel& P = plan.levels[k - 1];
        // Young's lattice links and block offsets
        for (size_t s = 0; s < parts.size(); ++s) {
            HostShape& S = L.shapes[s];
            int bo = 0;
            int slot = 0;
            for (const Part& c : children_of(S.rows)) {
                int id = P.index.at(c);
                S.child.push_back(id);
                S.boff.push_back(bo);
                bo += P.shapes[id].dim;
                P.shapes[id].parent.push_back((int)s);
                P.shapes[id].parent_block.push_back(slot++);
            }
            if (bo != S.dim) throw std::logic_error("branching rule violated");
        }
        // YOR tables
        for (size_t s = 0; s < parts.size(); ++s) {
            HostShape& S = L.shapes[s];
            const int d = S.dim;
            S.partner.assign(k - 1, std::vector<int32_t>(d));
            S.diag.assign(k - 1, std::vector<double>(d, 0.0));
            S.off.assign(k - 1, std::vector<double>(d, 0.0));
            S.dist.assign(k - 1, std::vector<int8_t>(d, 0));
            // j <= k-3: block-diagonal stack of the children's tables
            for (int j = 0; j + 3 <= k; ++j) {
                for (size_t b = 0; b < S.child.size(); ++b) {
                    const HostShape& C = P.shapes[S.child[b]];
                    for (int t = 0; t < C.dim; ++t) {
                        S.partner[j][S.boff[b] + t] = C.partner[j][t] + S.boff[b];
                        S.diag[j][S.boff[b] + t] = C.diag[j][t];
                        S.off[j][S.boff[b] + t] = C.off[j][t];
                        S.dist[j][S.boff[b] + t] = C.dist[j][t];
                    }
                }
            }
            // j = k-2: the new transposition (k-2 k-1).  Row t = (mu, nu, rest):
            // entry k-1 in box lambda/mu, entry k-2 in box mu/nu.
            const int j = k - 2;
            for (int t = 0; t < d; ++t) S.partner[j][t] = t;
            if (k == 2) {
                // (2): same row, d = +1 ; (1,1): same column, d = -1
                S.diag[j][0] = (S.rows.size() == 1) ? 1.0 : -1.0;
                S.dist[j][0] = (S.rows.size() == 1) ? 1 : -1;
                continue;
            }
            const HostLevel& Q = plan.levels[k - 2];
            for (size_t b = 0; b < S.child.size(); ++b) {
                const HostShape& M = P.shapes[S.child[b]];
                int rb, cb;
                removed_box(S.rows, M.rows, rb, cb);
                for (size_t c = 0; c < M.child.size(); ++c) {
                    const HostShape& N = Q.shapes[M.child[c]];
                    int ra, ca;
                    removed_box(M.rows, N.rows, ra, ca);
                    const int dist = (ra - rb) + (cb - ca);     // tableau.py:214-220
                    const int lo = S.boff[b] + M.boff[c];
                    for (int t = 0; t < N.dim; ++t) {
                        S.diag[j][lo + t] = 1.0 / dist;
                        S.dist[j][lo + t] = (int8_t)dist;
                    }
                    if (std::abs(dist) > 1) {
                        // swapped tableau: k-1 in (ra,ca), k-2 in (rb,cb)
                        Part m2 = S.rows;
                        m2[ra] -= 1;
                        if (m2.back() == 0) m2.pop_back();
                        const int id2 = P.index.at(m2);
                        const HostShape& M2 = P.shapes[id2];
                        size_t b2 = 0, c2 = 0;
                        while (S.child[b2] != id2) ++b2;
                        while (M2.child[c2] != M.child[c]) ++c2;
                        const int lo2 = S.boff[b2] + M2.boff[c2];
                        const double o = std::sqrt(1.0 - 1.0 / ((double)dist * dist));
                        for (int t = 0; t < N.dim; ++t) {
                            S.partner[j][lo + t] = lo2 + t;
                            S.off[j][lo + t] = o;
                        }
                    }
                }
            }
        }
    }
}

}  // namespace snfft
