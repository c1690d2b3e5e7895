// Device kernels of the S_n FFT (sm_100a).  Included by snfft.cu only.
//
// Data model:
//   level-k array X_k : for every partition mu of k (reference order) a
//   [d_mu][d_mu][NINST_k] block at element offset coef_off(mu) * NINST_k; the
//   instance index is the fastest dimension.  An instance of level k is
//   (chain prefix p_k, batch b): inst_k = p_k * B + b, and the k cosets of
//   S_{k-1} in S_k (fourier.py:63-80) are the planes
//   inst_{k-1} = i * NINST_k + inst_k.
//   X_n is stored in the public packed layout [lambda][B][d][d] instead.
//
// One forward level step (fourier.py:253-273) for output column c of block b
// of F_lambda:
//     F_lambda[:, boff_b + c] = sum_i  rho(s_i) ... rho(s_{k-2}) . embed_b( F^(i)_mu[:, c] )
// Every thread owns one column-instance (c, inst); the d_lambda-vector lives
// in shared memory as work[row * W + lane]; a sweep rho(s_j) is a list of
// disjoint 2x2 Young's-orthogonal-form rotations (irreps.py:91-109) applied
// in place.  G row-groups share one column tile and split the rotations.
#pragma once
#include <stdint.h>

namespace snfft {

struct FwdTask {          // one (lambda, child block b) pair of level k
    long long lam_off;    // coef_off(lambda)
    long long mu_off;     // coef_off(mu_b) in level k-1
    int d_lam, d_mu, boff, lam;
};
struct OpRange {          // sweep rho_lambda(s_j): index lam * (k-1) + j
    int pair_off, npairs, neg_off, nneg;
};
struct InvTask {          // one mu of level k-1
    long long mu_off;
    int d_mu, par_off, npar, pad;
};
struct InvParent {        // lambda covering mu
    long long lam_off;
    int d_lam, boff, lam, pad;
    double scale;         // d_lambda / (k d_mu), fourier.py:191-199
};

template <typename T>
struct LevelArgs {
    const T* in;
    T* out;
    const FwdTask* ftasks;
    const InvTask* itasks;
    const InvParent* parents;
    const OpRange* ops;
    const uint32_t* pair_rows;    // t1 | t2 << 16
    const T* pair_coef;           // (a, b) per pair:  x1' = a x1 + b x2 ; x2' = b x1 - a x2
    const uint32_t* neg_rows;     // rows with rho[t][t] = -1
    int k;
    int cw, ci;                   // column x instance tile of the packed-layout level (cw * ci == W)
    int gstride;                  // elements of shared memory per column group
    long long ninst;              // NINST_k
    long long batch;              // B
};

struct IndexArgs {
    int n;
    long long nfact;
    long long batch;
    long long fact[21];
};

// ---------------------------------------------------------------------------
// integer helpers (bit exact)
// ---------------------------------------------------------------------------

// r-th permutation of range(n) in lexicographic order (utils.py:61-63).
__host__ __device__ inline void unrank_perm(int n, long long r, const long long* fact, int* sig) {
    int vals[SNFFT_MAX_N];
    for (int i = 0; i < n; ++i) vals[i] = i;
    for (int i = 0; i < n; ++i) {
        const long long f = fact[n - 1 - i];
        const int dgt = (int)(r / f);
        r -= dgt * f;
        sig[i] = vals[dgt];
        for (int q = dgt; q + 1 < n - i; ++q) vals[q] = vals[q + 1];
    }
}

// Lexicographic rank (permutations.py:187-192).
__host__ __device__ inline long long rank_perm(int n, const int* sig, const long long* fact) {
    long long r = 0;
    for (int i = 0; i < n; ++i) {
        int smaller = 0;
        for (int q = i + 1; q < n; ++q) smaller += sig[q] < sig[i];
        r += smaller * fact[n - 1 - i];
    }
    return r;
}

// Level-1 instance number of the permutation with rank r: i_k = position of
// value k-1 once the larger values are deleted (the coset of fourier.py:77-80
// at level k), p_1 = sum_k i_k * n!/k!.
__host__ __device__ inline long long chain_index(int n, long long r, const long long* fact) {
    int sig[SNFFT_MAX_N];
    unrank_perm(n, r, fact, sig);
    long long p1 = 0;
    for (int k = n; k >= 2; --k) {
        const int v = k - 1;
        int cnt = 0;
        for (int p = 0; p < n; ++p) {
            if (sig[p] == v) break;
            cnt += sig[p] < v;
        }
        p1 += cnt * (fact[n] / fact[k]);
    }
    return p1;
}

// ---------------------------------------------------------------------------
// sweep: work <- rho_lambda(s_j) work   (in place, rotations are disjoint)
// ---------------------------------------------------------------------------
template <typename T, int W, int G>
__device__ __forceinline__ void sweep(T* __restrict__ work, const LevelArgs<T>& a, int lam, int j,
                                      int rg, int lane) {
    const OpRange r = a.ops[lam * (a.k - 1) + j];
    const uint32_t* __restrict__ rows = a.pair_rows + r.pair_off;
    const T* __restrict__ coef = a.pair_coef + 2 * (size_t)r.pair_off;
    int p = rg;
    // two rotations per trip: independent loads first, then the math
    for (; p + G < r.npairs; p += 2 * G) {
        const uint32_t ra = rows[p], rb = rows[p + G];
        const T a0 = coef[2 * p], b0 = coef[2 * p + 1];
        const T a1 = coef[2 * (p + G)], b1 = coef[2 * (p + G) + 1];
        T* pa1 = work + (ra & 0xffffu) * W + lane;
        T* pa2 = work + (ra >> 16) * W + lane;
        T* pb1 = work + (rb & 0xffffu) * W + lane;
        T* pb2 = work + (rb >> 16) * W + lane;
        const T x1 = *pa1, x2 = *pa2, y1 = *pb1, y2 = *pb2;
        *pa1 = a0 * x1 + b0 * x2;
        *pa2 = b0 * x1 - a0 * x2;
        *pb1 = a1 * y1 + b1 * y2;
        *pb2 = b1 * y1 - a1 * y2;
    }
    for (; p < r.npairs; p += G) {
        const uint32_t ra = rows[p];
        const T a0 = coef[2 * p], b0 = coef[2 * p + 1];
        T* pa1 = work + (ra & 0xffffu) * W + lane;
        T* pa2 = work + (ra >> 16) * W + lane;
        const T x1 = *pa1, x2 = *pa2;
        *pa1 = a0 * x1 + b0 * x2;
        *pa2 = b0 * x1 - a0 * x2;
    }
    const uint32_t* __restrict__ neg = a.neg_rows + r.neg_off;
    for (int q = rg; q < r.nneg; q += G) {
        T* px = work + neg[q] * W + lane;
        *px = -*px;
    }
}

template <int G>
__device__ __forceinline__ void group_sync() {
    if (G > 1) __syncthreads();
}

// ---------------------------------------------------------------------------
// forward level step S_{k-1} -> S_k      grid (chunks, tasks)
// ---------------------------------------------------------------------------
template <typename T, int W, int G, int RPT, int NG, bool FINAL>
__global__ void __launch_bounds__(W* G* NG) fwd_level_kernel(const LevelArgs<T> a) {
    SNFFT_DYN_SMEM(smem_raw);
    static_assert(G == 1 || NG == 1, "row groups need the whole CTA barrier");
    const FwdTask tk = a.ftasks[blockIdx.y];
    const int tid = threadIdx.x;
    const int grp = tid / (W * G), lt = tid % (W * G), rg = lt / W, lane = lt % W;
    const long long chunk = (long long)blockIdx.x * NG + grp;
    const int k = a.k;
    const long long ninst = a.ninst, ninst_prev = a.ninst * k;

    long long c_in, inst_in, nchunks;
    long long cx = 0, bx = 0;
    if (FINAL) {
        const long long ncx = (tk.d_mu + a.cw - 1) / a.cw, nbx = (a.batch + a.ci - 1) / a.ci;
        nchunks = ncx * nbx;
        cx = chunk % ncx;
        bx = chunk / ncx;
        c_in = cx * a.cw + lane / a.ci;          // load mapping: instance fastest
        inst_in = bx * a.ci + lane % a.ci;
    } else {
        const long long qn = (long long)tk.d_mu * ninst;
        nchunks = (qn + W - 1) / W;
        const long long q = chunk * W + lane;
        c_in = q / ninst;
        inst_in = q % ninst;
    }
    if ((long long)blockIdx.x * NG >= nchunks) return;          // whole CTA idle
    const bool active = chunk < nchunks && c_in < tk.d_mu && inst_in < ninst;

    T* __restrict__ work = reinterpret_cast<T*>(smem_raw) + (size_t)grp * a.gstride;
    const T* __restrict__ src = a.in + (tk.mu_off + c_in) * ninst_prev + inst_in;
    const long long rstride = (long long)tk.d_mu * ninst_prev;

    T acc[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) acc[q] = T(0);

    for (int i = 0; i < k; ++i) {
        // work <- embed_b( F^(i)_mu[:, c] )
        for (int t = rg; t < tk.d_lam; t += G) {
            const int rr = t - tk.boff;
            T v = T(0);
            if (active && rr >= 0 && rr < tk.d_mu) v = src[rr * rstride + i * ninst];
            work[t * W + lane] = v;
        }
        group_sync<G>();
        for (int j = k - 2; j >= i; --j) {       // rho(s_i) ... rho(s_{k-2}) . work
            sweep<T, W, G>(work, a, tk.lam, j, rg, lane);
            group_sync<G>();
        }
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int t = rg + q * G;
            if (t < tk.d_lam) acc[q] += work[t * W + lane];
        }
        group_sync<G>();
    }

    if (FINAL) {
        // transpose the tile through shared memory so that stores run along columns
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int t = rg + q * G;
            if (t < tk.d_lam) work[t * W + lane] = acc[q];
        }
        __syncthreads();
        const int c_rel = lane % a.cw, i_rel = lane / a.cw;
        const long long c_out = cx * a.cw + c_rel, inst_out = bx * a.ci + i_rel;
        const int slot = c_rel * a.ci + i_rel;
        if (chunk < nchunks && c_out < tk.d_mu && inst_out < a.batch) {
            T* __restrict__ dst = a.out + tk.lam_off * a.batch +
                                  inst_out * ((long long)tk.d_lam * tk.d_lam) + tk.boff + c_out;
            for (int t = rg; t < tk.d_lam; t += G) dst[(long long)t * tk.d_lam] = work[t * W + slot];
        }
    } else if (active) {
        T* __restrict__ dst = a.out + (tk.lam_off + tk.boff + c_in) * ninst + inst_in;
        const long long ostride = (long long)tk.d_lam * ninst;
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int t = rg + q * G;
            if (t < tk.d_lam) dst[t * ostride] = acc[q];
        }
    }
}

// ---------------------------------------------------------------------------
// inverse (adjoint) level step S_k -> S_{k-1}     grid (chunks, tasks, cosets)
//   G^(i)_mu[:, c] = sum_{lambda covers mu} d_lambda/(k d_mu) *
//                    [ rho(s_{k-2}) ... rho(s_i) F_lambda[:, boff + c] ]_{rows of block mu}
// ---------------------------------------------------------------------------
template <typename T, int W, int G, int RPT, int NG, bool FIRST>
__global__ void __launch_bounds__(W* G* NG) inv_level_kernel(const LevelArgs<T> a) {
    SNFFT_DYN_SMEM(smem_raw);
    static_assert(G == 1 || NG == 1, "row groups need the whole CTA barrier");
    const InvTask tk = a.itasks[blockIdx.y];
    const int i = blockIdx.z;
    const int tid = threadIdx.x;
    const int grp = tid / (W * G), lt = tid % (W * G), rg = lt / W, lane = lt % W;
    const long long chunk = (long long)blockIdx.x * NG + grp;
    const int k = a.k;
    const long long ninst = a.ninst, ninst_prev = a.ninst * k;

    long long c_in, inst_in, nchunks;
    long long cx = 0, bx = 0;
    if (FIRST) {
        const long long ncx = (tk.d_mu + a.cw - 1) / a.cw, nbx = (a.batch + a.ci - 1) / a.ci;
        nchunks = ncx * nbx;
        cx = chunk % ncx;
        bx = chunk / ncx;
        c_in = cx * a.cw + lane % a.cw;          // load mapping: column fastest
        inst_in = bx * a.ci + lane / a.cw;
    } else {
        const long long qn = (long long)tk.d_mu * ninst;
        nchunks = (qn + W - 1) / W;
        const long long q = chunk * W + lane;
        c_in = q / ninst;
        inst_in = q % ninst;
    }
    if ((long long)blockIdx.x * NG >= nchunks) return;
    const bool active = chunk < nchunks && c_in < tk.d_mu && inst_in < ninst;

    T* __restrict__ work = reinterpret_cast<T*>(smem_raw) + (size_t)grp * a.gstride;

    T acc[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) acc[q] = T(0);

    for (int p = 0; p < tk.npar; ++p) {
        const InvParent par = a.parents[tk.par_off + p];
        const T* __restrict__ src;
        long long rstride;
        if (FIRST) {
            src = a.in + par.lam_off * a.batch + inst_in * ((long long)par.d_lam * par.d_lam) +
                  par.boff + c_in;
            rstride = par.d_lam;
        } else {
            src = a.in + (par.lam_off + par.boff + c_in) * ninst + inst_in;
            rstride = (long long)par.d_lam * ninst;
        }
        for (int t = rg; t < par.d_lam; t += G) work[t * W + lane] = active ? src[t * rstride] : T(0);
        group_sync<G>();
        for (int j = i; j <= k - 2; ++j) {       // rho(c_i)^T = rho(s_{k-2}) ... rho(s_i)
            sweep<T, W, G>(work, a, par.lam, j, rg, lane);
            group_sync<G>();
        }
        const T scale = (T)par.scale;
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int t = rg + q * G;
            if (t < tk.d_mu) acc[q] += scale * work[(par.boff + t) * W + lane];
        }
        group_sync<G>();
    }

    if (FIRST) {
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int t = rg + q * G;
            if (t < tk.d_mu) work[t * W + lane] = acc[q];
        }
        __syncthreads();
        const int i_rel = lane % a.ci, c_rel = lane / a.ci;     // store mapping: instance fastest
        const long long c_out = cx * a.cw + c_rel, inst_out = bx * a.ci + i_rel;
        const int slot = i_rel * a.cw + c_rel;
        if (chunk < nchunks && c_out < tk.d_mu && inst_out < a.batch) {
            T* __restrict__ dst = a.out + (tk.mu_off + c_out) * ninst_prev + i * ninst + inst_out;
            const long long ostride = (long long)tk.d_mu * ninst_prev;
            for (int t = rg; t < tk.d_mu; t += G) dst[t * ostride] = work[t * W + slot];
        }
    } else if (active) {
        T* __restrict__ dst = a.out + (tk.mu_off + c_in) * ninst_prev + i * ninst + inst_in;
        const long long ostride = (long long)tk.d_mu * ninst_prev;
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int t = rg + q * G;
            if (t < tk.d_mu) dst[t * ostride] = acc[q];
        }
    }
}

// ---------------------------------------------------------------------------
// input gather / output scatter: the composition of restrict_to_coset over all
// levels (fourier.py:253) as one tiled transpose.  grid (ranks/64, B/32), 256 thr
//   GATHER :  x1[p1(r) * B + b] = f[b * n! + r]
//   SCATTER:  f[b * n! + r]     = x1[p1(r) * B + b]
// ---------------------------------------------------------------------------
constexpr int GS_TR = 64, GS_TB = 32, GS_THREADS = 256;

template <typename T>
__host__ __device__ constexpr size_t gather_smem_bytes() {
    return GS_TR * sizeof(long long) + (size_t)GS_TB * (GS_TR + 1) * sizeof(T);
}

template <typename T, bool SCATTER>
__global__ void __launch_bounds__(GS_THREADS) gather_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                            const IndexArgs ia) {
    SNFFT_DYN_SMEM(smem_raw);
    long long* p1 = reinterpret_cast<long long*>(smem_raw);
    T* tile = reinterpret_cast<T*>(smem_raw + GS_TR * sizeof(long long));
    const long long r0 = (long long)blockIdx.x * GS_TR, b0 = (long long)blockIdx.y * GS_TB;
    const int tid = threadIdx.x;
    if (tid < GS_TR) {
        const long long r = r0 + tid;
        p1[tid] = r < ia.nfact ? chain_index(ia.n, r, ia.fact) : -1;
    }
    if (SCATTER) __syncthreads();
    if (!SCATTER) {
        for (int e = tid; e < GS_TB * GS_TR; e += GS_THREADS) {
            const int bb = e / GS_TR, rr = e % GS_TR;
            if (b0 + bb < ia.batch && r0 + rr < ia.nfact)
                tile[bb * (GS_TR + 1) + rr] = in[(b0 + bb) * ia.nfact + r0 + rr];
        }
    } else {
        for (int e = tid; e < GS_TB * GS_TR; e += GS_THREADS) {
            const int rr = e / GS_TB, bb = e % GS_TB;
            if (b0 + bb < ia.batch && r0 + rr < ia.nfact)
                tile[bb * (GS_TR + 1) + rr] = in[p1[rr] * ia.batch + b0 + bb];
        }
    }
    __syncthreads();
    if (!SCATTER) {
        for (int e = tid; e < GS_TB * GS_TR; e += GS_THREADS) {
            const int rr = e / GS_TB, bb = e % GS_TB;
            if (b0 + bb < ia.batch && r0 + rr < ia.nfact)
                out[p1[rr] * ia.batch + b0 + bb] = tile[bb * (GS_TR + 1) + rr];
        }
    } else {
        for (int e = tid; e < GS_TB * GS_TR; e += GS_THREADS) {
            const int bb = e / GS_TR, rr = e % GS_TR;
            if (b0 + bb < ia.batch && r0 + rr < ia.nfact)
                out[(b0 + bb) * ia.nfact + r0 + rr] = tile[bb * (GS_TR + 1) + rr];
        }
    }
}

// ---------------------------------------------------------------------------
// index kernels (bit exact)
// ---------------------------------------------------------------------------
__global__ void perm_rank_kernel(const int8_t* __restrict__ perms, long long m, long long* __restrict__ ranks,
                                 const IndexArgs ia) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    int sig[SNFFT_MAX_N];
    for (int i = 0; i < ia.n; ++i) sig[i] = perms[e * ia.n + i];
    ranks[e] = rank_perm(ia.n, sig, ia.fact);
}

__global__ void perm_unrank_kernel(const long long* __restrict__ ranks, long long m, int8_t* __restrict__ perms,
                                   const IndexArgs ia) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    int sig[SNFFT_MAX_N];
    unrank_perm(ia.n, ranks[e], ia.fact, sig);
    for (int i = 0; i < ia.n; ++i) perms[e * ia.n + i] = (int8_t)sig[i];
}

// idx[r] = rank in S_n of the element of coset `coset` whose S_{n-1} rank is r
// (closed form of argwhere(sn_perms[:, coset] == n-1), fourier.py:79)
__global__ void coset_index_kernel(int coset, long long* __restrict__ idx, const IndexArgs ia) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = ia.n;
    if (r >= ia.fact[n - 1]) return;
    long long out = (long long)(n - 1 - coset) * ia.fact[n - 1 - coset];
    long long rem = r;
    for (int j = 0; j < n - 1; ++j) {
        const long long f = ia.fact[n - 2 - j];
        const long long lj = rem / f;
        rem -= lj * f;
        const int pos = j < coset ? j : j + 1;
        out += lj * ia.fact[n - 1 - pos];
    }
    idx[r] = out;
}

__global__ void chain_index_kernel(long long* __restrict__ p1, const IndexArgs ia) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ia.nfact) return;
    p1[r] = chain_index(ia.n, r, ia.fact);
}


// ---------------------------------------------------------------------------
// dense representation matrices rho_lambda(sigma) for every sigma in S_n
// (SnIrrep.matrix_tensor, irreps.py:122-148), used by slow_sn_ft / slow_sn_ift.
// sigma = c^(n)_{i_n} * ... * c^(2)_{i_2}  =>  rho(sigma) e_c is obtained by
// applying the words s_i ... s_{k-2} for k = 2..n to the unit vector e_c.
// One lane per (sigma, column); block of 32 lanes, column vectors in shared memory.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(32) dense_rep_kernel(const LevelArgs<T> a, int lam, int d, T* __restrict__ out,
                                                         const IndexArgs ia) {
    SNFFT_DYN_SMEM(smem_raw);
    T* work = reinterpret_cast<T*>(smem_raw);
    const int lane = threadIdx.x;
    const long long q = (long long)blockIdx.x * 32 + lane;
    const long long rank = q / d;
    const int c = (int)(q % d);
    if (rank >= ia.nfact) return;
    for (int t = 0; t < d; ++t) work[t * 32 + lane] = (t == c) ? T(1) : T(0);
    int sig[SNFFT_MAX_N];
    unrank_perm(ia.n, rank, ia.fact, sig);
    for (int k = 2; k <= ia.n; ++k) {
        const int v = k - 1;
        int i = 0;
        for (int p = 0; p < ia.n; ++p) {
            if (sig[p] == v) break;
            i += sig[p] < v;
        }
        for (int j = k - 2; j >= i; --j) sweep<T, 32, 1>(work, a, lam, j, 0, lane);
    }
    for (int t = 0; t < d; ++t) out[(rank * d + t) * d + c] = work[t * 32 + lane];
}

}  // namespace snfft
