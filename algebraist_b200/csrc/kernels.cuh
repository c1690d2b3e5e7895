// Device kernels of the S_n FFT (sm_100a).  Included by snfft.cu only.
//
// Data model:
//   level-k array X_k : for every partition mu of k (reference order) a
//   [d_mu][d_mu][NINST_k] block at element offset coef_off(mu) * NINST_k; the
//   instance index is the fastest dimension.  An instance of level k is
//   (chain prefix p_k, batch b): inst_k = p_k * B + b, and the k cosets of
//   S_{k-1} in S_k (fourier.py:63-80) are the planes
//   inst_{k-1} = i * NINST_k + inst_k.
//   (X_n included; pack_kernel converts it to the public packed layout.)
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

#include "base_gen.cuh"
#include "vec2.cuh"

namespace snfft {

// Sweep tables.  Every (lambda, child block b) pair of a level owns one blob of
// 32-bit words in LevelArgs::tables:
//   [0, k-1)        word offset (inside the blob) of the entries of rho_lambda(s_j), j = 0..k-2
//   [k-1, 2(k-1))   npairs_j | nneg_j << 16
//   entries of j    npairs_j rotations  t1 | t2 << 13 | code << 26, then nneg_j rows to negate
// restricted to the rows the column can occupy before step j (see build_level).
// `code` indexes LevelArgs::lut: a = lut[code] = 1/d, b = lut[LUT + code] = sqrt(1 - 1/d^2) (interleaved in shared memory)
// for the signed axial distance d (irreps.py:91-109):  x1' = a x1 + b x2 ; x2' = b x1 - a x2.
constexpr int LUT = 24;

struct FwdTask {          // one (lambda, child block b) pair of level k
    long long lam_off;    // coef_off(lambda)
    long long mu_off;     // coef_off(mu_b) in level k-1
    int d_lam, d_mu, boff;
    int tab_off, tab_words, pad;
};
struct InvTask {          // one mu of level k-1
    long long mu_off;
    int d_mu, par_off, npar, pad;
};
struct InvParent {        // lambda covering mu
    long long lam_off;
    int d_lam, boff, tab_off, tab_words;
    double scale;         // d_lambda / (k d_mu), fourier.py:191-199
};

template <typename T>
struct LevelArgs {
    const T* in;
    T* out;
    const FwdTask* ftasks;
    const InvTask* itasks;
    const InvParent* parents;
    const uint32_t* tables;
    int k;
    int bstride;                  // elements of one work buffer (rows * TC)
    int tab_smem_words;           // shared-memory words reserved for one task table
    long long ninst;              // NINST_k
    long long batch;              // B
    // forward only: the chains (cosets of S_{k-1}) that are present.  The input holds nch planes per
    // coefficient (plane j = coset cid[j], ascending); the missing cosets contribute nothing.  A full
    // step has nch = k, cid[j] = j; the coset-sharded transform passes the cosets one rank owns.
    int nch;
    int cid[SNFFT_MAX_N];
    T lut[2 * LUT];
};

// per-row tables of the full sweeps of level n (dense_rep_kernel only)
template <typename T>
struct RowTables {
    const int* partner;           // [(lam_row_base + j * d) + t]
    const T* diag;
    const T* off;
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
    if (n <= 12) {                       // n! < 2^31: 32-bit divisions are several times cheaper on the GPU
        int vals[SNFFT_MAX_N];
        unsigned rr = (unsigned)r;
        for (int i = 0; i < n; ++i) vals[i] = i;
        for (int i = 0; i < n; ++i) {
            const unsigned f = (unsigned)fact[n - 1 - i];
            const unsigned dgt = rr / f;
            rr -= dgt * f;
            sig[i] = vals[dgt];
            for (int q = (int)dgt; q + 1 < n - i; ++q) vals[q] = vals[q + 1];
        }
    } else {
        unrank_perm(n, r, fact, sig);
    }
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
// sweep: work <- rho_lambda(s_j) work on the first `nact` work buffers
// (in place; the rotations of one sweep touch disjoint rows).  Thread (rg, lane)
// owns V adjacent columns and every G-th rotation of the list.
// ---------------------------------------------------------------------------
template <typename T, int V>
struct alignas(sizeof(T) * V) Pack {
    T v[V];
};

template <typename T, int V, int TC, int G, int NBUF>
__device__ __forceinline__ void sweep(T* __restrict__ work, int bstride, int nact, const uint32_t* __restrict__ tab,
                                      const T* __restrict__ lut, int k, int j, int rg, int lane) {
    using P = Pack<T, V>;
    const uint32_t cnt = tab[k - 1 + j];
    const int npairs = (int)(cnt & 0xffffu), nneg = (int)(cnt >> 16);
    const uint32_t* __restrict__ ent = tab + tab[j];
    T* __restrict__ base = work + lane * V;
    for (int p = rg; p < npairs; p += G) {
        const uint32_t e = ent[p];
        const int o1 = (int)(e & 0x1fffu) * TC, o2 = (int)((e >> 13) & 0x1fffu) * TC;
        const Pack<T, 2> cab = *reinterpret_cast<const Pack<T, 2>*>(lut + 2 * (e >> 26));     // interleaved (a, b)
        const T ca = cab.v[0], cb = cab.v[1];
        P x1[NBUF], x2[NBUF];
#pragma unroll
        for (int m = 0; m < NBUF; ++m)
            if (m < nact) {
                x1[m] = *reinterpret_cast<const P*>(base + m * bstride + o1);
                x2[m] = *reinterpret_cast<const P*>(base + m * bstride + o2);
            }
#pragma unroll
        for (int m = 0; m < NBUF; ++m)
            if (m < nact) {
                P y1, y2;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    y1.v[v] = ca * x1[m].v[v] + cb * x2[m].v[v];
                    y2.v[v] = cb * x1[m].v[v] - ca * x2[m].v[v];
                }
                *reinterpret_cast<P*>(base + m * bstride + o1) = y1;
                *reinterpret_cast<P*>(base + m * bstride + o2) = y2;
            }
    }
    for (int q = rg; q < nneg; q += G) {
        const int o = (int)ent[npairs + q] * TC;
#pragma unroll
        for (int m = 0; m < NBUF; ++m)
            if (m < nact) {
                P x = *reinterpret_cast<const P*>(base + m * bstride + o);
#pragma unroll
                for (int v = 0; v < V; ++v) x.v[v] = -x.v[v];
                *reinterpret_cast<P*>(base + m * bstride + o) = x;
            }
    }
}

// Shared memory of a level kernel: [work buffers][lut][task table]
template <typename T>
__host__ __device__ inline size_t level_smem_bytes(int ng, int nbuf, int bstride, int tab_words) {
    return ((size_t)ng * nbuf * bstride + 2 * LUT) * sizeof(T) + (size_t)tab_words * sizeof(uint32_t);
}

// Copy one task table into shared memory (STAGE) or use it in place.
template <bool STAGE>
__device__ __forceinline__ const uint32_t* stage_table(uint32_t* tab_s, const uint32_t* __restrict__ tables,
                                                      int tab_off, int tab_words, int tid, int nthreads) {
    if (!STAGE) return tables + tab_off;
    for (int w = tid; w < tab_words; w += nthreads) tab_s[w] = tables[tab_off + w];
    return tab_s;
}

// 4/8-byte asynchronous global -> shared copy (LDGSTS): the fill of a work buffer
// issues all its copies back to back instead of waiting on each load.
template <typename T>
__device__ __forceinline__ void copy_async(T* dst_smem, const T* src_global) {
    SNFFT_CP_ASYNC(dst_smem, src_global, sizeof(T));
}

template <int G>
__device__ __forceinline__ void group_sync() {
    if (G > 1) __syncthreads();
}

// Column tile bookkeeping shared by the forward and inverse kernels.
// A tile is TC = W * V "slots"; a slot is one (column c, instance) pair of the
// task's child block, enumerated instance-fastest over the flattened (c, inst)
// range.  Threads have two roles:
//   transfer role: moves rows between global and shared memory.  Scalar form:
//                  slot fs = lt % TC, rows fr, fr + FR, ...; vector form (NINST % TV == 0):
//                  TV = 16 / sizeof(T) adjacent slots (= adjacent instances of one
//                  column, 16 aligned bytes) per thread.
//   sweep role   : V slots lane*V.., rows / rotations rg, rg + G, ...
struct SlotMap {
    long long c, inst;      // (column, instance) of this thread's first transfer slot
    bool ok;                // inside the range
};

template <int TC>
__device__ __forceinline__ SlotMap make_slots(long long chunk, int fs, int d_mu, long long ninst, long long* nchunks) {
    SlotMap s;
    const long long qn = (long long)d_mu * ninst;
    *nchunks = (qn + TC - 1) / TC;
    const long long q = chunk * TC + fs;
    s.c = q / ninst;
    s.inst = q % ninst;
    s.ok = q < qn;
    return s;
}

// ---------------------------------------------------------------------------
// forward level step S_{k-1} -> S_k      grid (chunks, tasks)
// ---------------------------------------------------------------------------
template <typename T, int V, int W, int G, int RPT, int NG, int NBUF, int MINB, bool STAGE>
__global__ void __launch_bounds__(W* G* NG, MINB) fwd_level_kernel(const LevelArgs<T> a) {
    SNFFT_DYN_SMEM(smem_raw);
    constexpr int TC = W * V, GT = W * G, FR = GT / TC;
    constexpr int TV = 16 / (int)sizeof(T), NVS = TC / TV, FRV = GT / NVS;    // vector transfer role
    static_assert(G == 1 || NG == 1, "row groups need the whole CTA barrier");
    static_assert(GT % TC == 0 && TC % TV == 0, "transfer role needs whole rows of threads");
    using P = Pack<T, V>;
    using PV = Pack<T, TV>;
    const FwdTask tk = a.ftasks[blockIdx.y];
    const int tid = threadIdx.x;
    const int grp = tid / GT, lt = tid % GT, rg = lt / W, lane = lt % W;
    const long long chunk = (long long)blockIdx.x * NG + grp;
    const int k = a.k, nch = a.nch;
    const long long ninst = a.ninst, ninst_prev = a.ninst * nch;
    // 16-byte transfers need alignment; with G == 1 nothing synchronises the lanes, so every
    // thread may only touch its own slot (scalar role)
    const bool vec = G > 1 && (ninst % TV == 0);
    const int fs = vec ? (lt % NVS) * TV : lt % TC;     // first slot this thread transfers
    const int fr = vec ? lt / NVS : lt / TC, frs = vec ? FRV : FR;

    long long nchunks;
    const SlotMap sm = make_slots<TC>(chunk, fs, tk.d_mu, ninst, &nchunks);
    if ((long long)blockIdx.x * NG >= nchunks) return;          // whole CTA idle

    const int bstride = a.bstride;
    T* __restrict__ work = reinterpret_cast<T*>(smem_raw) + (size_t)grp * NBUF * bstride;
    T* __restrict__ lut = reinterpret_cast<T*>(smem_raw) + (size_t)NG * NBUF * bstride;
    if (tid < 2 * LUT) lut[tid] = a.lut[(tid & 1) * LUT + (tid >> 1)];      // shared copy interleaves (a, b) per code
    const uint32_t* __restrict__ tab = stage_table<STAGE>(reinterpret_cast<uint32_t*>(lut + 2 * LUT), a.tables,
                                                          tk.tab_off, tk.tab_words, tid, W * G * NG);
    __syncthreads();
    const T* __restrict__ src = a.in + (tk.mu_off + sm.c) * ninst_prev + sm.inst;
    const long long rstride = (long long)tk.d_mu * ninst_prev;

    P acc[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[q].v[v] = T(0);

    for (int i0 = 0; i0 < nch; i0 += NBUF) {      // i0: first plane of the group; its coset is a.cid[i0]
        const int nb = (nch - i0 < NBUF) ? (nch - i0) : NBUF;
        // work[m] <- embed_b( F^(cid[i0+m])_mu[:, c] ): zero rows outside the block, then the block rows
        if (G > 1) {
            PV zero;
#pragma unroll
            for (int v = 0; v < TV; ++v) zero.v[v] = T(0);
            const int zs = (lt % NVS) * TV;
            for (int t = lt / NVS; t < tk.d_lam; t += FRV) {
                const int rr = t - tk.boff;
                if (rr >= 0 && rr < tk.d_mu && sm.ok && vec) continue;      // filled by the copies below
                if (rr >= 0 && rr < tk.d_mu && !vec) continue;
#pragma unroll
                for (int m = 0; m < NBUF; ++m)
                    if (m < nb) *reinterpret_cast<PV*>(work + m * bstride + t * TC + zs) = zero;
            }
        } else {
            for (int t = fr; t < tk.d_lam; t += FR) {
                const int rr = t - tk.boff;
                if (rr >= 0 && rr < tk.d_mu) continue;
#pragma unroll
                for (int m = 0; m < NBUF; ++m)
                    if (m < nb) work[m * bstride + t * TC + fs] = T(0);
            }
        }
        if (vec) {
            if (sm.ok) {
                const T* __restrict__ p = src + (long long)fr * rstride + (long long)i0 * ninst;
                T* __restrict__ w = work + (tk.boff + fr) * TC + fs;
                for (int r = fr; r < tk.d_mu; r += FRV, p += (long long)FRV * rstride, w += FRV * TC) {
#pragma unroll
                    for (int m = 0; m < NBUF; ++m)
                        if (m < nb) SNFFT_CP_ASYNC(w + m * bstride, p + m * ninst, 16);
                }
            }
        } else {
            for (int r = fr; r < tk.d_mu; r += FR) {
#pragma unroll
                for (int m = 0; m < NBUF; ++m)
                    if (m < nb) {
                        T* w = work + m * bstride + (tk.boff + r) * TC + fs;
                        if (sm.ok) copy_async(w, src + r * rstride + (i0 + m) * ninst);
                        else *w = T(0);
                    }
            }
        }
        SNFFT_CP_ASYNC_WAIT();
        group_sync<G>();
        for (int j = k - 2; j >= a.cid[i0]; --j) {       // rho(s_i) ... rho(s_{k-2}) . work, i = cid[i0 + m] <= j
            int nact = 1;
#pragma unroll
            for (int m = 1; m < NBUF; ++m)
                if (m < nb && a.cid[i0 + m] <= j) nact = m + 1;
            sweep<T, V, TC, G, NBUF>(work, bstride, nact, tab, lut, k, j, rg, lane);
            group_sync<G>();
        }
#pragma unroll
        for (int m = 0; m < NBUF; ++m)
            if (m < nb) {
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
                    const int t = rg + q * G;
                    if (t < tk.d_lam) {
                        const P x = *reinterpret_cast<const P*>(work + m * bstride + t * TC + lane * V);
#pragma unroll
                        for (int v = 0; v < V; ++v) acc[q].v[v] += x.v[v];
                    }
                }
            }
        group_sync<G>();
    }

    // results leave through shared memory so that the global stores are contiguous
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        const int t = rg + q * G;
        if (t < tk.d_lam) *reinterpret_cast<P*>(work + t * TC + lane * V) = acc[q];
    }
    __syncthreads();
    if (sm.ok) {
        T* __restrict__ dst = a.out + (tk.lam_off + tk.boff + sm.c) * ninst + sm.inst;
        const long long ostride = (long long)tk.d_lam * ninst;
        if (vec) {
            dst += (long long)fr * ostride;
            const T* __restrict__ w = work + fr * TC + fs;
            for (int t = fr; t < tk.d_lam; t += FRV, dst += (long long)FRV * ostride, w += FRV * TC)
                *reinterpret_cast<PV*>(dst) = *reinterpret_cast<const PV*>(w);
        } else {
            for (int t = fr; t < tk.d_lam; t += FR) dst[t * ostride] = work[t * TC + fs];
        }
    }
}

// ---------------------------------------------------------------------------
// inverse (adjoint) level step S_k -> S_{k-1}     grid (chunks * coset groups, tasks)
//   G^(i)_mu[:, c] = sum_{lambda covers mu} d_lambda/(k d_mu) *
//                    [ rho(s_{k-2}) ... rho(s_i) F_lambda[:, boff + c] ]_{rows of block mu}
// The coset group index is the fastest grid dimension so that the CTAs that
// re-read one column tile of F_lambda run together and hit in L2.
// ---------------------------------------------------------------------------
template <typename T, int V, int W, int G, int RPT, int NG, int NBUF, int MINB, bool STAGE>
__global__ void __launch_bounds__(W* G* NG, MINB) inv_level_kernel(const LevelArgs<T> a) {
    SNFFT_DYN_SMEM(smem_raw);
    constexpr int TC = W * V, GT = W * G, FR = GT / TC;
    constexpr int TV = 16 / (int)sizeof(T), NVS = TC / TV, FRV = GT / NVS;
    static_assert(G == 1 || NG == 1, "row groups need the whole CTA barrier");
    static_assert(GT % TC == 0 && TC % TV == 0, "transfer role needs whole rows of threads");
    using P = Pack<T, V>;
    using PV = Pack<T, TV>;
    const InvTask tk = a.itasks[blockIdx.y];
    const int k = a.k;
    const int ngroups = (k + NBUF - 1) / NBUF;
    const int i0 = (int)(blockIdx.x % ngroups) * NBUF;
    const int nb = (k - i0 < NBUF) ? (k - i0) : NBUF;
    const int tid = threadIdx.x;
    const int grp = tid / GT, lt = tid % GT, rg = lt / W, lane = lt % W;
    const long long cblock = blockIdx.x / ngroups;
    const long long chunk = cblock * NG + grp;
    const long long ninst = a.ninst, ninst_prev = a.ninst * k;
    const bool vec = G > 1 && (ninst % TV == 0);
    const int fs = vec ? (lt % NVS) * TV : lt % TC;
    const int fr = vec ? lt / NVS : lt / TC;

    long long nchunks;
    const SlotMap sm = make_slots<TC>(chunk, fs, tk.d_mu, ninst, &nchunks);
    if (cblock * NG >= nchunks) return;

    const int bstride = a.bstride;
    T* __restrict__ work = reinterpret_cast<T*>(smem_raw) + (size_t)grp * NBUF * bstride;
    T* __restrict__ lut = reinterpret_cast<T*>(smem_raw) + (size_t)NG * NBUF * bstride;
    if (tid < 2 * LUT) lut[tid] = a.lut[(tid & 1) * LUT + (tid >> 1)];      // shared copy interleaves (a, b) per code

    P acc[NBUF][RPT];
#pragma unroll
    for (int m = 0; m < NBUF; ++m)
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
            for (int v = 0; v < V; ++v) acc[m][q].v[v] = T(0);

    for (int p = 0; p < tk.npar; ++p) {
        const InvParent par = a.parents[tk.par_off + p];
        const T* __restrict__ src = a.in + (par.lam_off + par.boff + sm.c) * ninst + sm.inst;
        const long long rstride = (long long)par.d_lam * ninst;
        const uint32_t* __restrict__ tab = stage_table<STAGE>(reinterpret_cast<uint32_t*>(lut + 2 * LUT), a.tables,
                                                              par.tab_off, par.tab_words, tid, W * G * NG);
        if (vec) {
            PV zero;
#pragma unroll
            for (int v = 0; v < TV; ++v) zero.v[v] = T(0);
            const T* __restrict__ q = src + (long long)fr * rstride;
            T* __restrict__ w = work + fr * TC + fs;
            for (int t = fr; t < par.d_lam; t += FRV, q += (long long)FRV * rstride, w += FRV * TC) {
#pragma unroll
                for (int m = 0; m < NBUF; ++m)
                    if (m < nb) {
                        if (sm.ok) SNFFT_CP_ASYNC(w + m * bstride, q, 16);
                        else *reinterpret_cast<PV*>(w + m * bstride) = zero;
                    }
            }
        } else {
            for (int t = fr; t < par.d_lam; t += FR) {
#pragma unroll
                for (int m = 0; m < NBUF; ++m)
                    if (m < nb) {
                        if (sm.ok) copy_async(work + m * bstride + t * TC + fs, src + t * rstride);
                        else work[m * bstride + t * TC + fs] = T(0);
                    }
            }
        }
        SNFFT_CP_ASYNC_WAIT();
        __syncthreads();
        for (int j = i0; j <= k - 2; ++j) {      // rho(c_i)^T = rho(s_{k-2}) ... rho(s_i)
            const int nact = (j - i0 + 1 < nb) ? (j - i0 + 1) : nb;
            sweep<T, V, TC, G, NBUF>(work, bstride, nact, tab, lut, k, j, rg, lane);
            group_sync<G>();
        }
        const T scale = (T)par.scale;
#pragma unroll
        for (int m = 0; m < NBUF; ++m)
            if (m < nb) {
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
                    const int t = rg + q * G;
                    if (t < tk.d_mu) {
                        const P x = *reinterpret_cast<const P*>(work + m * bstride + (par.boff + t) * TC + lane * V);
#pragma unroll
                        for (int v = 0; v < V; ++v) acc[m][q].v[v] += scale * x.v[v];
                    }
                }
            }
        __syncthreads();          // the next parent overwrites the work buffers and the staged table
    }

    // stage the results of all chains in shared memory, then store contiguously
#pragma unroll
    for (int m = 0; m < NBUF; ++m)
        if (m < nb) {
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                const int t = rg + q * G;
                if (t < tk.d_mu) *reinterpret_cast<P*>(work + m * bstride + t * TC + lane * V) = acc[m][q];
            }
        }
    __syncthreads();
    if (sm.ok) {
        const long long ostride = (long long)tk.d_mu * ninst_prev;
#pragma unroll
        for (int m = 0; m < NBUF; ++m)
            if (m < nb) {
                T* __restrict__ dst = a.out + (tk.mu_off + sm.c) * ninst_prev + (long long)(i0 + m) * ninst + sm.inst;
                if (vec) {
                    dst += (long long)fr * ostride;
                    const T* __restrict__ w = work + m * bstride + fr * TC + fs;
                    for (int t = fr; t < tk.d_mu; t += FRV, dst += (long long)FRV * ostride, w += FRV * TC)
                        *reinterpret_cast<PV*>(dst) = *reinterpret_cast<const PV*>(w);
                } else {
                    for (int t = fr; t < tk.d_mu; t += FR) dst[t * ostride] = work[m * bstride + t * TC + fs];
                }
            }
    }
}

// ---------------------------------------------------------------------------
// generated register kernels for the lowest levels (base_gen.cuh, gen_base.py):
// one instance per thread, lanes over consecutive instances so every global
// access is a contiguous 128-byte row.
//   s5_kernel : X_1 [120][NINST_5] -> X_5 [120][NINST_5]   (levels 2..5; inverse: 5..2)
// ---------------------------------------------------------------------------
constexpr int GEN_THREADS = 128;

template <typename T, bool INVERSE>
__global__ void __launch_bounds__(GEN_THREADS) s5_kernel(const T* __restrict__ in, T* __restrict__ out, long long ninst5) {
    const long long inst = (long long)blockIdx.x * GEN_THREADS + threadIdx.x;
    if (inst >= ninst5) return;
    T x[120], y[120];
#pragma unroll
    for (int p = 0; p < 120; ++p) x[p] = in[(size_t)p * ninst5 + inst];
    if (INVERSE) s5_inverse<T>(x, y);
    else s5_forward<T>(x, y);
#pragma unroll
    for (int e = 0; e < 120; ++e) out[(size_t)e * ninst5 + inst] = y[e];
}

// ---------------------------------------------------------------------------
// layout change of the top level: X_n instance-fastest [coef e][B]  <->  public
// packed layout [lambda][B][d][d].  Per partition a [d^2][B] -> [B][d^2] tiled
// transpose through shared memory.  grid (e-tiles * b-tiles, partitions)
// ---------------------------------------------------------------------------
constexpr int PK_T = 32;

template <typename T, bool TO_PACKED>
__global__ void __launch_bounds__(PK_T * 8) pack_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                         const long long* __restrict__ coef_off,
                                                         const long long* __restrict__ tile_prefix, int nparts,
                                                         long long batch) {
    __shared__ T tile[PK_T][PK_T + 1];
    // blockIdx.x = (global e-tile) * nbt + b-tile ; tile_prefix[l] = first e-tile of partition l
    const long long nbt = (batch + PK_T - 1) / PK_T;
    const long long etg = blockIdx.x / nbt, bt = blockIdx.x % nbt;
    int lam = 0, hi = nparts - 1;
    while (lam < hi) {
        const int mid = (lam + hi + 1) / 2;
        if (tile_prefix[mid] <= etg) lam = mid;
        else hi = mid - 1;
    }
    const long long lo = coef_off[lam], dd = coef_off[lam + 1] - lo;
    const long long et = etg - tile_prefix[lam];
    const int tx = threadIdx.x % PK_T, ty = threadIdx.x / PK_T;       // 32 x 8
    const long long e0 = et * PK_T, b0 = bt * PK_T;
    // instance-fastest side: element (e, b) at (lo + e) * batch + b ; packed side: lo * batch + b * dd + e
    for (int r = ty; r < PK_T; r += 8) {
        if (TO_PACKED) {
            const long long e = e0 + r, b = b0 + tx;
            if (e < dd && b < batch) tile[r][tx] = in[(lo + e) * batch + b];
        } else {
            const long long b = b0 + r, e = e0 + tx;
            if (e < dd && b < batch) tile[tx][r] = in[lo * batch + b * dd + e];
        }
    }
    __syncthreads();
    for (int r = ty; r < PK_T; r += 8) {
        if (TO_PACKED) {
            const long long b = b0 + r, e = e0 + tx;
            if (e < dd && b < batch) out[lo * batch + b * dd + e] = tile[tx][r];
        } else {
            const long long e = e0 + r, b = b0 + tx;
            if (e < dd && b < batch) out[(lo + e) * batch + b] = tile[r][tx];
        }
    }
}

// ---------------------------------------------------------------------------
// Vectorised pack / unpack (B % VT == 0): same register-block transposition as gather_vec_kernel.  A CTA
// owns TR = 32 * VT consecutive coefficients e of one partition and walks over the batch.  The
// instance-fastest side moves 16-byte vectors along b; the packed side moves 16-byte vectors along e when
// d^2 is a multiple of VT and single elements otherwise (odd d).   grid (e-tiles of all partitions), 256 threads
// ---------------------------------------------------------------------------
template <typename T, bool TO_PACKED>
__global__ void __launch_bounds__(256) pack_vec_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                       const long long* __restrict__ coef_off,
                                                       const long long* __restrict__ tile_prefix, int nparts,
                                                       long long batch) {
    constexpr int VT = 16 / (int)sizeof(T), TB = 8 * VT, TR = 32 * VT;
    using PV = Pack<T, VT>;
    alignas(16) __shared__ T tile[TR * TB];
    const int tid = threadIdx.x;
    int lam = 0, hi = nparts - 1;
    while (lam < hi) {
        const int mid = (lam + hi + 1) / 2;
        if (tile_prefix[mid] <= (long long)blockIdx.x) lam = mid;
        else hi = mid - 1;
    }
    const long long lo = coef_off[lam], dd = coef_off[lam + 1] - lo;
    const long long e0 = ((long long)blockIdx.x - tile_prefix[lam]) * TR;
    const bool evec = dd % VT == 0;
    const int cb = tid & 31, bq = tid >> 5;       // packed side: coefficients e0 + VT*cb .., batch rows VT*bq ..
    const int ch = tid & 7, rq = tid >> 3;        // instance side: coefficient e0 + rq + 32*it, batch chunk ch
    const T* __restrict__ pk_in = in + lo * batch;
    T* __restrict__ pk_out = out + lo * batch;
    for (long long b0 = 0; b0 < batch; b0 += TB) {
        if (TO_PACKED) {
#pragma unroll
            for (int it = 0; it < VT; ++it) {
                const int rr = rq + 32 * it;
                const long long e = e0 + rr, b = b0 + VT * ch;
                PV v;
                if (e < dd && b < batch) v = *reinterpret_cast<const PV*>(in + (lo + e) * batch + b);
                else {
#pragma unroll
                    for (int q = 0; q < VT; ++q) v.v[q] = T(0);
                }
                *reinterpret_cast<PV*>(tile + rr * TB + ((ch ^ ((rr / VT) & 7)) * VT)) = v;
            }
            __syncthreads();
            PV col[VT];
#pragma unroll
            for (int j = 0; j < VT; ++j)
                col[j] = *reinterpret_cast<const PV*>(tile + (VT * cb + j) * TB + ((bq ^ (cb & 7)) * VT));
#pragma unroll
            for (int i = 0; i < VT; ++i) {
                const long long b = b0 + VT * bq + i, e = e0 + VT * cb;
                if (b >= batch) continue;
                T* __restrict__ dst = pk_out + b * dd + e;
                if (evec) {
                    if (e < dd) {
                        PV row;
#pragma unroll
                        for (int j = 0; j < VT; ++j) row.v[j] = col[j].v[i];
                        *reinterpret_cast<PV*>(dst) = row;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < VT; ++j)
                        if (e + j < dd) dst[j] = col[j].v[i];
                }
            }
            __syncthreads();
        } else {
            PV blk[VT];
#pragma unroll
            for (int i = 0; i < VT; ++i) {
                const long long b = b0 + VT * bq + i, e = e0 + VT * cb;
                const T* __restrict__ src = pk_in + b * dd + e;
#pragma unroll
                for (int v = 0; v < VT; ++v) blk[i].v[v] = T(0);
                if (b < batch) {
                    if (evec) {
                        if (e < dd) blk[i] = *reinterpret_cast<const PV*>(src);
                    } else {
#pragma unroll
                        for (int j = 0; j < VT; ++j)
                            if (e + j < dd) blk[i].v[j] = src[j];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < VT; ++j) {
                PV col;
#pragma unroll
                for (int i = 0; i < VT; ++i) col.v[i] = blk[i].v[j];
                *reinterpret_cast<PV*>(tile + (VT * cb + j) * TB + ((bq ^ (cb & 7)) * VT)) = col;
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < VT; ++it) {
                const int rr = rq + 32 * it;
                const long long e = e0 + rr, b = b0 + VT * ch;
                if (e < dd && b < batch)
                    *reinterpret_cast<PV*>(out + (lo + e) * batch + b) =
                        *reinterpret_cast<const PV*>(tile + rr * TB + ((ch ^ ((rr / VT) & 7)) * VT));
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------
// input gather / output scatter: the composition of restrict_to_coset over all
// levels (fourier.py:253) as one tiled transpose.  grid (ranks/64, B/32), 256 thr
//   GATHER :  x1[p1(r) * B + b] = f[b * n! + r]
//   SCATTER:  f[b * n! + r]     = x1[p1(r) * B + b]
// ---------------------------------------------------------------------------
constexpr int GS_TR = 128, GS_TB = 32, GS_THREADS = 256;

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
    const long long r0 = (long long)blockIdx.x * GS_TR;
    const int tid = threadIdx.x;
    if (tid < GS_TR) {
        const long long r = r0 + tid;
        p1[tid] = r < ia.nfact ? chain_index(ia.n, r, ia.fact) : -1;
    }
    __syncthreads();
    // batch tiles with a grid stride: grid.y is capped at 65535 (any batch size works)
    for (long long b0 = (long long)blockIdx.y * GS_TB; b0 < ia.batch; b0 += (long long)gridDim.y * GS_TB) {
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
    __syncthreads();            // the tile is reused by the next batch tile
    }
}

// ---------------------------------------------------------------------------
// Vectorised gather / scatter (B % VT == 0, VT = 16 bytes / sizeof(T)): every global access is a 16-byte
// vector, the transposition happens in VT x VT register blocks, and the shared tile [TR][TB] keeps rows of
// TB = 8 * VT batch entries (eight 16-byte chunks, XOR-swizzled by the rank so that both sides are
// conflict free).  A CTA owns TR = 32 * VT consecutive ranks and walks over the whole batch; p1tab (the
// chain index of every rank, 32-bit, built once per plan) replaces the per-rank index computation.
//   grid (ranks / TR), 256 threads
// ---------------------------------------------------------------------------
template <typename T, bool SCATTER>
__global__ void __launch_bounds__(256) gather_vec_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                         const uint32_t* __restrict__ p1tab, const IndexArgs ia) {
    constexpr int VT = 16 / (int)sizeof(T), TB = 8 * VT, TR = 32 * VT;
    using PV = Pack<T, VT>;
    alignas(16) __shared__ T tile[TR * TB];
    const int tid = threadIdx.x;
    const long long r0 = (long long)blockIdx.x * TR;
    // function side (f[b][r]): thread (cb, bq) owns ranks r0 + VT*cb .. and batch rows VT*bq ..
    const int cb = tid & 31, bq = tid >> 5;
    // level side (x1[p1(r)][b]): thread (rq, ch) owns rank rq + 32*it and the 16-byte batch chunk ch
    const int ch = tid & 7, rq = tid >> 3;
    const bool f_ok = r0 + (long long)VT * cb < ia.nfact;       // n! is a multiple of VT for n >= 4
    long long p1[VT];
#pragma unroll
    for (int it = 0; it < VT; ++it) {
        const long long r = r0 + rq + 32 * it;
        p1[it] = r < ia.nfact ? (p1tab ? (long long)p1tab[r] : chain_index(ia.n, r, ia.fact)) : -1;
    }
    // the loads of the NEXT batch tile are issued before the barrier of the current one (one DRAM latency per tile off the
    // critical path of a CTA)
    PV nxt[VT];
    auto load_f = [&](long long b0, PV (&dst)[VT]) {             // dst[i] = f[b0 + VT*bq + i][r0 + VT*cb ..]
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const long long b = b0 + VT * bq + i;
            if (f_ok && b < ia.batch) dst[i] = *reinterpret_cast<const PV*>(in + b * ia.nfact + r0 + VT * cb);
            else {
#pragma unroll
                for (int v = 0; v < VT; ++v) dst[i].v[v] = T(0);
            }
        }
    };
    auto load_x = [&](long long b0, PV (&dst)[VT]) {             // dst[it] = x1[p1(rank rq + 32*it)][b0 + VT*ch ..]
#pragma unroll
        for (int it = 0; it < VT; ++it) {
            const long long b = b0 + VT * ch;
            if (p1[it] >= 0 && b < ia.batch) dst[it] = *reinterpret_cast<const PV*>(in + p1[it] * ia.batch + b);
            else {
#pragma unroll
                for (int q = 0; q < VT; ++q) dst[it].v[q] = T(0);
            }
        }
    };
    if (ia.batch > 0) {
        if (!SCATTER) load_f(0, nxt);
        else load_x(0, nxt);
    }
    for (long long b0 = 0; b0 < ia.batch; b0 += TB) {
        if (!SCATTER) {
            PV blk[VT];
#pragma unroll
            for (int i = 0; i < VT; ++i) blk[i] = nxt[i];
#pragma unroll
            for (int j = 0; j < VT; ++j) {                       // rank VT*cb + j, batch chunk bq
                PV col;
#pragma unroll
                for (int i = 0; i < VT; ++i) col.v[i] = blk[i].v[j];
                *reinterpret_cast<PV*>(tile + (VT * cb + j) * TB + ((bq ^ (cb & 7)) * VT)) = col;
            }
            if (b0 + TB < ia.batch) load_f(b0 + TB, nxt);
            __syncthreads();
#pragma unroll
            for (int it = 0; it < VT; ++it) {
                const int rr = rq + 32 * it;
                const long long b = b0 + VT * ch;
                if (p1[it] >= 0 && b < ia.batch)
                    *reinterpret_cast<PV*>(out + p1[it] * ia.batch + b) =
                        *reinterpret_cast<const PV*>(tile + rr * TB + ((ch ^ ((rr / VT) & 7)) * VT));
            }
            __syncthreads();
        } else {
#pragma unroll
            for (int it = 0; it < VT; ++it) {
                const int rr = rq + 32 * it;
                *reinterpret_cast<PV*>(tile + rr * TB + ((ch ^ ((rr / VT) & 7)) * VT)) = nxt[it];
            }
            if (b0 + TB < ia.batch) load_x(b0 + TB, nxt);
            __syncthreads();
            PV col[VT];
#pragma unroll
            for (int j = 0; j < VT; ++j)
                col[j] = *reinterpret_cast<const PV*>(tile + (VT * cb + j) * TB + ((bq ^ (cb & 7)) * VT));
#pragma unroll
            for (int i = 0; i < VT; ++i) {
                const long long b = b0 + VT * bq + i;
                if (f_ok && b < ia.batch) {
                    PV row;
#pragma unroll
                    for (int j = 0; j < VT; ++j) row.v[j] = col[j].v[i];
                    *reinterpret_cast<PV*>(out + b * ia.nfact + r0 + VT * cb) = row;
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------
// calc_power (fourier.py:362-393): power[lambda][b] = d_lambda / n! * || F_lambda[b] ||_F^2 on the packed
// coefficients [lambda][B][d][d].   grid (partitions, B), 256 threads, fp64 accumulation
// ---------------------------------------------------------------------------
// Sum over the 256 threads of a CTA: warp-shuffle tree inside each warp, one shared-memory word per warp, then the
// first warp.  (The host-emulation build runs one fiber per thread and has no shuffles: plain shared-memory tree.)
__device__ __forceinline__ double block_sum_256(double v, double* part) {
#ifdef SNFFT_HOST_EMU
    part[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    const double r = part[0];
    __syncthreads();
    return r;
#else
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) part[warp] = v;
    __syncthreads();
    v = lane < 8 ? part[lane] : 0.0;
    if (warp == 0)
        for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) part[0] = v;
    __syncthreads();
    const double r = part[0];
    __syncthreads();
    return r;
#endif
}

template <typename T>
__global__ void __launch_bounds__(256) power_kernel(const T* __restrict__ F, const long long* __restrict__ coef_off,
                                                   const int* __restrict__ dims, long long batch, double inv_order,
                                                   T* __restrict__ out) {
    __shared__ double part[256];
    const int lam = blockIdx.x;
    const long long lo = coef_off[lam], dd = coef_off[lam + 1] - lo;
    for (long long b = blockIdx.y; b < batch; b += gridDim.y) {        // grid.y is capped at 65535
        const T* __restrict__ src = F + lo * batch + b * dd;
        double acc = 0.0;
        for (long long e = threadIdx.x; e < dd; e += 256) {
            const double v = (double)src[e];
            acc += v * v;
        }
        const double tot = block_sum_256(acc, part);
        if (threadIdx.x == 0) out[(long long)lam * batch + b] = (T)(tot * (double)dims[lam] * inv_order);
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

// restrict_to_coset (fourier.py:63-80) for a list of cosets and a batch of functions:
//   out[(j * B + b) * (n-1)! + r'] = f[b * n! + idx_{coset[j]}(r')],  idx = the closed form of coset_index_kernel.
// The (j, b) order makes the planes of the next level step contiguous (plane j = coset[j], B instances each).
struct RestrictArgs {
    IndexArgs ia;
    long long batch;
    int ncos;
    int coset[SNFFT_MAX_N];
};

template <typename T>
__global__ void restrict_kernel(const T* __restrict__ f, T* __restrict__ out, const RestrictArgs ra) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = ra.ia.n;
    const long long fm = ra.ia.fact[n - 1];
    if (e >= (long long)ra.ncos * ra.batch * fm) return;
    const long long r = e % fm, jb = e / fm;
    const long long b = jb % ra.batch;
    const int coset = ra.coset[jb / ra.batch];
    long long src = (long long)(n - 1 - coset) * ra.ia.fact[n - 1 - coset];
    long long rem = r;
    for (int j = 0; j < n - 1; ++j) {
        const long long fj = ra.ia.fact[n - 2 - j];
        const long long lj = rem / fj;
        rem -= lj * fj;
        const int pos = j < coset ? j : j + 1;
        src += lj * ra.ia.fact[n - 1 - pos];
    }
    out[e] = f[b * ra.ia.nfact + src];
}

// Two nested restrictions (coset i_n of S_n, then coset i_{n-1} of S_{n-1}) for a list of pairs:
//   out[(p * B + b) * (n-2)! + r''] = f[b * n! + idx^n_{i_n[p]}( idx^{n-1}_{i_{n-1}[p]}(r'') )]
// The pairs of ONE function are the unit of work of the coset-sharded transform (n (n-1) two-level cosets).
constexpr int RESTRICT2_MAXPAIRS = SNFFT_MAX_N * (SNFFT_MAX_N - 1);
struct Restrict2Args {
    IndexArgs ia;
    long long batch;
    int npairs;
    unsigned char top[RESTRICT2_MAXPAIRS], sub[RESTRICT2_MAXPAIRS];
};

// 32-bit arithmetic: n! < 2^32 for n <= 12 (SNFFT_MAX_N); the 64-bit divisions of coset_index_kernel cost
// about ten times as much and this runs once per element of the function
__device__ __forceinline__ unsigned coset_lift(int n, int coset, unsigned r, const long long* fact) {
    unsigned out = (unsigned)(n - 1 - coset) * (unsigned)fact[n - 1 - coset];
    unsigned rem = r;
    for (int j = 0; j < n - 1; ++j) {
        const unsigned fj = (unsigned)fact[n - 2 - j];
        const unsigned lj = rem / fj;
        rem -= lj * fj;
        const int pos = j < coset ? j : j + 1;
        out += lj * (unsigned)fact[n - 1 - pos];
    }
    return out;
}

template <typename T>
__global__ void restrict2_kernel(const T* __restrict__ f, T* __restrict__ out, const Restrict2Args ra) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = ra.ia.n;
    const long long fm = ra.ia.fact[n - 2];
    if (e >= (long long)ra.npairs * ra.batch * fm) return;
    const long long pb = e / fm;
    const unsigned r = (unsigned)(e - pb * fm);
    const long long b = pb % ra.batch;
    const int p = (int)(pb / ra.batch);
    const unsigned mid = coset_lift(n - 1, ra.sub[p], r, ra.ia.fact);
    out[e] = f[b * ra.ia.nfact + coset_lift(n, ra.top[p], mid, ra.ia.fact)];
}

// Level array of the pairs' sub-transforms -> level array n-2 of the full transform, zero where a pair is absent:
//   dst[e * ninst_full + slot[p] + b] = src[e * (npairs * B) + p * B + b],   slot[p] = (i_{n-1} n + i_n) B
// (instance of level n-2 = i_{n-1} NINST_{n-1} + i_n B + b).  dst is zeroed beforehand.
struct ExpandArgs {
    long long batch, ncoef, ninst_full;
    int npairs;
    int slot[RESTRICT2_MAXPAIRS];
};

template <typename T>
__global__ void expand_pairs_kernel(const T* __restrict__ src, T* __restrict__ dst, const ExpandArgs ea) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ni = (long long)ea.npairs * ea.batch;
    if (q >= ea.ncoef * ni) return;
    const long long e = q / ni, pb = q - e * ni;
    const int p = (int)(pb / ea.batch);
    const long long b = pb - (long long)p * ea.batch;
    dst[e * ea.ninst_full + ea.slot[p] + b] = src[q];
}

// Input permutation / output scatter for SMALL batches (restrict_to_coset composed over all levels,
// fourier.py:63-80,253; inverse lift_from_coset, fourier.py:46-60): a thread owns one rank r and walks over the
// batch, x1[p1(r) * B + b] <-> f[b * n! + r].  The function side is coalesced across threads, the level side moves
// B consecutive elements per rank.  One function (B = 1, the coset-sharded transform of a large function) runs
// at the speed of a permutation with 4-byte granularity instead of the tile transposes sized for large batches.
template <typename T, bool SCATTER>
__global__ void __launch_bounds__(256) gather_small_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                           const uint32_t* __restrict__ p1tab, const IndexArgs ia) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ia.nfact) return;
    const long long p1 = p1tab ? (long long)p1tab[r] : chain_index(ia.n, r, ia.fact);
    const long long B = ia.batch;
    for (long long b = 0; b < B; ++b) {
        if (!SCATTER) out[p1 * B + b] = in[b * ia.nfact + r];
        else out[b * ia.nfact + r] = in[p1 * B + b];
    }
}

__global__ void chain_index32_kernel(uint32_t* __restrict__ p1, const IndexArgs ia) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ia.nfact) return;
    p1[r] = (uint32_t)chain_index(ia.n, r, ia.fact);
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
// applying the words s_i ... s_{k-2} for k = 2..n to the unit vector e_c, each
// s_j in gather form  y[t] = diag[t] x[t] + off[t] x[partner[t]]  (irreps.py:91-109).
// One lane per (sigma, column); block of 32 lanes, two column buffers in shared memory.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(32) dense_rep_kernel(const RowTables<T> tb, int d, T* __restrict__ out,
                                                         const IndexArgs ia) {
    SNFFT_DYN_SMEM(smem_raw);
    T* x = reinterpret_cast<T*>(smem_raw);
    T* y = x + d * 32;
    const int lane = threadIdx.x;
    const long long q = (long long)blockIdx.x * 32 + lane;
    const long long rank = q / d;
    const int c = (int)(q % d);
    if (rank >= ia.nfact) return;
    for (int t = 0; t < d; ++t) x[t * 32 + lane] = (t == c) ? T(1) : T(0);
    int sig[SNFFT_MAX_N];
    unrank_perm(ia.n, rank, ia.fact, sig);
    for (int k = 2; k <= ia.n; ++k) {
        const int v = k - 1;
        int i = 0;
        for (int p = 0; p < ia.n; ++p) {
            if (sig[p] == v) break;
            i += sig[p] < v;
        }
        for (int j = k - 2; j >= i; --j) {
            const int* __restrict__ pt = tb.partner + (size_t)j * d;
            const T* __restrict__ dg = tb.diag + (size_t)j * d;
            const T* __restrict__ of = tb.off + (size_t)j * d;
            for (int t = 0; t < d; ++t) y[t * 32 + lane] = dg[t] * x[t * 32 + lane] + of[t] * x[pt[t] * 32 + lane];
            T* tmp = x;
            x = y;
            y = tmp;
        }
    }
    for (int t = 0; t < d; ++t) out[(rank * d + t) * d + c] = x[t * 32 + lane];
}


// rho_lambda(sigma) for ONE permutation (SnIrrep.matrix_representations[sigma], irreps.py:122-141): the same word
// of adjacent transpositions as dense_rep_kernel, applied to the unit vector e_c by CTA c with the column in
// shared memory (any d: 2310 rows at n = 11 are 18 KB in fp64).   grid (d), 128 threads, 2 d elements of shared memory
struct PermArg {
    int sig[SNFFT_MAX_N];
};

template <typename T>
__global__ void __launch_bounds__(128) rho_kernel(const RowTables<T> tb, int d, int n, const PermArg pa, T* __restrict__ out) {
    SNFFT_DYN_SMEM(smem_raw);
    T* x = reinterpret_cast<T*>(smem_raw);
    T* y = x + d;
    const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    for (int t = tid; t < d; t += nt) x[t] = (t == c) ? T(1) : T(0);
    __syncthreads();
    for (int k = 2; k <= n; ++k) {
        const int v = k - 1;
        int i = 0;
        for (int p = 0; p < n; ++p) {
            if (pa.sig[p] == v) break;
            i += pa.sig[p] < v;
        }
        for (int j = k - 2; j >= i; --j) {
            const int* __restrict__ pt = tb.partner + (size_t)j * d;
            const T* __restrict__ dg = tb.diag + (size_t)j * d;
            const T* __restrict__ of = tb.off + (size_t)j * d;
            for (int t = tid; t < d; t += nt) y[t] = dg[t] * x[t] + of[t] * x[pt[t]];
            __syncthreads();
            T* tmp = x;
            x = y;
            y = tmp;
        }
    }
    for (int t = tid; t < d; t += nt) out[(size_t)t * d + c] = x[t];
}

// Translation of functions on S_n by a fixed permutation (reference tests/test_fourier.py:112-121):
//   left : out[b][k] = f[b][rank(pi^-1 * p_k)],   (a * b)[i] = b[a[i]]  (permutations.py:85-88)
//   right: out[b][k] = f[b][rank(p_k * pi^-1)]
// A thread owns one rank k (unrank, compose, rank: bit exact integer work) and walks over the batch.
struct TranslateArgs {
    IndexArgs ia;
    int pinv[SNFFT_MAX_N];
    int right;
};

template <typename T>
__global__ void __launch_bounds__(256) translate_kernel(const T* __restrict__ f, T* __restrict__ out, const TranslateArgs ta) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = ta.ia.n;
    if (k >= ta.ia.nfact) return;
    int p[SNFFT_MAX_N], q[SNFFT_MAX_N];
    unrank_perm(n, k, ta.ia.fact, p);
    for (int i = 0; i < n; ++i) q[i] = ta.right ? ta.pinv[p[i]] : p[ta.pinv[i]];
    const long long src = rank_perm(n, q, ta.ia.fact);
    for (long long b = 0; b < ta.ia.batch; ++b) out[b * ta.ia.nfact + k] = f[b * ta.ia.nfact + src];
}

// ---------------------------------------------------------------------------
// Per-partition products of two packed coefficient arrays: C_lambda[b] = A_lambda[b] . B_lambda[b]
// (d x d, row major), the Fourier side of the convolution theorem (reference tests/test_fourier.py:87-109).
// One launch for all partitions; a CTA owns a 64 x 64 tile of one batch row of one partition, 256 threads x (4 x 4) outputs,
// K in steps of 16 through shared memory (A stored transposed so both operand reads are conflict free); fp32 partitions
// with d >= MMParts::big_min: 128 x 128 tiles, 8 x 8 outputs per thread (block_matmul_big).
// Plain FFMA / DFMA: the parity contract is fp32 1e-5 / fp64 1e-12, which rules out the TF32 tensor path.
// ---------------------------------------------------------------------------
constexpr int MM_TILE = 64, MM_K = 16, MM_THREADS = 256;

// ONE launch for every partition: the partitions are listed largest first (the long CTAs start early), blk0 = first
// CTA of each, off = its offset in the packed arrays (B * coef_off)
constexpr int MM_MAXPARTS = 80;          // 77 partitions at n = 12
struct MMParts {
    int nparts;
    int big_min;                         // fp32: partitions with d >= big_min run the 128 x 128 tile path
    int d[MM_MAXPARTS];
    long long blk0[MM_MAXPARTS + 1];
    long long off[MM_MAXPARTS];
};
#ifndef SNFFT_MM_BIG_K
#define SNFFT_MM_BIG_K 16
#endif
constexpr int MM_BIG = 128, MM_BIG_K = SNFFT_MM_BIG_K;      // K step of the large tiles: 8 or 16
__host__ __device__ inline int mm_tile_of(int d, int big_min, bool f32) { return f32 && d >= big_min ? MM_BIG : MM_TILE; }

// fp32, large partitions: a CTA owns a 128 x 128 tile, a thread 8 x 8 outputs as four 4 x 4 blocks (rows ty*4.., 64+ty*4..,
// columns tx*4.., 64+tx*4..: every operand read is one conflict-free 16-byte shared load, 64 FMA per 4 loads), K in steps
// of 8, the next K step's global loads issued before the arithmetic of the current one.
// RH / CH: rows m0 + 64.. / columns n0 + 64.. of the tile exist (edge tiles of a partition skip the arithmetic of an
// absent half: d = 288 costs 6.25 tile units instead of 9).
template <bool RH, bool CH>
__device__ __forceinline__ void block_matmul_big(const float* __restrict__ a, const float* __restrict__ bm, float* __restrict__ c,
                                                 int d, int m0, int n0, float (&As)[MM_BIG_K][MM_BIG + 4],
                                                 float (&Bs)[MM_BIG_K][MM_BIG + 4]) {
    constexpr int NP = RH ? 4 : 2, NJ = CH ? 8 : 4;        // row pairs and columns this thread accumulates
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;
    // loaders: A tile 128 rows x MM_BIG_K k -> thread (row = tid / 2, k = KH (tid % 2) ..+KH-1, KH = MM_BIG_K / 2); B tile
    // MM_BIG_K k x 128 columns -> thread (k = tid / 32 + 8 h, column = 4 (tid % 32) ..+3).  Scalar loads: d is not a
    // multiple of 4 in general.
    constexpr int KH = MM_BIG_K / 2, NB = MM_BIG_K / 8;
    const int ar = tid / 2, ak = (tid % 2) * KH;
    const int bk = tid / 32, bc = (tid % 32) * 4;
    float pa[KH], pb[NB][4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < KH; ++q)
            pa[q] = (m0 + ar < d && k0 + ak + q < d) ? a[(size_t)(m0 + ar) * d + k0 + ak + q] : 0.f;
#pragma unroll
        for (int h = 0; h < NB; ++h)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                pb[h][q] = (k0 + bk + 8 * h < d && n0 + bc + q < d) ? bm[(size_t)(k0 + bk + 8 * h) * d + n0 + bc + q] : 0.f;
    };
    F2 acc[NP][NJ];
#pragma unroll
    for (int i = 0; i < NP; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = F2(0.0);
    fetch(0);
    for (int k0 = 0; k0 < d; k0 += MM_BIG_K) {
#pragma unroll
        for (int q = 0; q < KH; ++q) As[ak + q][ar] = pa[q];
#pragma unroll
        for (int h = 0; h < NB; ++h)
#pragma unroll
            for (int q = 0; q < 4; ++q) Bs[bk + 8 * h][bc + q] = pb[h][q];
        __syncthreads();
        if (k0 + MM_BIG_K < d) fetch(k0 + MM_BIG_K);
#pragma unroll
        for (int kk = 0; kk < MM_BIG_K; ++kk) {
            float av[8], bv[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                av[i] = As[kk][ty * 4 + i];
                if (RH) av[4 + i] = As[kk][64 + ty * 4 + i];
                bv[i] = Bs[kk][tx * 4 + i];
                if (CH) bv[4 + i] = Bs[kk][64 + tx * 4 + i];
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                F2 b2;
                b2.x = bv[j];
                b2.y = bv[j];
#pragma unroll
                for (int pr = 0; pr < NP; ++pr) {
                    F2 a2;
                    a2.x = av[2 * pr];
                    a2.y = av[2 * pr + 1];
                    acc[pr][j] = sn_fma(a2, b2, acc[pr][j]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4), cc = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
            if (r < d && cc < d) c[(size_t)r * d + cc] = (i & 1) ? acc[i / 2][j].y : acc[i / 2][j].x;
        }
}

template <typename T>
__global__ void __launch_bounds__(MM_THREADS, 2) block_matmul_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                                  T* __restrict__ C, const MMParts mp) {
    const int tid = threadIdx.x;
    int part = 0;
    while (part + 1 < mp.nparts && (long long)blockIdx.x >= mp.blk0[part + 1]) ++part;
    const int d = mp.d[part];
    const int tile = mm_tile_of(d, mp.big_min, sizeof(T) == 4), tiles = (d + tile - 1) / tile;
    A += mp.off[part];
    B += mp.off[part];
    C += mp.off[part];
    const long long blk = (long long)blockIdx.x - mp.blk0[part];
    const long long b = blk / ((long long)tiles * tiles);
    const int rem = (int)(blk - b * (long long)tiles * tiles);
    const int m0 = (rem / tiles) * tile, n0 = (rem % tiles) * tile;
    const T* __restrict__ a = A + (size_t)b * d * d;
    const T* __restrict__ bm = B + (size_t)b * d * d;
    if constexpr (sizeof(T) == 4) {
        __shared__ float Ab[MM_BIG_K][MM_BIG + 4];
        __shared__ float Bb[MM_BIG_K][MM_BIG + 4];
        if (tile == MM_BIG) {             // uniform per CTA
            const float* af = (const float*)a;
            const float* bf = (const float*)bm;
            float* cf = (float*)(C + (size_t)b * d * d);
            const bool rh = m0 + 64 < d, ch = n0 + 64 < d;
            if (rh && ch) block_matmul_big<true, true>(af, bf, cf, d, m0, n0, Ab, Bb);
            else if (rh) block_matmul_big<true, false>(af, bf, cf, d, m0, n0, Ab, Bb);
            else if (ch) block_matmul_big<false, true>(af, bf, cf, d, m0, n0, Ab, Bb);
            else block_matmul_big<false, false>(af, bf, cf, d, m0, n0, Ab, Bb);
            return;
        }
    }
    __shared__ T As[MM_K][MM_TILE + 4];
    __shared__ T Bs[MM_K][MM_TILE + 4];
    const int ty = tid / 16, tx = tid % 16;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    for (int k0 = 0; k0 < d; k0 += MM_K) {
        // A tile: 64 rows x 16 k, thread -> (row = e / 16, k = e % 16); B tile: 16 k x 64 columns
#pragma unroll
        for (int e = tid; e < MM_TILE * MM_K; e += MM_THREADS) {
            const int r = e / MM_K, kk = e % MM_K;
            As[kk][r] = (m0 + r < d && k0 + kk < d) ? a[(size_t)(m0 + r) * d + k0 + kk] : T(0);
            const int kb = e / MM_TILE, cc = e % MM_TILE;
            Bs[kb][cc] = (k0 + kb < d && n0 + cc < d) ? bm[(size_t)(k0 + kb) * d + n0 + cc] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < MM_K; ++kk) {
            T av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
        __syncthreads();
    }
    T* __restrict__ c = C + (size_t)b * d * d;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = m0 + ty * 4 + i, cc = n0 + tx * 4 + j;
            if (r < d && cc < d) c[(size_t)r * d + cc] = acc[i][j];
        }
}

}  // namespace snfft
