// "Top / bottom" kernels of the level steps k = 9, 10 (generated code tz_gen_*.cuh, gen_tz.py has the derivation
// and the reference citations).  A level step (fourier.py:253-273, inverse fourier.py:191-199) is split in two
// streaming kernels with an intermediate array Z of MB/k of the level's size (MB = 7):
//   forward:  X_{k-1} --Z kernel (level-7 column operators on pieces of the inputs, summed over the cosets i < 7)--> Z
//             Z, X_{k-1}[cosets i >= 7] --T kernel (top sweeps s_6 .. s_{k-2} as one small map per (mu, alpha))--> X_k
//   inverse:  X_k --T' kernel--> Z', X_{k-1}[cosets i >= 7];   Z' --Z' kernel--> X_{k-1}[cosets i < 7]
// No shared memory, no barriers, no tables in the inner code: the Young's-orthogonal-form coefficients
// (irreps.py:91-109) are FFMA immediates of straight-line code, lanes run over consecutive instances so every
// global access is a contiguous run, every element of X_{k-1}, Z and X_k moves once per kernel that needs it.
//
// Level 11 uses MB = 8 (level-8 column operators, three top sweeps like level 10).
// One translation unit per (kind, dtype, direction): SNFFT_TZ_KIND = 7 / 8 (Z kernel with that MB, shared by the
// levels that use it) or the level k = 9, 10, 11 of a T kernel.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <type_traits>

#ifndef SNFFT_TZ_KIND
#error "tz_kernels.cuh: define SNFFT_TZ_KIND (7 / 8 = Z kernel with that MB, 9 / 10 / 11 = T kernel of that level)"
#endif
#include "vec2.cuh"

namespace snfft {       // not the anonymous namespace: sn_fma must overload with vec2.cuh's for the generated code

// ---- short vectors of instances (T kernels) ------------------------------------------------------------------
template <typename S, int NV>
struct alignas(sizeof(S) * NV) TzVec {
    S v[NV];
    TzVec() = default;
    __host__ __device__ __forceinline__ explicit TzVec(double x) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = (S)x;
    }
};
// Elementwise arithmetic; fp32 vectors of even length go through the packed F2 operations of vec2.cuh (FFMA2 / FADD2 /
// FMUL2: half the floating-point instructions of the straight-line code, same results).
template <typename S, int NV>
struct TzPairs {
    static constexpr bool value = std::is_same<S, float>::value && NV % 2 == 0;
};
template <typename S, int NV>
__host__ __device__ __forceinline__ F2 tz_pair(const TzVec<S, NV>& a, int i) {
    F2 p;
    p.x = (float)a.v[2 * i];
    p.y = (float)a.v[2 * i + 1];
    return p;
}
template <typename S, int NV>
__host__ __device__ __forceinline__ void tz_set_pair(TzVec<S, NV>& a, int i, F2 p) {
    a.v[2 * i] = (S)p.x;
    a.v[2 * i + 1] = (S)p.y;
}
template <typename S, int NV>
__host__ __device__ __forceinline__ TzVec<S, NV> sn_fma(const TzVec<S, NV>& a, const TzVec<S, NV>& b, const TzVec<S, NV>& c) {
    TzVec<S, NV> r;
    if constexpr (TzPairs<S, NV>::value) {
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) tz_set_pair(r, i, sn_fma(tz_pair(a, i), tz_pair(b, i), tz_pair(c, i)));
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) r.v[i] = ::snfft::sn_fma(a.v[i], b.v[i], c.v[i]);
    }
    return r;
}
template <typename S, int NV>
__host__ __device__ __forceinline__ TzVec<S, NV> operator*(const TzVec<S, NV>& a, const TzVec<S, NV>& b) {
    TzVec<S, NV> r;
    if constexpr (TzPairs<S, NV>::value) {
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) tz_set_pair(r, i, tz_pair(a, i) * tz_pair(b, i));
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) r.v[i] = a.v[i] * b.v[i];
    }
    return r;
}
template <typename S, int NV>
__host__ __device__ __forceinline__ TzVec<S, NV> operator+(const TzVec<S, NV>& a, const TzVec<S, NV>& b) {
    TzVec<S, NV> r;
    if constexpr (TzPairs<S, NV>::value) {
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) tz_set_pair(r, i, tz_pair(a, i) + tz_pair(b, i));
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) r.v[i] = a.v[i] + b.v[i];
    }
    return r;
}
template <typename S, int NV>
__host__ __device__ __forceinline__ TzVec<S, NV> operator-(const TzVec<S, NV>& a, const TzVec<S, NV>& b) {
    TzVec<S, NV> r;
    if constexpr (TzPairs<S, NV>::value) {
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) tz_set_pair(r, i, tz_pair(a, i) - tz_pair(b, i));
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) r.v[i] = a.v[i] - b.v[i];
    }
    return r;
}
template <typename S, int NV>
__host__ __device__ __forceinline__ TzVec<S, NV> operator-(const TzVec<S, NV>& a) {
    TzVec<S, NV> r;
    if constexpr (TzPairs<S, NV>::value) {
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) tz_set_pair(r, i, -tz_pair(a, i));
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) r.v[i] = -a.v[i];
    }
    return r;
}

// Accessor of the level-k side of a T kernel for one parent lambda (outputs of the forward kernel, inputs of the inverse
// one); rel = row * d_lambda + column of the element inside lambda, relative to the thread's (r, c).
//   PK = false: the level array X_k, instance fastest: element e of lambda at (coef_off + e) * s + inst.
//   PK = true (top level k = n only, NINST_n = B): the PUBLIC packed layout [lambda][B][d][d] (the (B, d, d) tensors
//        sn_fft returns, fourier.py:291-295): element e of batch row b at coef_off * B + b d^2 + e.  The thread's NV
//        instances are NV batch rows: NV scalar accesses d^2 apart; the warp tile of t_kernel (TZ_PK_CW columns x
//        32 / TZ_PK_CW instance vectors) makes every such access a run of TZ_PK_CW consecutive elements per batch row.
//        This replaces the separate pack / unpack pass over the coefficients (2 n! B sizeof of HBM traffic each).
template <typename T, bool PK, bool CONST>
struct TzAcc {
    typename std::conditional<CONST, const T*, T*>::type p;
    unsigned s;
    __host__ __device__ __forceinline__ TzAcc(typename std::conditional<CONST, const T*, T*>::type lb, unsigned co, unsigned dl,
                                              unsigned r, unsigned c, unsigned s_, unsigned inst)
        : p(lb + ((size_t)co + r * dl + c) * s_ + inst), s(s_) {}
    __host__ __device__ __forceinline__ void st(unsigned rel, const T& v) const { p[(size_t)rel * s] = v; }
    __host__ __device__ __forceinline__ T ld(unsigned rel) const { return p[(size_t)rel * s]; }
};
template <typename S, int NV, bool CONST>
struct TzAcc<TzVec<S, NV>, true, CONST> {
    typedef TzVec<S, NV> T;
    typename std::conditional<CONST, const S*, S*>::type p;
    unsigned dd;
    __host__ __device__ __forceinline__ TzAcc(typename std::conditional<CONST, const T*, T*>::type lb, unsigned co, unsigned dl,
                                              unsigned r, unsigned c, unsigned s_, unsigned inst)
        : p((typename std::conditional<CONST, const S*, S*>::type)lb + (size_t)co * NV * s_ + (size_t)(NV * inst) * (dl * dl) + r * dl + c),
          dd(dl * dl) {}
    __host__ __device__ __forceinline__ void st(unsigned rel, const T& v) const {
#pragma unroll
        for (int i = 0; i < NV; ++i) p[rel + (size_t)i * dd] = v.v[i];
    }
    __host__ __device__ __forceinline__ T ld(unsigned rel) const {
        T v;
#pragma unroll
        for (int i = 0; i < NV; ++i) v.v[i] = p[rel + (size_t)i * dd];
        return v;
    }
};

}  // namespace snfft

#include "tz_gen_tab.cuh"
#if SNFFT_TZ_KIND == 7
#include "tz_gen_z7.cuh"
#elif SNFFT_TZ_KIND == 8
#include "tz_gen_z8.cuh"
#elif SNFFT_TZ_KIND == 9
#include "tz_gen_t9.cuh"
#elif SNFFT_TZ_KIND == 10
#include "tz_gen_t10.cuh"
#elif SNFFT_TZ_KIND == 11
#include "tz_gen_t11.cuh"
#endif

namespace snfft {
namespace {

#ifndef SNFFT_Z_THREADS
#define SNFFT_Z_THREADS 128
#endif
constexpr int Z_THREADS = SNFFT_Z_THREADS;
// threads per CTA of the T kernels; measured (S_10, B = 128, profiles r03o): 256 / 128 / 64 -> level 9 0.71 / 0.66 / 0.66 ms,
// level 10 on the packed layout 1.08 / 1.05 / 1.05 ms forward, 0.97 / 0.94 / 0.94 ms inverse (smaller CTAs retire and
// refill more evenly: every thread runs one straight-line function and exits)
#ifndef SNFFT_T_THREADS
#define SNFFT_T_THREADS 128
#endif
constexpr int T_THREADS = SNFFT_T_THREADS;
constexpr int Z_MAXGROUPS = 15;          // partitions of MB - 1 = 6 (11) or 7 (15)
constexpr int T_MAXTASKS = 512;

// ---- Z kernel --------------------------------------------------------------------------------------------
// A group is a shape zeta of S_{MB-1}; its tasks are the parents alpha (forward; MB = 8: x their child blocks) or the cosets i < MB (inverse); its
// lanes are (segment, column c, instance): a segment is one piece (mu, path mu -> zeta) of level k-1.  Block
// order inside a group as in col_kernels.cuh: `chunk` tiles of one task, then the same tiles of the next task,
// so that the sibling tasks re-read their inputs from L2 while the CTAs of an SM share code.
struct ZSeg {
    int ustart;            // first unit (segment, c) of this segment inside its group
    int dmu;
    int in_off;            // (coef_off(mu) + off d_mu) k     : element offset / NINST_k of row 0, column 0, coset 0 in X_{k-1}
    int z_off;             // MB (coef_off(mu) + off d_mu)    : the same in Z
};
struct ZGroups {
    int ngroups, chunk, k;
    unsigned blk0[Z_MAXGROUPS + 1];
    int task0[Z_MAXGROUPS], ntask[Z_MAXGROUPS];
    unsigned ubase[Z_MAXGROUPS];                 // first entry of the group in unit2seg
    unsigned long long lanes[Z_MAXGROUPS];       // units of the group * NINST_k
};

#if SNFFT_TZ_KIND == 7 || SNFFT_TZ_KIND == 8
template <typename T, bool INV>
#ifndef SNFFT_Z_V2_REGS_THREADS
#define SNFFT_Z_V2_REGS_THREADS 384        // resident threads per SM the F2 / fp64 instantiations are compiled for (register cap 65536 / this)
#endif
__global__ void __launch_bounds__(Z_THREADS, (sizeof(T) == 4 ? 640 : SNFFT_Z_V2_REGS_THREADS) / Z_THREADS)
z_kernel(const T* __restrict__ src, T* __restrict__ dst, unsigned s, int s_shift, const ZGroups g,
         const ZSeg* __restrict__ segs, const unsigned* __restrict__ unit2seg) {
    const unsigned b = blockIdx.x;
    int gi = 0;
    while (gi + 1 < g.ngroups && b >= g.blk0[gi + 1]) ++gi;
    const unsigned local = b - g.blk0[gi];
    const unsigned nt = (unsigned)g.ntask[gi];
    const unsigned long long qn = g.lanes[gi];
    const unsigned ntiles = (unsigned)((qn + Z_THREADS - 1) / Z_THREADS);
    const unsigned per = (unsigned)g.chunk * nt;
    const unsigned ch = local / per, r = local - ch * per;
    const unsigned left = ntiles - ch * (unsigned)g.chunk;
    const unsigned cht = left < (unsigned)g.chunk ? left : (unsigned)g.chunk;
    const unsigned tk = r / cht;
    const unsigned tile = ch * (unsigned)g.chunk + (r - tk * cht);
    const int task = g.task0[gi] + (int)tk;
    const unsigned long long q = (unsigned long long)tile * Z_THREADS + threadIdx.x;
    if (q >= qn) return;
    unsigned u, inst;
    if (s_shift >= 0) {                 // NINST_k a power of two
        u = (unsigned)(q >> s_shift);
        inst = (unsigned)q & (s - 1u);
    } else if ((q >> 32) == 0) {
        u = (unsigned)q / s;
        inst = (unsigned)q - u * s;
    } else {
        u = (unsigned)(q / s);
        inst = (unsigned)(q - (unsigned long long)u * s);
    }
    const ZSeg sg = segs[unit2seg[g.ubase[gi] + u]];
    const unsigned c = u - (unsigned)sg.ustart;
    const size_t xo = ((size_t)sg.in_off + (size_t)c * (unsigned)g.k) * s + inst;
    const size_t zo = ((size_t)sg.z_off + c) * s + inst;
    const unsigned rs = (unsigned)sg.dmu * (unsigned)g.k * s, zs = (unsigned)sg.dmu * s;
#if SNFFT_TZ_KIND == 7
    if constexpr (!INV) zb7f_run<T>(task, src + xo, dst + zo, rs, s, zs);
    else zb7i_run<T>(task, src + zo, dst + xo, rs, s, zs);
#else
    if constexpr (!INV) zb8f_run<T>(task, src + xo, dst + zo, rs, s, zs);
    else zb8i_run<T>(task, src + zo, dst + xo, rs, s, zs);
#endif
}
#endif

// ---- T kernel --------------------------------------------------------------------------------------------
// A task is (mu of k-1, alpha of 7); a thread is (row r of alpha, column c of mu, NV adjacent instances).
// The generated code is elementwise in the instance, so it runs unchanged on a short vector of instances:
// 16-byte (8-byte) global transfers, NV times the bytes in flight per thread and 1/NV of the index arithmetic
// (with scalar threads the kernel is latency bound: <= 10 loads of 4 bytes per thread at level 9).
constexpr int T_HINTS = 512;
struct TTasks {
    int ntasks, hint_shift;
    int pk_ct_fastest;                   // packed kernels: consecutive warps walk the column tiles (1) or the instance tiles (0) first
    unsigned blk0[T_MAXTASKS + 1];       // first block of every task (d_alpha, d_mu per task: TzDev<K>, generated)
    unsigned short hint[T_HINTS + 1];    // hint[j] = the task that owns block j << hint_shift: the search starts there
};

#if SNFFT_TZ_KIND >= 9
// T = S (scalar threads) or TzVec<S, NV>; s = NINST_k / NV; s_shift = log2(s) when s is a power of two, else -1.
// (A two-dimensional grid with the task in blockIdx.y and early exit of the tiles past a small task's end was
// measured: 0.78 -> 0.97 ms at k = 9 -- 400 k empty CTAs are not free -- so the task is found by binary search.)
// PK = true (top level, T = TzVec<S, NV>): the level-k side is the public packed layout (TzAcc).  A warp is then a tile
// of TZ_PK_CW (template parameter) columns x 32 / TZ_PK_CW instance vectors of one row r: the instance-fastest side (Z, X_{k-1}) moves runs of
// NV * 32 / TZ_PK_CW instances per column, the packed side runs of TZ_PK_CW columns per batch row; consecutive warps
// walk the column tiles of one (r, instance tile) first (TTasks::pk_ct_fastest), i.e. along the rows of F_lambda.
// cw = 0: not packed
__host__ __device__ __forceinline__ unsigned long long t_task_lanes(unsigned dal, unsigned dmu, unsigned s, unsigned cw) {
    if (cw == 0) return (unsigned long long)dal * dmu * s;
    const unsigned iw = 32 / cw;
    return (unsigned long long)dal * ((dmu + cw - 1) / cw) * ((s + iw - 1) / iw) * 32ull;
}

template <typename T, int K, bool INV, bool PK, int TZ_PK_CW = 16>
__global__ void __launch_bounds__(T_THREADS, 512 / T_THREADS)
t_kernel(typename std::conditional<INV, T*, const T*>::type zb, typename std::conditional<INV, T*, const T*>::type xb,
         typename std::conditional<INV, const T*, T*>::type lb, unsigned s, int s_shift, const TTasks g) {
    const unsigned b = blockIdx.x;
    // last task with blk0 <= b, searched between the hints of this block's bucket (usually zero or one step:
    // a full binary search over the tasks was a third of this kernel's instructions)
    int lo = g.hint[b >> g.hint_shift], hi = g.hint[(b >> g.hint_shift) + 1];
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (g.blk0[mid] <= b) lo = mid;
        else hi = mid - 1;
    }
    const int task = lo;
    const unsigned dmu = TzDev<K>::dmu(task);
    const unsigned long long qn = t_task_lanes(TzDev<K>::dal(task), dmu, s, PK ? TZ_PK_CW : 0);
    const unsigned long long q = (unsigned long long)(b - g.blk0[task]) * T_THREADS + threadIdx.x;
    if (q >= qn) return;
    unsigned r, c, inst;
    if constexpr (PK) {
        constexpr unsigned IW = 32 / TZ_PK_CW;
        const unsigned nip = (s + IW - 1) / IW, nct = (dmu + TZ_PK_CW - 1) / TZ_PK_CW;
        const unsigned lane = (unsigned)q & 31u, w = (unsigned)(q >> 5);
        unsigned ct, ip;
        if (g.pk_ct_fastest) {
            const unsigned ri = w / nct;
            ct = w - ri * nct;
            r = ri / nip;
            ip = ri - r * nip;
        } else {
            const unsigned rt = w / nip;
            ip = w - rt * nip;
            r = rt / nct;
            ct = rt - r * nct;
        }
        c = ct * TZ_PK_CW + (lane % TZ_PK_CW);
        inst = ip * IW + lane / TZ_PK_CW;
        if (c >= dmu || inst >= s) return;
    } else {
        unsigned rc;
        if (s_shift >= 0) {
            rc = (unsigned)(q >> s_shift);
            inst = (unsigned)q & (s - 1u);
        } else if ((q >> 32) == 0) {
            rc = (unsigned)q / s;
            inst = (unsigned)q - rc * s;
        } else {
            rc = (unsigned)(q / s);
            inst = (unsigned)(q - (unsigned long long)rc * s);
        }
        r = rc / dmu;
        c = rc - r * dmu;
    }
    TzDev<K>::template run<T, INV, PK>(task, zb + inst, xb + inst, lb, r, c, s, inst);
}
#endif

}  // namespace
}  // namespace snfft
