// Level kernels, second generation ("TB"): fused top sweeps + register bottom sweeps.
// Included by snfft.cu only.
//
// One forward level step (reference fourier.py:253-273) for column c of block b of F_lambda:
//     F_lambda[:, boff_b + c] = sum_i  rho(s_i) ... rho(s_{k-2}) . embed_b( F^(i)_mu[:, c] )
// Row t of lambda is a path lambda = mu_k -> mu_{k-1} -> ... -> mu_1 in Young's lattice
// (restriction-adapted order, tableau.py:171-198) and rho(s_j) only changes mu_{j+1}.  Hence
//   * the s TOP sweeps s_{k-2} .. s_{k-1-s} act on the first s coordinates of the path: for a fixed
//     zeta = mu_{k-1-s} and a fixed tableau t of zeta they map the <= s! rows (b, path b -> zeta, t)
//     of the embedded column to the <= (s+1)! rows (path lambda -> zeta, t) with a small dense
//     matrix C that does not depend on t ("orbit" of zeta).  The T phase evaluates C straight from
//     global memory into shared memory: out = C . in, no sweep ever runs in shared memory;
//   * the BOTTOM sweeps s_{MB-2} .. s_0, MB = k - s, act inside the sub-blocks (fixed path
//     lambda -> alpha, alpha a partition of MB, d_alpha <= 35 rows) as the irrep alpha of S_MB.
//     The B phase keeps one sub-block of one column-instance in the registers of a thread and runs
//     the generated code of bot_gen.cuh on it; the sum over the k cosets stays in registers too.
// Chain i uses the sweeps j >= i only: its T matrix is the product of the first min(s, k-1-i)
// top sweeps ("class" cls) and its bottom part skips the sweeps j < i.
// The inverse (adjoint, fourier.py:191-199) runs the same tables backwards: B' from global memory
// into shared memory, then T' = C^T from shared memory into register accumulators over the parents.
#pragma once
#include <stdint.h>

#include "bot_gen.cuh"

namespace snfft {

constexpr int TB_MAXS = 3;                         // fused top sweeps
constexpr int TB_MAXIN = 6;                        // TB_MAXS!
constexpr int TB_ORB_WORDS = 3 + TB_MAXIN;         // start, len, nin, in_row[6]
constexpr int TB_BT_STRIDE = 1 + 2 * (TB_MAXS + 1);   // base | type << 16, then (lo, hi) row masks per class
constexpr int TB_TT_HEAD = 4;                      // norb, R, s, pad

// T table of one (lambda, b) (32-bit words, offsets relative to the table):
//   [0] norb  [1] R = sum of orbit lengths  [2] s  [3] offset of the byte table orbit-row -> orbit
//   orbit o at TB_TT_HEAD + o * TB_ORB_WORDS: start (prefix sum of len), len = d_zeta, nin, in_row[TB_MAXIN]
//            (in_row: row inside block b of the embedded column, i.e. row of F_mu)
//   directory at TB_TT_HEAD + norb * TB_ORB_WORDS: for cls = 0..s, orbit o: {out list offset, coefficient offset}
//   out list: nout, out_row[nout] (rows of lambda at t = 0); coefficients C[e][q] at ct[coef offset + e * s! + q]
//            (rows padded to s! entries)
struct TbTask {            // one (lambda, child block b)
    long long lam_off;     // coef_off(lambda)
    long long mu_off;      // coef_off(mu_b) in level k-1
    int d_lam, d_mu, boff, nsb;
    int tt_off, tt_words, ct_off, ct_len, bt_off, bt_words, R, pad;
    double scale;          // d_lambda / (k d_mu), used by the inverse (fourier.py:191-199)
};
struct TbInvTask {         // one mu of level k-1; its parents are TbTask entries
    long long mu_off;
    int d_mu, par_off, npar, R;
};

template <typename T>
struct TbArgs {
    const T* in;
    T* out;
    const TbTask* tasks;
    const TbInvTask* itasks;
    const uint32_t* tt;
    const T* ct;
    const uint32_t* bt;
    int k, s;
    int tc_log2;                  // column-instance tile TC = 1 << tc_log2 slots
    int bstride;                  // elements of one work buffer (rows * TC)
    int xstride;                  // elements of one input staging buffer of the forward kernel (d_mu * TC)
    int tt_smem_words, ct_smem_len, bt_smem_words;
    int vec;                      // 16-byte global transfers allowed (alignment checked by the host)
    int skip;                     // timing experiments only (SNFFT_TB_SKIP): 1 no T phase, 2 no B phase, 4 no input copies
    long long ninst;              // NINST_k
};

// forward: work[nbuf] + xin[2][nbuf] + one table set; inverse: work[nbuf] + one tile + two table sets
template <typename T>
__host__ __device__ inline size_t tb_smem_bytes(bool inverse, int nbuf, int bstride, int xstride, int ct_len, int tt_words,
                                                int bt_words) {
    const size_t tab = (size_t)ct_len * sizeof(T) + ((size_t)tt_words + bt_words) * sizeof(uint32_t);
    if (inverse) return ((size_t)nbuf + 1) * bstride * sizeof(T) + 2 * tab;
    return (size_t)nbuf * ((size_t)bstride + 2 * (size_t)xstride) * sizeof(T) + tab;
}

template <typename T, int VT>
struct alignas(sizeof(T) * VT) TbVec {
    T v[VT];
};

// B-phase row access of the forward kernel: masked loads from the work buffer, sum into registers.
template <typename T>
struct TbFwdIO {
    const T* p;
    int stride;
    uint32_t lo, hi;
    T* acc;
    template <int D>
    __device__ __forceinline__ void load(T (&x)[D]) const {
#pragma unroll
        for (int r = 0; r < D; ++r) {
            const uint32_t bit = r < 32 ? (lo >> (r & 31)) & 1u : (hi >> ((r - 32) & 31)) & 1u;
            x[r] = bit ? p[r * stride] : T(0);
        }
    }
    template <int D>
    __device__ __forceinline__ void store(const T (&x)[D]) const {
#pragma unroll
        for (int r = 0; r < D; ++r) acc[r] += x[r];
    }
};

// B'-phase row access of the inverse kernel: rows of F_lambda from global memory, masked stores to the work buffer.
template <typename T>
struct TbInvIO {
    const T* src;
    long long istride;
    T* dst;
    int stride;
    uint32_t lo, hi;
    template <int D>
    __device__ __forceinline__ void load(T (&x)[D]) const {
#pragma unroll
        for (int r = 0; r < D; ++r) x[r] = src[r * istride];
    }
    template <int D>
    __device__ __forceinline__ void store(const T (&x)[D]) const {
#pragma unroll
        for (int r = 0; r < D; ++r) {
            const uint32_t bit = r < 32 ? (lo >> (r & 31)) & 1u : (hi >> ((r - 32) & 31)) & 1u;
            if (bit) dst[r * stride] = x[r];
        }
    }
};

// y += c * x on VT adjacent column-instances.  fp32 uses the packed FFMA2 of sm_100a.
template <typename T, int VT>
__device__ __forceinline__ void tb_axpy(TbVec<T, VT>& y, T c, const TbVec<T, VT>& x) {
#pragma unroll
    for (int v = 0; v < VT; ++v) y.v[v] += c * x.v[v];
}
#if !defined(SNFFT_HOST_EMU)
template <>
__device__ __forceinline__ void tb_axpy<float, 4>(TbVec<float, 4>& y, float c, const TbVec<float, 4>& x) {
    unsigned long long cc, x0, x1, y0, y1;
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x0) : "f"(x.v[0]), "f"(x.v[1]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x1) : "f"(x.v[2]), "f"(x.v[3]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(y0) : "f"(y.v[0]), "f"(y.v[1]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(y1) : "f"(y.v[2]), "f"(y.v[3]));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(y0) : "l"(cc), "l"(x0));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(y1) : "l"(cc), "l"(x1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(y.v[0]), "=f"(y.v[1]) : "l"(y0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(y.v[2]), "=f"(y.v[3]) : "l"(y1));
}
#endif

// Asynchronous copy of the tables of one task into shared memory (table offsets and lengths are
// multiples of 16 bytes, see build_tb_level).
template <typename T>
__device__ __forceinline__ void tb_stage_async(const TbArgs<T>& a, const TbTask& tk, T* ct, uint32_t* tt, uint32_t* bt,
                                               int tid, int nthreads) {
    constexpr int VT = 16 / (int)sizeof(T);
    for (int w = tid * VT; w < tk.ct_len; w += nthreads * VT) SNFFT_CP_ASYNC(ct + w, a.ct + tk.ct_off + w, 16);
    for (int w = tid * 4; w < tk.tt_words; w += nthreads * 4) SNFFT_CP_ASYNC(tt + w, a.tt + tk.tt_off + w, 16);
    for (int w = tid * 4; w < tk.bt_words; w += nthreads * 4) SNFFT_CP_ASYNC(bt + w, a.bt + tk.bt_off + w, 16);
}

// Transfer role of a thread: which slots of the column-instance tile it moves between global and shared
// memory, resolved once per CTA (the 64-bit divisions are not repeated per copy).
//   vec : 16 bytes = VT adjacent slots of one column, rows r0, r0 + rstep, ...
//   !vec: one slot, rows r0, r0 + rstep, ...
template <typename T>
struct TbXfer {
    long long goff;       // c * cstride + inst of the first slot
    int soff;             // slot offset inside a tile row
    int r0, rstep;
    bool ok;              // slot(s) inside the flattened (column, instance) range
    long long g0, gstep;  // global element offset of row r0 (set_rows) and step between this thread's rows
    int w0, wstep;        // same in the shared-memory tile
};
template <typename T>
__device__ __forceinline__ void tb_xfer_rows(TbXfer<T>& x, long long rstride, int tcl) {
    x.g0 = x.goff + (long long)x.r0 * rstride;
    x.gstep = (long long)x.rstep * rstride;
    x.w0 = (x.r0 << tcl) + x.soff;
    x.wstep = x.rstep << tcl;
}

template <typename T>
__device__ __forceinline__ TbXfer<T> tb_xfer(long long q0, long long qn, long long ninst, long long cstride, int tcl,
                                             bool vec, int tid, int nthreads) {
    constexpr int VT = 16 / (int)sizeof(T);
    TbXfer<T> x;
    int vtl = 0;
    while ((1 << vtl) < VT) ++vtl;
    const int sh = vec ? tcl - vtl : tcl;                 // log2 of transfers per row
    x.soff = (tid & ((1 << sh) - 1)) * (vec ? VT : 1);
    x.r0 = tid >> sh;
    x.rstep = nthreads >> sh;
    const long long q = q0 + x.soff;
    const long long c = q / ninst;
    x.goff = c * cstride + (q - c * ninst);
    x.ok = q < qn;
    return x;
}

// Asynchronous copy of a [rows][TC] tile of a level array (row r of this thread's slots at
// src + r * rstride + x.goff) into shared memory; slots past the end of the range are zero.
template <typename T>
__device__ __forceinline__ void tb_fill_async(T* __restrict__ dst, const T* __restrict__ src, int rows, const TbXfer<T>& x,
                                              bool vec) {
    constexpr int VT = 16 / (int)sizeof(T);
    using PV = TbVec<T, VT>;
    const T* __restrict__ p = src + x.g0;
    T* __restrict__ w = dst + x.w0;
    const long long pstep = x.gstep;
    const int wstep = x.wstep;
    if (vec) {
        if (x.ok) {
            for (int r = x.r0; r < rows; r += x.rstep, p += pstep, w += wstep) SNFFT_CP_ASYNC(w, p, 16);
        } else {
            PV zero;
#pragma unroll
            for (int v = 0; v < VT; ++v) zero.v[v] = T(0);
            for (int r = x.r0; r < rows; r += x.rstep, w += wstep) *reinterpret_cast<PV*>(w) = zero;
        }
    } else {
        for (int r = x.r0; r < rows; r += x.rstep, p += pstep, w += wstep) {
            if (x.ok) SNFFT_CP_ASYNC(w, p, sizeof(T));
            else *w = T(0);
        }
    }
}

// One T-phase item of the forward kernel: orbit row `om` (orbit o, tableau t), slot vector sv, chain m.
template <typename T, int MAXIN>
__device__ __forceinline__ void tb_t_item(const uint32_t* __restrict__ tt, const T* __restrict__ ct,
                                          const T* __restrict__ xs, T* __restrict__ w, int om, int cls, int tcl) {
    constexpr int VT = 16 / (int)sizeof(T);
    using PV = TbVec<T, VT>;
    const int norb = (int)tt[0];
    const uint8_t* __restrict__ r2o = reinterpret_cast<const uint8_t*>(tt + tt[3]);
    const int o = r2o[om];
    const uint32_t* __restrict__ orb = tt + TB_TT_HEAD + o * TB_ORB_WORDS;
    const int t = om - (int)orb[0], nin = (int)orb[2];
    const uint32_t* __restrict__ dir = tt + TB_TT_HEAD + norb * TB_ORB_WORDS + (cls * norb + o) * 2;
    const uint32_t* __restrict__ ol = tt + dir[0];
    const T* __restrict__ cf = ct + dir[1];
    const int nout = (int)ol[0];
    xs += (size_t)t << tcl;
    w += (size_t)t << tcl;
    PV in[MAXIN];
#pragma unroll
    for (int qi = 0; qi < MAXIN; ++qi) {
        if (qi < nin) in[qi] = *reinterpret_cast<const PV*>(xs + ((size_t)orb[3 + qi] << tcl));
        else {
#pragma unroll
            for (int v = 0; v < VT; ++v) in[qi].v[v] = T(0);
        }
    }
    for (int e = 0; e < nout; ++e, cf += MAXIN) {
        PV y;
#pragma unroll
        for (int v = 0; v < VT; ++v) y.v[v] = T(0);
#pragma unroll
        for (int qi = 0; qi < MAXIN; ++qi)
            if (qi < nin) tb_axpy<T, VT>(y, cf[qi], in[qi]);
        *reinterpret_cast<PV*>(w + ((size_t)ol[1 + e] << tcl)) = y;
    }
}

// ---------------------------------------------------------------------------
// forward level step S_{k-1} -> S_k      grid (column-instance tiles, tasks)
// shared memory: work[NBUF][d_lam][TC] | xin[2][NBUF][d_mu][TC] | ct | tt | bt
// ---------------------------------------------------------------------------
template <typename T, int NT, int MINB, int MB, int MAXIN, int NBUF>
__global__ void __launch_bounds__(NT, MINB) fwd_tb_kernel(const TbArgs<T> a) {
    SNFFT_DYN_SMEM(smem_raw);
    constexpr int VT = 16 / (int)sizeof(T), MAXD = BotMaxDim<MB>::value;
    constexpr int VTL = VT == 4 ? 2 : (VT == 2 ? 1 : 0);
    const TbTask tk = a.tasks[blockIdx.y];
    const int tid = threadIdx.x, k = a.k, s = a.s;
    const int tcl = a.tc_log2, TC = 1 << tcl, tcvl = tcl - VTL, TCV = 1 << tcvl;
    const long long ninst = a.ninst, ninst_prev = ninst * k;
    const long long qn = (long long)tk.d_mu * ninst;
    const long long ntiles = (qn + TC - 1) >> tcl;
    // persistent over the tiles of this task: tile = blockIdx.x, blockIdx.x + gridDim.x, ...  The tables are
    // staged once and the first inputs of the next tile are prefetched behind the last B phase of this one.
    long long tile = blockIdx.x;
    if (tile >= ntiles) return;
    const int bstride = a.bstride, xstride = a.xstride;
    T* __restrict__ work = reinterpret_cast<T*>(smem_raw);
    T* __restrict__ xin = work + (size_t)NBUF * bstride;
    T* __restrict__ ct = xin + (size_t)2 * NBUF * xstride;
    uint32_t* __restrict__ tt = reinterpret_cast<uint32_t*>(ct + a.ct_smem_len);
    uint32_t* __restrict__ bt = tt + a.tt_smem_words;
    const bool vec = a.vec != 0;
    const long long rstride = (long long)tk.d_mu * ninst_prev;      // one row of F_mu
    const T* __restrict__ src0 = a.in + tk.mu_off * ninst_prev;     // element (row r, column c, coset i, inst) at
                                                                    // src0 + r * rstride + c * ninst_prev + i * ninst + inst
    TbXfer<T> xf = tb_xfer<T>(tile << tcl, qn, ninst, ninst_prev, tcl, vec, tid, NT);
    tb_xfer_rows(xf, rstride, tcl);
    tb_stage_async(a, tk, ct, tt, bt, tid, NT);
    const int nb0 = k < NBUF ? k : NBUF;
    for (int m = 0; m < nb0; ++m)
        tb_fill_async(xin + (size_t)m * xstride, src0 + (long long)m * ninst, tk.d_mu, xf, vec);

    // B role: sub-block sb of column-instance slot bslot
    const int sb = tid >> tcl, bslot = tid & (TC - 1);
    int par = 0;                      // staging buffer of the current group
    for (;;) {
        const long long q0 = tile << tcl;
        const long long bq = q0 + bslot;
        const bool b_valid = sb < tk.nsb && bq < qn;
        const long long tile_next = tile + gridDim.x;
        int bbase = 0, btype = 0;
        T acc[MAXD];
#pragma unroll
        for (int r = 0; r < MAXD; ++r) acc[r] = T(0);

        for (int i0 = 0; i0 < k; i0 += NBUF, par ^= 1) {
            const int nb = (k - i0 < NBUF) ? (k - i0) : NBUF;
            SNFFT_CP_ASYNC_WAIT();
            __syncthreads();          // inputs of this group (and the tables) have landed; the previous group is done
            // prefetch the next group's inputs (of this tile, or the first group of the next tile) behind this
            // group's T and B phases
            if (i0 + NBUF < k) {
                const int nbn = (k - i0 - NBUF < NBUF) ? (k - i0 - NBUF) : NBUF;
                for (int m = 0; m < nbn; ++m)
                    tb_fill_async(xin + (size_t)((par ^ 1) * NBUF + m) * xstride, src0 + (long long)(i0 + NBUF + m) * ninst,
                                  tk.d_mu, xf, vec);
            } else if (tile_next < ntiles) {
                xf = tb_xfer<T>(tile_next << tcl, qn, ninst, ninst_prev, tcl, vec, tid, NT);
                tb_xfer_rows(xf, rstride, tcl);
                for (int m = 0; m < nb0; ++m)
                    tb_fill_async(xin + (size_t)((par ^ 1) * NBUF + m) * xstride, src0 + (long long)m * ninst, tk.d_mu, xf,
                                  vec);
            }
            // ---- T phase: work[m] <- C_cls . (rows of F^(i0+m)_mu[:, c]) on the support of class cls
            const int R = (int)tt[1];
            const int rt = (a.skip & 1) ? 0 : (R << tcvl);
            for (int item = tid; item < nb * rt; item += NT) {
                int it = item, m = 0;
                while (it >= rt) {
                    it -= rt;
                    ++m;
                }
                const int i = i0 + m;
                const int cls = (k - 1 - i < s) ? (k - 1 - i) : s;
                const int svo = (it & (TCV - 1)) * VT;
                tb_t_item<T, MAXIN>(tt, ct, xin + (size_t)(par * NBUF + m) * xstride + svo, work + (size_t)m * bstride + svo,
                                    it >> tcvl, cls, tcl);
            }
            __syncthreads();          // work buffers complete
            // ---- B phase: bottom sweeps of every chain of the group in registers, summed into acc
            if (b_valid) {
                const uint32_t bw0 = bt[sb * TB_BT_STRIDE];
                bbase = (int)(bw0 & 0xffffu);
                btype = (int)(bw0 >> 16);
#pragma unroll 1
                for (int m = 0; m < ((a.skip & 2) ? 0 : nb); ++m) {
                    const int i = i0 + m;
                    const int cls = (k - 1 - i < s) ? (k - 1 - i) : s;
                    TbFwdIO<T> io;
                    io.lo = bt[sb * TB_BT_STRIDE + 1 + 2 * cls];
                    io.hi = bt[sb * TB_BT_STRIDE + 2 + 2 * cls];
                    if ((io.lo | io.hi) == 0u) continue;
                    io.p = work + (size_t)m * bstride + ((size_t)bbase << tcl) + bslot;
                    io.stride = TC;
                    io.acc = acc;
                    bot_apply<T, MB, false>(btype, i, io);
                }
            }
        }

        if (b_valid) {
            const long long c = bq / ninst, inst = bq - c * ninst;
            T* __restrict__ dst = a.out + (tk.lam_off + (long long)bbase * tk.d_lam + tk.boff + c) * ninst + inst;
            const long long ostride = (long long)tk.d_lam * ninst;
            const int D = bot_dim<MB>(btype);
#pragma unroll
            for (int r = 0; r < MAXD; ++r)
                if (r < D) dst[r * ostride] = acc[r];
        }
        tile = tile_next;
        if (tile >= ntiles) break;
    }
}

// ---------------------------------------------------------------------------
// inverse (adjoint) level step S_k -> S_{k-1}     grid (tiles * coset groups, tasks)
//   G^(i)_mu[:, c] = sum_{lambda covers mu} d_lambda/(k d_mu) *
//                    [ rho(s_{k-2}) ... rho(s_i) F_lambda[:, boff + c] ]_{rows of block mu}
// The coset group index is the fastest grid dimension so that the CTAs that re-read one column
// tile of F_lambda run together and hit in L2.
// shared memory: work[NBUF][d_lam][TC] | fin[d_lam][TC] | 2 x (ct | tt | bt)
// ---------------------------------------------------------------------------
template <typename T, int NT, int MINB, int MB, int MAXIN, int NBUF>
__global__ void __launch_bounds__(NT, MINB) inv_tb_kernel(const TbArgs<T> a) {
    SNFFT_DYN_SMEM(smem_raw);
    constexpr int VT = 16 / (int)sizeof(T);
    constexpr int VTL = VT == 4 ? 2 : (VT == 2 ? 1 : 0);
    using PV = TbVec<T, VT>;
    const TbInvTask tk = a.itasks[blockIdx.y];
    const int tid = threadIdx.x, k = a.k, s = a.s;
    const int tcl = a.tc_log2, TC = 1 << tcl, tcvl = tcl - VTL, TCV = 1 << tcvl;
    const int ngroups = (k + NBUF - 1) / NBUF;
    const int i0 = (int)(blockIdx.x % ngroups) * NBUF;
    const int nb = (k - i0 < NBUF) ? (k - i0) : NBUF;
    const long long ninst = a.ninst, ninst_prev = ninst * k;
    const long long qn = (long long)tk.d_mu * ninst;
    const long long q0 = (long long)(blockIdx.x / ngroups) << tcl;
    if (q0 >= qn) return;
    const int bstride = a.bstride;
    T* __restrict__ work = reinterpret_cast<T*>(smem_raw);
    T* __restrict__ fin = work + (size_t)NBUF * bstride;
    unsigned char* tab0 = reinterpret_cast<unsigned char*>(fin + bstride);
    const size_t tab_bytes = (size_t)a.ct_smem_len * sizeof(T) + ((size_t)a.tt_smem_words + a.bt_smem_words) * sizeof(uint32_t);
    const bool vec = a.vec != 0;
    TbXfer<T> xf = tb_xfer<T>(q0, qn, ninst, ninst, tcl, vec, tid, NT);

    // T' role: one (orbit row, slot vector) per thread, fixed for all parents
    const int orow = tid >> tcvl, svo = (tid & (TCV - 1)) * VT;
    const bool t_role = orow < tk.R;
    const long long tq = q0 + svo;

    PV acc[NBUF][MAXIN];
#pragma unroll
    for (int m = 0; m < NBUF; ++m)
#pragma unroll
        for (int qi = 0; qi < MAXIN; ++qi)
#pragma unroll
            for (int v = 0; v < VT; ++v) acc[m][qi].v[v] = T(0);

    // prefetch of parent p: its tables into table buffer p & 1 and its column tile into fin
    auto tables_of = [&](int p, T*& ct, uint32_t*& tt, uint32_t*& bt) {
        ct = reinterpret_cast<T*>(tab0 + (size_t)(p & 1) * tab_bytes);
        tt = reinterpret_cast<uint32_t*>(ct + a.ct_smem_len);
        bt = tt + a.tt_smem_words;
    };
    auto prefetch = [&](int p) {
        const TbTask par = a.tasks[tk.par_off + p];
        T* ct;
        uint32_t *tt, *bt;
        tables_of(p, ct, tt, bt);
        tb_stage_async(a, par, ct, tt, bt, tid, NT);
        // column c of block b of F_lambda: element (row r, column c, inst) at (lam_off + r * d_lam + boff + c) * ninst + inst
        tb_xfer_rows(xf, (long long)par.d_lam * ninst, tcl);
        tb_fill_async(fin, a.in + (par.lam_off + par.boff) * ninst, par.d_lam, xf, vec);
    };
    prefetch(0);

    const uint32_t* tt_last = nullptr;
    for (int p = 0; p < tk.npar; ++p) {
        const TbTask par = a.tasks[tk.par_off + p];
        T* ct;
        uint32_t *tt, *bt;
        tables_of(p, ct, tt, bt);
        tt_last = tt;
        SNFFT_CP_ASYNC_WAIT();
        __syncthreads();                    // tile and tables of parent p have landed; work buffers are free
        // ---- B' phase: rows of F_lambda[:, boff + c] -> bottom sweeps (ascending) -> work[m]
        // flat loop over (chain, sub-block, slot): the sub-blocks are sorted by size, so the rounds are balanced
        {
            const int per = par.nsb << tcl;
            for (int item = tid; item < nb * per; item += NT) {
                int r = item >> tcl, m = 0;
                while (r >= par.nsb) {
                    r -= par.nsb;
                    ++m;
                }
                const int bslot = item & (TC - 1);
                if (q0 + bslot >= qn) continue;
                const int i = i0 + m;
                const int cls = (k - 1 - i < s) ? (k - 1 - i) : s;
                TbInvIO<T> io;
                io.lo = bt[r * TB_BT_STRIDE + 1 + 2 * cls];
                io.hi = bt[r * TB_BT_STRIDE + 2 + 2 * cls];
                if ((io.lo | io.hi) == 0u) continue;
                const uint32_t bw0 = bt[r * TB_BT_STRIDE];
                const int bbase = (int)(bw0 & 0xffffu), btype = (int)(bw0 >> 16);
                io.src = fin + ((size_t)bbase << tcl) + bslot;
                io.istride = TC;
                io.stride = TC;
                io.dst = work + (size_t)m * bstride + ((size_t)bbase << tcl) + bslot;
                bot_apply<T, MB, true>(btype, i, io);
            }
        }
        __syncthreads();                    // work buffers complete, tile buffer free
        if (p + 1 < tk.npar) prefetch(p + 1);
        // ---- T' phase: acc[m][q] += scale * sum_e C[e][q] work[m][out_row[e] + t]
        if (t_role) {
            const int norb = (int)tt[0];
            const uint8_t* __restrict__ r2o = reinterpret_cast<const uint8_t*>(tt + tt[3]);
            const int o = r2o[orow];
            const uint32_t* __restrict__ orb = tt + TB_TT_HEAD + o * TB_ORB_WORDS;
            const int t = orow - (int)orb[0], nin = (int)orb[2];
            const T scale = (T)par.scale;
#pragma unroll
            for (int m = 0; m < NBUF; ++m)
                if (m < nb) {
                    const int i = i0 + m;
                    const int cls = (k - 1 - i < s) ? (k - 1 - i) : s;
                    const uint32_t* __restrict__ dir = tt + TB_TT_HEAD + norb * TB_ORB_WORDS + (cls * norb + o) * 2;
                    const uint32_t* __restrict__ ol = tt + dir[0];
                    const T* __restrict__ cf = ct + dir[1];
                    const int nout = (int)ol[0];
                    const T* __restrict__ w = work + (size_t)m * bstride + ((size_t)t << tcl) + svo;
                    for (int e = 0; e < nout; ++e, cf += MAXIN) {
                        PV x = *reinterpret_cast<const PV*>(w + ((size_t)ol[1 + e] << tcl));
#pragma unroll
                        for (int v = 0; v < VT; ++v) x.v[v] *= scale;
#pragma unroll
                        for (int qi = 0; qi < MAXIN; ++qi)
                            if (qi < nin) tb_axpy<T, VT>(acc[m][qi], cf[qi], x);
                    }
                }
        }
    }

    // store: G^(i)_mu[in_row[q] + t, c]; the orbit records of mu are the same in every parent's table
    if (t_role) {
        const uint32_t* __restrict__ tt = tt_last;
        const uint8_t* __restrict__ r2o = reinterpret_cast<const uint8_t*>(tt + tt[3]);
        const uint32_t* __restrict__ orb = tt + TB_TT_HEAD + r2o[orow] * TB_ORB_WORDS;
        const int t = orow - (int)orb[0], nin = (int)orb[2];
        long long c_e[VT], inst_e[VT];
        bool ok_e[VT];
#pragma unroll
        for (int e = 0; e < VT; ++e) {
            if (e == 0 || !vec) {
                c_e[e] = (tq + e) / ninst;
                inst_e[e] = (tq + e) - c_e[e] * ninst;
                ok_e[e] = tq + e < qn;
            }
        }
#pragma unroll
        for (int m = 0; m < NBUF; ++m)
            if (m < nb) {
#pragma unroll
                for (int qi = 0; qi < MAXIN; ++qi)
                    if (qi < nin) {
                        const long long row = (long long)orb[3 + qi] + t;
                        T* __restrict__ dst = a.out + (tk.mu_off + row * tk.d_mu) * ninst_prev + (long long)(i0 + m) * ninst;
                        if (vec) {
                            if (ok_e[0]) *reinterpret_cast<PV*>(dst + c_e[0] * ninst_prev + inst_e[0]) = acc[m][qi];
                        } else {
#pragma unroll
                            for (int e = 0; e < VT; ++e)
                                if (ok_e[e]) dst[c_e[e] * ninst_prev + inst_e[e]] = acc[m][qi].v[e];
                        }
                    }
            }
    }
}

}  // namespace snfft
