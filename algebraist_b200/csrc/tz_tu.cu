// One translation unit per (kind, dtype, direction) of the top / bottom kernels (tz_kernels.cuh):
//   -DSNFFT_TZ_KIND=7|8        Z kernel (level-7 / level-8 column operators with run-time strides)
//   -DSNFFT_TZ_KIND=9|10|11    T kernel of that level
//   -DSNFFT_TZ_F64=0|1  -DSNFFT_TZ_INV=0|1
// Exports (0 or a cudaError_t):
//   int snfft_tz_z<MB>_<f32|f64>_<fwd|inv>(int k, const void* src, void* dst, long long ninst, stream)
//        forward: src = X_{k-1}, dst = Z;  inverse: src = Z', dst = X_{k-1};  k must be a level with that MB
//   int snfft_tz_z<MB>_prepare_<f32|f64>_<fwd|inv>(int k)    uploads the level's segment tables to the current device
//   int snfft_tz_t<k>_<f32|f64>_<fwd|inv>(void* z, void* x, void* xk, long long ninst, stream)
//        forward: reads z, x (= X_{k-1}), writes xk (= X_k);  inverse: reads xk, writes z and the cosets i >= 7 of x
//   int snfft_tz_tpk<k>_<f32|f64>_<fwd|inv>(...)   the same with xk = the packed coefficients [lambda][B][d][d] (k = n)
#ifdef SNFFT_HOST_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define SNFFT_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

#include "tz_kernels.cuh"

#if SNFFT_TZ_F64
typedef double tz_t;
#define TZ_DT f64
#else
typedef float tz_t;
#define TZ_DT f32
#endif
#if SNFFT_TZ_INV
#define TZ_DIR inv
#else
#define TZ_DIR fwd
#endif
#ifndef TZ_VEC_DEFAULT
#define TZ_VEC_DEFAULT 4
#endif
#ifndef TZ_PK_CW_DEFAULT
// measured at S_10, B = 128, fp32 (profiles r02v): forward (scalar STORES to the packed side) 1.38 / 1.21 / 1.11 ms with
// 8 / 16 / 32 columns per warp tile, inverse (scalar LOADS from it) 1.03 / 0.97 / 1.19 ms
#if SNFFT_TZ_INV
#define TZ_PK_CW_DEFAULT 16
#else
#define TZ_PK_CW_DEFAULT 32
#endif
#endif
#define TZ_CAT3_(a, b, c) a##b##_##c
#define TZ_CAT3(a, b, c) TZ_CAT3_(a, b, c)

namespace {
using namespace snfft;

#if SNFFT_TZ_KIND == 7 || SNFFT_TZ_KIND == 8
constexpr int kMaxDev = 64;
struct ZDev {                      // device copy of one level's segment tables
    ZSeg* segs = nullptr;
    unsigned* unit2seg = nullptr;
    unsigned ubase[Z_MAXGROUPS] = {};
    unsigned units[Z_MAXGROUPS] = {};
    bool ready = false;
};
ZDev g_zdev[kMaxDev][3];           // [device][k - 9]
std::mutex g_zmu;

template <int K>
int z_upload(ZDev& d) {
    using L = TzLevel<K>;
    std::vector<ZSeg> segs(L::nsegs);
    std::vector<unsigned> u2s;
    int g_prev = -1;
    unsigned ustart = 0;
    for (int i = 0; i < Z_MAXGROUPS; ++i) d.ubase[i] = d.units[i] = 0;
    for (int i = 0; i < L::nsegs; ++i) {
        const int g = L::seg_group()[i];
        if (g != g_prev) {          // segments are sorted by group
            d.ubase[g] = (unsigned)u2s.size();
            ustart = 0;
            g_prev = g;
        }
        segs[i].ustart = (int)ustart;
        segs[i].dmu = L::seg_dmu()[i];
        segs[i].in_off = L::seg_in()[i];
        segs[i].z_off = L::seg_z()[i];
        for (int c = 0; c < segs[i].dmu; ++c) u2s.push_back((unsigned)i);
        ustart += (unsigned)segs[i].dmu;
        d.units[g] = ustart;
    }
    cudaError_t e;
    if ((e = cudaMalloc((void**)&d.segs, segs.size() * sizeof(ZSeg))) != cudaSuccess) return (int)e;
    if ((e = cudaMalloc((void**)&d.unit2seg, u2s.size() * sizeof(unsigned))) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpy(d.segs, segs.data(), segs.size() * sizeof(ZSeg), cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpy(d.unit2seg, u2s.data(), u2s.size() * sizeof(unsigned), cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
    d.ready = true;
    return 0;
}

int z_tables(int k, ZDev** out) {
#if SNFFT_TZ_KIND == 7
    if (k != 9 && k != 10) return (int)cudaErrorInvalidValue;
#else
    if (k != 11) return (int)cudaErrorInvalidValue;
#endif
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= kMaxDev) return (int)cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lock(g_zmu);
    ZDev& d = g_zdev[dev][k - 9];
    if (!d.ready) {
#if SNFFT_TZ_KIND == 7
        const int rc = k == 9 ? z_upload<9>(d) : z_upload<10>(d);
#else
        const int rc = z_upload<11>(d);
#endif
        if (rc) return rc;
    }
    *out = &d;
    return 0;
}
#endif

}  // namespace

#if SNFFT_TZ_KIND == 7 || SNFFT_TZ_KIND == 8

#if SNFFT_TZ_KIND == 7
#define TZ_ZNAME TZ_CAT3(snfft_tz_z7_, TZ_DT, TZ_DIR)
#define TZ_ZPREP TZ_CAT3(snfft_tz_z7_prepare_, TZ_DT, TZ_DIR)
#else
#define TZ_ZNAME TZ_CAT3(snfft_tz_z8_, TZ_DT, TZ_DIR)
#define TZ_ZPREP TZ_CAT3(snfft_tz_z8_prepare_, TZ_DT, TZ_DIR)
#endif

extern "C" int TZ_ZPREP(int k) {
    ZDev* d = nullptr;
    return z_tables(k, &d);
}

extern "C" int TZ_ZNAME(int k, const void* src, void* dst, long long ninst, cudaStream_t st) {
    using namespace snfft;
    static const int chunk = [] {
        // tiles of one task that run before the next task of its group (see col_kernels.cuh); SNFFT_TZ_CHUNK overrides
        const char* e = std::getenv("SNFFT_TZ_CHUNK");
        return e ? std::atoi(e) : (SNFFT_TZ_INV ? 74 : 148);
    }();
    ZDev* d = nullptr;
    int rc = z_tables(k, &d);
    if (rc) return rc;
    using I = ZInfo<SNFFT_TZ_KIND, SNFFT_TZ_INV != 0>;
    static_assert(I::ngroups <= Z_MAXGROUPS, "too many groups");
    // 32-bit strides: row stride d_mu k NINST of X_{k-1} (d_mu <= 216 at k = 10, <= 768 at k = 11)
    if (ninst <= 0 || (unsigned long long)ninst * (k == 11 ? 768ull : 216ull) * (unsigned long long)k >= (1ull << 32))
        return (int)cudaErrorInvalidValue;
    bool v2 = false;
    int s_shift2 = -1;
#if !SNFFT_TZ_F64
    // fp32: two adjacent instances per thread as one F2 (vec2.cuh) when the instance count and the buffers allow it;
    // SNFFT_TZ_ZV2=0 keeps the scalar threads (bit-identical results)
    static const bool v2_on = [] {
        const char* e = std::getenv("SNFFT_TZ_ZV2");
        return !e || std::atoi(e) != 0;
    }();
    static const int chunk2 = [] {
        const char* e = std::getenv("SNFFT_TZ_CHUNK2");
        // measured (profiles r03n, Z of level 9): forward 0.56 / 0.56 / 0.55 / 0.54 / 0.56 ms at 18 / 37 / 74 / 148 / 296
        // tiles, inverse 0.52 / 0.52 / 0.52 / 0.54 / 0.55 ms
        return e ? std::atoi(e) : (SNFFT_TZ_INV ? 37 : 148);
    }();
    v2 = v2_on && ninst % 2 == 0 && ((uintptr_t)src | (uintptr_t)dst) % sizeof(F2) == 0;
    if (v2) {
        ninst /= 2;
        if ((ninst & (ninst - 1)) == 0)
            for (s_shift2 = 0; (1ll << s_shift2) < ninst; ++s_shift2) {
            }
    }
    const int chunk_used = v2 ? chunk2 : chunk;
#else
    const int chunk_used = chunk;
#endif
    ZGroups g;
    g.ngroups = I::ngroups;
    g.chunk = chunk_used < 1 ? 1 : chunk_used;
    g.k = k;
    unsigned long long blk = 0;
    for (int i = 0; i < I::ngroups; ++i) {
        g.blk0[i] = (unsigned)blk;
        g.task0[i] = I::task0()[i];
        g.ntask[i] = I::ntask()[i];
        g.ubase[i] = d->ubase[i];
        g.lanes[i] = (unsigned long long)d->units[i] * (unsigned long long)ninst;
        blk += ((g.lanes[i] + Z_THREADS - 1) / Z_THREADS) * (unsigned long long)g.ntask[i];
    }
    g.blk0[I::ngroups] = (unsigned)blk;
    if (blk > 0x7fffffffull) return (int)cudaErrorInvalidValue;
    if (blk == 0) return 0;
    dim3 grid((unsigned)blk, 1, 1), block(Z_THREADS, 1, 1);
#if !SNFFT_TZ_F64
    if (v2) {
        auto kfn2 = z_kernel<F2, SNFFT_TZ_INV != 0>;
        SNFFT_LAUNCH(kfn2, grid, block, 0, st, (const F2*)src, (F2*)dst, (unsigned)ninst, s_shift2, g, (const ZSeg*)d->segs,
                     (const unsigned*)d->unit2seg);
        return (int)cudaGetLastError();
    }
#endif
    auto kfn = z_kernel<tz_t, SNFFT_TZ_INV != 0>;
    int s_shift = -1;
    if ((ninst & (ninst - 1)) == 0)
        for (s_shift = 0; (1ll << s_shift) < ninst; ++s_shift) {
        }
    SNFFT_LAUNCH(kfn, grid, block, 0, st, (const tz_t*)src, (tz_t*)dst, (unsigned)ninst, s_shift, g, (const ZSeg*)d->segs,
                 (const unsigned*)d->unit2seg);
    return (int)cudaGetLastError();
}

#else

#define TZ_TNAME_(k, dt, dir) snfft_tz_t##k##_##dt##_##dir
#define TZ_TNAME__(k, dt, dir) TZ_TNAME_(k, dt, dir)
#define TZ_TNAME TZ_TNAME__(SNFFT_TZ_KIND, TZ_DT, TZ_DIR)
#define TZ_TPKNAME_(k, dt, dir) snfft_tz_tpk##k##_##dt##_##dir
#define TZ_TPKNAME__(k, dt, dir) TZ_TPKNAME_(k, dt, dir)
#define TZ_TPKNAME TZ_TPKNAME__(SNFFT_TZ_KIND, TZ_DT, TZ_DIR)

// pk != 0 (top level only, NINST = B): xk is the public packed coefficient array [lambda][B][d][d] instead of the level
// array X_k -- the pack (forward) / unpack (inverse) pass is fused into this kernel (TzAcc, tz_kernels.cuh).  Needs the
// widest instance vector (B % 4 == 0 for fp32, B % 2 == 0 for fp64, 16-byte aligned z / x); returns
// cudaErrorInvalidValue otherwise (snfft.cu: tz_pack_ok decides before the call).
static int tz_t_launch(void* z, void* x, void* xk, long long ninst, int pk, cudaStream_t st) {
    using namespace snfft;
    using L = TzLevel<SNFFT_TZ_KIND>;
    static_assert(L::ntasks <= T_MAXTASKS, "too many T tasks");
    if (ninst <= 0 || ninst >= (1ll << 32)) return (int)cudaErrorInvalidValue;
    // instances per thread: 16-byte transfers when the instance count and the buffers allow it (SNFFT_TZ_VEC=1|2|4 caps it)
    static const int vec_cap = [] {
        const char* e = std::getenv("SNFFT_TZ_VEC");
        return e ? std::atoi(e) : TZ_VEC_DEFAULT;
    }();
    constexpr int nv_max = 16 / (int)sizeof(tz_t);
    int nv = nv_max;
    if (nv > vec_cap) nv = vec_cap < 1 ? 1 : vec_cap;
    const uintptr_t align_bits = pk ? ((uintptr_t)z | (uintptr_t)x) : ((uintptr_t)z | (uintptr_t)x | (uintptr_t)xk);
    while (nv > 1 && (ninst % nv != 0 || align_bits % (sizeof(tz_t) * nv) != 0)) nv >>= 1;
    if (pk && (nv != nv_max || (uintptr_t)xk % sizeof(tz_t) != 0)) return (int)cudaErrorInvalidValue;
    ninst /= nv;
    // columns per warp tile of the packed kernels (tz_kernels.cuh).  The tuning build (-DSNFFT_TZ_PK_TUNE) holds all three
    // widths and SNFFT_TZ_PK_CW = 8 | 16 | 32 picks one; the product build has the measured default only.
    static const int pk_cw = [] {
#ifdef SNFFT_TZ_PK_TUNE
        const char* e = std::getenv("SNFFT_TZ_PK_CW");
        const int v = e ? std::atoi(e) : TZ_PK_CW_DEFAULT;
        return v == 8 || v == 16 || v == 32 ? v : TZ_PK_CW_DEFAULT;
#else
        return TZ_PK_CW_DEFAULT;
#endif
    }();
    static const int pk_order = [] {
        // measured (S_10, B = 128): column tiles first 1.08 ms forward / 0.97 ms inverse, instance tiles first 1.12 / 0.97
        const char* e = std::getenv("SNFFT_TZ_PK_ORDER");
        return e ? std::atoi(e) : 1;
    }();
    TTasks g;
    g.ntasks = L::ntasks;
    g.pk_ct_fastest = pk_order;
    unsigned long long blk = 0;
    for (int t = 0; t < L::ntasks; ++t) {
        g.blk0[t] = (unsigned)blk;
        const unsigned long long q = t_task_lanes((unsigned)L::t_dal()[t], (unsigned)L::t_dmu()[t], (unsigned)ninst, pk ? (unsigned)pk_cw : 0u);
        blk += (q + T_THREADS - 1) / T_THREADS;
    }
    for (int t = L::ntasks; t <= T_MAXTASKS; ++t) g.blk0[t] = (unsigned)blk;
    if (blk > 0x7fffffffull) return (int)cudaErrorInvalidValue;
    g.hint_shift = 0;
    while ((blk >> g.hint_shift) >= (unsigned long long)T_HINTS) ++g.hint_shift;
    for (int j = 0, t = 0; j <= T_HINTS; ++j) {          // owner of block j << hint_shift (clamped to the last task)
        const unsigned long long b0 = (unsigned long long)j << g.hint_shift;
        while (t + 1 < L::ntasks && g.blk0[t + 1] <= b0) ++t;
        g.hint[j] = (unsigned short)t;
    }
    int s_shift = -1;
    if ((ninst & (ninst - 1)) == 0)
        for (s_shift = 0; (1ll << s_shift) < ninst; ++s_shift) {
        }
    dim3 grid((unsigned)blk, 1, 1), block(T_THREADS, 1, 1);
#if SNFFT_TZ_INV
#define TZ_T_LAUNCH(VT, PK, CW)                                                                           \
    {                                                                                                     \
        auto kfn = t_kernel<VT, SNFFT_TZ_KIND, true, PK, CW>;                                             \
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (VT*)z, (VT*)x, (const VT*)xk, (unsigned)ninst, s_shift, g);         \
    }
#else
#define TZ_T_LAUNCH(VT, PK, CW)                                                                           \
    {                                                                                                     \
        auto kfn = t_kernel<VT, SNFFT_TZ_KIND, false, PK, CW>;                                            \
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const VT*)z, (const VT*)x, (VT*)xk, (unsigned)ninst, s_shift, g);   \
    }
#endif
    using VMAX = TzVec<tz_t, nv_max>;
    if (pk) {
#ifdef SNFFT_TZ_PK_TUNE
        if (pk_cw == 8) TZ_T_LAUNCH(VMAX, true, 8)
        else if (pk_cw == 16) TZ_T_LAUNCH(VMAX, true, 16)
        else TZ_T_LAUNCH(VMAX, true, 32)
#else
        TZ_T_LAUNCH(VMAX, true, TZ_PK_CW_DEFAULT)
#endif
    } else if (nv == nv_max) {
        TZ_T_LAUNCH(VMAX, false, 16)
    } else if (nv == 2) {
        using V2 = TzVec<tz_t, 2>;
        TZ_T_LAUNCH(V2, false, 16)
    } else {
        TZ_T_LAUNCH(tz_t, false, 16)
    }
#undef TZ_T_LAUNCH
    return (int)cudaGetLastError();
}

extern "C" int TZ_TNAME(void* z, void* x, void* xk, long long ninst, cudaStream_t st) { return tz_t_launch(z, x, xk, ninst, 0, st); }
extern "C" int TZ_TPKNAME(void* z, void* x, void* xk, long long ninst, cudaStream_t st) { return tz_t_launch(z, x, xk, ninst, 1, st); }

#endif
