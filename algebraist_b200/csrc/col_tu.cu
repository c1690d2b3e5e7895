// One translation unit per (level, dtype, direction) of the column kernels (col_kernels.cuh): the generated
// code is large, so the build compiles the eight combinations in parallel.
//   -DSNFFT_COL_K=6|7|8|9  -DSNFFT_COL_F64=0|1  -DSNFFT_COL_INV=0|1  [-DSNFFT_COL_PART=0..3 for level 9, forward only]
// Exports  int snfft_col_launch_<K>[p<part>]_<f32|f64>_<fwd|inv>(in, out, ninst, stream)  (0 or a cudaError_t).
#ifdef SNFFT_HOST_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define SNFFT_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "col_kernels.cuh"

#if SNFFT_COL_F64
typedef double col_t;
#define COL_DT f64
#else
typedef float col_t;
#define COL_DT f32
#endif
#if SNFFT_COL_INV
#define COL_DIR inv
#else
#define COL_DIR fwd
#endif
#ifdef SNFFT_COL_PART
#define COL_CAT_(k, part, dt, dir) snfft_col_launch_##k##p##part##_##dt##_##dir
#define COL_CAT(k, part, dt, dir) COL_CAT_(k, part, dt, dir)
#define COL_NAME COL_CAT(SNFFT_COL_K, SNFFT_COL_PART, COL_DT, COL_DIR)
#else
#define COL_CAT_(k, dt, dir) snfft_col_launch_##k##_##dt##_##dir
#define COL_CAT(k, dt, dir) COL_CAT_(k, dt, dir)
#define COL_NAME COL_CAT(SNFFT_COL_K, COL_DT, COL_DIR)
#endif

extern "C" int COL_NAME(const void* in, void* out, long long ninst, cudaStream_t st) {
    using namespace snfft;
    static const int chunk = [] {
        // tiles of one task that run before the next task of the group (col_kernels.cuh); tuning switches:
        // SNFFT_COL_CHUNK_<K> for one level, SNFFT_COL_CHUNK for all
        char name[32];
        std::snprintf(name, sizeof name, "SNFFT_COL_CHUNK_%d", SNFFT_COL_K);
        const char* e = std::getenv(name);
        if (!e) e = std::getenv("SNFFT_COL_CHUNK");
        if (e) return std::atoi(e);
        // measured (profiles r01l/r01m), with the CTA sizes of _build.py (640 / 384 threads at levels 8 and 9, 128 below):
        // level 9 forward 4 tiles, level 8 16 (fp64: 8), level 7 inverse 74, otherwise one wave of 148
        return SNFFT_COL_K == 9 ? 4 : (SNFFT_COL_K == 8 ? (SNFFT_COL_F64 ? 8 : 16) : (SNFFT_COL_K == 7 && SNFFT_COL_INV ? 74 : 148));
    }();
#if !SNFFT_COL_F64
    // fp32: two adjacent instances per thread as one F2 (vec2.cuh: 8-byte transfers, FFMA2) when the instance count and
    // the buffers allow it; SNFFT_COL_V2=0 keeps the scalar threads (bit-identical results)
    static const bool v2_on = [] {
        const char* e = std::getenv("SNFFT_COL_V2");
        return !e || std::atoi(e) != 0;
    }();
    if (v2_on && ninst % 2 == 0 && ((uintptr_t)in | (uintptr_t)out) % sizeof(F2) == 0) {
        static const int chunk2 = [] {
            char name[32];
            std::snprintf(name, sizeof name, "SNFFT_COL_CHUNK2_%d", SNFFT_COL_K);
            const char* e = std::getenv(name);
            if (e) return std::atoi(e);
            // measured (profiles r03n): level 7 forward 18..296 tiles within 1 %, inverse 0.65 / 0.65 / 0.66 / 0.67 / 0.69 ms at
            // 18 / 37 / 74 / 148 / 296; level 8: 4..32 within 2 %
            return SNFFT_COL_K == 8 ? 8 : (SNFFT_COL_K == 7 && SNFFT_COL_INV ? 37 : 148);
        }();
        const long long n2 = ninst / 2;
        const ColGroups g2 = col_groups<SNFFT_COL_K, SNFFT_COL_INV != 0, COL_THREADS_V2>(n2, chunk2);
        if (n2 >= (1ll << 32) || g2.total > 0x7fffffffull) return (int)cudaErrorInvalidValue;
        if (g2.total == 0) return 0;
        dim3 grid2((unsigned)g2.total, 1, 1), block2(COL_THREADS_V2, 1, 1);
        auto kfn2 = col_kernel<F2, SNFFT_COL_K, SNFFT_COL_INV != 0, COL_THREADS_V2>;
        SNFFT_LAUNCH(kfn2, grid2, block2, 0, st, (const F2*)in, (F2*)out, n2, g2);
        return (int)cudaGetLastError();
    }
#endif
    const ColGroups g = col_groups<SNFFT_COL_K, SNFFT_COL_INV != 0>(ninst, chunk);
    // the kernels use a 32-bit row stride and a one-dimensional grid
    if (ninst >= (1ll << 32) || g.total > 0x7fffffffull) return (int)cudaErrorInvalidValue;
    const unsigned nblk = (unsigned)g.total;
    if (nblk == 0) return 0;
    dim3 grid(nblk, 1, 1), block(COL_THREADS, 1, 1);
    auto kfn = col_kernel<col_t, SNFFT_COL_K, SNFFT_COL_INV != 0>;
    SNFFT_LAUNCH(kfn, grid, block, 0, st, (const col_t*)in, (col_t*)out, ninst, g);
    return (int)cudaGetLastError();
}
