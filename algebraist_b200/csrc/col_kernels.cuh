// Column kernels for the level steps k = 6, 7, 8 (generated code col_gen_<k>.cuh, gen_col.py).
// A thread is (task, column c, instance): it owns the rows of one output column that the bottom sweeps keep
// together (<= 35 values) in registers, reads its inputs straight from the level array below (lanes over
// consecutive instances: every global access is a contiguous run) and writes its rows once.  Forward task:
// rows R of F_lambda[:, boff_b + c] (fourier.py:253-273); inverse task: rows R of G^(i)_mu[:, c] for every
// coset i (fourier.py:191-199).  No shared memory, no barriers, no tables.
//
// Block order: the tasks that read the same child block mu ("group") are interleaved in chunks of `chunk` tiles:
// the CTAs of one chunk of one task come first, then the same tiles for the next task of the group, ...  A chunk
// is small enough for its inputs to stay in L2 until the other tasks of the group re-read them and large enough
// for the CTAs that are resident on one SM to run the same straight-line code (instruction cache reuse).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef SNFFT_COL_K
#error "col_kernels.cuh: define SNFFT_COL_K (6..8) -- one translation unit per level"
#endif
#if SNFFT_COL_K == 6
#include "col_gen_6.cuh"
#elif SNFFT_COL_K == 7
#include "col_gen_7.cuh"
#elif SNFFT_COL_K == 8
#include "col_gen_8.cuh"
#endif

namespace snfft {
namespace {        // internal linkage, see col_gen_<k>.cuh

#ifndef SNFFT_COL_THREADS
#define SNFFT_COL_THREADS 128
#endif
constexpr int COL_THREADS = SNFFT_COL_THREADS;     // per level (build flag): large CTAs keep the warps of an SM on one task's code
constexpr int COL_MAXGROUPS = 24;

struct ColGroups {
    int ngroups, chunk;
    unsigned long long total;             // blocks of the launch (must fit a 31-bit grid)
    unsigned blk0[COL_MAXGROUPS + 1];     // first block of each group
    int task0[COL_MAXGROUPS], ntask[COL_MAXGROUPS], dmu[COL_MAXGROUPS];
};

// threads per CTA of the F2 instantiation (two instances per thread, the register footprint of the fp64 one)
#ifndef SNFFT_COL_THREADS_V2
#define SNFFT_COL_THREADS_V2 (SNFFT_COL_THREADS > 384 ? 384 : SNFFT_COL_THREADS)
#endif
constexpr int COL_THREADS_V2 = SNFFT_COL_THREADS_V2;

template <int K, bool INV, int NT = COL_THREADS>
inline ColGroups col_groups(long long ninst, int chunk) {
    using I = ColInfo<K, INV>;
    static_assert(I::ngroups <= COL_MAXGROUPS, "too many groups");
    ColGroups g;
    g.ngroups = I::ngroups;
    g.chunk = chunk < 1 ? 1 : chunk;
    unsigned long long blk = 0;
    for (int i = 0; i < I::ngroups; ++i) {
        g.blk0[i] = (unsigned)blk;
        g.task0[i] = I::task0()[i];
        g.ntask[i] = I::ntask()[i];
        g.dmu[i] = I::dmu()[i];
        const long long tiles = ((long long)g.dmu[i] * ninst + NT - 1) / NT;
        blk += (unsigned long long)tiles * g.ntask[i];
    }
    for (int i = I::ngroups; i <= COL_MAXGROUPS; ++i) g.blk0[i] = (unsigned)blk;
    g.total = blk;
    return g;
}

// Register cap: ptxas hoists the (read-only) loads of a whole task as far as the register file allows, which without
// a cap means 255 registers and spills; MINB CTAs per SM bound it (fp32: 102 registers, fp64: 170).
#ifndef SNFFT_COL_V2_REGS_THREADS
#define SNFFT_COL_V2_REGS_THREADS 384      // resident threads per SM the F2 / fp64 instantiations are compiled for (register cap 65536 / this)
#endif
template <typename T, int K, bool INV, int NT = COL_THREADS>
__global__ void __launch_bounds__(NT, (sizeof(T) == 4 ? 640 : SNFFT_COL_V2_REGS_THREADS) / NT) col_kernel(const T* __restrict__ in, T* __restrict__ out, long long ninst,
                                                          const ColGroups g) {
    const unsigned b = blockIdx.x;
    int gi = 0;
    while (gi + 1 < g.ngroups && b >= g.blk0[gi + 1]) ++gi;
    const unsigned local = b - g.blk0[gi];
    const unsigned nt = (unsigned)g.ntask[gi];
    const long long qn = (long long)g.dmu[gi] * ninst;
    const unsigned ntiles = (unsigned)((qn + NT - 1) / NT);
    const unsigned per = (unsigned)g.chunk * nt;              // blocks of a full chunk
    const unsigned ch = local / per, r = local - ch * per;
    const unsigned left = ntiles - ch * (unsigned)g.chunk;    // tiles of this chunk (the last one may be short)
    const unsigned cht = left < (unsigned)g.chunk ? left : (unsigned)g.chunk;
    const unsigned tk = r / cht;
    const unsigned tile = ch * (unsigned)g.chunk + (r - tk * cht);
    const int task = g.task0[gi] + (int)tk;
    const long long q = (long long)tile * NT + threadIdx.x;
    if (q >= qn) return;
    const int c = (int)(q / ninst);
    const long long inst = q - (long long)c * ninst;
    // 32-bit row stride: every element address is ONE wide multiply-add (stride * constant + base)
    col_run<T, K, INV>(task, c, in + inst, out + inst, (unsigned)ninst);
}

}  // namespace
}  // namespace snfft
