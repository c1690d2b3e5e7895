// Levels 2..6 in ONE kernel (fp32): the generated levels-2..5 code (base_gen.cuh: one S_5 instance per thread, 120 values
// in registers) and the level-6 child-major column tasks (s6_gen.cuh, gen_col.py) exchange the S_5 coefficients through a
// shared-memory tile instead of a level array in HBM -- one pass over the data instead of two
// (reference: the recursion of fourier.py:253-273 between the base case and level 6; inverse fourier.py:191-199).
//
// A CTA owns S6_LANES = 32 consecutive instances of level 6 and has 6 warps:
//   forward : warp i runs levels 2..5 of coset i of those instances (lane = instance; global reads are 128-byte rows of
//             X_1) and writes its 120 coefficients to the tile  sm[(e * 6 + i) * 32 + lane]   (e = coefficient of S_5);
//             after one barrier warp w runs the level-6 tasks (mu, c) number w, w + 6, ... (sorted by cost) from the
//             tile and writes its rows of X_6 (128-byte rows again).
//   inverse : the same in the opposite order (level-6 adjoint tasks fill the tile, then levels 5..2 per coset).
// The tile is 720 * 32 * 4 = 90 KB: two CTAs per SM, one in each phase most of the time.  Every global element moves
// once; shared-memory accesses are conflict free (lanes are consecutive words).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "base_gen.cuh"
#include "s6_gen.cuh"

namespace snfft {
namespace {

constexpr int S6_THREADS = 6 * S6_LANES;
constexpr size_t S6_SMEM_ELEMS = 720 * (size_t)S6_LANES;

template <typename T, bool INV>
__global__ void __launch_bounds__(S6_THREADS, 2) s6_kernel(const T* __restrict__ in, T* __restrict__ out, unsigned ninst6) {
    SNFFT_DYN_SMEM(smem_raw);
    T* sm = reinterpret_cast<T*>(smem_raw);
    const unsigned lane = threadIdx.x % S6_LANES, w = threadIdx.x / S6_LANES;
    const unsigned inst6 = blockIdx.x * S6_LANES + lane;
    const bool ok = inst6 < ninst6;
    const size_t ninst5 = 6 * (size_t)ninst6;
    if constexpr (!INV) {
        if (ok) {
            T x[120], y[120];
            const T* __restrict__ src = in + (size_t)w * ninst6 + inst6;        // coset i = w: inst5 = i * NINST_6 + inst6
#pragma unroll
            for (int p = 0; p < 120; ++p) x[p] = src[(size_t)p * ninst5];
            s5_forward<T>(x, y);
#pragma unroll
            for (int e = 0; e < 120; ++e) sm[(e * 6 + w) * S6_LANES + lane] = y[e];
        }
        __syncthreads();
        if (ok) {
            for (int j = w; j < S6_NITEMS; j += 6) {
                const unsigned c = s6f_item_c[j], off = s6f_item_off[j];
                s6f_run<T>(s6f_item_task[j], (const T*)(sm + off * 6 * S6_LANES + lane), out + (size_t)c * ninst6 + inst6, ninst6);
            }
        }
    } else {
        // (one straight-line function per warp with the next column's loads issued ahead was measured: 1.57 ms against
        // 1.02 ms -- six different code streams per CTA do not fit the instruction caches)
        if (ok) {
            for (int j = w; j < S6_NITEMS; j += 6) {
                const unsigned c = s6i_item_c[j], off = s6i_item_off[j];
                s6i_run<T>(s6i_item_task[j], in + (size_t)c * ninst6 + inst6, sm + off * 6 * S6_LANES + lane, ninst6);
            }
        }
        __syncthreads();
        if (ok) {
            T x[120], y[120];
#pragma unroll
            for (int e = 0; e < 120; ++e) x[e] = sm[(e * 6 + w) * S6_LANES + lane];
            s5_inverse<T>(x, y);
            T* __restrict__ dst = out + (size_t)w * ninst6 + inst6;
#pragma unroll
            for (int p = 0; p < 120; ++p) dst[(size_t)p * ninst5] = y[p];
        }
    }
}

}  // namespace
}  // namespace snfft
