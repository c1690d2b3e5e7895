// Translation unit of the fused levels-2..6 kernel (s6_kernels.cuh), fp32:  -DSNFFT_S6_INV=0|1
// Exports  int snfft_s6_launch_f32_<fwd|inv>(in, out, ninst6, stream)  (0 or a cudaError_t):
//   forward: in = X_1 (gathered input, [120][NINST_5]), out = X_6;  inverse: in = X_6, out = X_1.
#ifdef SNFFT_HOST_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define SNFFT_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SNFFT_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

#include "s6_kernels.cuh"

#if SNFFT_S6_INV
#define S6_NAME snfft_s6_launch_f32_inv
#else
#define S6_NAME snfft_s6_launch_f32_fwd
#endif

extern "C" int S6_NAME(const void* in, void* out, long long ninst6, cudaStream_t st) {
    using namespace snfft;
    if (ninst6 <= 0) return 0;
    // 32-bit instance stride and grid
    if (ninst6 >= (1ll << 31)) return (int)cudaErrorInvalidValue;
    auto kfn = s6_kernel<float, SNFFT_S6_INV != 0>;
    const size_t smem = S6_SMEM_ELEMS * sizeof(float);
#ifndef SNFFT_HOST_EMU
    {   // the attribute is per device; setting it twice is harmless (two threads may race to the flag)
        static bool done[64] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return (int)e;
        if (dev < 0 || dev >= 64 || !done[dev]) {
            if ((e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
    }
#endif
    dim3 grid((unsigned)((ninst6 + S6_LANES - 1) / S6_LANES), 1, 1), block(S6_THREADS, 1, 1);
    SNFFT_LAUNCH(kfn, grid, block, smem, st, (const float*)in, (float*)out, (unsigned)ninst6);
    return (int)cudaGetLastError();
}
