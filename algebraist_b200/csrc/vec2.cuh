// Fused multiply-add spelling of the generated code, and the packed pair of fp32 instances.
//
// The generators (gen_base.py, gen_col.py, gen_tz.py) emit every rotation output as  sn_fma(T(coef), a, b).  For
// T = float / double that is the FFMA / DFMA the compiler contracted `coef * a + b` into before.  For T = F2 -- two
// ADJACENT INSTANCES of a level array in one 64-bit register pair -- it is ONE `fma.rn.f32x2` (SASS FFMA2, sm_100+;
// the coefficient stays an immediate): the level kernels of k >= 7 are bound by instruction issue, not by HBM
// (profiles r02h: issue slots 45-48 % busy at 29 % occupancy, 48 warp instructions per output element at level 8),
// and with F2 threads every load, store, address computation and FMA instruction does the work of two.
// The arithmetic per instance is unchanged (same operations, same order, round-to-nearest): results are bit-identical
// to the scalar instantiation.
#pragma once
#include <math.h>

#ifndef SNFFT_HD
#if defined(__CUDACC__) || defined(SNFFT_HOST_EMU)
#define SNFFT_HD __host__ __device__ __forceinline__
#else
#define SNFFT_HD inline
#endif
#endif

namespace snfft {

SNFFT_HD float sn_fma(float a, float b, float c) { return fmaf(a, b, c); }
SNFFT_HD double sn_fma(double a, double b, double c) { return fma(a, b, c); }

struct alignas(8) F2 {
    float x, y;
    F2() = default;
    SNFFT_HD explicit F2(double v) : x((float)v), y((float)v) {}
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float2 f2_raw(F2 a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ F2 f2_wrap(float2 r) {
    F2 o;
    o.x = r.x;
    o.y = r.y;
    return o;
}
__device__ __forceinline__ F2 sn_fma(F2 a, F2 b, F2 c) { return f2_wrap(__ffma2_rn(f2_raw(a), f2_raw(b), f2_raw(c))); }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { return f2_wrap(__fadd2_rn(f2_raw(a), f2_raw(b))); }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return f2_wrap(__ffma2_rn(f2_raw(b), make_float2(-1.f, -1.f), f2_raw(a))); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { return f2_wrap(__fmul2_rn(f2_raw(a), f2_raw(b))); }
__device__ __forceinline__ F2 operator-(F2 a) { return f2_wrap(__fmul2_rn(f2_raw(a), make_float2(-1.f, -1.f))); }
#else
SNFFT_HD F2 f2_make(float x, float y) {
    F2 o;
    o.x = x;
    o.y = y;
    return o;
}
SNFFT_HD F2 sn_fma(F2 a, F2 b, F2 c) { return f2_make(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
SNFFT_HD F2 operator+(F2 a, F2 b) { return f2_make(a.x + b.x, a.y + b.y); }
SNFFT_HD F2 operator-(F2 a, F2 b) { return f2_make(a.x - b.x, a.y - b.y); }
SNFFT_HD F2 operator*(F2 a, F2 b) { return f2_make(a.x * b.x, a.y * b.y); }
SNFFT_HD F2 operator-(F2 a) { return f2_make(-a.x, -a.y); }
#endif

}  // namespace snfft
