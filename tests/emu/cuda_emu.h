// Minimal CUDA-on-CPU shim used ONLY by the CPU test-suite (tests/): it lets
// g++ compile algebraist_b200/csrc/snfft.cu unchanged so that the kernel logic
// (indexing, tables, barriers) is checked against the oracle without a GPU.
// It is not part of the product and the product never loads a library built
// with it.  Threads of a block are ucontext fibers; __syncthreads() yields to
// a round-robin scheduler, which gives exact barrier semantics on one OS thread.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

inline dim3 threadIdx, blockIdx, blockDim, gridDim;

#define __shared__ static
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(x)

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };

inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
typedef int cudaEvent_t;
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = 0; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }

namespace emu {

inline unsigned char* dyn_smem = nullptr;

struct Sched {
    ucontext_t main_ctx;
    std::vector<ucontext_t> ctx;
    std::vector<unsigned char> stacks;
    std::vector<char> done;
    const std::function<void()>* body = nullptr;
    int cur = 0;
    // per emulated thread.  The generated straight-line task functions hold every temporary in their frame at -O1 (and
    // AddressSanitizer puts redzones around each: an F2 level-8 task overflowed 64 KB stacks in the sanitizer build, which
    // corrupted the heap next to the stack array before ASan reported it)
#if defined(__SANITIZE_ADDRESS__)
    static constexpr size_t kStack = 1024 * 1024;
#else
    static constexpr size_t kStack = 256 * 1024;
#endif
};
inline Sched g;

inline void trampoline() {
    (*g.body)();
    g.done[g.cur] = 1;
    swapcontext(&g.ctx[g.cur], &g.main_ctx);
}

inline void yield_barrier() { swapcontext(&g.ctx[g.cur], &g.main_ctx); }

// named barriers (bar.sync / bar.arrive id, count): a waiting fiber yields until the barrier's generation
// changes.  Kernels that use them must not rely on the one-yield __syncthreads afterwards.
struct NamedBar {
    int arrived = 0;
    unsigned gen = 0;
};
inline NamedBar named[16];
inline void named_reset() {
    for (NamedBar& b : named) b = NamedBar();
}
inline void named_arrive(int id, int count) {
    NamedBar& b = named[id];
    if (++b.arrived == count) {
        b.arrived = 0;
        ++b.gen;
    }
}
inline void named_sync(int id, int count) {
    NamedBar& b = named[id];
    const unsigned my = b.gen;
    if (++b.arrived == count) {
        b.arrived = 0;
        ++b.gen;
        return;
    }
    while (b.gen == my) yield_barrier();
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nt = (int)(block.x * block.y * block.z);
    std::vector<unsigned char> shared(smem + 64);
    dyn_smem = shared.data();
    if ((int)g.ctx.size() < nt) {
        g.ctx.resize(nt);
        g.stacks.resize((size_t)nt * Sched::kStack);
    }
    g.done.assign(nt, 0);
    g.body = &body;
    gridDim = grid;
    blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = dim3(bx, by, bz);
                named_reset();
                for (int t = 0; t < nt; ++t) {
                    getcontext(&g.ctx[t]);
                    g.ctx[t].uc_stack.ss_sp = g.stacks.data() + (size_t)t * Sched::kStack;
                    g.ctx[t].uc_stack.ss_size = Sched::kStack;
                    g.ctx[t].uc_link = &g.main_ctx;
                    makecontext(&g.ctx[t], trampoline, 0);
                    g.done[t] = 0;
                }
                int alive = nt;
                while (alive > 0) {
                    alive = 0;
                    for (int t = 0; t < nt; ++t) {
                        if (g.done[t]) continue;
                        g.cur = t;
                        threadIdx = dim3((unsigned)t % block.x, ((unsigned)t / block.x) % block.y,
                                         (unsigned)t / (block.x * block.y));
                        swapcontext(&g.main_ctx, &g.ctx[t]);
                        if (!g.done[t]) ++alive;
                    }
                }
            }
    dyn_smem = nullptr;
}

}  // namespace emu

inline void __syncthreads() { emu::yield_barrier(); }

#define SNFFT_CP_ASYNC(dst, src, bytes) std::memcpy(dst, src, bytes)
#define SNFFT_CP_ASYNC_WAIT() ((void)0)
#define SNFFT_CP_ASYNC_COMMIT() ((void)0)
#define SNFFT_CP_ASYNC_WAIT_PRIOR(n) ((void)0)
#define SNFFT_BAR_SYNC(id, count) emu::named_sync(id, count)
#define SNFFT_BAR_ARRIVE(id, count) emu::named_arrive(id, count)
#define SNFFT_FENCE_BLOCK() ((void)0)
#define SNFFT_DYN_SMEM(name) unsigned char* name = emu::dyn_smem
#define SNFFT_LAUNCH(kfn, grid, block, smem, stream, ...) \
    emu::launch(grid, block, smem, [&]() { kfn(__VA_ARGS__); })
