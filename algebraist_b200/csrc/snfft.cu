// C ABI of the B200-native S_n FFT (include/snfft.h): plan upload, launch
// logic and the extern "C" entry points.  Compiled by nvcc for sm_100a; the
// same file is compiled by g++ against tests/emu/cuda_emu.h (SNFFT_HOST_EMU)
// by the CPU test-suite only, to check the kernel logic without a GPU.
#include "../../include/snfft.h"

#ifdef SNFFT_HOST_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define SNFFT_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SNFFT_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "plan.h"

namespace snfft {

// ---------------------------------------------------------------------------
// kernel classes: (id, dmax, W lanes, G row groups, RPT rows per thread, NG column groups per CTA)
// A shape of dimension d runs in the first class with dmax >= d.
// ---------------------------------------------------------------------------
#define SNFFT_CLASSES_F32(X)                                                                   \
    X(0, 16, 32, 1, 16, 8) X(1, 96, 32, 4, 24, 1) X(2, 256, 32, 8, 32, 1) X(3, 768, 32, 16, 48, 1) \
    X(4, 2310, 16, 32, 73, 1) X(5, 7700, 4, 128, 61, 1)
#define SNFFT_CLASSES_F64(X)                                                                   \
    X(0, 16, 32, 1, 16, 8) X(1, 96, 32, 4, 24, 1) X(2, 256, 32, 8, 32, 1) X(3, 768, 32, 16, 48, 1) \
    X(4, 2310, 8, 64, 37, 1) X(5, 7700, 2, 256, 31, 1)

struct KClass {
    int id, dmax, W, G, RPT, NG;
};
#define SNFFT_ROW(id, dmax, W, G, RPT, NG) {id, dmax, W, G, RPT, NG},
static const KClass kClasses32[] = {SNFFT_CLASSES_F32(SNFFT_ROW)};
static const KClass kClasses64[] = {SNFFT_CLASSES_F64(SNFFT_ROW)};
static const int kNumClasses = 6;

static std::atomic<int> g_force_class{-1};
static std::atomic<long long> g_launches{0};
static thread_local std::string g_error;

static int set_error(int code, const std::string& msg) {
    g_error = msg;
    return code;
}

#define SNFFT_CUDA(call)                                                                      \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return set_error(SNFFT_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

static const KClass* class_table(int dtype) { return dtype == SNFFT_F32 ? kClasses32 : kClasses64; }

static int pick_class(int dtype, int d) {
    const KClass* tab = class_table(dtype);
    const int forced = g_force_class.load();
    if (forced >= 0 && forced < kNumClasses && d <= tab[forced].dmax) return forced;
    for (int c = 0; c < kNumClasses; ++c)
        if (d <= tab[c].dmax) return c;
    return -1;
}

struct TaskRange {
    int cls, begin, count, max_dmu, max_dlam;
    long long elems;      // coefficients this launch produces per instance (its share of k!)
};

struct DevLevel {
    void *ops = nullptr, *pair_rows = nullptr, *pair_coef = nullptr, *neg_rows = nullptr;
    void *ftasks = nullptr, *itasks = nullptr, *parents = nullptr;
    std::vector<TaskRange> franges, iranges;
};

}  // namespace snfft

using namespace snfft;

struct ProfEvent {
    cudaEvent_t e0, e1;
    int kind, k, cls;
    long long elems;
};

struct snfft_plan {
    bool prof_on = false;
    std::vector<ProfEvent> prof;
    int n = 0, dtype = 0, device = 0;
    size_t esize = 4;
    HostPlan host;
    std::vector<DevLevel> lev;      // lev[k], k = 2..n
    std::vector<void*> allocs;      // every device allocation of the tables
    IndexArgs ia;
    // scratch of the *_host entry points
    void *h_in = nullptr, *h_out = nullptr, *h_ws = nullptr;
    size_t h_io_bytes = 0, h_ws_bytes = 0;
};

namespace snfft {

template <typename V>
static int upload(snfft_plan* p, const std::vector<V>& v, void** out) {
    *out = nullptr;
    const size_t bytes = std::max<size_t>(v.size() * sizeof(V), 16);
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return set_error(SNFFT_ENOMEM, "cudaMalloc of plan table failed");
    p->allocs.push_back(d);
    if (!v.empty()) SNFFT_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
    *out = d;
    return SNFFT_OK;
}

// Build and upload the device tables of level k.
template <typename T>
static int build_level(snfft_plan* p, int k) {
    const HostLevel& L = p->host.levels[k];
    const HostLevel& P = p->host.levels[k - 1];
    DevLevel& D = p->lev[k];
    const int ns = (int)L.shapes.size();

    std::vector<OpRange> ops((size_t)ns * (k - 1));
    std::vector<uint32_t> pair_rows, neg_rows;
    std::vector<T> pair_coef;
    for (int s = 0; s < ns; ++s) {
        const HostShape& S = L.shapes[s];
        for (int j = 0; j <= k - 2; ++j) {
            OpRange r;
            r.pair_off = (int)pair_rows.size();
            r.neg_off = (int)neg_rows.size();
            for (int t = 0; t < S.dim; ++t) {
                const int q = S.partner[j][t];
                if (q > t) {
                    pair_rows.push_back((uint32_t)t | ((uint32_t)q << 16));
                    pair_coef.push_back((T)S.diag[j][t]);
                    pair_coef.push_back((T)S.off[j][t]);
                } else if (q == t) {
                    if (S.diag[j][t] == -1.0) neg_rows.push_back((uint32_t)t);
                    else if (S.diag[j][t] != 1.0) return set_error(SNFFT_EINVAL, "YOR table: unpaired row with |diag| != 1");
                }
            }
            r.npairs = (int)pair_rows.size() - r.pair_off;
            r.nneg = (int)neg_rows.size() - r.neg_off;
            ops[(size_t)s * (k - 1) + j] = r;
        }
    }

    // forward tasks: one per (lambda, child block), grouped by kernel class of d_lambda
    std::vector<std::vector<FwdTask>> fby(kNumClasses);
    for (int s = 0; s < ns; ++s) {
        const HostShape& S = L.shapes[s];
        const int cls = pick_class(p->dtype, S.dim);
        if (cls < 0) return set_error(SNFFT_ENOTSUP, "irrep dimension exceeds the largest kernel class");
        for (size_t b = 0; b < S.child.size(); ++b) {
            const HostShape& M = P.shapes[S.child[b]];
            FwdTask t;
            t.lam_off = S.coef_off;
            t.mu_off = M.coef_off;
            t.d_lam = S.dim;
            t.d_mu = M.dim;
            t.boff = S.boff[b];
            t.lam = s;
            fby[cls].push_back(t);
        }
    }
    std::vector<FwdTask> ftasks;
    for (int c = 0; c < kNumClasses; ++c) {
        if (fby[c].empty()) continue;
        // big columns first: the long CTAs start early
        std::stable_sort(fby[c].begin(), fby[c].end(),
                         [](const FwdTask& x, const FwdTask& y) { return x.d_lam > y.d_lam; });
        TaskRange r{c, (int)ftasks.size(), (int)fby[c].size(), 0, 0, 0};
        for (const FwdTask& t : fby[c]) {
            r.elems += (long long)t.d_lam * t.d_mu;
            r.max_dmu = std::max(r.max_dmu, t.d_mu);
            r.max_dlam = std::max(r.max_dlam, t.d_lam);
            ftasks.push_back(t);
        }
        D.franges.push_back(r);
    }

    // inverse tasks: one per mu of level k-1, class by the largest covering lambda
    std::vector<std::vector<InvTask>> iby(kNumClasses);
    std::vector<std::vector<int>> iby_maxd(kNumClasses);
    std::vector<InvParent> parents;
    for (size_t m = 0; m < P.shapes.size(); ++m) {
        const HostShape& M = P.shapes[m];
        InvTask t;
        t.mu_off = M.coef_off;
        t.d_mu = M.dim;
        t.par_off = (int)parents.size();
        t.npar = (int)M.parent.size();
        t.pad = 0;
        int maxd = 0;
        for (size_t q = 0; q < M.parent.size(); ++q) {
            const HostShape& S = L.shapes[M.parent[q]];
            InvParent ip;
            ip.lam_off = S.coef_off;
            ip.d_lam = S.dim;
            ip.boff = S.boff[M.parent_block[q]];
            ip.lam = M.parent[q];
            ip.pad = 0;
            ip.scale = (double)S.dim / ((double)k * (double)M.dim);
            parents.push_back(ip);
            maxd = std::max(maxd, S.dim);
        }
        const int cls = pick_class(p->dtype, maxd);
        if (cls < 0) return set_error(SNFFT_ENOTSUP, "irrep dimension exceeds the largest kernel class");
        iby[cls].push_back(t);
        iby_maxd[cls].push_back(maxd);
    }
    std::vector<InvTask> itasks;
    for (int c = 0; c < kNumClasses; ++c) {
        if (iby[c].empty()) continue;
        TaskRange r{c, (int)itasks.size(), (int)iby[c].size(), 0, 0, 0};
        for (size_t q = 0; q < iby[c].size(); ++q) {
            r.elems += (long long)k * iby[c][q].d_mu * iby[c][q].d_mu;
            r.max_dmu = std::max(r.max_dmu, iby[c][q].d_mu);
            r.max_dlam = std::max(r.max_dlam, iby_maxd[c][q]);
            itasks.push_back(iby[c][q]);
        }
        D.iranges.push_back(r);
    }

    int rc;
    if ((rc = upload(p, ops, &D.ops))) return rc;
    if ((rc = upload(p, pair_rows, &D.pair_rows))) return rc;
    if ((rc = upload(p, pair_coef, &D.pair_coef))) return rc;
    if ((rc = upload(p, neg_rows, &D.neg_rows))) return rc;
    if ((rc = upload(p, ftasks, &D.ftasks))) return rc;
    if ((rc = upload(p, itasks, &D.itasks))) return rc;
    if ((rc = upload(p, parents, &D.parents))) return rc;
    return SNFFT_OK;
}

// Optional per-launch CUDA-event timing (snfft_profile_begin / snfft_profile_end).
struct ProfScope {
    const snfft_plan* p;
    cudaStream_t st;
    ProfEvent ev;
    bool on;
    ProfScope(const snfft_plan* p_, cudaStream_t st_, int kind, int k, int cls, long long elems)
        : p(p_), st(st_), on(p_->prof_on) {
        if (!on) return;
        ev.kind = kind;
        ev.k = k;
        ev.cls = cls;
        ev.elems = elems;
        if (cudaEventCreate(&ev.e0) != cudaSuccess || cudaEventCreate(&ev.e1) != cudaSuccess) {
            on = false;
            return;
        }
        cudaEventRecord(ev.e0, st);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(ev.e1, st);
        const_cast<snfft_plan*>(p)->prof.push_back(ev);
    }
};

static void tile_shape(int W, long long B, int* cw, int* ci) {
    int c = 1;
    while (c * 2 <= std::min(4, W) && c * 2 <= B) c *= 2;
    *ci = c;
    *cw = W / c;
}

template <typename KF>
static int set_smem(KF kfn, size_t bytes) {
#ifndef SNFFT_HOST_EMU
    if (bytes > 48 * 1024)
        SNFFT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
#else
    (void)kfn;
    (void)bytes;
#endif
    return SNFFT_OK;
}

template <typename T, int W, int G, int RPT, int NG>
static int launch_fwd_range(const TaskRange& r, bool final_level, LevelArgs<T> a, cudaStream_t st) {
    long long chunks;
    if (final_level) {
        tile_shape(W, a.batch, &a.cw, &a.ci);
        chunks = ((r.max_dmu + a.cw - 1) / a.cw) * ((a.batch + a.ci - 1) / a.ci);
    } else {
        a.cw = W;
        a.ci = 1;
        chunks = ((long long)r.max_dmu * a.ninst + W - 1) / W;
    }
    a.gstride = r.max_dlam * W;
    a.ftasks += r.begin;
    const size_t smem = (size_t)NG * a.gstride * sizeof(T);
    dim3 grid((unsigned)((chunks + NG - 1) / NG), (unsigned)r.count, 1), block(W * G * NG, 1, 1);
    int rc;
    if (final_level) {
        auto kfn = fwd_level_kernel<T, W, G, RPT, NG, true>;
        if ((rc = set_smem(kfn, smem))) return rc;
        SNFFT_LAUNCH(kfn, grid, block, smem, st, a);
    } else {
        auto kfn = fwd_level_kernel<T, W, G, RPT, NG, false>;
        if ((rc = set_smem(kfn, smem))) return rc;
        SNFFT_LAUNCH(kfn, grid, block, smem, st, a);
    }
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

template <typename T, int W, int G, int RPT, int NG>
static int launch_inv_range(const TaskRange& r, bool first_level, LevelArgs<T> a, cudaStream_t st) {
    long long chunks;
    if (first_level) {
        tile_shape(W, a.batch, &a.cw, &a.ci);
        chunks = ((r.max_dmu + a.cw - 1) / a.cw) * ((a.batch + a.ci - 1) / a.ci);
    } else {
        a.cw = W;
        a.ci = 1;
        chunks = ((long long)r.max_dmu * a.ninst + W - 1) / W;
    }
    a.gstride = r.max_dlam * W;
    a.itasks += r.begin;
    const size_t smem = (size_t)NG * a.gstride * sizeof(T);
    dim3 grid((unsigned)((chunks + NG - 1) / NG), (unsigned)r.count, (unsigned)a.k), block(W * G * NG, 1, 1);
    int rc;
    if (first_level) {
        auto kfn = inv_level_kernel<T, W, G, RPT, NG, true>;
        if ((rc = set_smem(kfn, smem))) return rc;
        SNFFT_LAUNCH(kfn, grid, block, smem, st, a);
    } else {
        auto kfn = inv_level_kernel<T, W, G, RPT, NG, false>;
        if ((rc = set_smem(kfn, smem))) return rc;
        SNFFT_LAUNCH(kfn, grid, block, smem, st, a);
    }
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

template <typename T>
static LevelArgs<T> level_args(const snfft_plan* p, int k, const void* in, void* out, long long B) {
    const DevLevel& D = p->lev[k];
    LevelArgs<T> a;
    a.in = (const T*)in;
    a.out = (T*)out;
    a.ftasks = (const FwdTask*)D.ftasks;
    a.itasks = (const InvTask*)D.itasks;
    a.parents = (const InvParent*)D.parents;
    a.ops = (const OpRange*)D.ops;
    a.pair_rows = (const uint32_t*)D.pair_rows;
    a.pair_coef = (const T*)D.pair_coef;
    a.neg_rows = (const uint32_t*)D.neg_rows;
    a.k = k;
    a.cw = a.ci = a.gstride = 0;
    a.ninst = B * (p->host.factorial[p->n] / p->host.factorial[k]);
    a.batch = B;
    return a;
}

template <typename T>
static int run_fwd_level(const snfft_plan* p, int k, const void* in, void* out, long long B, cudaStream_t st) {
    const LevelArgs<T> a = level_args<T>(p, k, in, out, B);
    const bool fin = (k == p->n);
    for (const TaskRange& r : p->lev[k].franges) {
        int rc = SNFFT_ENOTSUP;
        ProfScope prof(p, st, 0, k, r.cls, r.elems);
        if constexpr (sizeof(T) == 4) {
            switch (r.cls) {
#define SNFFT_CASE(id, dmax, W, G, RPT, NG) \
    case id: rc = launch_fwd_range<T, W, G, RPT, NG>(r, fin, a, st); break;
                SNFFT_CLASSES_F32(SNFFT_CASE)
            }
        } else {
            switch (r.cls) { SNFFT_CLASSES_F64(SNFFT_CASE) }
#undef SNFFT_CASE
        }
        if (rc) return rc;
    }
    return SNFFT_OK;
}

template <typename T>
static int run_inv_level(const snfft_plan* p, int k, const void* in, void* out, long long B, cudaStream_t st) {
    const LevelArgs<T> a = level_args<T>(p, k, in, out, B);
    const bool first = (k == p->n);
    for (const TaskRange& r : p->lev[k].iranges) {
        int rc = SNFFT_ENOTSUP;
        ProfScope prof(p, st, 1, k, r.cls, r.elems);
        if constexpr (sizeof(T) == 4) {
            switch (r.cls) {
#define SNFFT_CASE(id, dmax, W, G, RPT, NG) \
    case id: rc = launch_inv_range<T, W, G, RPT, NG>(r, first, a, st); break;
                SNFFT_CLASSES_F32(SNFFT_CASE)
            }
        } else {
            switch (r.cls) { SNFFT_CLASSES_F64(SNFFT_CASE) }
#undef SNFFT_CASE
        }
        if (rc) return rc;
    }
    return SNFFT_OK;
}

template <typename T, bool SCATTER>
static int run_gather(const snfft_plan* p, const void* in, void* out, long long B, cudaStream_t st) {
    IndexArgs ia = p->ia;
    ia.batch = B;
    dim3 grid((unsigned)((ia.nfact + GS_TR - 1) / GS_TR), (unsigned)((B + GS_TB - 1) / GS_TB), 1);
    dim3 block(GS_THREADS, 1, 1);
    ProfScope prof(p, st, SCATTER ? 3 : 2, 1, -1, ia.nfact);
    auto kfn = gather_kernel<T, SCATTER>;
    SNFFT_LAUNCH(kfn, grid, block, gather_smem_bytes<T>(), st, (const T*)in, (T*)out, ia);
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

template <typename T>
static int forward_impl(const snfft_plan* p, const void* f, long long B, void* F, void* ws, cudaStream_t st) {
    const size_t half = (size_t)B * p->host.factorial[p->n] * sizeof(T);
    void* cur = ws;
    void* nxt = (char*)ws + half;
    int rc;
    if ((rc = run_gather<T, false>(p, f, cur, B, st))) return rc;
    for (int k = 2; k <= p->n; ++k) {
        void* out = (k == p->n) ? F : nxt;
        if ((rc = run_fwd_level<T>(p, k, cur, out, B, st))) return rc;
        std::swap(cur, nxt);
    }
    return SNFFT_OK;
}

template <typename T>
static int inverse_impl(const snfft_plan* p, const void* F, long long B, void* f, void* ws, cudaStream_t st) {
    const size_t half = (size_t)B * p->host.factorial[p->n] * sizeof(T);
    const void* cur = F;
    void* bufs[2] = {ws, (char*)ws + half};
    int which = 0, rc;
    for (int k = p->n; k >= 2; --k) {
        if ((rc = run_inv_level<T>(p, k, cur, bufs[which], B, st))) return rc;
        cur = bufs[which];
        which ^= 1;
    }
    return run_gather<T, true>(p, cur, f, B, st);
}

static IndexArgs make_index_args(int n) {
    IndexArgs ia;
    ia.n = n;
    ia.fact[0] = 1;
    for (int i = 1; i <= 20; ++i) ia.fact[i] = ia.fact[i - 1] * i;
    ia.nfact = ia.fact[n];
    ia.batch = 1;
    return ia;
}

}  // namespace snfft

// ---------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------
extern "C" {

const char* snfft_build_info(void) {
#ifdef SNFFT_HOST_EMU
    return "snfft host-emulation build (tests only)";
#else
    return "snfft sm_100a";
#endif
}

const char* snfft_last_error(void) { return g_error.c_str(); }

int64_t snfft_launch_count(void) { return (int64_t)g_launches.load(); }

// test hook: run every shape whose dimension fits in kernel class `cls` (0..5; -1 = automatic);
// affects plans created afterwards.
void snfft_debug_force_class(int cls) { g_force_class.store(cls); }

int snfft_plan_create(int n, int dtype, int device, snfft_plan_t** out) {
    if (!out) return set_error(SNFFT_EINVAL, "out is NULL");
    *out = nullptr;
    if (n < 2 || n > SNFFT_MAX_N) return set_error(SNFFT_EINVAL, "n must be in [2, 12]");
    if (dtype != SNFFT_F32 && dtype != SNFFT_F64) return set_error(SNFFT_EINVAL, "dtype must be 0 (f32) or 1 (f64)");
    snfft_plan* p = nullptr;
    try {
        SNFFT_CUDA(cudaSetDevice(device));
        p = new snfft_plan();
        p->n = n;
        p->dtype = dtype;
        p->device = device;
        p->esize = dtype == SNFFT_F32 ? 4 : 8;
        build_host_plan(n, p->host);
        p->ia = make_index_args(n);
        p->lev.resize(n + 1);
        for (int k = 2; k <= n; ++k) {
            const int rc = dtype == SNFFT_F32 ? build_level<float>(p, k) : build_level<double>(p, k);
            if (rc) {
                snfft_plan_destroy(p);
                return rc;
            }
        }
    } catch (const std::bad_alloc&) {
        if (p) snfft_plan_destroy(p);
        return set_error(SNFFT_ENOMEM, "out of host memory while building the plan");
    } catch (const std::exception& e) {
        if (p) snfft_plan_destroy(p);
        return set_error(SNFFT_EINVAL, e.what());
    }
    *out = p;
    return SNFFT_OK;
}

void snfft_plan_destroy(snfft_plan_t* p) {
    if (!p) return;
    for (void* d : p->allocs) cudaFree(d);
    if (p->h_in) cudaFree(p->h_in);
    if (p->h_out) cudaFree(p->h_out);
    if (p->h_ws) cudaFree(p->h_ws);
    delete p;
}

int snfft_plan_info(const snfft_plan_t* p, int* n, int* dtype, int* n_partitions,
                    const int32_t** partitions_flat, const int32_t** dims, const int64_t** coef_offsets) {
    if (!p) return set_error(SNFFT_EINVAL, "plan is NULL");
    const HostLevel& L = p->host.levels[p->n];
    if (n) *n = p->n;
    if (dtype) *dtype = p->dtype;
    if (n_partitions) *n_partitions = (int)L.shapes.size();
    if (partitions_flat) *partitions_flat = L.partitions_flat.data();
    if (dims) *dims = L.dims.data();
    if (coef_offsets) *coef_offsets = L.coef_offsets.data();
    return SNFFT_OK;
}

int snfft_plan_level_info(const snfft_plan_t* p, int k, int* n_partitions, const int32_t** partitions_flat,
                          const int32_t** dims) {
    if (!p || k < 1 || k > p->n) return set_error(SNFFT_EINVAL, "bad plan or level");
    const HostLevel& L = p->host.levels[k];
    if (n_partitions) *n_partitions = (int)L.shapes.size();
    if (partitions_flat) *partitions_flat = L.partitions_flat.data();
    if (dims) *dims = L.dims.data();
    return SNFFT_OK;
}

int snfft_plan_yor_table(const snfft_plan_t* p, int k, int part, int j, int* d, const int32_t** partner,
                         const double** diag, const double** offdiag) {
    if (!p || k < 2 || k > p->n) return set_error(SNFFT_EINVAL, "bad plan or level");
    const HostLevel& L = p->host.levels[k];
    if (part < 0 || part >= (int)L.shapes.size() || j < 0 || j > k - 2)
        return set_error(SNFFT_EINVAL, "bad partition or transposition index");
    const HostShape& S = L.shapes[part];
    if (d) *d = S.dim;
    if (partner) *partner = S.partner[j].data();
    if (diag) *diag = S.diag[j].data();
    if (offdiag) *offdiag = S.off[j].data();
    return SNFFT_OK;
}

size_t snfft_workspace_bytes(const snfft_plan_t* p, int64_t B) {
    if (!p || B < 0) return 0;
    return 2 * (size_t)B * (size_t)p->host.factorial[p->n] * p->esize;
}

static int check_io(const snfft_plan_t* p, const void* a, const void* b, int64_t B, const void* ws, size_t ws_bytes) {
    if (!p) return set_error(SNFFT_EINVAL, "plan is NULL");
    if (B < 0) return set_error(SNFFT_EINVAL, "negative batch");
    if (B == 0) return SNFFT_OK;
    if (!a || !b) return set_error(SNFFT_EINVAL, "NULL data pointer");
    if (!ws || ws_bytes < snfft_workspace_bytes(p, B)) return set_error(SNFFT_EWORKSPACE, "workspace too small");
    return SNFFT_OK;
}

int snfft_forward(const snfft_plan_t* p, const void* f, int64_t B, void* F, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_io(p, f, F, B, ws, ws_bytes);
    if (rc || B == 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return p->dtype == SNFFT_F32 ? forward_impl<float>(p, f, B, F, ws, st) : forward_impl<double>(p, f, B, F, ws, st);
}

int snfft_inverse(const snfft_plan_t* p, const void* F, int64_t B, void* f, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_io(p, F, f, B, ws, ws_bytes);
    if (rc || B == 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return p->dtype == SNFFT_F32 ? inverse_impl<float>(p, F, B, f, ws, st) : inverse_impl<double>(p, F, B, f, ws, st);
}

static int host_scratch(snfft_plan_t* p, int64_t B) {
    const size_t io = (size_t)B * (size_t)p->host.factorial[p->n] * p->esize;
    const size_t ws = snfft_workspace_bytes(p, B);
    if (io > p->h_io_bytes) {
        if (p->h_in) cudaFree(p->h_in);
        if (p->h_out) cudaFree(p->h_out);
        p->h_in = p->h_out = nullptr;
        p->h_io_bytes = 0;
        if (cudaMalloc(&p->h_in, io) != cudaSuccess || cudaMalloc(&p->h_out, io) != cudaSuccess)
            return set_error(SNFFT_ENOMEM, "device allocation of host-call staging failed");
        p->h_io_bytes = io;
    }
    if (ws > p->h_ws_bytes) {
        if (p->h_ws) cudaFree(p->h_ws);
        p->h_ws = nullptr;
        p->h_ws_bytes = 0;
        if (cudaMalloc(&p->h_ws, ws) != cudaSuccess)
            return set_error(SNFFT_ENOMEM, "device allocation of host-call workspace failed");
        p->h_ws_bytes = ws;
    }
    return SNFFT_OK;
}

static int host_call(snfft_plan_t* p, const void* in, int64_t B, void* out, void* stream, bool fwd) {
    if (!p) return set_error(SNFFT_EINVAL, "plan is NULL");
    if (B < 0) return set_error(SNFFT_EINVAL, "negative batch");
    if (B == 0) return SNFFT_OK;
    if (!in || !out) return set_error(SNFFT_EINVAL, "NULL data pointer");
    SNFFT_CUDA(cudaSetDevice(p->device));
    int rc = host_scratch(p, B);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t io = (size_t)B * (size_t)p->host.factorial[p->n] * p->esize;
    SNFFT_CUDA(cudaMemcpyAsync(p->h_in, in, io, cudaMemcpyHostToDevice, st));
    rc = fwd ? snfft_forward(p, p->h_in, B, p->h_out, p->h_ws, p->h_ws_bytes, stream)
             : snfft_inverse(p, p->h_in, B, p->h_out, p->h_ws, p->h_ws_bytes, stream);
    if (rc) return rc;
    SNFFT_CUDA(cudaMemcpyAsync(out, p->h_out, io, cudaMemcpyDeviceToHost, st));
    SNFFT_CUDA(cudaStreamSynchronize(st));
    return SNFFT_OK;
}

int snfft_forward_host(snfft_plan_t* p, const void* f_host, int64_t B, void* F_host, void* stream) {
    return host_call(p, f_host, B, F_host, stream, true);
}

int snfft_inverse_host(snfft_plan_t* p, const void* F_host, int64_t B, void* f_host, void* stream) {
    return host_call(p, F_host, B, f_host, stream, false);
}

int snfft_perm_rank(int n, const int8_t* perms, int64_t m, int64_t* ranks, void* stream) {
    if (n < 1 || n > SNFFT_MAX_N || m < 0) return set_error(SNFFT_EINVAL, "bad n or m");
    if (m == 0) return SNFFT_OK;
    if (!perms || !ranks) return set_error(SNFFT_EINVAL, "NULL pointer");
    const IndexArgs ia = make_index_args(n);
    dim3 grid((unsigned)((m + 255) / 256), 1, 1), block(256, 1, 1);
    SNFFT_LAUNCH(perm_rank_kernel, grid, block, 0, (cudaStream_t)stream, perms, (long long)m, (long long*)ranks, ia);
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

int snfft_perm_unrank(int n, const int64_t* ranks, int64_t m, int8_t* perms, void* stream) {
    if (n < 1 || n > SNFFT_MAX_N || m < 0) return set_error(SNFFT_EINVAL, "bad n or m");
    if (m == 0) return SNFFT_OK;
    if (!perms || !ranks) return set_error(SNFFT_EINVAL, "NULL pointer");
    const IndexArgs ia = make_index_args(n);
    dim3 grid((unsigned)((m + 255) / 256), 1, 1), block(256, 1, 1);
    SNFFT_LAUNCH(perm_unrank_kernel, grid, block, 0, (cudaStream_t)stream, (const long long*)ranks, (long long)m, perms, ia);
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

int snfft_coset_index(int n, int coset, int64_t* idx, void* stream) {
    if (n < 2 || n > SNFFT_MAX_N || coset < 0 || coset >= n) return set_error(SNFFT_EINVAL, "bad n or coset");
    if (!idx) return set_error(SNFFT_EINVAL, "NULL pointer");
    const IndexArgs ia = make_index_args(n);
    const long long m = ia.fact[n - 1];
    dim3 grid((unsigned)((m + 255) / 256), 1, 1), block(256, 1, 1);
    SNFFT_LAUNCH(coset_index_kernel, grid, block, 0, (cudaStream_t)stream, coset, (long long*)idx, ia);
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

int snfft_chain_index(int n, int64_t* p1, void* stream) {
    if (n < 1 || n > SNFFT_MAX_N) return set_error(SNFFT_EINVAL, "bad n");
    if (!p1) return set_error(SNFFT_EINVAL, "NULL pointer");
    const IndexArgs ia = make_index_args(n);
    dim3 grid((unsigned)((ia.nfact + 255) / 256), 1, 1), block(256, 1, 1);
    SNFFT_LAUNCH(chain_index_kernel, grid, block, 0, (cudaStream_t)stream, (long long*)p1, ia);
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

int snfft_dense_rep(const snfft_plan_t* p, int part, void* out, void* stream) {
    if (!p || !out) return set_error(SNFFT_EINVAL, "NULL plan or output");
    const HostLevel& L = p->host.levels[p->n];
    if (part < 0 || part >= (int)L.shapes.size()) return set_error(SNFFT_EINVAL, "bad partition index");
    const int d = L.shapes[part].dim;
    const long long q = p->ia.nfact * d;
    dim3 grid((unsigned)((q + 31) / 32), 1, 1), block(32, 1, 1);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (p->dtype == SNFFT_F32) {
        const LevelArgs<float> a = level_args<float>(p, p->n, nullptr, nullptr, 1);
        auto kfn = dense_rep_kernel<float>;
        const size_t smem = (size_t)d * 32 * sizeof(float);
        if ((rc = set_smem(kfn, smem))) return rc;
        SNFFT_LAUNCH(kfn, grid, block, smem, st, a, part, d, (float*)out, p->ia);
    } else {
        const LevelArgs<double> a = level_args<double>(p, p->n, nullptr, nullptr, 1);
        auto kfn = dense_rep_kernel<double>;
        const size_t smem = (size_t)d * 32 * sizeof(double);
        if ((rc = set_smem(kfn, smem))) return rc;
        SNFFT_LAUNCH(kfn, grid, block, smem, st, a, part, d, (double*)out, p->ia);
    }
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

int snfft_profile_begin(snfft_plan_t* p) {
    if (!p) return set_error(SNFFT_EINVAL, "plan is NULL");
    for (ProfEvent& e : p->prof) {
        cudaEventDestroy(e.e0);
        cudaEventDestroy(e.e1);
    }
    p->prof.clear();
    p->prof_on = true;
    return SNFFT_OK;
}

int snfft_profile_end(snfft_plan_t* p, snfft_profile_record_t* records, int max_records, int* n_records) {
    if (!p || !n_records) return set_error(SNFFT_EINVAL, "NULL plan or n_records");
    p->prof_on = false;
    std::vector<snfft_profile_record_t> agg;
    for (ProfEvent& e : p->prof) {
        float ms = 0.f;
        SNFFT_CUDA(cudaEventSynchronize(e.e1));
        SNFFT_CUDA(cudaEventElapsedTime(&ms, e.e0, e.e1));
        cudaEventDestroy(e.e0);
        cudaEventDestroy(e.e1);
        bool found = false;
        for (auto& r : agg)
            if (r.kind == e.kind && r.level == e.k && r.kernel_class == e.cls) {
                r.launches += 1;
                r.total_ms += ms;
                found = true;
            }
        if (!found) agg.push_back(snfft_profile_record_t{e.kind, e.k, e.cls, 1, (int64_t)e.elems, (double)ms});
    }
    p->prof.clear();
    *n_records = (int)agg.size();
    for (int i = 0; i < (int)agg.size() && i < max_records && records; ++i) records[i] = agg[i];
    return SNFFT_OK;
}

}  // extern "C"
