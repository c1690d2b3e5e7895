// C ABI of the B200-native S_n FFT (include/snfft.h): plan upload, launch
// logic and the extern "C" entry points.  Compiled by nvcc for sm_100a; the
// same file is compiled by g++ against tests/emu/cuda_emu.h (SNFFT_HOST_EMU)
// by the CPU test-suite only, to check the kernel logic without a GPU.
#include "../../include/snfft.h"

#ifdef SNFFT_HOST_EMU
#include "cuda_emu.h"
#else
#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#define SNFFT_CP_ASYNC(dst, src, bytes) __pipeline_memcpy_async(dst, src, bytes)
#define SNFFT_CP_ASYNC_WAIT()      \
    do {                           \
        __pipeline_commit();       \
        __pipeline_wait_prior(0);  \
    } while (0)
#define SNFFT_CP_ASYNC_COMMIT() __pipeline_commit()
#define SNFFT_CP_ASYNC_WAIT_PRIOR(n) __pipeline_wait_prior(n)
// named barriers of warp-specialised kernels: `count` threads (a multiple of 32) arrive or wait
#define SNFFT_BAR_SYNC(id, count) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory")
#define SNFFT_BAR_ARRIVE(id, count) asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory")
#define SNFFT_FENCE_BLOCK() __threadfence_block()
#define SNFFT_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SNFFT_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "plan.h"

namespace snfft {

// ---------------------------------------------------------------------------
// kernel classes: (id, dmax, V columns per thread, W lanes, G row groups, RPT rows per thread,
//                  NG column groups per CTA, NBUF chains per pass, MINB CTAs per SM wanted,
//                  STAGE the task table in shared memory, RPTI rows per thread of the inverse kernel)
// A shape of dimension d runs in the first class with dmax >= d.  The column
// tile is TC = V * W slots; shared memory = NG * NBUF * d * TC * sizeof(T) + table.
// The forward kernel keeps RPT * G >= d_lambda output rows in registers; the inverse kernel keeps the rows of
// the child block, and the children of the shapes of one class are much smaller than dmax (n <= 12: at most
// 35 / 90 / 216 / 768 / 2310 rows in classes 1..5), so RPTI * G only has to cover those (checked at launch).
// ---------------------------------------------------------------------------
#define SNFFT_CLASSES_F32(X)                                                                            \
    X(0, 16, 1, 32, 1, 16, 8, 1, 4, true, 16) X(1, 96, 4, 8, 8, 12, 1, 2, 8, true, 5)                   \
    X(2, 256, 4, 8, 16, 16, 1, 2, 3, true, 6) X(3, 768, 4, 4, 64, 12, 1, 2, 2, true, 4)                  \
    X(4, 2310, 4, 2, 512, 5, 1, 2, 1, true, 2) X(5, 7700, 4, 1, 1024, 8, 1, 1, 1, false, 3)
#define SNFFT_CLASSES_F64(X)                                                                            \
    X(0, 16, 1, 32, 1, 16, 8, 1, 4, true, 16) X(1, 96, 2, 8, 8, 12, 1, 2, 8, true, 5)                   \
    X(2, 256, 2, 8, 16, 16, 1, 2, 3, true, 6) X(3, 768, 2, 4, 64, 12, 1, 2, 2, true, 4)                  \
    X(4, 2310, 2, 2, 512, 5, 1, 2, 1, true, 2) X(5, 7700, 2, 1, 1024, 8, 1, 1, 1, false, 3)

struct KClass {
    int id, dmax, V, W, G, RPT, NG, NBUF, MINB;
    bool STAGE;
};
#define SNFFT_ROW(id, dmax, V, W, G, RPT, NG, NBUF, MINB, STAGE, RPTI) {id, dmax, V, W, G, RPT, NG, NBUF, MINB, STAGE},
static const KClass kClasses32[] = {SNFFT_CLASSES_F32(SNFFT_ROW)};
static const KClass kClasses64[] = {SNFFT_CLASSES_F64(SNFFT_ROW)};
static const int kNumClasses = 6;

static std::atomic<int> g_force_class{-1};
static std::atomic<int> g_no_base{0};
static std::atomic<int> g_no_vec_gather{0};   // test hook: 1 = always the scalar gather / scatter kernel
static std::atomic<int> g_tb_mode{-1};      // -1 automatic kernel choice, >= 0 table-driven sweep kernels at every level >= 6 (cross-check)
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

// Makes `dev` the current CUDA device for a scope and restores the caller's device afterwards (plan creation and
// the *_host entry points must not change the process-wide current device as a side effect).
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        if (prev == dev) {
            prev = -1;
            ok = true;
        } else {
            ok = cudaSetDevice(dev) == cudaSuccess;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

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
    int cls, begin, count, max_dmu, max_dlam, max_tab;
    long long elems;      // coefficients this launch produces per instance (its share of k!)
};

struct DevLevel {
    // column kernels (col_kernels.cuh, levels 6..8): registers only; chosen when the plan is created
    bool col_fwd = false, col_inv = false;
    // top / bottom kernels (tz_kernels.cuh, levels 9 and 10): two streaming launches per step through the third
    // workspace buffer; chosen when the plan is created
    bool tz_fwd = false, tz_inv = false;
    void* tables = nullptr;
    void *ftasks = nullptr, *itasks = nullptr, *parents = nullptr;
    int nftasks = 0, nitasks = 0, tab_words = 0;
    std::vector<FwdTask> h_ftasks;            // host copies (snfft_subplan_create filters them)
    std::vector<InvTask> h_itasks;
    std::vector<InvParent> h_parents;
    double lut_a[LUT], lut_b[LUT];
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
    void* d_p1 = nullptr;               // chain index of every rank (uint32), n <= 11; larger n compute it per tile
    void* d_coef_off = nullptr;         // device copy of the level-n coefficient offsets (pack / unpack)
    void* d_dims = nullptr;             // device copy of d_lambda per partition of level n (snfft_power)
    void* d_tile_prefix = nullptr;      // first 32-coefficient tile of every partition (pack / unpack grid)
    long long n_pack_tiles = 0;
    void *d_vtile_prefix32 = nullptr, *d_vtile_prefix64 = nullptr;   // the same for pack_vec_kernel (128 / 64 coefficients)
    long long n_vtiles32 = 0, n_vtiles64 = 0;
    // per-row YOR tables of level n for snfft_dense_rep, uploaded on first use
    void *rt_partner = nullptr, *rt_diag = nullptr, *rt_off = nullptr;
    std::vector<long long> rt_base;     // per partition: offset of its [n-1][d] block
    // scratch of the *_host entry points
    void *h_in = nullptr, *h_out = nullptr, *h_ws = nullptr;
    size_t h_io_bytes = 0, h_ws_bytes = 0;
};

// A subset of the partitions of n and what the pruned pipeline needs to compute only those (SURVEY 8f.4:
// band-limited transforms): the level steps k >= k_first run the table-driven sweep kernels on task lists
// filtered to the down-closure of the subset in Young's lattice (a block of F_lambda only needs the blocks of the
// shapes contained in lambda, fourier.py:259-265); the levels below run the plan's ordinary kernels.
struct SubLevel {
    void *ftasks = nullptr, *itasks = nullptr, *parents = nullptr;
    std::vector<snfft::TaskRange> franges, iranges;
};
struct snfft_subplan {
    const snfft_plan* plan = nullptr;
    int k_first = 0;                          // first pruned level (n + 1: nothing is pruned)
    std::vector<std::vector<char>> keep;      // keep[k][shape]
    std::vector<SubLevel> lev;
    std::vector<void*> allocs;
    double kept_fraction = 1.0;               // sum of d^2 over the kept partitions of n / n!
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

    // Support-restricted sweeps.  A column of block b starts with support on the rows of
    // block b; rho(s_{k-2}), rho(s_{k-3}), ... spread it.  The table of (lambda, b) lists, for
    // every j, the rotations / sign flips of rho_lambda(s_j) that touch the support before
    // step j (S_{j+1}); the inverse needs exactly the same lists (rows needed after step j).
    // Coefficient codes: d in [2, k-1] -> d - 2 ; d in [-(k-1), -2] -> (k - 2) + (-d - 2).
    for (int c = 0; c < LUT; ++c) D.lut_a[c] = D.lut_b[c] = 0.0;
    for (int dd = 2; dd <= k - 1; ++dd) {
        const double o = std::sqrt(1.0 - 1.0 / ((double)dd * dd));       // irreps.py:103-104
        D.lut_a[dd - 2] = 1.0 / dd;
        D.lut_b[dd - 2] = o;
        D.lut_a[(k - 2) + dd - 2] = 1.0 / (-dd);
        D.lut_b[(k - 2) + dd - 2] = o;
    }
    if (2 * (k - 2) > LUT) return set_error(SNFFT_ENOTSUP, "coefficient table too small for this n");
    std::vector<uint32_t> tables;
    std::vector<std::vector<int>> tab_off(ns), tab_words(ns);        // [lambda][b]
    for (int s = 0; s < ns; ++s) {
        const HostShape& S = L.shapes[s];
        if (S.dim >= 8192) return set_error(SNFFT_ENOTSUP, "irrep dimension exceeds the 13-bit row field");
        tab_off[s].resize(S.child.size());
        tab_words[s].resize(S.child.size());
        for (size_t b = 0; b < S.child.size(); ++b) {
            const size_t base = tables.size();
            tables.resize(base + 2 * (k - 1));
            std::vector<char> sup(S.dim, 0);
            const int dmu = P.shapes[S.child[b]].dim;
            for (int t = 0; t < dmu; ++t) sup[S.boff[b] + t] = 1;
            for (int j = k - 2; j >= 0; --j) {
                const size_t start = tables.size();
                std::vector<uint32_t> negs;
                std::vector<int> grow;
                for (int t = 0; t < S.dim; ++t) {
                    const int q = S.partner[j][t];
                    if (q > t) {
                        if (!sup[t] && !sup[q]) continue;
                        const int dd = S.dist[j][t];
                        const int code = dd > 0 ? dd - 2 : (k - 2) + (-dd - 2);
                        if (dd == 0 || dd == 1 || dd == -1 || code < 0 || code >= LUT ||
                            D.lut_a[code] != S.diag[j][t] || D.lut_b[code] != S.off[j][t])
                            return set_error(SNFFT_EINVAL, "YOR table: inconsistent axial distance");
                        tables.push_back((uint32_t)t | ((uint32_t)q << 13) | ((uint32_t)code << 26));
                        grow.push_back(t);
                        grow.push_back(q);
                    } else if (q == t) {
                        if (S.diag[j][t] == -1.0) {
                            if (sup[t]) negs.push_back((uint32_t)t);
                        } else if (S.diag[j][t] != 1.0) {
                            return set_error(SNFFT_EINVAL, "YOR table: unpaired row with |diag| != 1");
                        }
                    }
                }
                const size_t npairs = tables.size() - start;
                if (npairs > 0xffff || negs.size() > 0xffff) return set_error(SNFFT_ENOTSUP, "sweep list too long");
                tables.insert(tables.end(), negs.begin(), negs.end());
                for (int t : grow) sup[t] = 1;
                tables[base + j] = (uint32_t)(start - base);
                tables[base + (k - 1) + j] = (uint32_t)npairs | ((uint32_t)negs.size() << 16);
            }
            tab_off[s][b] = (int)base;
            tab_words[s][b] = (int)(tables.size() - base);
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
            t.tab_off = tab_off[s][b];
            t.tab_words = tab_words[s][b];
            t.pad = 0;
            fby[cls].push_back(t);
        }
    }
    std::vector<FwdTask> ftasks;
    for (int c = 0; c < kNumClasses; ++c) {
        if (fby[c].empty()) continue;
        // big columns first: the long CTAs start early
        std::stable_sort(fby[c].begin(), fby[c].end(),
                         [](const FwdTask& x, const FwdTask& y) { return x.d_lam > y.d_lam; });
        TaskRange r{c, (int)ftasks.size(), (int)fby[c].size(), 0, 0, 0, 0};
        for (const FwdTask& t : fby[c]) {
            r.elems += (long long)t.d_lam * t.d_mu;
            r.max_tab = std::max(r.max_tab, t.tab_words);
            r.max_dmu = std::max(r.max_dmu, t.d_mu);
            r.max_dlam = std::max(r.max_dlam, t.d_lam);
            ftasks.push_back(t);
        }
        D.franges.push_back(r);
    }

    // inverse tasks: one per mu of level k-1, class by the largest covering lambda
    std::vector<std::vector<InvTask>> iby(kNumClasses);
    std::vector<std::vector<int>> iby_maxd(kNumClasses), iby_maxtab(kNumClasses);
    std::vector<InvParent> parents;
    for (size_t m = 0; m < P.shapes.size(); ++m) {
        const HostShape& M = P.shapes[m];
        InvTask t;
        t.mu_off = M.coef_off;
        t.d_mu = M.dim;
        t.par_off = (int)parents.size();
        t.npar = (int)M.parent.size();
        t.pad = 0;
        int maxd = 0, maxtab = 0;
        for (size_t q = 0; q < M.parent.size(); ++q) {
            const HostShape& S = L.shapes[M.parent[q]];
            InvParent ip;
            ip.lam_off = S.coef_off;
            ip.d_lam = S.dim;
            ip.boff = S.boff[M.parent_block[q]];
            ip.tab_off = tab_off[M.parent[q]][M.parent_block[q]];
            ip.tab_words = tab_words[M.parent[q]][M.parent_block[q]];
            maxtab = std::max(maxtab, ip.tab_words);
            ip.scale = (double)S.dim / ((double)k * (double)M.dim);
            parents.push_back(ip);
            maxd = std::max(maxd, S.dim);
        }
        const int cls = pick_class(p->dtype, maxd);
        if (cls < 0) return set_error(SNFFT_ENOTSUP, "irrep dimension exceeds the largest kernel class");
        iby[cls].push_back(t);
        iby_maxd[cls].push_back(maxd);
        iby_maxtab[cls].push_back(maxtab);
    }
    std::vector<InvTask> itasks;
    for (int c = 0; c < kNumClasses; ++c) {
        if (iby[c].empty()) continue;
        TaskRange r{c, (int)itasks.size(), (int)iby[c].size(), 0, 0, 0, 0};
        for (size_t q = 0; q < iby[c].size(); ++q) {
            r.max_tab = std::max(r.max_tab, iby_maxtab[c][q]);
            r.elems += (long long)k * iby[c][q].d_mu * iby[c][q].d_mu;
            r.max_dmu = std::max(r.max_dmu, iby[c][q].d_mu);
            r.max_dlam = std::max(r.max_dlam, iby_maxd[c][q]);
            itasks.push_back(iby[c][q]);
        }
        D.iranges.push_back(r);
    }

    D.nftasks = (int)ftasks.size();
    D.nitasks = (int)itasks.size();
    D.tab_words = (int)tables.size();
    D.h_ftasks = ftasks;
    D.h_itasks = itasks;
    D.h_parents = parents;
    int rc;
    if ((rc = upload(p, tables, &D.tables))) return rc;
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

// Tuning / test switches: environment variables that list levels ("7,8,9"; "" = none), read when a plan is created.
static bool level_listed(const char* env, int k, bool dflt) {
    const char* e = std::getenv(env);
    if (!e) return dflt;
    for (const char* q = e; *q;) {
        char* end;
        const long v = std::strtol(q, &end, 10);
        if (end == q) break;
        if (v == k) return true;
        q = (*end == ',') ? end + 1 : end;
    }
    return false;
}

template <typename T, int V, int W, int G, int RPT, int NG, int NBUF, int MINB, bool STAGE>
static int launch_fwd_range(const TaskRange& r, LevelArgs<T> a, cudaStream_t st) {
    constexpr int TC = V * W;
    const long long chunks = ((long long)r.max_dmu * a.ninst + TC - 1) / TC;
    a.bstride = r.max_dlam * TC;
    a.ftasks += r.begin;
    a.tab_smem_words = STAGE ? r.max_tab : 0;
    const size_t smem = level_smem_bytes<T>(NG, NBUF, a.bstride, a.tab_smem_words);
    dim3 grid((unsigned)((chunks + NG - 1) / NG), (unsigned)r.count, 1), block(W * G * NG, 1, 1);
    int rc;
    {
        auto kfn = fwd_level_kernel<T, V, W, G, RPT, NG, NBUF, MINB, STAGE>;
        if ((rc = set_smem(kfn, smem))) return rc;
        SNFFT_LAUNCH(kfn, grid, block, smem, st, a);
    }
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

template <typename T, int V, int W, int G, int RPT, int NG, int NBUF, int MINB, bool STAGE>
static int launch_inv_range(const TaskRange& r, LevelArgs<T> a, cudaStream_t st) {
    constexpr int TC = V * W;
    const long long chunks = ((long long)r.max_dmu * a.ninst + TC - 1) / TC;
    a.bstride = r.max_dlam * TC;
    a.itasks += r.begin;
    a.tab_smem_words = STAGE ? r.max_tab : 0;
    const size_t smem = level_smem_bytes<T>(NG, NBUF, a.bstride, a.tab_smem_words);
    const long long ngroups = (a.k + NBUF - 1) / NBUF;      // coset groups, fastest grid dimension
    dim3 grid((unsigned)(((chunks + NG - 1) / NG) * ngroups), (unsigned)r.count, 1), block(W * G * NG, 1, 1);
    int rc;
    {
        auto kfn = inv_level_kernel<T, V, W, G, RPT, NG, NBUF, MINB, STAGE>;
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
    a.tables = (const uint32_t*)D.tables;
    for (int c = 0; c < LUT; ++c) {
        a.lut[c] = (T)D.lut_a[c];
        a.lut[LUT + c] = (T)D.lut_b[c];
    }
    a.k = k;
    a.bstride = a.tab_smem_words = 0;
    a.ninst = B * (p->host.factorial[p->n] / p->host.factorial[k]);
    a.batch = B;
    a.nch = k;
    for (int j = 0; j < SNFFT_MAX_N; ++j) a.cid[j] = j;
    return a;
}

// nch > 0: only the cosets cid[0..nch) (ascending) are present in `in` (nch planes per coefficient)
// `sub_tasks` / `sub_ranges`: a filtered task list of the same level (snfft_subplan), else the plan's own
template <typename T>
static int run_fwd_level(const snfft_plan* p, int k, const void* in, void* out, long long B, cudaStream_t st,
                         int nch = 0, const int* cid = nullptr, const void* sub_tasks = nullptr,
                         const std::vector<TaskRange>* sub_ranges = nullptr) {
    LevelArgs<T> a = level_args<T>(p, k, in, out, B);
    if (nch > 0) {
        a.nch = nch;
        for (int j = 0; j < nch; ++j) a.cid[j] = cid[j];
    }
    if (sub_ranges) a.ftasks = (const FwdTask*)sub_tasks;
    for (const TaskRange& r : (sub_ranges ? *sub_ranges : p->lev[k].franges)) {
        int rc = SNFFT_ENOTSUP;
        ProfScope prof(p, st, 0, k, r.cls, r.elems);
        if constexpr (sizeof(T) == 4) {
            switch (r.cls) {
#define SNFFT_CASE(id, dmax, V, W, G, RPT, NG, NBUF, MINB, STAGE, RPTI) \
    case id: rc = launch_fwd_range<T, V, W, G, RPT, NG, NBUF, MINB, STAGE>(r, a, st); break;
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
static int run_inv_level(const snfft_plan* p, int k, const void* in, void* out, long long B, cudaStream_t st,
                         const void* sub_tasks = nullptr, const void* sub_parents = nullptr,
                         const std::vector<TaskRange>* sub_ranges = nullptr) {
    LevelArgs<T> a = level_args<T>(p, k, in, out, B);
    if (sub_ranges) {
        a.itasks = (const InvTask*)sub_tasks;
        a.parents = (const InvParent*)sub_parents;
    }
    for (const TaskRange& r : (sub_ranges ? *sub_ranges : p->lev[k].iranges)) {
        int rc = SNFFT_ENOTSUP;
        ProfScope prof(p, st, 1, k, r.cls, r.elems);
        if constexpr (sizeof(T) == 4) {
            switch (r.cls) {
#define SNFFT_CASE(id, dmax, V, W, G, RPT, NG, NBUF, MINB, STAGE, RPTI)                                   \
    case id:                                                                                             \
        if (r.max_dmu > RPTI * G) return set_error(SNFFT_ENOTSUP, "child block larger than the inverse class allows"); \
        rc = launch_inv_range<T, V, W, G, RPTI, NG, NBUF, MINB, STAGE>(r, a, st);                        \
        break;
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
    constexpr int VT = 16 / (int)sizeof(T);
    if (p->n >= 4 && B % VT == 0 && (uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0 && g_no_vec_gather.load() == 0) {
        dim3 grid((unsigned)((ia.nfact + 32 * VT - 1) / (32 * VT)), 1, 1), block(256, 1, 1);
        ProfScope prof(p, st, SCATTER ? 3 : 2, 1, -2, ia.nfact);
        auto kfn = gather_vec_kernel<T, SCATTER>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const T*)in, (T*)out, (const uint32_t*)p->d_p1, ia);
        g_launches.fetch_add(1);
        SNFFT_CUDA(cudaGetLastError());
        return SNFFT_OK;
    }
    if (B <= 8 && g_no_vec_gather.load() == 0) {         // small batches: one thread per rank (kernel_class -6)
        dim3 grid((unsigned)((ia.nfact + 255) / 256), 1, 1), block(256, 1, 1);
        ProfScope prof(p, st, SCATTER ? 3 : 2, 1, -6, ia.nfact);
        auto kfn = gather_small_kernel<T, SCATTER>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const T*)in, (T*)out, (const uint32_t*)p->d_p1, ia);
        g_launches.fetch_add(1);
        SNFFT_CUDA(cudaGetLastError());
        return SNFFT_OK;
    }
    dim3 grid((unsigned)((ia.nfact + GS_TR - 1) / GS_TR), (unsigned)std::min<long long>((B + GS_TB - 1) / GS_TB, 65535), 1);
    dim3 block(GS_THREADS, 1, 1);
    ProfScope prof(p, st, SCATTER ? 3 : 2, 1, -1, ia.nfact);
    auto kfn = gather_kernel<T, SCATTER>;
    SNFFT_LAUNCH(kfn, grid, block, gather_smem_bytes<T>(), st, (const T*)in, (T*)out, ia);
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

// X_n instance-fastest <-> packed [lambda][B][d][d]
template <typename T, bool TO_PACKED>
static int run_pack(const snfft_plan* p, const void* in, void* out, long long B, cudaStream_t st) {
    const HostLevel& L = p->host.levels[p->n];
    constexpr int VT = 16 / (int)sizeof(T);
    if (B == 1) {        // one function: [d][d][1] per partition IS the packed layout, a plain copy
        ProfScope prof(p, st, TO_PACKED ? 6 : 7, p->n, -6, p->host.factorial[p->n]);
        SNFFT_CUDA(cudaMemcpyAsync(out, in, (size_t)p->host.factorial[p->n] * sizeof(T), cudaMemcpyDeviceToDevice, st));
        return SNFFT_OK;
    }
    if (B % VT == 0 && (uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0 && g_no_vec_gather.load() == 0) {
        const void* prefix = sizeof(T) == 4 ? p->d_vtile_prefix32 : p->d_vtile_prefix64;
        const long long vt = sizeof(T) == 4 ? p->n_vtiles32 : p->n_vtiles64;
        dim3 grid((unsigned)vt, 1, 1), block(256, 1, 1);
        ProfScope prof(p, st, TO_PACKED ? 6 : 7, p->n, -2, p->host.factorial[p->n]);
        auto kfn = pack_vec_kernel<T, TO_PACKED>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const T*)in, (T*)out, (const long long*)p->d_coef_off,
                     (const long long*)prefix, (int)L.shapes.size(), B);
        g_launches.fetch_add(1);
        SNFFT_CUDA(cudaGetLastError());
        return SNFFT_OK;
    }
    const long long tiles = p->n_pack_tiles * ((B + PK_T - 1) / PK_T);
    dim3 grid((unsigned)tiles, 1, 1), block(PK_T * 8, 1, 1);
    ProfScope prof(p, st, TO_PACKED ? 6 : 7, p->n, -1, p->host.factorial[p->n]);
    auto kfn = pack_kernel<T, TO_PACKED>;
    SNFFT_LAUNCH(kfn, grid, block, 0, st, (const T*)in, (T*)out, (const long long*)p->d_coef_off,
                 (const long long*)p->d_tile_prefix, (int)L.shapes.size(), B);
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

// Generated register kernel: levels 2..5 of every S_5 instance.
template <typename T, bool INVERSE>
static int run_s5(const snfft_plan* p, const void* in, void* out, long long B, cudaStream_t st) {
    const long long ninst = B * (p->host.factorial[p->n] / 120);
    dim3 grid((unsigned)((ninst + GEN_THREADS - 1) / GEN_THREADS), 1, 1), block(GEN_THREADS, 1, 1);
    ProfScope prof(p, st, INVERSE ? 5 : 4, 5, -1, 120);
    auto kfn = s5_kernel<T, INVERSE>;
    SNFFT_LAUNCH(kfn, grid, block, 0, st, (const T*)in, (T*)out, ninst);
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

// Levels 2..6 in one kernel (s6_kernels.cuh / s6_tu.cu, fp32): the S_5 coefficients go from the generated levels-2..5
// code to the level-6 column tasks through shared memory instead of a level array.  SNFFT_S6=0 keeps the two kernels.
extern "C" {
int snfft_s6_launch_f32_fwd(const void*, void*, long long, cudaStream_t);
int snfft_s6_launch_f32_inv(const void*, void*, long long, cudaStream_t);
}
template <typename T>
static bool s6_ok(const snfft_plan* p, bool inverse) {
    static const bool on = [] {
        const char* e = std::getenv("SNFFT_S6");
        return !e || std::atoi(e) != 0;
    }();
    return on && sizeof(T) == 4 && p->n >= 6 && g_no_base.load() == 0 && (inverse ? p->lev[6].col_inv : p->lev[6].col_fwd);
}
// forward: in = X_1 (the gathered input), out = X_6; inverse: in = X_6, out = X_1.
// Profile record: kind 4 / 5 (base kernel), level 6, kernel_class -8.
template <typename T, bool INVERSE>
static int run_s6(const snfft_plan* p, const void* in, void* out, long long B, cudaStream_t st) {
    const long long ninst6 = B * (p->host.factorial[p->n] / 720);
    ProfScope prof(p, st, INVERSE ? 5 : 4, 6, -8, 720);
    const int rc = INVERSE ? snfft_s6_launch_f32_inv(in, out, ninst6, st) : snfft_s6_launch_f32_fwd(in, out, ninst6, st);
    g_launches.fetch_add(1);
    if (rc) return set_error(SNFFT_ECUDA, std::string("levels 2..6 kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
    return SNFFT_OK;
}

// Column kernels (col_kernels.cuh / col_tu.cu: one translation unit per level, dtype and direction).
extern "C" {
#define SNFFT_COL_DECL(k, dt) \
    int snfft_col_launch_##k##_##dt##_fwd(const void*, void*, long long, cudaStream_t); \
    int snfft_col_launch_##k##_##dt##_inv(const void*, void*, long long, cudaStream_t);
SNFFT_COL_DECL(6, f32) SNFFT_COL_DECL(6, f64) SNFFT_COL_DECL(7, f32) SNFFT_COL_DECL(7, f64) SNFFT_COL_DECL(8, f32)
SNFFT_COL_DECL(8, f64)
#undef SNFFT_COL_DECL
}
constexpr int kColFirst = 6, kColLast = 8, kColLastInv = 8;

// Levels that use the column kernels: automatic kernel choice only (the forced TB / sweep modes of
// snfft_debug_tb_mode keep testing the sweep kernels), SNFFT_COL_FWD / SNFFT_COL_INV (lists of levels) override.
static void col_pick_dirs(int k, bool* fwd, bool* inv) {
    *fwd = *inv = false;
    if (k < kColFirst || k > kColLast || g_tb_mode.load() != -1) return;
    *fwd = level_listed("SNFFT_COL_FWD", k, true);
    *inv = k <= kColLastInv && level_listed("SNFFT_COL_INV", k, true);
}

template <typename T, bool INVERSE>
static int run_col_level(const snfft_plan* p, int k, const void* in, void* out, long long B, cudaStream_t st) {
    const long long ninst = B * (p->host.factorial[p->n] / p->host.factorial[k]);
    ProfScope prof(p, st, INVERSE ? 1 : 0, k, -3, p->host.factorial[k]);
    int rc;
#define SNFFT_COL_CALL(kk, dt) (INVERSE ? snfft_col_launch_##kk##_##dt##_inv(in, out, ninst, st) : snfft_col_launch_##kk##_##dt##_fwd(in, out, ninst, st))
    if (sizeof(T) == 4) rc = k == 6 ? SNFFT_COL_CALL(6, f32) : (k == 7 ? SNFFT_COL_CALL(7, f32) : SNFFT_COL_CALL(8, f32));
    else rc = k == 6 ? SNFFT_COL_CALL(6, f64) : (k == 7 ? SNFFT_COL_CALL(7, f64) : SNFFT_COL_CALL(8, f64));
#undef SNFFT_COL_CALL
    g_launches.fetch_add(1);
    if (rc) return set_error(SNFFT_ECUDA, std::string("column kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
    return SNFFT_OK;
}

// Top / bottom kernels (tz_kernels.cuh / tz_tu.cu: one translation unit per kernel, dtype and direction).
extern "C" {
#define SNFFT_TZ_DECL(dt, dir)                                                                 \
    int snfft_tz_z7_##dt##_##dir(int, const void*, void*, long long, cudaStream_t);            \
    int snfft_tz_z8_##dt##_##dir(int, const void*, void*, long long, cudaStream_t);            \
    int snfft_tz_z7_prepare_##dt##_##dir(int);                                                 \
    int snfft_tz_z8_prepare_##dt##_##dir(int);                                                 \
    int snfft_tz_t9_##dt##_##dir(void*, void*, void*, long long, cudaStream_t);                \
    int snfft_tz_t10_##dt##_##dir(void*, void*, void*, long long, cudaStream_t);               \
    int snfft_tz_t11_##dt##_##dir(void*, void*, void*, long long, cudaStream_t);               \
    int snfft_tz_tpk9_##dt##_##dir(void*, void*, void*, long long, cudaStream_t);              \
    int snfft_tz_tpk10_##dt##_##dir(void*, void*, void*, long long, cudaStream_t);             \
    int snfft_tz_tpk11_##dt##_##dir(void*, void*, void*, long long, cudaStream_t);
SNFFT_TZ_DECL(f32, fwd) SNFFT_TZ_DECL(f32, inv) SNFFT_TZ_DECL(f64, fwd) SNFFT_TZ_DECL(f64, inv)
#undef SNFFT_TZ_DECL
}
constexpr int kTzFirst = 9, kTzLast = 11;
static int tz_mb(int k) { return k == 11 ? 8 : 7; }       // gen_tz.py LEVEL_MB

// Levels that use the top / bottom kernels: automatic kernel choice only; SNFFT_TZ_FWD / SNFFT_TZ_INV (lists of
// levels, "" = none) override when a plan is created.  They take precedence over the column kernels.
static void tz_pick_dirs(int k, bool* fwd, bool* inv) {
    *fwd = *inv = false;
    if (k < kTzFirst || k > kTzLast || g_tb_mode.load() != -1) return;
    *fwd = level_listed("SNFFT_TZ_FWD", k, true);
    *inv = level_listed("SNFFT_TZ_INV", k, true);
}

static int tz_prepare(int dtype, int k, bool fwd, bool inv) {
    int rc = 0;
    const bool f32 = dtype == SNFFT_F32;
    if (tz_mb(k) == 7) {
        if (fwd) rc = f32 ? snfft_tz_z7_prepare_f32_fwd(k) : snfft_tz_z7_prepare_f64_fwd(k);
        if (!rc && inv) rc = f32 ? snfft_tz_z7_prepare_f32_inv(k) : snfft_tz_z7_prepare_f64_inv(k);
    } else {
        if (fwd) rc = f32 ? snfft_tz_z8_prepare_f32_fwd(k) : snfft_tz_z8_prepare_f64_fwd(k);
        if (!rc && inv) rc = f32 ? snfft_tz_z8_prepare_f32_inv(k) : snfft_tz_z8_prepare_f64_inv(k);
    }
    if (rc) return set_error(SNFFT_ECUDA, std::string("top/bottom kernel tables: ") + cudaGetErrorString((cudaError_t)rc));
    return SNFFT_OK;
}

// The top level's T kernel can read / write the public packed coefficients directly (tz_tu.cu: snfft_tz_tpk*): no
// separate pack / unpack pass.  Needs full instance vectors; SNFFT_TZ_PACK=0 keeps the two-pass route.
template <typename T>
static bool tz_pack_ok(const snfft_plan* p, int k, long long B, const void* level_buf, const void* z, const void* packed) {
    static const bool on = [] {
        const char* e = std::getenv("SNFFT_TZ_PACK");
        return !e || std::atoi(e) != 0;
    }();
    const long long nv = 16 / (long long)sizeof(T);
    return on && k == p->n && B % nv == 0 && ((uintptr_t)level_buf | (uintptr_t)z) % 16 == 0 && (uintptr_t)packed % sizeof(T) == 0;
}

// One level step through the top / bottom kernels.  forward: in = X_{k-1}, out = X_k; inverse: in = X_k,
// out = X_{k-1}; z = the third workspace buffer (MB/k of a level array is used).  packed (k = n, tz_pack_ok): the
// level-n side (forward: out, inverse: in) is the public packed layout [lambda][B][d][d].
// Profile records: kernel_class -4 = Z kernel, -5 = T kernel (-7 = T kernel on the packed layout).
template <typename T, bool INVERSE>
static int run_tz_level(const snfft_plan* p, int k, const void* in, void* out, void* z, long long B, cudaStream_t st,
                        bool packed = false) {
    const long long ninst = B * (p->host.factorial[p->n] / p->host.factorial[k]);
    const long long fk = p->host.factorial[k];
    int rc;
    auto fail = [](int e) { return set_error(SNFFT_ECUDA, std::string("top/bottom kernel launch: ") + cudaGetErrorString((cudaError_t)e)); };
    constexpr bool F32 = sizeof(T) == 4;
    const int mb = tz_mb(k);
    typedef int (*zfn_t)(int, const void*, void*, long long, cudaStream_t);
    typedef int (*tfn_t)(void*, void*, void*, long long, cudaStream_t);
    zfn_t zfn;
    tfn_t tfn;
    if (!INVERSE) {
        zfn = mb == 7 ? (F32 ? snfft_tz_z7_f32_fwd : snfft_tz_z7_f64_fwd) : (F32 ? snfft_tz_z8_f32_fwd : snfft_tz_z8_f64_fwd);
        tfn = k == 9 ? (F32 ? snfft_tz_t9_f32_fwd : snfft_tz_t9_f64_fwd)
                     : (k == 10 ? (F32 ? snfft_tz_t10_f32_fwd : snfft_tz_t10_f64_fwd) : (F32 ? snfft_tz_t11_f32_fwd : snfft_tz_t11_f64_fwd));
        if (packed)
            tfn = k == 9 ? (F32 ? snfft_tz_tpk9_f32_fwd : snfft_tz_tpk9_f64_fwd)
                         : (k == 10 ? (F32 ? snfft_tz_tpk10_f32_fwd : snfft_tz_tpk10_f64_fwd) : (F32 ? snfft_tz_tpk11_f32_fwd : snfft_tz_tpk11_f64_fwd));
        {
            ProfScope prof(p, st, 0, k, -4, fk * mb / k);
            rc = zfn(k, in, z, ninst, st);
            g_launches.fetch_add(1);
            if (rc) return fail(rc);
        }
        ProfScope prof(p, st, 0, k, packed ? -7 : -5, fk);
        rc = tfn(z, (void*)in, out, ninst, st);
        g_launches.fetch_add(1);
        if (rc) return fail(rc);
    } else {
        zfn = mb == 7 ? (F32 ? snfft_tz_z7_f32_inv : snfft_tz_z7_f64_inv) : (F32 ? snfft_tz_z8_f32_inv : snfft_tz_z8_f64_inv);
        tfn = k == 9 ? (F32 ? snfft_tz_t9_f32_inv : snfft_tz_t9_f64_inv)
                     : (k == 10 ? (F32 ? snfft_tz_t10_f32_inv : snfft_tz_t10_f64_inv) : (F32 ? snfft_tz_t11_f32_inv : snfft_tz_t11_f64_inv));
        if (packed)
            tfn = k == 9 ? (F32 ? snfft_tz_tpk9_f32_inv : snfft_tz_tpk9_f64_inv)
                         : (k == 10 ? (F32 ? snfft_tz_tpk10_f32_inv : snfft_tz_tpk10_f64_inv) : (F32 ? snfft_tz_tpk11_f32_inv : snfft_tz_tpk11_f64_inv));
        {
            ProfScope prof(p, st, 1, k, packed ? -7 : -5, fk);
            rc = tfn(z, out, (void*)in, ninst, st);
            g_launches.fetch_add(1);
            if (rc) return fail(rc);
        }
        ProfScope prof(p, st, 1, k, -4, fk * mb / k);
        rc = zfn(k, z, out, ninst, st);
        g_launches.fetch_add(1);
        if (rc) return fail(rc);
    }
    return SNFFT_OK;
}

static IndexArgs make_index_args(int n);

// level steps k_first .. n of the forward pipeline: cur = X_{k_first - 1} on entry, X_n on return
// F_packed != nullptr: the caller wants the packed coefficients there; when the top level runs the top / bottom kernels
// its T kernel writes them directly and *packed_done is set (otherwise the caller packs X_n itself).
template <typename T>
static int forward_level_range(const snfft_plan* p, int k_first, void*& cur, void*& nxt, void* zbuf, long long B, cudaStream_t st,
                               int k_stop = -1, void* F_packed = nullptr, bool* packed_done = nullptr) {
    const int base = g_no_base.load();
    const int k_end = k_stop < 0 ? p->n : k_stop - 1;
    for (int k = k_first; k <= k_end; ++k) {
        int rc;
        if (base == 0 && p->lev[k].tz_fwd && F_packed && packed_done && tz_pack_ok<T>(p, k, B, cur, zbuf, F_packed)) {
            if ((rc = run_tz_level<T, false>(p, k, cur, F_packed, zbuf, B, st, true))) return rc;
            *packed_done = true;
            return SNFFT_OK;
        }
        if (base == 0 && p->lev[k].tz_fwd) rc = run_tz_level<T, false>(p, k, cur, nxt, zbuf, B, st);
        else if (base == 0 && p->lev[k].col_fwd) rc = run_col_level<T, false>(p, k, cur, nxt, B, st);
        else rc = run_fwd_level<T>(p, k, cur, nxt, B, st);
        if (rc) return rc;
        std::swap(cur, nxt);
    }
    return SNFFT_OK;
}

// levels 1..n of the forward pipeline; *result points at X_n (instance fastest) inside the workspace
template <typename T>
static int forward_levels(const snfft_plan* p, const void* f, long long B, void* ws, cudaStream_t st, void** result,
                          int k_stop = -1, void** other = nullptr,          // k_stop: stop before that level step
                          void* F_packed = nullptr, bool* packed_done = nullptr) {
    const size_t half = (size_t)B * p->host.factorial[p->n] * sizeof(T);
    void* cur = ws;
    void* nxt = (char*)ws + half;
    void* zbuf = (char*)ws + 2 * half;          // third buffer: only when the plan has top / bottom levels
    int rc;
    if ((rc = run_gather<T, false>(p, f, cur, B, st))) return rc;
    int k_first = 2;
    const int base = g_no_base.load();          // 0: generated register kernels for levels 2..5, else table-driven level kernels
    if (s6_ok<T>(p, false) && (k_stop < 0 || k_stop > 6)) {
        if ((rc = run_s6<T, false>(p, cur, nxt, B, st))) return rc;
        std::swap(cur, nxt);
        k_first = 7;
    } else if (base == 0 && p->n > 5) {
        if ((rc = run_s5<T, false>(p, cur, nxt, B, st))) return rc;
        std::swap(cur, nxt);
        k_first = 6;
    }
    if ((rc = forward_level_range<T>(p, k_first, cur, nxt, zbuf, B, st, k_stop, F_packed, packed_done))) return rc;
    *result = cur;
    if (other) *other = nxt;
    return SNFFT_OK;
}

template <typename T>
static int forward_impl(const snfft_plan* p, const void* f, long long B, void* F, void* ws, cudaStream_t st) {
    void* xn = nullptr;
    bool packed = false;
    const int rc = forward_levels<T>(p, f, B, ws, st, &xn, -1, nullptr, F, &packed);
    if (rc) return rc;
    if (packed) return SNFFT_OK;
    return run_pack<T, true>(p, xn, F, B, st);
}

// Partial forward transform over a subset of the top-level cosets {sigma : sigma[i] = n-1}:
//   F_partial(lambda) = sum_{i in cosets} rho_lambda(c_i) . blockdiag_mu F^(i)_mu      (fourier.py:253-273)
// so that the sum of F_partial over a partition of range(n) is the full transform.  The restricted functions
// f^(i) (restrict_to_coset, fourier.py:63-80) go through the S_{n-1} plan as one batch of ncos * B functions.
// workspace: [restricted input][S_{n-1} pipeline scratch][X_n]
template <typename T>
static int forward_cosets_impl(const snfft_plan* pn, const snfft_plan* pm, const void* f, long long B, const int* cosets,
                               int ncos, void* F, void* ws, cudaStream_t st) {
    const int n = pn->n;
    const long long fm = pn->host.factorial[n - 1], fn = pn->host.factorial[n];
    const long long Bm = (long long)ncos * B;
    T* rin = (T*)ws;
    void* wsm = (char*)ws + (size_t)Bm * fm * sizeof(T);
    T* xn = (T*)((char*)wsm + snfft_workspace_bytes(pm, Bm));
    {
        const long long total = Bm * fm;
        RestrictArgs ra;
        ra.ia = make_index_args(n);
        ra.batch = B;
        ra.ncos = ncos;
        for (int j = 0; j < ncos; ++j) ra.coset[j] = cosets[j];
        dim3 grid((unsigned)((total + 255) / 256), 1, 1), block(256, 1, 1);
        auto kfn = restrict_kernel<T>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const T*)f, rin, ra);
        g_launches.fetch_add(1);
        SNFFT_CUDA(cudaGetLastError());
    }
    void* xm = nullptr;
    int rc = forward_levels<T>(pm, rin, Bm, wsm, st, &xm);
    if (rc) return rc;
    if ((rc = run_fwd_level<T>(pn, n, xm, xn, B, st, ncos, cosets))) return rc;
    (void)fn;
    return run_pack<T, true>(pn, xn, F, B, st);
}

// The same partial transform over a subset of the n (n-1) TWO-level cosets {sigma : sigma[i] = n-1, then the
// restriction's [j] = n-2} (finer units for 8 ranks than the n top-level cosets).  The pairs'
// restrictions go through the S_{n-2} plan as one batch; their level arrays are placed at their instance slots
// of a zeroed level-(n-2) array of the full transform, and the last two level steps run at full size: the terms
// of absent pairs are zero, so the result is the partial sum (fourier.py:253-273 applied twice, by linearity).
// workspace: [restricted input][S_{n-2} pipeline scratch][three level arrays of the full transform]
template <typename T>
static int forward_cosets2_impl(const snfft_plan* pn, const snfft_plan* pm, const void* f, long long B, const int* top,
                                const int* sub, int npairs, void* F, void* ws, cudaStream_t st) {
    const int n = pn->n;
    const long long fm = pn->host.factorial[n - 2], fn = pn->host.factorial[n];
    const long long Bm = (long long)npairs * B;
    T* rin = (T*)ws;
    void* wsm = (char*)ws + (size_t)Bm * fm * sizeof(T);
    const size_t full = (size_t)B * fn * sizeof(T);
    void* cur = (char*)wsm + snfft_workspace_bytes(pm, Bm);
    void* nxt = (char*)cur + full;
    void* zbuf = (char*)nxt + full;
    {
        Restrict2Args ra;
        ra.ia = make_index_args(n);
        ra.batch = B;
        ra.npairs = npairs;
        for (int j = 0; j < npairs; ++j) {
            ra.top[j] = (unsigned char)top[j];
            ra.sub[j] = (unsigned char)sub[j];
        }
        const long long total = Bm * fm;
        dim3 grid((unsigned)((total + 255) / 256), 1, 1), block(256, 1, 1);
        auto kfn = restrict2_kernel<T>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const T*)f, rin, ra);
        g_launches.fetch_add(1);
        SNFFT_CUDA(cudaGetLastError());
    }
    void* xm = nullptr;
    int rc = forward_levels<T>(pm, rin, Bm, wsm, st, &xm);
    if (rc) return rc;
    SNFFT_CUDA(cudaMemsetAsync(cur, 0, full, st));
    {
        ExpandArgs ea;
        ea.batch = B;
        ea.ncoef = fm;
        ea.ninst_full = B * (long long)n * (n - 1);
        ea.npairs = npairs;
        for (int j = 0; j < npairs; ++j) ea.slot[j] = (int)(((long long)sub[j] * n + top[j]) * B);
        const long long total = fm * Bm;
        dim3 grid((unsigned)((total + 255) / 256), 1, 1), block(256, 1, 1);
        auto kfn = expand_pairs_kernel<T>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const T*)xm, (T*)cur, ea);
        g_launches.fetch_add(1);
        SNFFT_CUDA(cudaGetLastError());
    }
    bool packed = false;
    if ((rc = forward_level_range<T>(pn, n - 1, cur, nxt, zbuf, B, st, -1, F, &packed))) return rc;
    if (packed) return SNFFT_OK;
    return run_pack<T, true>(pn, cur, F, B, st);
}

template <typename T>
static int inverse_tail(const snfft_plan* p, int k_top, const void* cur, void** bufs, int which, void* zbuf, long long B,
                        void* f, cudaStream_t st);

template <typename T>
static int inverse_impl(const snfft_plan* p, const void* F, long long B, void* f, void* ws, cudaStream_t st) {
    const size_t half = (size_t)B * p->host.factorial[p->n] * sizeof(T);
    void* bufs[2] = {ws, (char*)ws + half};
    void* zbuf = (char*)ws + 2 * half;
    int rc;
    if (g_no_base.load() == 0 && p->lev[p->n].tz_inv && tz_pack_ok<T>(p, p->n, B, bufs[1], zbuf, F)) {
        // the top level's T kernel reads the packed coefficients itself
        if ((rc = run_tz_level<T, true>(p, p->n, F, bufs[1], zbuf, B, st, true))) return rc;
        return inverse_tail<T>(p, p->n - 1, bufs[1], bufs, 0, zbuf, B, f, st);
    }
    if ((rc = run_pack<T, false>(p, F, bufs[0], B, st))) return rc;
    return inverse_tail<T>(p, p->n, bufs[0], bufs, 1, zbuf, B, f, st);
}

// level steps k_top .. 2 of the inverse pipeline and the final scatter: cur = X_{k_top} on entry
template <typename T>
static int inverse_tail(const snfft_plan* p, int k_top, const void* cur, void** bufs, int which, void* zbuf, long long B,
                        void* f, cudaStream_t st) {
    int rc;
    const int base = g_no_base.load();
    const bool s6 = s6_ok<T>(p, true) && k_top >= 6;
    const int k_last = s6 ? 7 : ((base == 0 && p->n > 5) ? 6 : 2);      // lowest level the per-level kernels handle
    for (int k = k_top; k >= k_last; --k) {
        if (base == 0 && p->lev[k].tz_inv) rc = run_tz_level<T, true>(p, k, cur, bufs[which], zbuf, B, st);
        else if (base == 0 && p->lev[k].col_inv) rc = run_col_level<T, true>(p, k, cur, bufs[which], B, st);
        else rc = run_inv_level<T>(p, k, cur, bufs[which], B, st);
        if (rc) return rc;
        cur = bufs[which];
        which ^= 1;
    }
    if (s6) {
        if ((rc = run_s6<T, true>(p, cur, bufs[which], B, st))) return rc;
        cur = bufs[which];
        which ^= 1;
    } else if (base == 0 && p->n > 5) {
        if ((rc = run_s5<T, true>(p, cur, bufs[which], B, st))) return rc;
        cur = bufs[which];
        which ^= 1;
    }
    return run_gather<T, true>(p, cur, f, B, st);
}

constexpr int kSubFirst = 8;        // lowest level the subset transforms prune (levels below run the column kernels)

template <typename T>
static int forward_subset_impl(const snfft_plan* p, const snfft_subplan* sp, const void* f, long long B, void* F, void* ws,
                               cudaStream_t st) {
    if (sp->k_first > p->n || g_no_base.load() != 0) return forward_impl<T>(p, f, B, F, ws, st);
    const size_t half = (size_t)B * p->host.factorial[p->n] * sizeof(T);
    void *cur = nullptr, *nxt = nullptr;
    int rc = forward_levels<T>(p, f, B, ws, st, &cur, sp->k_first, &nxt);
    if (rc) return rc;
    for (int k = sp->k_first; k <= p->n; ++k) {
        // the partitions outside the subset come out as zeros
        if (k == p->n) SNFFT_CUDA(cudaMemsetAsync(nxt, 0, half, st));
        const SubLevel& L = sp->lev[k];
        if ((rc = run_fwd_level<T>(p, k, cur, nxt, B, st, 0, nullptr, L.ftasks, &L.franges))) return rc;
        std::swap(cur, nxt);
    }
    return run_pack<T, true>(p, cur, F, B, st);
}

template <typename T>
static int inverse_subset_impl(const snfft_plan* p, const snfft_subplan* sp, const void* F, long long B, void* f, void* ws,
                               cudaStream_t st) {
    if (sp->k_first > p->n || g_no_base.load() != 0) return inverse_impl<T>(p, F, B, f, ws, st);
    const size_t half = (size_t)B * p->host.factorial[p->n] * sizeof(T);
    void* bufs[2] = {ws, (char*)ws + half};
    void* zbuf = (char*)ws + 2 * half;
    int which = 1, rc;
    if ((rc = run_pack<T, false>(p, F, bufs[0], B, st))) return rc;
    const void* cur = bufs[0];
    for (int k = p->n; k >= sp->k_first; --k) {
        // the level below the last pruned step runs the ordinary kernels: it must see zeros outside the subset
        if (k == sp->k_first) SNFFT_CUDA(cudaMemsetAsync(bufs[which], 0, half, st));
        const SubLevel& L = sp->lev[k];
        if ((rc = run_inv_level<T>(p, k, cur, bufs[which], B, st, L.itasks, L.parents, &L.iranges))) return rc;
        cur = bufs[which];
        which ^= 1;
    }
    return inverse_tail<T>(p, sp->k_first - 1, cur, bufs, which, zbuf, B, f, st);
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

// Per-row YOR tables of level n ([n-1][d] per partition: partner, diagonal, off-diagonal in gather form), used by
// snfft_dense_rep and snfft_rho.  Uploaded on first use under a lock; the pointers are published only after all
// three uploads succeeded, so a failed or concurrent first call cannot leave a half-initialised plan behind.
static std::mutex g_rt_mutex;

template <typename T>
static int ensure_row_tables(snfft_plan* p) {
    std::lock_guard<std::mutex> lock(g_rt_mutex);
    if (p->rt_partner) return SNFFT_OK;
    const HostLevel& L = p->host.levels[p->n];
    const int n = p->n;
    std::vector<int32_t> pt;
    std::vector<T> dg, of;
    std::vector<long long> base;
    for (const HostShape& S : L.shapes) {
        base.push_back((long long)pt.size());
        for (int j = 0; j <= n - 2; ++j)
            for (int t = 0; t < S.dim; ++t) {
                pt.push_back(S.partner[j][t]);
                dg.push_back((T)S.diag[j][t]);
                of.push_back(S.partner[j][t] == t ? (T)0 : (T)S.off[j][t]);
            }
    }
    void *d_pt = nullptr, *d_dg = nullptr, *d_of = nullptr;
    int rc;
    if ((rc = upload(p, pt, &d_pt))) return rc;
    if ((rc = upload(p, dg, &d_dg))) return rc;
    if ((rc = upload(p, of, &d_of))) return rc;
    p->rt_base = base;
    p->rt_diag = d_dg;
    p->rt_off = d_of;
    p->rt_partner = d_pt;           // last: the readiness flag
    return SNFFT_OK;
}

template <typename T>
static RowTables<T> row_tables(const snfft_plan* p, int part) {
    RowTables<T> tb;
    tb.partner = (const int*)p->rt_partner + p->rt_base[part];
    tb.diag = (const T*)p->rt_diag + p->rt_base[part];
    tb.off = (const T*)p->rt_off + p->rt_base[part];
    return tb;
}

template <typename T>
static int rho_impl(snfft_plan* p, int part, const int* sigma, void* out, cudaStream_t st) {
    int rc = ensure_row_tables<T>(p);
    if (rc) return rc;
    const int d = p->host.levels[p->n].shapes[part].dim;
    PermArg pa;
    for (int i = 0; i < SNFFT_MAX_N; ++i) pa.sig[i] = i < p->n ? sigma[i] : 0;
    const size_t smem = (size_t)2 * d * sizeof(T);
    dim3 grid((unsigned)d, 1, 1), block(128, 1, 1);
    auto kfn = rho_kernel<T>;
    if ((rc = set_smem(kfn, smem))) return rc;
    SNFFT_LAUNCH(kfn, grid, block, smem, st, row_tables<T>(p, part), d, p->n, pa, (T*)out);
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

template <typename T>
static int dense_rep_impl(snfft_plan* p, int part, void* out, cudaStream_t st) {
    const HostLevel& L = p->host.levels[p->n];
    {
        const int rc0 = ensure_row_tables<T>(p);
        if (rc0) return rc0;
    }
    const int d = L.shapes[part].dim;
    const RowTables<T> tb = row_tables<T>(p, part);
    const long long q = p->ia.nfact * d;
    dim3 grid((unsigned)((q + 31) / 32), 1, 1), block(32, 1, 1);
    auto kfn = dense_rep_kernel<T>;
    const size_t smem = (size_t)2 * d * 32 * sizeof(T);
    int rc;
    if ((rc = set_smem(kfn, smem))) return rc;
    SNFFT_LAUNCH(kfn, grid, block, smem, st, tb, d, (T*)out, p->ia);
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

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

// test hook: how the levels run.  0 = generated kernels (default), != 0 = the table-driven sweep kernels at EVERY
// level (second implementation, used as a cross-check)
void snfft_debug_disable_base(int mode) { g_no_base.store(mode); }
void snfft_debug_tb_mode(int mode) { g_tb_mode.store(mode); }
void snfft_debug_scalar_gather(int on) { g_no_vec_gather.store(on); }

int snfft_plan_create(int n, int dtype, int device, snfft_plan_t** out) {
    if (!out) return set_error(SNFFT_EINVAL, "out is NULL");
    *out = nullptr;
    if (n < 2 || n > SNFFT_MAX_N) return set_error(SNFFT_EINVAL, "n must be in [2, 12]");
    if (dtype != SNFFT_F32 && dtype != SNFFT_F64) return set_error(SNFFT_EINVAL, "dtype must be 0 (f32) or 1 (f64)");
    snfft_plan* p = nullptr;
    DeviceGuard guard(device);          // the caller's current device is restored on every exit path
    if (!guard.ok) return set_error(SNFFT_ECUDA, "cudaSetDevice failed");
    try {
        p = new snfft_plan();
        p->n = n;
        p->dtype = dtype;
        p->device = device;
        p->esize = dtype == SNFFT_F32 ? 4 : 8;
        build_host_plan(n, p->host);
        p->ia = make_index_args(n);
        p->lev.resize(n + 1);
        {
            const HostLevel& L = p->host.levels[n];
            std::vector<long long> offs(L.coef_offsets.begin(), L.coef_offsets.end());
            std::vector<long long> pre;
            long long t = 0;
            for (const HostShape& S : L.shapes) {
                pre.push_back(t);
                t += ((long long)S.dim * S.dim + PK_T - 1) / PK_T;
            }
            p->n_pack_tiles = t;
            int rc = upload(p, offs, &p->d_coef_off);
            if (!rc) {
                std::vector<int> dm(L.dims.begin(), L.dims.end());
                rc = upload(p, dm, &p->d_dims);
            }
            if (!rc) rc = upload(p, pre, &p->d_tile_prefix);
            for (int tr : {128, 64}) {
                std::vector<long long> vpre;
                long long vt = 0;
                for (const HostShape& S : L.shapes) {
                    vpre.push_back(vt);
                    vt += ((long long)S.dim * S.dim + tr - 1) / tr;
                }
                if (!rc) rc = upload(p, vpre, tr == 128 ? &p->d_vtile_prefix32 : &p->d_vtile_prefix64);
                (tr == 128 ? p->n_vtiles32 : p->n_vtiles64) = vt;
            }
#ifdef SNFFT_HOST_EMU
            const int p1_max_n = 7;               // the emulation runs one fiber per rank: keep the table small there
#else
            const int p1_max_n = 11;
#endif
            if (!rc && n >= 4 && n <= p1_max_n) {       // 4 * n! bytes: 160 MB at n = 11
                void* d = nullptr;
                if (cudaMalloc(&d, (size_t)p->ia.nfact * sizeof(uint32_t)) == cudaSuccess) {
                    p->allocs.push_back(d);
                    p->d_p1 = d;
                    dim3 grid((unsigned)((p->ia.nfact + 255) / 256), 1, 1), block(256, 1, 1);
                    SNFFT_LAUNCH(chain_index32_kernel, grid, block, 0, (cudaStream_t)0, (uint32_t*)d, p->ia);
                    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize((cudaStream_t)0) != cudaSuccess) rc = SNFFT_ECUDA;
                }                                  // allocation failure: the kernels fall back to computing the index
            }
            if (rc) {
                snfft_plan_destroy(p);
                return rc;
            }
        }
        for (int k = 2; k <= n; ++k) {
            int rc = dtype == SNFFT_F32 ? build_level<float>(p, k) : build_level<double>(p, k);
            col_pick_dirs(k, &p->lev[k].col_fwd, &p->lev[k].col_inv);
            tz_pick_dirs(k, &p->lev[k].tz_fwd, &p->lev[k].tz_inv);
            if (!rc) rc = tz_prepare(dtype, k, p->lev[k].tz_fwd, p->lev[k].tz_inv);
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
    // two ping-pong level arrays, plus the intermediate array of the top / bottom kernels when a level uses them
    return (p->n >= kTzFirst ? 3 : 2) * (size_t)B * (size_t)p->host.factorial[p->n] * p->esize;
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

size_t snfft_forward_cosets_workspace_bytes(const snfft_plan_t* pn, const snfft_plan_t* pm, int64_t B, int ncosets) {
    if (!pn || !pm || B < 0 || ncosets < 0) return 0;
    const size_t Bm = (size_t)B * (size_t)ncosets;
    return Bm * (size_t)pm->host.factorial[pm->n] * pm->esize + snfft_workspace_bytes(pm, (int64_t)Bm) +
           (size_t)B * (size_t)pn->host.factorial[pn->n] * pn->esize;
}

int snfft_forward_cosets(const snfft_plan_t* pn, const snfft_plan_t* pm, const void* f, int64_t B, const int* cosets,
                         int ncosets, void* F, void* ws, size_t ws_bytes, void* stream) {
    if (!pn || !pm) return set_error(SNFFT_EINVAL, "plan is NULL");
    if (pm->n != pn->n - 1 || pm->dtype != pn->dtype || pm->device != pn->device)
        return set_error(SNFFT_EINVAL, "the second plan must be the S_{n-1} plan of the same dtype and device");
    if (pn->n < 3) return set_error(SNFFT_EINVAL, "coset sharding needs n >= 3");
    if (B < 0 || ncosets < 0 || ncosets > pn->n) return set_error(SNFFT_EINVAL, "bad batch or coset count");
    if (B == 0) return SNFFT_OK;
    if (!f || !F || (ncosets && !cosets)) return set_error(SNFFT_EINVAL, "NULL data pointer");
    for (int j = 0; j < ncosets; ++j)
        if (cosets[j] < 0 || cosets[j] >= pn->n || (j > 0 && cosets[j] <= cosets[j - 1]))
            return set_error(SNFFT_EINVAL, "cosets must be strictly ascending indices in [0, n)");
    const size_t out_bytes = (size_t)B * (size_t)pn->host.factorial[pn->n] * pn->esize;
    cudaStream_t st = (cudaStream_t)stream;
    if (ncosets == 0) {              // a rank that owns nothing contributes zero
        SNFFT_CUDA(cudaMemsetAsync(F, 0, out_bytes, st));
        return SNFFT_OK;
    }
    if (!ws || ws_bytes < snfft_forward_cosets_workspace_bytes(pn, pm, B, ncosets))
        return set_error(SNFFT_EWORKSPACE, "workspace too small");
    return pn->dtype == SNFFT_F32 ? forward_cosets_impl<float>(pn, pm, f, B, cosets, ncosets, F, ws, st)
                                  : forward_cosets_impl<double>(pn, pm, f, B, cosets, ncosets, F, ws, st);
}

size_t snfft_forward_cosets2_workspace_bytes(const snfft_plan_t* pn, const snfft_plan_t* pm, int64_t B, int npairs) {
    if (!pn || !pm || B < 0 || npairs < 0) return 0;
    const size_t Bm = (size_t)B * (size_t)npairs;
    return Bm * (size_t)pm->host.factorial[pm->n] * pm->esize + snfft_workspace_bytes(pm, (int64_t)Bm) +
           3 * (size_t)B * (size_t)pn->host.factorial[pn->n] * pn->esize;
}

int snfft_forward_cosets2(const snfft_plan_t* pn, const snfft_plan_t* pm, const void* f, int64_t B, const int* top,
                          const int* sub, int npairs, void* F, void* ws, size_t ws_bytes, void* stream) {
    if (!pn || !pm) return set_error(SNFFT_EINVAL, "plan is NULL");
    if (pm->n != pn->n - 2 || pm->dtype != pn->dtype || pm->device != pn->device)
        return set_error(SNFFT_EINVAL, "the second plan must be the S_{n-2} plan of the same dtype and device");
    const int n = pn->n;
    if (n < 4) return set_error(SNFFT_EINVAL, "two-level coset sharding needs n >= 4");
    if (B < 0 || npairs < 0 || npairs > n * (n - 1)) return set_error(SNFFT_EINVAL, "bad batch or pair count");
    if (B == 0) return SNFFT_OK;
    if (!f || !F || (npairs && (!top || !sub))) return set_error(SNFFT_EINVAL, "NULL data pointer");
    std::vector<char> seen((size_t)n * (n - 1), 0);
    for (int j = 0; j < npairs; ++j) {
        if (top[j] < 0 || top[j] >= n || sub[j] < 0 || sub[j] >= n - 1)
            return set_error(SNFFT_EINVAL, "pairs must be (i_n in [0, n), i_{n-1} in [0, n-1))");
        char& s = seen[(size_t)top[j] * (n - 1) + sub[j]];
        if (s) return set_error(SNFFT_EINVAL, "duplicate two-level coset");
        s = 1;
    }
    const size_t out_bytes = (size_t)B * (size_t)pn->host.factorial[n] * pn->esize;
    cudaStream_t st = (cudaStream_t)stream;
    if (npairs == 0) {
        SNFFT_CUDA(cudaMemsetAsync(F, 0, out_bytes, st));
        return SNFFT_OK;
    }
    if ((long long)B * n * (n - 1) >= (1ll << 31)) return set_error(SNFFT_ENOTSUP, "batch too large for the pair slots");
    if (!ws || ws_bytes < snfft_forward_cosets2_workspace_bytes(pn, pm, B, npairs))
        return set_error(SNFFT_EWORKSPACE, "workspace too small");
    return pn->dtype == SNFFT_F32 ? forward_cosets2_impl<float>(pn, pm, f, B, top, sub, npairs, F, ws, st)
                                  : forward_cosets2_impl<double>(pn, pm, f, B, top, sub, npairs, F, ws, st);
}

extern "C++" {
template <typename V>
static int sub_upload(snfft_subplan* sp, const std::vector<V>& v, void** out) {
    *out = nullptr;
    void* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(v.size() * sizeof(V), 16)) != cudaSuccess)
        return set_error(SNFFT_ENOMEM, "cudaMalloc of subset-plan table failed");
    sp->allocs.push_back(d);
    if (!v.empty()) SNFFT_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
    *out = d;
    return SNFFT_OK;
}
}  // extern "C++"

int snfft_subplan_create(const snfft_plan_t* p, const int32_t* keep, snfft_subplan_t** out) {
    if (!p || !keep || !out) return set_error(SNFFT_EINVAL, "NULL plan, mask or output");
    *out = nullptr;
    const int n = p->n;
    DeviceGuard guard(p->device);
    if (!guard.ok) return set_error(SNFFT_ECUDA, "cudaSetDevice failed");
    snfft_subplan* sp = new (std::nothrow) snfft_subplan();
    if (!sp) return set_error(SNFFT_ENOMEM, "out of host memory");
    sp->plan = p;
    sp->keep.resize(n + 1);
    sp->lev.resize(n + 1);
    const HostLevel& Ln = p->host.levels[n];
    sp->keep[n].assign(Ln.shapes.size(), 0);
    double kept = 0.0;
    bool any = false;
    for (size_t q = 0; q < Ln.shapes.size(); ++q)
        if (keep[q]) {
            sp->keep[n][q] = 1;
            kept += (double)Ln.shapes[q].dim * Ln.shapes[q].dim;
            any = true;
        }
    if (!any) {
        delete sp;
        return set_error(SNFFT_EINVAL, "the subset of partitions is empty");
    }
    sp->kept_fraction = kept / (double)p->host.factorial[n];
    sp->k_first = n >= kSubFirst ? kSubFirst : n + 1;
    for (int k = n; k >= sp->k_first; --k) {           // down-closure: the children of every kept shape
        const HostLevel& L = p->host.levels[k];
        sp->keep[k - 1].assign(p->host.levels[k - 1].shapes.size(), 0);
        for (size_t q = 0; q < L.shapes.size(); ++q)
            if (sp->keep[k][q])
                for (int c : L.shapes[q].child) sp->keep[k - 1][c] = 1;
    }
    int rc = SNFFT_OK;
    for (int k = sp->k_first; k <= n && !rc; ++k) {
        const HostLevel& L = p->host.levels[k];
        const HostLevel& P = p->host.levels[k - 1];
        const DevLevel& D = p->lev[k];
        auto shape_of = [](const HostLevel& H, long long coef_off) {
            for (size_t q = 0; q < H.shapes.size(); ++q)
                if (H.shapes[q].coef_off == coef_off) return (int)q;
            return -1;
        };
        SubLevel& S = sp->lev[k];
        std::vector<FwdTask> ft;
        for (const TaskRange& r : D.franges) {
            TaskRange nr{r.cls, (int)ft.size(), 0, 0, 0, 0, 0};
            for (int t = r.begin; t < r.begin + r.count; ++t) {
                const FwdTask& task = D.h_ftasks[t];
                const int q = shape_of(L, task.lam_off);
                if (q < 0 || !sp->keep[k][q]) continue;
                ft.push_back(task);
                nr.count += 1;
                nr.elems += (long long)task.d_lam * task.d_mu;
                nr.max_tab = std::max(nr.max_tab, task.tab_words);
                nr.max_dmu = std::max(nr.max_dmu, task.d_mu);
                nr.max_dlam = std::max(nr.max_dlam, task.d_lam);
            }
            if (nr.count) S.franges.push_back(nr);
        }
        std::vector<InvTask> it;
        std::vector<InvParent> par;
        for (const TaskRange& r : D.iranges) {
            TaskRange nr{r.cls, (int)it.size(), 0, 0, 0, 0, 0};
            for (int t = r.begin; t < r.begin + r.count; ++t) {
                InvTask task = D.h_itasks[t];
                const int m = shape_of(P, task.mu_off);
                if (m < 0 || !sp->keep[k - 1][m]) continue;
                const int par0 = (int)par.size();
                int maxd = 0, maxtab = 0;
                for (int q = task.par_off; q < task.par_off + task.npar; ++q) {
                    const InvParent& ip = D.h_parents[q];
                    const int lam = shape_of(L, ip.lam_off);
                    if (lam < 0 || !sp->keep[k][lam]) continue;
                    par.push_back(ip);
                    maxd = std::max(maxd, ip.d_lam);
                    maxtab = std::max(maxtab, ip.tab_words);
                }
                task.par_off = par0;
                task.npar = (int)par.size() - par0;
                if (task.npar == 0) continue;             // cannot happen: mu is the child of a kept shape
                it.push_back(task);
                nr.count += 1;
                nr.elems += (long long)k * task.d_mu * task.d_mu;
                nr.max_tab = std::max(nr.max_tab, maxtab);
                nr.max_dmu = std::max(nr.max_dmu, task.d_mu);
                nr.max_dlam = std::max(nr.max_dlam, maxd);
            }
            if (nr.count) S.iranges.push_back(nr);
        }
        if (!rc) rc = sub_upload(sp, ft, &S.ftasks);
        if (!rc) rc = sub_upload(sp, it, &S.itasks);
        if (!rc) rc = sub_upload(sp, par, &S.parents);
    }
    if (rc) {
        snfft_subplan_destroy(sp);
        return rc;
    }
    *out = sp;
    return SNFFT_OK;
}

void snfft_subplan_destroy(snfft_subplan_t* sp) {
    if (!sp) return;
    for (void* d : sp->allocs) cudaFree(d);
    delete sp;
}

int snfft_subplan_info(const snfft_subplan_t* sp, int* first_pruned_level, double* kept_fraction) {
    if (!sp) return set_error(SNFFT_EINVAL, "subset plan is NULL");
    if (first_pruned_level) *first_pruned_level = sp->k_first;
    if (kept_fraction) *kept_fraction = sp->kept_fraction;
    return SNFFT_OK;
}

int snfft_forward_subset(const snfft_plan_t* p, const snfft_subplan_t* sp, const void* f, int64_t B, void* F, void* ws,
                         size_t ws_bytes, void* stream) {
    int rc = check_io(p, f, F, B, ws, ws_bytes);
    if (rc || B == 0) return rc;
    if (!sp || sp->plan != p) return set_error(SNFFT_EINVAL, "the subset plan does not belong to this plan");
    cudaStream_t st = (cudaStream_t)stream;
    return p->dtype == SNFFT_F32 ? forward_subset_impl<float>(p, sp, f, B, F, ws, st)
                                 : forward_subset_impl<double>(p, sp, f, B, F, ws, st);
}

int snfft_inverse_subset(const snfft_plan_t* p, const snfft_subplan_t* sp, const void* F, int64_t B, void* f, void* ws,
                         size_t ws_bytes, void* stream) {
    int rc = check_io(p, F, f, B, ws, ws_bytes);
    if (rc || B == 0) return rc;
    if (!sp || sp->plan != p) return set_error(SNFFT_EINVAL, "the subset plan does not belong to this plan");
    cudaStream_t st = (cudaStream_t)stream;
    return p->dtype == SNFFT_F32 ? inverse_subset_impl<float>(p, sp, F, B, f, ws, st)
                                 : inverse_subset_impl<double>(p, sp, F, B, f, ws, st);
}

int snfft_block_matmul(const snfft_plan_t* p, const void* A, const void* Bm, void* C, int64_t B, void* stream) {
    if (!p) return set_error(SNFFT_EINVAL, "plan is NULL");
    if (B < 0) return set_error(SNFFT_EINVAL, "negative batch");
    if (B == 0) return SNFFT_OK;
    if (!A || !Bm || !C) return set_error(SNFFT_EINVAL, "NULL data pointer");
    const HostLevel& L = p->host.levels[p->n];
    cudaStream_t st = (cudaStream_t)stream;
    if (L.shapes.size() > (size_t)MM_MAXPARTS) return set_error(SNFFT_ENOTSUP, "snfft_block_matmul: too many partitions");
    std::vector<int> order(L.shapes.size());
    for (size_t q = 0; q < order.size(); ++q) order[q] = (int)q;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return L.shapes[x].dim > L.shapes[y].dim; });
    MMParts mp;
    mp.nparts = (int)order.size();
    // smallest d that runs the 128 x 128 tile path (fp32); SNFFT_MM_BIG_MIN overrides (1 = every partition: tests)
    mp.big_min = [] {
        const char* e = std::getenv("SNFFT_MM_BIG_MIN");
        return e ? std::atoi(e) : 160;
    }();
    long long nblk = 0;
    for (size_t i = 0; i < order.size(); ++i) {
        const int d = L.shapes[order[i]].dim;
        const int tile = mm_tile_of(d, mp.big_min, p->dtype == SNFFT_F32);
        const long long tiles = (d + tile - 1) / tile;
        mp.d[i] = d;
        mp.blk0[i] = nblk;
        mp.off[i] = (long long)L.coef_offsets[order[i]] * (long long)B;
        nblk += (long long)B * tiles * tiles;
    }
    mp.blk0[order.size()] = nblk;
    if (nblk > 0x7fffffffll) return set_error(SNFFT_ENOTSUP, "snfft_block_matmul: batch too large");
    dim3 grid((unsigned)nblk, 1, 1), block(MM_THREADS, 1, 1);
    if (p->dtype == SNFFT_F32) {
        auto kfn = block_matmul_kernel<float>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const float*)A, (const float*)Bm, (float*)C, mp);
    } else {
        auto kfn = block_matmul_kernel<double>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const double*)A, (const double*)Bm, (double*)C, mp);
    }
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
}

int snfft_power(const snfft_plan_t* p, const void* F, int64_t B, void* out, void* stream) {
    if (!p) return set_error(SNFFT_EINVAL, "plan is NULL");
    if (B < 0) return set_error(SNFFT_EINVAL, "negative batch");
    if (B == 0) return SNFFT_OK;
    if (!F || !out) return set_error(SNFFT_EINVAL, "NULL data pointer");
    const HostLevel& L = p->host.levels[p->n];
    dim3 grid((unsigned)L.shapes.size(), (unsigned)std::min<int64_t>(B, 65535), 1), block(256, 1, 1);
    const double inv_order = 1.0 / (double)p->host.factorial[p->n];
    cudaStream_t st = (cudaStream_t)stream;
    if (p->dtype == SNFFT_F32) {
        auto kfn = power_kernel<float>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const float*)F, (const long long*)p->d_coef_off, (const int*)p->d_dims,
                     (long long)B, inv_order, (float*)out);
    } else {
        auto kfn = power_kernel<double>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const double*)F, (const long long*)p->d_coef_off, (const int*)p->d_dims,
                     (long long)B, inv_order, (double*)out);
    }
    g_launches.fetch_add(1);
    SNFFT_CUDA(cudaGetLastError());
    return SNFFT_OK;
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
    DeviceGuard guard(p->device);
    if (!guard.ok) return set_error(SNFFT_ECUDA, "cudaSetDevice failed");
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

int snfft_dense_rep(snfft_plan_t* p, int part, void* out, void* stream) {
    if (!p || !out) return set_error(SNFFT_EINVAL, "NULL plan or output");
    const HostLevel& L = p->host.levels[p->n];
    if (part < 0 || part >= (int)L.shapes.size()) return set_error(SNFFT_EINVAL, "bad partition index");
    if ((size_t)2 * L.shapes[part].dim * 32 * p->esize > 200 * 1024)
        return set_error(SNFFT_ENOTSUP, "snfft_dense_rep: irrep too large for the dense cross-check");
    cudaStream_t st = (cudaStream_t)stream;
    return p->dtype == SNFFT_F32 ? dense_rep_impl<float>(p, part, out, st) : dense_rep_impl<double>(p, part, out, st);
}

static bool is_permutation(int n, const int* sigma) {
    if (!sigma) return false;
    bool seen[SNFFT_MAX_N] = {};
    for (int i = 0; i < n; ++i) {
        if (sigma[i] < 0 || sigma[i] >= n || seen[sigma[i]]) return false;
        seen[sigma[i]] = true;
    }
    return true;
}

int snfft_rho(snfft_plan_t* p, int part, const int* sigma, void* out, void* stream) {
    if (!p || !out) return set_error(SNFFT_EINVAL, "NULL plan or output");
    const HostLevel& L = p->host.levels[p->n];
    if (part < 0 || part >= (int)L.shapes.size()) return set_error(SNFFT_EINVAL, "bad partition index");
    if (!is_permutation(p->n, sigma)) return set_error(SNFFT_EINVAL, "sigma is not a one-line permutation of range(n)");
    cudaStream_t st = (cudaStream_t)stream;
    return p->dtype == SNFFT_F32 ? rho_impl<float>(p, part, sigma, out, st) : rho_impl<double>(p, part, sigma, out, st);
}

int snfft_translate(int n, int dtype, const void* f, int64_t B, const int* perm, int right, void* out, void* stream) {
    if (n < 1 || n > SNFFT_MAX_N) return set_error(SNFFT_EINVAL, "bad n");
    if (dtype != SNFFT_F32 && dtype != SNFFT_F64) return set_error(SNFFT_EINVAL, "dtype must be 0 (f32) or 1 (f64)");
    if (B < 0) return set_error(SNFFT_EINVAL, "negative batch");
    if (!is_permutation(n, perm)) return set_error(SNFFT_EINVAL, "perm is not a one-line permutation of range(n)");
    if (B == 0) return SNFFT_OK;
    if (!f || !out || f == out) return set_error(SNFFT_EINVAL, "NULL data pointer or in-place translation");
    TranslateArgs ta;
    ta.ia = make_index_args(n);
    ta.ia.batch = B;
    ta.right = right ? 1 : 0;
    for (int i = 0; i < SNFFT_MAX_N; ++i) ta.pinv[i] = 0;
    for (int i = 0; i < n; ++i) ta.pinv[perm[i]] = i;
    dim3 grid((unsigned)((ta.ia.nfact + 255) / 256), 1, 1), block(256, 1, 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SNFFT_F32) {
        auto kfn = translate_kernel<float>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const float*)f, (float*)out, ta);
    } else {
        auto kfn = translate_kernel<double>;
        SNFFT_LAUNCH(kfn, grid, block, 0, st, (const double*)f, (double*)out, ta);
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
