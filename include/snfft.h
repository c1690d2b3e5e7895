/*
 * snfft.h -- C ABI of the B200-native S_n fast Fourier transform.
 *
 * This is the drop-in boundary for the hot path of dashstander/algebraist
 * (`sn_fft` / `sn_ifft`, reference src/algebraist/fourier.py).  The reference
 * is pure Python and has no FFI of its own; each entry point below names the
 * reference function(s) whose work it replaces.  The Python package
 * `algebraist_b200` binds this header with ctypes and
 * re-creates the reference's Python signatures on top of it.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary
 *   - every function returns 0 on success or a negative SNFFT_E* status and
 *     never throws; snfft_last_error() gives a thread-local message
 *   - all data pointers are DEVICE pointers unless the name says `_host`
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *     every kernel is launched asynchronously on it, no host synchronisation
 *   - dtype: 0 = float32, 1 = float64
 *
 * Data layout
 *   function values  f : [B][n!]      row r = f(sigma_r), sigma_r the r-th
 *                                     permutation of range(n) in lexicographic
 *                                     order (utils.py:61-63)
 *   coefficients     F : packed; for each partition lambda in the reference's
 *                        generate_partitions(n) order (tableau.py:103-123) one
 *                        contiguous [B][d][d] row-major block that starts at
 *                        element  B * coef_offset[lambda]   (sum_lambda d^2 = n!)
 */
#ifndef SNFFT_H
#define SNFFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNFFT_OK            0
#define SNFFT_EINVAL       -1   /* bad argument (n out of range, NULL, bad dtype, B < 0) */
#define SNFFT_ENOMEM       -2   /* host or device allocation failed                       */
#define SNFFT_ECUDA        -3   /* a CUDA runtime call or kernel launch failed            */
#define SNFFT_EWORKSPACE   -4   /* workspace smaller than snfft_workspace_bytes()         */
#define SNFFT_ENOTSUP      -5   /* valid request this build cannot serve                  */

#define SNFFT_F32 0
#define SNFFT_F64 1

#define SNFFT_MAX_N 12

typedef struct snfft_plan snfft_plan_t;

/* Library / build identification ("sm_100a" for the product build). */
const char* snfft_build_info(void);

/* Thread-local description of the last failure on this thread. */
const char* snfft_last_error(void);

/*
 * Build the per-n tables once (host combinatorics + device upload).
 * Replaces, at plan time, everything the reference rebuilds on every call:
 *   generate_partitions           tableau.py:103-123
 *   SnIrrep.__init__/dim          irreps.py:40-44, tableau.py:236-265
 *   split_partition / get_block_indices   irreps.py:64-83
 *   basis order                   tableau.py:171-198, irreps.py:60-62
 *   adj_transposition_matrix      irreps.py:91-109  (kept as sparse per-row tables)
 *   coset_rep_matrices            irreps.py:150-155 (never materialised; words s_i..s_{n-2})
 * 2 <= n <= SNFFT_MAX_N.  `device` is the CUDA device ordinal.
 */
int snfft_plan_create(int n, int dtype, int device, snfft_plan_t** out);
void snfft_plan_destroy(snfft_plan_t* plan);

/*
 * Plan introspection.  Any out pointer may be NULL.  The returned arrays are
 * owned by the plan (host memory).
 *   partitions_flat : [n_partitions][n] rows zero padded, reference order
 *   dims            : [n_partitions]    d_lambda           (tableau.py:236-265)
 *   coef_offsets    : [n_partitions+1]  prefix sums of d^2 (last entry = n!)
 */
int snfft_plan_info(const snfft_plan_t* plan, int* n, int* dtype, int* n_partitions,
                    const int32_t** partitions_flat, const int32_t** dims,
                    const int64_t** coef_offsets);

/*
 * Table read-back for tests (host arrays owned by the plan):
 * per-row Young's-orthogonal-form table of s_j = (j j+1) on partition number
 * `part` of level k (irreps.py:91-109):  rho[t][t] = diag[t],
 * rho[t][partner[t]] = offdiag[t]  (partner[t] == t when there is none).
 */
int snfft_plan_level_info(const snfft_plan_t* plan, int k, int* n_partitions,
                          const int32_t** partitions_flat, const int32_t** dims);
int snfft_plan_yor_table(const snfft_plan_t* plan, int k, int part, int j,
                         int* d, const int32_t** partner, const double** diag,
                         const double** offdiag);

/* Bytes of device scratch snfft_forward / snfft_inverse need for batch B. */
size_t snfft_workspace_bytes(const snfft_plan_t* plan, int64_t B);

/*
 * Forward transform of B functions: F(lambda) = sum_sigma f(sigma) rho_lambda(sigma).
 * Replaces sn_fft -> fourier_projection -> _fourier_projection /
 * restrict_to_coset (fourier.py:281-288, 223-278, 111-131, 63-80).
 */
int snfft_forward(const snfft_plan_t* plan, const void* f, int64_t B, void* F,
                  void* workspace, size_t workspace_bytes, void* stream);

/*
 * Inverse transform: f(sigma) = (1/n!) sum_lambda d_lambda <F(lambda), rho_lambda(sigma)>.
 * Replaces sn_ifft -> sn_fourier_decomposition -> inverse_fourier_projection /
 * _inverse_fourier_projection / lift_from_coset (fourier.py:298-299, 291-295,
 * 162-219, 134-159, 46-60).  The reference's recursive inverse raises for
 * n > 5 (fourier.py:202-214); this computes the mathematical inverse at every n.
 */
int snfft_inverse(const snfft_plan_t* plan, const void* F, int64_t B, void* f,
                  void* workspace, size_t workspace_bytes, void* stream);

/*
 * Power spectrum of packed coefficients: out[lambda][b] = d_lambda / n! * ||F_lambda[b]||_F^2
 * (calc_power / _calc_power, fourier.py:349-393; the entries add up to ||f_b||^2, Plancherel).
 * out: [n_partitions][B] in the plan's dtype.  B <= 65535.
 */
int snfft_power(const snfft_plan_t* plan, const void* F, int64_t B, void* out, void* stream);

/*
 * Per-partition matrix products of two packed coefficient arrays (layout of snfft_forward):
 * C_lambda[b] = A_lambda[b] . B_lambda[b] for every partition lambda and batch row b.  With A = FT(g), B = FT(f)
 * and snfft_inverse this is the group convolution sum_sigma f(sigma) g(pi sigma^-1) (convolution theorem, reference
 * tests/test_fourier.py:22-34, 87-109).  C must not alias A or B.
 */
int snfft_block_matmul(const snfft_plan_t* plan, const void* A, const void* B, void* C, int64_t batch, void* stream);

/*
 * Partial forward transform over a subset of the top-level cosets
 * {sigma : sigma[i] = n-1} (the coset-sharded transform of one large function,
 * SURVEY 8e / BASELINE config 5):
 *   F_partial(lambda) = sum_{i in cosets} rho_lambda(c_i) . blockdiag_mu F^(i)_mu
 * i.e. the terms of the top-level sum of fourier_projection (fourier.py:253-273)
 * that the given cosets own; restrict_to_coset (fourier.py:63-80) and the
 * S_{n-1} sub-transforms run inside.  The partial results of a partition of
 * range(n) add up to snfft_forward's output (same packed layout), so a rank
 * calls this with its cosets and the ranks finish with a sum reduction
 * (ncclReduceScatter) -- no other exchange.  `cosets` is a HOST array of
 * strictly ascending indices; plan_nm1 is the S_{n-1} plan of the same dtype
 * and device.  ncosets == 0 writes zeros.
 */
size_t snfft_forward_cosets_workspace_bytes(const snfft_plan_t* plan_n, const snfft_plan_t* plan_nm1, int64_t B,
                                            int ncosets);
int snfft_forward_cosets(const snfft_plan_t* plan_n, const snfft_plan_t* plan_nm1, const void* f, int64_t B,
                         const int* cosets, int ncosets, void* F, void* workspace, size_t workspace_bytes,
                         void* stream);

/*
 * The same partial transform over a subset of the n (n-1) TWO-level cosets
 * (sigma[top[j]] = n-1, and the restriction to that coset has value n-2 at
 * position sub[j]: restrict_to_coset applied twice, fourier.py:63-80,253):
 * finer units of work for sharding ONE function over 8 GPUs than the n
 * top-level cosets.  plan_nm2 is the S_{n-2} plan of the same dtype and
 * device; pairs must be distinct; npairs == 0 writes zeros.  The partial
 * results of a partition of all pairs add up to the full transform.
 */
size_t snfft_forward_cosets2_workspace_bytes(const snfft_plan_t* plan_n, const snfft_plan_t* plan_nm2, int64_t B,
                                             int npairs);
int snfft_forward_cosets2(const snfft_plan_t* plan_n, const snfft_plan_t* plan_nm2, const void* f, int64_t B,
                          const int* top, const int* sub, int npairs, void* F, void* workspace,
                          size_t workspace_bytes, void* stream);

/*
 * Transforms restricted to a subset of the partitions of n (band-limited transforms, e.g. the
 * partitions with lambda_1 >= n - k of ranking / grokking analyses, README.md:11,32).  `keep[q] != 0` selects
 * partition q of the plan (snfft_plan_info order).  The forward transform computes F(lambda) for the kept
 * partitions only (the other blocks of the packed output are zero); the inverse returns
 *   f(sigma) = 1/n! sum_{lambda kept} d_lambda <F(lambda), rho_lambda(sigma)>
 * and ignores the blocks of the other partitions -- with one kept partition that is a component of
 * sn_fourier_decomposition (fourier.py:291-295).  Level steps k >= 8 only touch the shapes contained in a kept
 * partition (Young's lattice down-closure); the levels below run the plan's ordinary kernels.  Same buffers
 * and workspace as snfft_forward / snfft_inverse.  A subset plan belongs to the plan it was created from and
 * must be destroyed before it.
 */
typedef struct snfft_subplan snfft_subplan_t;
int snfft_subplan_create(const snfft_plan_t* plan, const int32_t* keep, snfft_subplan_t** out);
void snfft_subplan_destroy(snfft_subplan_t* sub);
int snfft_subplan_info(const snfft_subplan_t* sub, int* first_pruned_level, double* kept_fraction);
int snfft_forward_subset(const snfft_plan_t* plan, const snfft_subplan_t* sub, const void* f, int64_t B, void* F,
                         void* workspace, size_t workspace_bytes, void* stream);
int snfft_inverse_subset(const snfft_plan_t* plan, const snfft_subplan_t* sub, const void* F, int64_t B, void* f,
                         void* workspace, size_t workspace_bytes, void* stream);

/*
 * Group-action utilities on the device.
 *   rho       : out[d][d] (row major, device) = rho_lambda(sigma) in Young's orthogonal form for the partition
 *               with index `part` of the plan's level n, sigma a one-line permutation of range(n) (host array)
 *               -- SnIrrep.matrix_representations[sigma], irreps.py:122-141; any dimension
 *   translate : out[b][k] = f[b][rank(perm^-1 * p_k)] (right == 0, the left translation of
 *               tests/test_fourier.py:112-121) or f[b][rank(p_k * perm^-1)] (right != 0); p_k = the k-th
 *               permutation in lexicographic order, (a * b)[i] = b[a[i]] (permutations.py:85-88).  f != out.
 */
int snfft_rho(snfft_plan_t* plan, int part, const int* sigma, void* out, void* stream);
int snfft_translate(int n, int dtype, const void* f, int64_t B, const int* perm, int right, void* out, void* stream);

/*
 * Same transforms with HOST buffers (pageable or pinned): copies in, runs the
 * device pipeline, copies out, and returns after the result is in `F_host` /
 * `f_host`.  Scratch is owned by the plan and grows on demand.
 */
int snfft_forward_host(snfft_plan_t* plan, const void* f_host, int64_t B, void* F_host, void* stream);
int snfft_inverse_host(snfft_plan_t* plan, const void* F_host, int64_t B, void* f_host, void* stream);

/*
 * Bit-exact integer kernels.
 *   perm_rank   : lexicographic rank of m one-line permutations of range(n)
 *                 (Permutation.permutation_index, permutations.py:187-192)
 *   perm_unrank : inverse; rows of generate_all_permutations(n) (utils.py:61-63)
 *   coset_index : ranks of {sigma : sigma[coset] = n-1} in increasing order,
 *                 (n-1)! entries  == argwhere(sn_perms[:, coset] == n-1)
 *                 (restrict_to_coset / lift_from_coset, fourier.py:59,79)
 *   chain_index : for every rank r in [0, n!) the level-1 instance number
 *                 p_1 = sum_k i_k * n!/k!, i_k = position of value k-1 after the
 *                 larger values are deleted: the composition of restrict_to_coset
 *                 over all levels (fourier.py:253), i.e. the input permutation
 *                 of the fast transform
 */
int snfft_perm_rank(int n, const int8_t* perms, int64_t m, int64_t* ranks, void* stream);
int snfft_perm_unrank(int n, const int64_t* ranks, int64_t m, int8_t* perms, void* stream);
int snfft_coset_index(int n, int coset, int64_t* idx, void* stream);
int snfft_chain_index(int n, int64_t* p1_of_rank, void* stream);

/*
 * Dense representation matrices of partition number `part` (reference order):
 * out[n!][d][d], out[r] = rho_lambda(sigma_r).  Replaces SnIrrep.matrix_tensor /
 * matrix_representations (irreps.py:122-148); used by slow_sn_ft / slow_sn_ift
 * (fourier.py:82-108, 302-327).  n!*d*d elements: practical to n = 8.
 */
int snfft_dense_rep(snfft_plan_t* plan, int part, void* out, void* stream);

/*
 * Per-launch device timing with CUDA events on the launching stream.
 * Between snfft_profile_begin and snfft_profile_end every kernel this plan
 * launches is bracketed by an event pair; _end synchronises on them and returns
 * one aggregated record per (kind, level, kernel class).
 *   kind: 0 forward level, 1 inverse level, 2 input gather, 3 output scatter,
 *         4 forward base kernel (level = 5: levels 2..5 in registers), 5 inverse base kernel,
 *         6 pack (top level -> [lambda][B][d][d]), 7 unpack
 *   kernel_class: 0..5 table-driven sweep-kernel classes, -1 generated base kernel / scalar data movement,
 *         -2 vectorised data movement, -3 register column kernels (levels 6..8: col_kernels.cuh),
 *         -4 / -5 the Z / T kernel of a top / bottom level step (levels 9..11: tz_kernels.cuh),
 *         -6 small-batch gather / scatter and the B = 1 pack copy,
 *         -7 the T kernel of the TOP level reading / writing the packed coefficients itself (no pack / unpack launch:
 *            batches with full instance vectors, B % 4 == 0 fp32, B % 2 == 0 fp64),
 *         -8 levels 2..6 in one kernel (fp32, n >= 6; kind 4 / 5, level 6: s6_kernels.cuh)
 *   elems_per_instance: coefficients the launch produces per function
 *     (its share of k!; the launch reads as many; the Z kernel: MB/k of k!), so its algorithmic traffic
 *     is 2 * elems_per_instance * NINST_level * sizeof(dtype) bytes.
 */
typedef struct snfft_profile_record {
    int32_t kind, level, kernel_class, launches;
    int64_t elems_per_instance;
    double total_ms;
} snfft_profile_record_t;
int snfft_profile_begin(snfft_plan_t* plan);
int snfft_profile_end(snfft_plan_t* plan, snfft_profile_record_t* records, int max_records, int* n_records);

/* Number of kernel launches issued by this library since load (all threads). */
int64_t snfft_launch_count(void);

/* Test hook: shapes whose dimension fits kernel class `cls` (0..5) run in that
 * class instead of the smallest fitting one; -1 restores the automatic choice.
 * Affects plans created afterwards.  Lets small n exercise every kernel class. */
void snfft_debug_force_class(int cls);

/* Test hook: how levels 2..5 run for n > 5.  0 = generated register kernel (default), != 0 = the per-level
 * table-driven sweep kernels at EVERY level; both must agree, so the second implementation stays covered.
 * Read at every call. */
void snfft_debug_disable_base(int mode);

/* Test hook: which kernels run the level steps k >= 6.  -1 = automatic (generated column kernels at k = 6..8,
 * top / bottom kernels at k = 9..11, table-driven sweep kernels above), >= 0 = sweep kernels only.  Affects plans
 * created afterwards (the Python plan cache: algebraist_b200.plan.clear_plan_cache); both must agree with the oracle. */
void snfft_debug_tb_mode(int mode);

/* Test hook: 1 = the input gather / output scatter always run the scalar tile kernel (the one used for
 * batches that are not a multiple of 16 bytes), 0 = vectorised kernel when the batch allows (default). */
void snfft_debug_scalar_gather(int on);

#ifdef __cplusplus
}
#endif
#endif /* SNFFT_H */
