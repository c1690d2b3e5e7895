/* CPU oracle, C inner loops  --  TEST INFRASTRUCTURE ONLY (see oracle/sn_oracle.py).
 *
 * Plain C restatement of one Clausen level step of the reference
 * (/root/reference/src/algebraist/fourier.py:253-273 forward; the adjoint with
 * the d_lambda/(k d_mu) weights of fourier.py:191-199 for the inverse).  The
 * Young's-orthogonal-form tables (irreps.py:91-109) are passed in from
 * oracle/sn_oracle.py; this file only does the arithmetic, in the same order
 * as the numpy restatement, so both give bit-identical fp64 results.
 * Pinned by tests/test_oracle_c.py against the numpy oracle and the reference
 * goldens.  Used as the CPU baseline of bench.py; never by the product.
 *
 * Data: a level-k array of partition lambda is (d, d, ninst) C-contiguous,
 * instance fastest; the k cosets are the planes [i*ninst, (i+1)*ninst) of the
 * level-(k-1) arrays.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DEFINE_LEVEL(T, SUF)                                                                      \
    /* w_out[t, :] = diag[t] * w[t, :] + off[t] * w[partner[t], :]   (rho(s_j) @ w) */           \
    static void apply_##SUF(int d, int64_t row, const int32_t* partner, const double* diag,       \
                            const double* off, const T* w, T* out) {                              \
        for (int t = 0; t < d; ++t) {                                                             \
            const T a = (T)diag[t], b = (T)off[t];                                                \
            const T* x = w + (int64_t)t * row;                                                    \
            const T* y = w + (int64_t)partner[t] * row;                                           \
            T* o = out + (int64_t)t * row;                                                        \
            if (b == (T)0) {                                                                      \
                for (int64_t q = 0; q < row; ++q) o[q] = a * x[q];                                \
            } else {                                                                              \
                for (int64_t q = 0; q < row; ++q) o[q] = a * x[q] + b * y[q];                     \
            }                                                                                     \
        }                                                                                         \
    }                                                                                             \
    /* F_lambda = sum_i rho(s_i)...rho(s_{k-2}) blockdiag_mu F^(i)_mu   (fourier.py:259-273) */   \
    int snft_level_forward_##SUF(int k, int d, int nblocks, const int32_t* boff,                  \
                                 const int32_t* dmu, const T* const* prev, int64_t ninst,         \
                                 const int32_t* partner, const double* diag, const double* off,   \
                                 T* out) {                                                        \
        const int64_t row = (int64_t)d * ninst, tot = (int64_t)d * row;                           \
        T* w = (T*)malloc(sizeof(T) * tot);                                                       \
        T* w2 = (T*)malloc(sizeof(T) * tot);                                                      \
        if (!w || !w2) { free(w); free(w2); return -2; }                                          \
        memset(out, 0, sizeof(T) * tot);                                                          \
        for (int i = 0; i < k; ++i) {                                                             \
            memset(w, 0, sizeof(T) * tot);                                                        \
            for (int b = 0; b < nblocks; ++b) {                                                   \
                const int dm = dmu[b], o = boff[b];                                               \
                for (int r = 0; r < dm; ++r)                                                      \
                    for (int c = 0; c < dm; ++c)                                                  \
                        memcpy(w + ((int64_t)(o + r) * d + o + c) * ninst,                        \
                               prev[b] + ((int64_t)r * dm + c) * (k * ninst) + i * ninst,         \
                               sizeof(T) * ninst);                                                \
            }                                                                                     \
            for (int j = k - 2; j >= i; --j) {                                                    \
                apply_##SUF(d, row, partner + (int64_t)j * d, diag + (int64_t)j * d,              \
                            off + (int64_t)j * d, w, w2);                                         \
                T* t_ = w; w = w2; w2 = t_;                                                       \
            }                                                                                     \
            for (int64_t q = 0; q < tot; ++q) out[q] += w[q];                                     \
        }                                                                                         \
        free(w); free(w2);                                                                        \
        return 0;                                                                                 \
    }                                                                                             \
    /* G^(i)_mu += d_lambda/(k d_mu) [rho(s_{k-2})...rho(s_i) F_lambda]_{mu block} */             \
    int snft_level_inverse_##SUF(int k, int d, int nblocks, const int32_t* boff,                  \
                                 const int32_t* dmu, T* const* next, int64_t ninst,               \
                                 const int32_t* partner, const double* diag, const double* off,   \
                                 const T* cur) {                                                  \
        const int64_t row = (int64_t)d * ninst, tot = (int64_t)d * row;                           \
        T* w = (T*)malloc(sizeof(T) * tot);                                                       \
        T* w2 = (T*)malloc(sizeof(T) * tot);                                                      \
        if (!w || !w2) { free(w); free(w2); return -2; }                                          \
        for (int i = 0; i < k; ++i) {                                                             \
            memcpy(w, cur, sizeof(T) * tot);                                                      \
            for (int j = i; j <= k - 2; ++j) {                                                    \
                apply_##SUF(d, row, partner + (int64_t)j * d, diag + (int64_t)j * d,              \
                            off + (int64_t)j * d, w, w2);                                         \
                T* t_ = w; w = w2; w2 = t_;                                                       \
            }                                                                                     \
            for (int b = 0; b < nblocks; ++b) {                                                   \
                const int dm = dmu[b], o = boff[b];                                               \
                const T scale = (T)((double)d / ((double)k * (double)dm));                        \
                for (int r = 0; r < dm; ++r)                                                      \
                    for (int c = 0; c < dm; ++c) {                                                \
                        T* dst = next[b] + ((int64_t)r * dm + c) * (k * ninst) + i * ninst;       \
                        const T* src = w + ((int64_t)(o + r) * d + o + c) * ninst;                \
                        for (int64_t q = 0; q < ninst; ++q) dst[q] += scale * src[q];             \
                    }                                                                             \
            }                                                                                     \
        }                                                                                         \
        free(w); free(w2);                                                                        \
        return 0;                                                                                 \
    }

DEFINE_LEVEL(double, f64)
DEFINE_LEVEL(float, f32)
