/*
 * flacenc_oracle.c -- scalar C restatement of flacenc-rs's per-frame analysis/encode path.
 *
 * TEST INFRASTRUCTURE ONLY (see flacenc_oracle.h).  Follows the reference's *stable* build
 * (scalar fakesimd, sequential order).  Every function cites the reference lines it restates
 * (paths relative to /root/reference/).  Compile with -ffp-contract=off: every fused
 * multiply-add below is an explicit fma()/fmaf() exactly where the reference calls mul_add.
 */
#define _GNU_SOURCE
#include "flacenc_oracle.h"

#include <assert.h>
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define FO_MIN(a, b) ((a) < (b) ? (a) : (b))
#define FO_MAX(a, b) ((a) > (b) ? (a) : (b))

/* ------------------------------------------------------------------------------------------
 * config.rs
 * ------------------------------------------------------------------------------------------ */

/* src/config.rs:97-107,143-151,180-191,218-222,257-264,287-297,352-358,411-417 */
void fo_config_default(fo_config *c) {
    memset(c, 0, sizeof(*c));
    c->block_size = 4096;
    c->multithread = 1;
    c->workers = 0;
    c->use_leftside = c->use_rightside = c->use_midside = 1;
    c->use_constant = c->use_fixed = c->use_lpc = 1;
    c->fixed_max_order = FO_MAX_FIXED_ORDER;
    c->fixed_order_sel = 1;
    c->approx_ent_partitions = 16;
    c->lpc_order = 10;
    c->quant_precision = 15;
    c->window_type = 1;
    c->tukey_alpha = 0.4f;
    c->prc_max_parameter = FO_MAX_RICE_PARAMETER;
    c->ext_lpc_order_search = 0;
    c->ext_lpc_precision_search = 0;
}

/* src/config.rs:109-130 (Encoder), :198-204 (SubFrameCoding: note fixed.verify() is NOT called),
 * :224-229 (Prc), :299-326 (Qlpc, non-experimental build), :371-387 (Window), :419-432 (OrderSel).
 * OrderSel lives inside Fixed, whose verify is never reached either, so only what the reference
 * actually checks is checked here. */
int fo_config_verify(const fo_config *c) {
    if (c->block_size < FO_MIN_BLOCK_SIZE || c->block_size > FO_MAX_BLOCK_SIZE) return 1;
    if (c->lpc_order < 1 || c->lpc_order > FO_MAX_LPC_ORDER) return 1;
    if (c->quant_precision < 1 || c->quant_precision > 15) return 1;
    /* use_direct_mse and mae_optimization_steps are accepted like a reference built with the `experimental`
     * feature (src/config.rs:305-321 checks them only in a build without it) */
    if (c->use_direct_mse != 0 && c->use_direct_mse != 1) return 1;
    if (c->mae_optimization_steps < 0) return 1;
    if (c->window_type == 1) {
        if (!(c->tukey_alpha >= 0.0f && c->tukey_alpha <= 1.0f)) return 1;
    } else if (c->window_type != 0) {
        return 1;
    }
    if (c->prc_max_parameter < 0 || c->prc_max_parameter > FO_MAX_RICE_PARAMETER) return 1;
    if (c->ext_lpc_order_search < 0 || c->ext_lpc_order_search > 8) return 1;
    if (c->ext_lpc_precision_search < 0 || c->ext_lpc_precision_search > 4) return 1;
    if (c->ext_lpc_order_search + c->ext_lpc_precision_search > 8) return 1;
    if ((c->ext_lpc_order_search > 0 || c->ext_lpc_precision_search > 0) && c->use_direct_mse) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * lpc.rs
 * ------------------------------------------------------------------------------------------ */

/* src/lpc.rs:96-120  window_weights (all arithmetic in f32, cosf from libm) */
void fo_window_weights(int window_type, float alpha, int len, float *out) {
    if (window_type == 0 || alpha == 0.0f) {
        for (int t = 0; t < len; t++) out[t] = 1.0f;
        return;
    }
    const float pi = 3.14159265358979323846f; /* std::f32::consts::PI */
    float max_t = (float)len - 1.0f;
    float alpha_len = alpha * max_t;
    for (int i = 0; i < len; i++) {
        float t = (float)i;
        float w;
        if (t < alpha_len / 2.0f) {
            w = 0.5f * (1.0f - cosf(2.0f * pi * t / alpha_len));
        } else if (t < max_t - alpha_len / 2.0f) {
            w = 1.0f;
        } else {
            w = 0.5f * (1.0f - cosf(2.0f * pi * (max_t - t) / alpha_len));
        }
        out[i] = w;
    }
}

/* src/lpc.rs:739-756  fill_windowed_signal: (f32)x * w */
void fo_fill_windowed_signal(const int32_t *signal, const float *window, int n, float *out) {
    for (int t = 0; t < n; t++) out[t] = (float)signal[t] * window[t];
}

/* src/lpc.rs:533-548  weighted_auto_correlation_nosimd with NoWeight, T = f64.
 * `order` is the number of lags (the caller passes lpc_order + 1); the sum starts at
 * t = order - 1 for EVERY lag; accumulation is a sequential f64 fma. */
void fo_auto_correlation_f64(int order, const float *signal, int n, double *dest) {
    for (int tau = 0; tau < order; tau++) dest[tau] = 0.0;
    for (int t = order - 1; t < n; t++) {
        double wy = (double)signal[t];
        for (int tau = 0; tau < order && tau < FO_MAX_LPC_ORDER + 1; tau++) {
            dest[tau] = fma((double)signal[t - tau], wy, dest[tau]);
        }
    }
}

/* same, T = f32 (only the reference's tests use it, src/lpc.rs:998-1022,1146-1169) */
void fo_auto_correlation_f32(int order, const float *signal, int n, float *dest) {
    for (int tau = 0; tau < order; tau++) dest[tau] = 0.0f;
    for (int t = order - 1; t < n; t++) {
        float wy = signal[t];
        for (int tau = 0; tau < order && tau < FO_MAX_LPC_ORDER + 1; tau++) {
            dest[tau] = fmaf(signal[t - tau], wy, dest[tau]);
        }
    }
}

/* src/lpc.rs:633-705  symmetric_levinson_recursion.  The `continue` on denom == 0 skips step n
 * (the diagonal-loading value is never used again: the loop is entered once). */
#define FO_DEFINE_LEVINSON(NAME, T, FMA, ZERO, ONE)                                          \
    void NAME(const T *coefs, const T *ys, int order, T *dest) {                              \
        for (int i = 0; i < order; i++) dest[i] = ZERO;                                       \
        if (order <= 0) return;                                                               \
        assert(coefs[0] >= ZERO);                                                             \
        if (coefs[0] == ZERO) return;                                                         \
        T forward[FO_MAX_LPC_ORDER + 1];                                                      \
        T forward_next[FO_MAX_LPC_ORDER + 1];                                                 \
        for (int i = 0; i <= FO_MAX_LPC_ORDER; i++) forward[i] = forward_next[i] = ZERO;      \
        T diagonal_loading = ZERO;                                                            \
        forward[0] = ONE / (coefs[0] + diagonal_loading);                                     \
        dest[0] = ys[0] / (coefs[0] + diagonal_loading);                                      \
        for (int n = 1; n < order; n++) {                                                     \
            T error = ZERO;                                                                   \
            for (int d = 0; d < n; d++) error = FMA(coefs[n - d], forward[d], error);         \
            T denom = FMA(error, -error, ONE);                                                \
            if (denom == ZERO) {                                                              \
                diagonal_loading = FO_MAX(ONE, diagonal_loading + diagonal_loading);          \
                continue;                                                                     \
            }                                                                                 \
            T alpha = ONE / denom;                                                            \
            T beta = -alpha * error;                                                          \
            for (int d = 0; d <= n; d++)                                                      \
                forward_next[d] = FMA(alpha, forward[d], beta * forward[n - d]);              \
            for (int d = 0; d <= n; d++) forward[d] = forward_next[d];                        \
            T delta = ZERO;                                                                   \
            for (int d = 0; d < n; d++) delta = FMA(coefs[n - d], dest[d], delta);            \
            for (int d = 0; d <= n; d++) dest[d] = FMA(ys[n] - delta, forward[n - d], dest[d]); \
        }                                                                                     \
        (void)diagonal_loading;                                                               \
    }
FO_DEFINE_LEVINSON(fo_levinson_f64, double, fma, 0.0, 1.0)
FO_DEFINE_LEVINSON(fo_levinson_f32, float, fmaf, 0.0f, 1.0f)

/* src/lpc.rs:234-255  find_shift */
int fo_find_shift(const double *coefs, int n, int precision) {
    assert(precision <= 15 && n > 0);
    double max_abs = fabs(coefs[0]);
    for (int i = 1; i < n; i++) max_abs = fmax(max_abs, fabs(coefs[i]));
    double l = ceil(log2(max_abs));
    double lo = (double)(INT16_MIN + 16);
    if (!(l > lo)) l = lo; /* Float::max(l, lo) */
    int abs_log2 = (l >= 32767.0) ? 32767 : (int)l; /* `as_()` saturating cast to i16 */
    int shift = (precision - 1) - abs_log2;
    if (shift < 0) shift = 0;   /* QLPC_MIN_SHIFT, src/constant.rs:131 */
    if (shift > 15) shift = 15; /* QLPC_MAX_SHIFT, src/constant.rs:124 */
    return shift;
}

void fo_find_shift_each(const double *values, uint64_t count, int precision, int32_t *shifts) {
    for (uint64_t i = 0; i < count; i++) shifts[i] = fo_find_shift(values + i, 1, precision);
}

struct fo_log2f_job { uint32_t first; uint64_t a, b; uint32_t *out; };
static void *fo_log2f_worker(void *arg) {
    struct fo_log2f_job *j = (struct fo_log2f_job *)arg;
    for (uint64_t i = j->a; i < j->b; i++) {
        uint32_t u = j->first + (uint32_t)i;
        float x, y;
        memcpy(&x, &u, 4);
        y = log2f(x); /* the libm call behind Rust's f32::log2 on linux-gnu (src/coding.rs:219-221) */
        memcpy(&j->out[i], &y, 4);
    }
    return NULL;
}
void fo_log2f_bits(uint32_t first, uint64_t count, int threads, uint32_t *out) {
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    pthread_t th[64];
    struct fo_log2f_job jobs[64];
    for (int w = 0; w < threads; w++) {
        jobs[w].first = first;
        jobs[w].a = count * (uint64_t)w / (uint64_t)threads;
        jobs[w].b = count * (uint64_t)(w + 1) / (uint64_t)threads;
        jobs[w].out = out;
        pthread_create(&th[w], NULL, fo_log2f_worker, &jobs[w]);
    }
    for (int w = 0; w < threads; w++) pthread_join(th[w], NULL);
}

/* src/lpc.rs:258-271  quantize_parameter */
static int16_t fo_quantize_parameter(double p, int shift) {
    double scalefac = ldexp(1.0, shift); /* powi(2, shift) */
    double scaled = round(p * scalefac); /* half away from zero */
    if (scaled < -32768.0) scaled = -32768.0;
    if (scaled > 32767.0) scaled = 32767.0;
    return (int16_t)scaled;
}

/* src/lpc.rs:273-302  quantize_parameters */
int fo_quantize_parameters(const double *coefs, int n, int precision, int16_t *q_out, int *shift_out) {
    for (int i = 0; i < FO_MAX_LPC_ORDER; i++) q_out[i] = 0;
    if (n == 0) {
        *shift_out = 0;
        return 0;
    }
    int shift = fo_find_shift(coefs, n, precision);
    int lo = -(1 << (precision - 1));
    int hi = (1 << (precision - 1)) - 1;
    for (int i = 0; i < n; i++) {
        int q = fo_quantize_parameter(coefs[i], shift);
        q = FO_MIN(FO_MAX(q, lo), hi);
        q_out[i] = (int16_t)q;
    }
    int order = FO_MAX_LPC_ORDER;
    while (order > 0 && q_out[order - 1] == 0) order--;
    if (order < 1) order = 1;
    *shift_out = shift;
    return order;
}

/* src/lpc.rs:306-390  compute_error (+ compute_error_impl for i32 and i64).
 * Integer adds/subs wrap like release-mode Rust / fakesimd (src/fakesimd.rs:46-110). */
void fo_compute_error(const int16_t *q, int order, int shift, const int32_t *signal, int n, int32_t *errors) {
    uint64_t maxabs_signal = 0;
    for (int t = 0; t < n; t++) {
        uint32_t a = signal[t] < 0 ? (uint32_t)0 - (uint32_t)signal[t] : (uint32_t)signal[t];
        if (a > maxabs_signal) maxabs_signal = a;
    }
    int64_t sumabs = 0;
    for (int j = 0; j < order; j++) sumabs += q[j] < 0 ? -(int64_t)q[j] : (int64_t)q[j];
    uint64_t maxabs = maxabs_signal * (uint64_t)sumabs;
    if (maxabs < (uint64_t)INT32_MAX) {
        /* i32 path: acc[t] = sum_j q[j]*x[t-1-j] (no overflow by the guard), e = x - (acc >> shift) */
        for (int t = 0; t < n; t++) errors[t] = 0;
        for (int j = 0; j < order; j++) {
            int32_t w = q[j];
            for (int t = 0; t + j + 1 < n; t++) {
                errors[t + j + 1] = (int32_t)((uint32_t)errors[t + j + 1] + (uint32_t)(w * signal[t]));
            }
        }
        for (int t = 0; t < n; t++) {
            errors[t] = (int32_t)((uint32_t)signal[t] - (uint32_t)(errors[t] >> shift));
        }
    } else {
        int64_t *acc = (int64_t *)calloc((size_t)n, sizeof(int64_t));
        for (int j = 0; j < order; j++) {
            int64_t w = q[j];
            for (int t = 0; t + j + 1 < n; t++) acc[t + j + 1] += w * (int64_t)signal[t];
        }
        for (int t = 0; t < n; t++) {
            int64_t v = (int64_t)signal[t] - (acc[t] >> shift);
            errors[t] = (int32_t)(uint32_t)(uint64_t)v; /* `v as i32` truncation */
        }
        free(acc);
    }
    for (int t = 0; t < order && t < n; t++) errors[t] = 0;
}

/* src/lpc.rs:760-801 weighted_lpc_from_auto_corr + :920-930 lpc_from_autocorr */
void fo_lpc_from_autocorr(const int32_t *signal, int n, int window_type, float alpha, int lpc_order,
                          double *coefs_out, double *corr_out) {
    for (int i = 0; i < lpc_order; i++) coefs_out[i] = 0.0;
    if (lpc_order == 0) return;
    float *window = (float *)malloc(sizeof(float) * (size_t)n);
    float *windowed = (float *)malloc(sizeof(float) * (size_t)n);
    double corr[FO_MAX_LPC_ORDER + 1];
    fo_window_weights(window_type, alpha, n, window);
    fo_fill_windowed_signal(signal, window, n, windowed);
    fo_auto_correlation_f64(lpc_order + 1, windowed, n, corr);
    for (int i = 0; i <= lpc_order; i++) assert(isfinite(corr[i]));
    fo_levinson_f64(corr, corr + 1, lpc_order, coefs_out);
    for (int i = 0; i < lpc_order; i++) assert(isfinite(coefs_out[i]));
    if (corr_out) memcpy(corr_out, corr, sizeof(double) * (size_t)(lpc_order + 1));
    free(window);
    free(windowed);
}

/* ---- `experimental` feature: direct-MSE (covariance method) LPC estimator -------------------
 * src/lpc.rs:573-600 weighted_lagged_outer_prod_sum with NoWeight: for t = order-1 .. len-1 of `signal`
 *   dest[i][j] = fma(signal[t-i], signal[t-j], dest[i][j]), j >= i, sequential in t, then mirrored. */
void fo_lagged_outer_prod_sum(int order, const float *signal, int len, double *dest /* order x order, row major */) {
    for (int i = 0; i < order * order; i++) dest[i] = 0.0;
    for (int t = order - 1; t < len; t++)
        for (int i = 0; i < order; i++)
            for (int j = i; j < order; j++)
                dest[i * order + j] = fma((double)signal[t - i], (double)signal[t - j], dest[i * order + j]);
    for (int i = 0; i < order; i++)
        for (int j = i + 1; j < order; j++) dest[j * order + i] = dest[i * order + j];
}

/* LpcFloat::solve_sym_mut (src/lpc.rs:76-87): `mat.clone().cholesky()` then `solve_mut(v)` of nalgebra 0.32.6
 * (flacenc-bin/Cargo.lock).  nalgebra is NOT part of /root/reference, so this restates its published algorithm and
 * PARITY IS UNPINNED for the last bits of this function:
 *   Cholesky::new      column by column; for k < j: col_j[j..] += (-L[j][k]) * col_k[j..] as separate multiply and add
 *                      (axpy: y = a*x + 1*y); the diagonal must be non-zero with a real square root, else not SPD;
 *                      L[j][j] = sqrt(d), col_j[j+1..] /= L[j][j]
 *   solve_mut          forward substitution, column oriented: b[i] /= L[i][i]; b[i+1..] += (-b[i]) * L[i+1.., i]
 *                      then back substitution with the transpose: b[i] = (b[i] - dot(L[i+1.., i], b[i+1..])) / L[i][i],
 *                      the dot product with nalgebra's eight interleaved accumulators (base/blas.rs dotx)
 * Returns 0 when the matrix is not positive definite. */
static double fo_na_dot(const double *a, int sa, const double *b, int n) {
    double res = 0.0, acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i = 0;
    while (n - i >= 8) {
        for (int k = 0; k < 8; k++) acc[k] += a[(i + k) * sa] * b[i + k];
        i += 8;
    }
    res += acc[0] + acc[4];
    res += acc[1] + acc[5];
    res += acc[2] + acc[6];
    res += acc[3] + acc[7];
    for (; i < n; i++) res += a[i * sa] * b[i];
    return res;
}

int fo_solve_sym(const double *mat, int n, double *v) {
    double L[FO_MAX_LPC_ORDER * FO_MAX_LPC_ORDER]; /* row major; only the lower triangle is used */
    for (int i = 0; i < n * n; i++) L[i] = mat[i];
    for (int j = 0; j < n; j++) {
        for (int k = 0; k < j; k++) {
            const double factor = -L[j * n + k];
            for (int r = j; r < n; r++) {
                const double prod = factor * L[r * n + k];
                L[r * n + j] = prod + L[r * n + j];
            }
        }
        const double diag = L[j * n + j];
        if (diag == 0.0 || !(diag >= 0.0)) return 0;
        const double denom = sqrt(diag);
        L[j * n + j] = denom;
        for (int r = j + 1; r < n; r++) L[r * n + j] = L[r * n + j] / denom;
    }
    for (int i = 0; i < n; i++) {
        const double diag = L[i * n + i];
        if (diag == 0.0) return 0;
        const double coeff = v[i] / diag;
        v[i] = coeff;
        for (int r = i + 1; r < n; r++) {
            const double prod = -coeff * L[r * n + i];
            v[r] = prod + v[r];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        const double dot = fo_na_dot(&L[(i + 1) * n + i], n, &v[i + 1], n - 1 - i);
        const double diag = L[i * n + i];
        if (diag == 0.0) return 0;
        v[i] = (v[i] - dot) / diag;
    }
    return 1;
}

/* src/lpc.rs:533-548 weighted_auto_correlation_nosimd with VecWeight (:194-199): the weight multiplies the
 * newest sample in f32 before the product is widened (`weight == NULL` is NoWeight: the sample itself) */
static void fo_weighted_auto_correlation(int order, const float *signal, int len, const float *weight, double *dest) {
    for (int i = 0; i < order; i++) dest[i] = 0.0;
    for (int t = order - 1; t < len; t++) {
        const float wyf = weight ? weight[t] * signal[t] : signal[t];
        const double wy = (double)wyf;
        for (int tau = 0; tau < order; tau++) dest[tau] = fma((double)signal[t - tau], wy, dest[tau]);
    }
}

/* src/lpc.rs:573-600 weighted_lagged_outer_prod_sum under ShiftedWeight<1, VecWeight> (:206-211): the weight of
 * time t + 1 multiplies signal[t - j] in f32 (`weight == NULL` is NoWeight) */
static void fo_weighted_lagged_outer_prod_sum(int order, const float *signal, int len, const float *weight, double *dest) {
    for (int i = 0; i < order * order; i++) dest[i] = 0.0;
    for (int t = order - 1; t < len; t++)
        for (int i = 0; i < order; i++)
            for (int j = i; j < order; j++) {
                const float wxf = weight ? weight[t + 1] * signal[t - j] : signal[t - j];
                dest[i * order + j] = fma((double)signal[t - i], (double)wxf, dest[i * order + j]);
            }
    for (int i = 0; i < order; i++)
        for (int j = i + 1; j < order; j++) dest[j * order + i] = dest[i * order + j];
}

/* src/lpc.rs:852-903 weighted_lpc_with_direct_mse on an already windowed signal */
static void fo_weighted_direct_mse(const float *windowed, int n, int lpc_order, const float *weight, double *coefs_out,
                                   double *corr_out, double *covar_out) {
    double corr[FO_MAX_LPC_ORDER + 1];
    double covar[FO_MAX_LPC_ORDER * FO_MAX_LPC_ORDER];
    fo_weighted_auto_correlation(lpc_order + 1, windowed, n, weight, corr);
    /* the statistics of the signal without its last sample, weights shifted by one */
    fo_weighted_lagged_outer_prod_sum(lpc_order, windowed, n - 1, weight, covar);
    if (corr_out) memcpy(corr_out, corr, sizeof(double) * (size_t)(lpc_order + 1));
    if (covar_out) memcpy(covar_out, covar, sizeof(double) * (size_t)(lpc_order * lpc_order));
    double xy[FO_MAX_LPC_ORDER];
    double regularizer = 0.0;
    for (;;) {
        for (int i = 0; i < lpc_order; i++) xy[i] = corr[1 + i];
        if (fo_solve_sym(covar, lpc_order, xy)) break;
        /* src/lpc.rs:889-896: the diagonal grows by 1, 1, 2, 4, ... until the factorisation succeeds.  (In the
         * reference a failed attempt leaves `xy` untouched: Cholesky::new fails before solve_mut runs.) */
        const double old = regularizer;
        regularizer = fmax(1.0, regularizer + regularizer);
        for (int i = 0; i < lpc_order; i++) covar[i * lpc_order + i] += regularizer - old;
    }
    for (int i = 0; i < lpc_order; i++) coefs_out[i] = xy[i];
}

/* src/lpc.rs:905-913 lpc_with_direct_mse (NoWeight) */
void fo_lpc_with_direct_mse(const int32_t *signal, int n, int window_type, float alpha, int lpc_order,
                            double *coefs_out, double *corr_out, double *covar_out) {
    for (int i = 0; i < lpc_order; i++) coefs_out[i] = 0.0;
    if (lpc_order == 0) return;
    float *window = (float *)malloc(sizeof(float) * (size_t)n);
    float *windowed = (float *)malloc(sizeof(float) * (size_t)n);
    fo_window_weights(window_type, alpha, n, window);
    fo_fill_windowed_signal(signal, window, n, windowed);
    fo_weighted_direct_mse(windowed, n, lpc_order, NULL, coefs_out, corr_out, covar_out);
    free(window);
    free(windowed);
}

/* src/lpc.rs:606-618 compute_raw_errors: "prediction - signal" in f32 with the coefficients rounded to f32, one fused
 * multiply-add per tap, taps in ascending order; errors[0..order) are left alone */
void fo_compute_raw_errors(const int32_t *signal, int n, const double *coefs, int lpc_order, float *errors) {
    for (int t = lpc_order; t < n; t++) {
        float e = (float)(-signal[t]);
        for (int j = 0; j < lpc_order; j++) e = fmaf((float)coefs[j], (float)signal[t - 1 - j], e);
        errors[t] = e;
    }
}

/* the weight of a raw error in src/lpc.rs:828: (|err|.max(1) / normalizer).max(0.01).powf(-1.2), all in f32; powf is
 * the libm the reference links (glibc here, like log2f in the entropy estimate) */
float fo_irls_weight(float err, float normalizer) {
    const float a = fmaxf(fabsf(err), 1.0f);
    return powf(fmaxf(a / normalizer, 0.01f), -1.2f);
}

/* bits of fo_irls_weight over the raw errors with float bit patterns [first, first + count), for the sweep tests */
struct fo_irlsw_job { uint32_t first; uint64_t a, b; float normalizer; uint32_t *out; };
static void *fo_irlsw_worker(void *arg) {
    struct fo_irlsw_job *j = (struct fo_irlsw_job *)arg;
    for (uint64_t i = j->a; i < j->b; i++) {
        uint32_t u = j->first + (uint32_t)i;
        float x, y;
        memcpy(&x, &u, 4);
        y = fo_irls_weight(x, j->normalizer);
        memcpy(&j->out[i], &y, 4);
    }
    return NULL;
}
void fo_irls_weight_bits(uint32_t first, uint64_t count, float normalizer, int threads, uint32_t *out) {
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    pthread_t th[64];
    struct fo_irlsw_job jobs[64];
    for (int w = 0; w < threads; w++) {
        jobs[w].first = first;
        jobs[w].a = count * (uint64_t)w / (uint64_t)threads;
        jobs[w].b = count * (uint64_t)(w + 1) / (uint64_t)threads;
        jobs[w].normalizer = normalizer;
        jobs[w].out = out;
        pthread_create(&th[w], NULL, fo_irlsw_worker, &jobs[w]);
    }
    for (int w = 0; w < threads; w++) pthread_join(th[w], NULL);
}

/* src/lpc.rs:814-850 lpc_with_irls_mae: `steps` + 1 weighted direct-MSE solutions, each re-weighted by the raw errors
 * of the one before; the solution with the smallest sequential f32 sum of |raw error| wins, the earliest on ties.
 * `sums_out` (may be NULL) receives the steps + 1 sums. */
void fo_lpc_with_irls_mae(const int32_t *signal, int n, int window_type, float alpha, int lpc_order, int steps,
                          double *coefs_out, float *sums_out) {
    for (int i = 0; i < lpc_order; i++) coefs_out[i] = 0.0;
    float *window = (float *)malloc(sizeof(float) * (size_t)n);
    float *windowed = (float *)malloc(sizeof(float) * (size_t)n);
    float *weights = (float *)malloc(sizeof(float) * (size_t)n);
    float *raw_errors = (float *)calloc((size_t)n, sizeof(float));
    fo_window_weights(window_type, alpha, n, window);
    fo_fill_windowed_signal(signal, window, n, windowed);
    for (int t = 0; t < n; t++) weights[t] = 1.0f;
    int32_t peak = 0;
    for (int t = 0; t < n; t++) {
        const int32_t a = signal[t] < 0 ? -signal[t] : signal[t];
        if (a > peak) peak = a;
    }
    const float normalizer = (float)peak;
    float best_error = FLT_MAX;
    int have_best = 0;
    double coefs[FO_MAX_LPC_ORDER];
    for (int step = 0; step <= steps; step++) {
        fo_weighted_direct_mse(windowed, n, lpc_order, weights, coefs, NULL, NULL);
        fo_compute_raw_errors(signal, n, coefs, lpc_order, raw_errors);
        float sum_abs_err = 0.0f;
        for (int t = 0; t < n; t++) sum_abs_err += fabsf(raw_errors[t]);
        if (sums_out) sums_out[step] = sum_abs_err;
        if (sum_abs_err < best_error) {
            best_error = sum_abs_err;
            have_best = 1;
            memcpy(coefs_out, coefs, sizeof(double) * (size_t)lpc_order);
        }
        for (int t = lpc_order; t < n; t++) weights[t] = fo_irls_weight(raw_errors[t], normalizer);
    }
    assert(have_best); /* the reference unwraps an Option here */
    (void)have_best;
    free(window);
    free(windowed);
    free(weights);
    free(raw_errors);
}

/* ------------------------------------------------------------------------------------------
 * rice.rs
 * ------------------------------------------------------------------------------------------ */

/* src/rice.rs:169-171 */
uint32_t fo_encode_signbit(int32_t v) {
    uint32_t a = v < 0 ? (uint32_t)0 - (uint32_t)v : (uint32_t)v;
    return (a << 1) - (uint32_t)(v < 0);
}

/* src/rice.rs:179-187 */
int32_t fo_decode_signbit(uint32_t v) {
    if (v & 1u) return -(int32_t)((v >> 1) + 1u);
    return (int32_t)(v >> 1);
}

/* src/rice.rs:157-165 */
int fo_finest_partition_order(int size, int min_part_size) {
    assert(min_part_size >= 1);
    uint32_t max_splits = (uint32_t)(size / min_part_size);
    int lz = max_splits ? __builtin_clz(max_splits) : 32;
    int max_order_for_min_part = 32 - lz - 1; /* the reference underflows (panics) on 0; callers never pass it */
    int tz = size ? __builtin_ctz((unsigned)size) : 32;
    int r = FO_MIN(max_order_for_min_part, tz);
    return FO_MIN(15, r);
}

/* src/rice.rs:65-105  PrcBitTable::from_errors: 32 lanes (p = 0..31), u32 wrapping adds,
 * clamp to 2^27-1 after every 16-sample chunk and after adding the offset. */
void fo_bit_table_from_errors(const uint32_t *errors, int n, uint32_t offset, uint32_t *table) {
    for (int p = 0; p < 32; p++) table[p] = 0;
    for (int c = 0; c < n; c += 16) {
        int len = FO_MIN(16, n - c);
        for (int k = 0; k < len; k++) {
            uint32_t v = errors[c + k];
            for (int p = 0; p < 32; p++) table[p] += v >> p;
        }
        for (int p = 0; p < 32; p++) table[p] = FO_MIN(table[p], FO_RICE_SAT);
    }
    for (int p = 0; p < 32; p++) {
        table[p] += offset + (uint32_t)n * (uint32_t)(p + 1);
        table[p] = FO_MIN(table[p], FO_RICE_SAT);
    }
}

/* src/rice.rs:117-141  minimizer: min over p <= max_p of (bits << 5 | p) */
void fo_bit_table_minimizer(const uint32_t *table, int max_p, int *p_out, uint32_t *bits_out) {
    uint32_t best = 0xFFFFFFFFu;
    for (int p = 0; p < 32; p++) {
        uint32_t v = (p <= max_p) ? table[p] : 0xFFFFFFFFu;
        uint32_t packed = (v << 5) | (uint32_t)p;
        if (packed < best) best = packed;
    }
    *p_out = (int)(best & 0x1F);
    *bits_out = best >> 5;
}

/* src/rice.rs:144-153 */
void fo_bit_table_merge(const uint32_t *a, const uint32_t *b, uint32_t offset, uint32_t *out) {
    for (int p = 0; p < 32; p++) {
        uint32_t v = a[p] + b[p] - offset;
        out[p] = FO_MIN(v, FO_RICE_SAT);
    }
}

/* src/rice.rs:246-298  PrcParameterFinder::find */
int fo_find_partitioned_rice_parameter(const int32_t *signal, int n, int warmup, int max_p,
                                       uint8_t *ps_out, uint64_t *code_bits_out) {
    int partition_order = fo_finest_partition_order(n, FO_MAX(64, warmup));
    int nparts = 1 << partition_order;
    uint32_t *errors = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)FO_MAX(n, 1));
    uint32_t(*tables)[32] = (uint32_t(*)[32])malloc(sizeof(uint32_t[32]) * (size_t)nparts);
    int *ps = (int *)malloc(sizeof(int) * (size_t)nparts);
    int *min_ps = (int *)malloc(sizeof(int) * (size_t)nparts);
    for (int t = 0; t < n; t++) errors[t] = fo_encode_signbit(signal[t]);

    int part_size = n / nparts;
    for (int p = 0; p < nparts; p++) {
        int start = FO_MAX(p * part_size, warmup);
        int end = (p + 1) * part_size;
        fo_bit_table_from_errors(errors + start, end - start, 4, tables[p]);
    }
    uint64_t min_bits = 0;
    for (int p = 0; p < nparts; p++) {
        uint32_t b;
        fo_bit_table_minimizer(tables[p], max_p, &min_ps[p], &b);
        min_bits += b;
    }
    int min_order = partition_order;
    while (nparts > 1) {
        int merged = nparts / 2;
        for (int i = 0; i < merged; i++) {
            uint32_t tmp[32];
            fo_bit_table_merge(tables[2 * i], tables[2 * i + 1], 4, tmp);
            memcpy(tables[i], tmp, sizeof(tmp));
        }
        nparts = merged;
        partition_order -= 1;
        uint64_t next_bits = 0;
        for (int p = 0; p < nparts; p++) {
            uint32_t b;
            fo_bit_table_minimizer(tables[p], max_p, &ps[p], &b);
            next_bits += b;
        }
        if (next_bits < min_bits) {
            min_bits = next_bits;
            memcpy(min_ps, ps, sizeof(int) * (size_t)nparts);
            min_order = partition_order;
        }
    }
    for (int p = 0; p < (1 << min_order); p++) ps_out[p] = (uint8_t)min_ps[p];
    *code_bits_out = min_bits;
    free(errors);
    free(tables);
    free(ps);
    free(min_ps);
    return min_order;
}

/* ------------------------------------------------------------------------------------------
 * coding.rs
 * ------------------------------------------------------------------------------------------ */

/* src/coding.rs:182-197  reset_fixed_lpc_errors: e[k+1][t] = e[k][t] - e[k][t-1], e[k][-1] = 0,
 * wrapping i32; errors5 is [5][n]. */
void fo_fixed_lpc_errors(const int32_t *signal, int n, int32_t *errors5) {
    memcpy(errors5, signal, sizeof(int32_t) * (size_t)n);
    for (int order = 0; order < FO_MAX_FIXED_ORDER; order++) {
        const int32_t *cur = errors5 + (size_t)order * n;
        int32_t *next = errors5 + (size_t)(order + 1) * n;
        int32_t carry = 0;
        for (int t = 0; t < n; t++) {
            next[t] = (int32_t)((uint32_t)cur[t] - (uint32_t)carry);
            carry = cur[t];
        }
    }
}

/* src/arrayutils.rs:496-506  find_sum_abs_f32 (stable: everything is "head", sequential f32) */
static float fo_sum_abs_f32(const int32_t *data, int n) {
    float acc = 0.0f;
    for (int i = 0; i < n; i++) {
        int32_t x = data[i];
        int32_t a = x < 0 ? (int32_t)((uint32_t)0 - (uint32_t)x) : x; /* i32::abs, wrapping */
        acc = (float)a + acc;
    }
    return acc;
}

/* Rust `f32 as usize`: saturating, NaN -> 0 */
static uint64_t fo_f32_as_usize(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)v;
}

/* src/coding.rs:200-227  estimate_entropy */
uint64_t fo_estimate_entropy(const int32_t *errors, int n, int warmup, int partitions) {
    int partition_size = (n + partitions - 1) / partitions;
    int offset = 0;
    uint64_t acc = 0;
    for (int p = 0; p < partitions; p++) {
        int end = FO_MIN(n, offset + partition_size);
        int partition_len = end - offset;
        if (end >= warmup) {
            int sample_count = FO_MIN(end - warmup, partition_len);
            float sum_errors = fo_sum_abs_f32(errors + offset, partition_len);
            float avg_errors = sum_errors * 2.0f / ((float)sample_count + 0.00001f);
            float geom_p = 1.0f / (avg_errors + 1.0f);
            float xent = fmaf(avg_errors, -log2f(1.0f - geom_p), -log2f(geom_p));
            acc += fo_f32_as_usize(xent * (float)sample_count);
        }
        offset = end;
    }
    return acc;
}

/* src/coding.rs:230-288  select_order_and_encode_residual (selection part): first minimum wins
 * (Iterator::min_by_key), accepted only if bits < baseline_bits. errors is [n_orders][n]. */
int fo_select_order(int order_sel, int partitions, int max_p, const int32_t *errors, int n_orders, int n,
                    int bps, uint64_t baseline_bits, uint64_t *bits_out) {
    int best = -1;
    uint64_t best_bits = 0;
    for (int order = 0; order < n_orders; order++) {
        const int32_t *err = errors + (size_t)order * n;
        uint64_t bits;
        if (order_sel == 0) {
            uint8_t ps[FO_MAX_RICE_PARTS];
            uint64_t code_bits;
            fo_find_partitioned_rice_parameter(err, n, order, max_p, ps, &code_bits);
            bits = (uint64_t)bps * order + code_bits;
        } else {
            bits = fo_estimate_entropy(err, n, order, partitions) + (uint64_t)bps * order;
        }
        if (best < 0 || bits < best_bits) {
            best = order;
            best_bits = bits;
        }
    }
    if (bits_out) *bits_out = best_bits;
    if (best >= 0 && best_bits < baseline_bits) return best;
    return -1;
}

/* src/coding.rs:140-176 encode_residual(_with_prc_parameter) + src/component/datatype.rs:2313-2344
 * (Residual::from_parts sums) + src/component/bitrepr.rs:532-544 (Residual::count_bits).
 * Fills the residual-related fields of `sf` and returns Residual::count_bits(). */
static uint64_t fo_encode_residual(int max_p, const int32_t *errors, int n, int warmup, fo_subframe *sf) {
    uint64_t code_bits;
    int order = fo_find_partitioned_rice_parameter(errors, n, warmup, max_p, sf->rice_params, &code_bits);
    int nparts = 1 << order;
    int part_size = n >> order;
    sf->part_order = order;
    sf->code_bits = code_bits;
    sf->residual = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    memcpy(sf->residual, errors, sizeof(int32_t) * (size_t)n);
    for (int t = 0; t < warmup && t < n; t++) sf->residual[t] = 0;

    uint64_t sum_q = 0, sum_p = 0;
    int offset = 0, rice2 = 0;
    for (int p = 0; p < nparts; p++) {
        int rice_p = sf->rice_params[p];
        int start = FO_MAX(offset, warmup);
        offset += part_size;
        for (int t = start; t < offset; t++) sum_q += fo_encode_signbit(errors[t]) >> rice_p;
        sum_p += (uint64_t)rice_p;
        if (rice_p > 14) rice2 = 1;
    }
    sf->sum_quotients = sum_q;
    sf->rice2 = rice2;
    uint64_t quotient_bits = sum_q + (uint64_t)n - (uint64_t)warmup;
    uint64_t remainder_bits = sum_p * (uint64_t)part_size - (uint64_t)warmup * sf->rice_params[0];
    return 2 + 4 + (uint64_t)nparts * (rice2 ? 5 : 4) + quotient_bits + remainder_bits;
}

/* src/arrayutils.rs:382-389 */
static int fo_is_constant(const int32_t *s, int n) {
    for (int t = 1; t < n; t++)
        if (s[0] != s[t]) return 0;
    return 1;
}

static void fo_subframe_reset(fo_subframe *sf, int type, const int32_t *samples, int n, int bps) {
    memset(sf, 0, sizeof(*sf));
    sf->type = type;
    sf->samples = samples;
    sf->n = n;
    sf->bps = bps;
}

/* src/coding.rs:298-331  fixed_lpc; returns 1 if Some */
static int fo_fixed_lpc(const fo_config *cfg, const int32_t *signal, int n, int bps, uint64_t baseline_bits,
                        fo_subframe *out) {
    assert(bps < 30);
    int32_t *errors = (int32_t *)malloc(sizeof(int32_t) * 5 * (size_t)n);
    fo_fixed_lpc_errors(signal, n, errors);
    int n_orders = FO_MIN(cfg->fixed_max_order + 1, FO_MAX_FIXED_ORDER + 1); /* .take(max_order + 1) of 5 */
    int order = fo_select_order(cfg->fixed_order_sel, cfg->approx_ent_partitions, cfg->prc_max_parameter,
                                errors, n_orders, n, bps, baseline_bits, NULL);
    if (order < 0) {
        free(errors);
        return 0;
    }
    fo_subframe_reset(out, FO_SF_FIXED, signal, n, bps);
    out->order = order;
    uint64_t rbits = fo_encode_residual(cfg->prc_max_parameter, errors + (size_t)order * n, n, order, out);
    out->bits = 8 + (uint64_t)bps * order + rbits; /* src/component/bitrepr.rs:475-477 */
    free(errors);
    return 1;
}

/* EXTENSION (not in the reference; config.ext_lpc_order_search = k > 0): the LPC orders tried besides P = lpc_order are
 * P - i * ceil(P / (k + 1)), i = 1..k, as long as they are >= 1.  Returns the number of orders written (P first). */
int fo_ext_lpc_orders(int lpc_order, int k, int *orders) {
    int n = 0;
    orders[n++] = lpc_order;
    if (k <= 0) return n;
    const int step = (lpc_order + k) / (k + 1);
    for (int i = 1; i <= k; i++) {
        const int o = lpc_order - i * step;
        if (o < 1) break;
        orders[n++] = o;
    }
    return n;
}

/* the LPC subframe of one set of unquantised coefficients (the second half of estimated_qlpc) */
static void fo_qlpc_subframe(const fo_config *cfg, const double *coefs, int lpc_order, int precision, const int32_t *signal,
                             int n, int bps, fo_subframe *out) {
    fo_subframe_reset(out, FO_SF_LPC, signal, n, bps);
    int16_t q[FO_MAX_LPC_ORDER];
    int shift;
    int order = fo_quantize_parameters(coefs, lpc_order, precision, q, &shift);
    for (int i = 0; i < FO_MAX_LPC_ORDER; i++) out->qlp[i] = i < order ? q[i] : 0;
    out->order = order;
    out->shift = shift;
    out->precision = precision;
    int32_t *errors = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    fo_compute_error(q, order, shift, signal, n, errors);
    uint64_t rbits = fo_encode_residual(cfg->prc_max_parameter, errors, n, order, out);
    /* src/component/bitrepr.rs:492-499 */
    out->bits = 8 + (uint64_t)bps * order + 4 + 5 + (uint64_t)precision * order + rbits;
    free(errors);
}

/* src/coding.rs:360-381  estimated_qlpc */
static void fo_estimated_qlpc(const fo_config *cfg, const int32_t *signal, int n, int bps, fo_subframe *out) {
    int lpc_order = cfg->lpc_order;
    double coefs[FO_MAX_LPC_ORDER];
    double corr[FO_MAX_LPC_ORDER + 1];
    /* src/coding.rs:333-351 perform_qlpc: the estimator the configuration names */
    if (cfg->use_direct_mse && cfg->mae_optimization_steps > 0)
        fo_lpc_with_irls_mae(signal, n, cfg->window_type, cfg->tukey_alpha, lpc_order, cfg->mae_optimization_steps, coefs, NULL);
    else if (cfg->use_direct_mse) fo_lpc_with_direct_mse(signal, n, cfg->window_type, cfg->tukey_alpha, lpc_order, coefs, NULL, NULL);
    else fo_lpc_from_autocorr(signal, n, cfg->window_type, cfg->tukey_alpha, lpc_order, coefs, corr);
    fo_qlpc_subframe(cfg, coefs, lpc_order, cfg->quant_precision, signal, n, bps, out);
    if (cfg->ext_lpc_order_search > 0 && !cfg->use_direct_mse) {
        /* EXTENSION: the Levinson solutions of lower orders on the same autocorrelation; fewest bits wins, the higher
         * order (the earlier candidate) on ties */
        int orders[16];
        const int no = fo_ext_lpc_orders(lpc_order, cfg->ext_lpc_order_search, orders);
        for (int k = 1; k < no; k++) {
            double ck[FO_MAX_LPC_ORDER];
            fo_levinson_f64(corr, corr + 1, orders[k], ck);
            fo_subframe cand;
            fo_qlpc_subframe(cfg, ck, orders[k], cfg->quant_precision, signal, n, bps, &cand);
            if (cand.bits < out->bits) {
                free(out->residual);
                *out = cand;
            } else {
                free(cand.residual);
            }
        }
    }
    if (cfg->ext_lpc_precision_search > 0 && !cfg->use_direct_mse) {
        /* EXTENSION: the order-P coefficients quantised with fewer bits (ranked after the lower orders; strictly fewer
         * subframe bits to win) */
        for (int k = 1; k <= cfg->ext_lpc_precision_search && cfg->quant_precision - k >= 1; k++) {
            fo_subframe cand;
            fo_qlpc_subframe(cfg, coefs, lpc_order, cfg->quant_precision - k, signal, n, bps, &cand);
            if (cand.bits < out->bits) {
                free(out->residual);
                *out = cand;
            } else {
                free(cand.residual);
            }
        }
    }
}

static void fo_subframe_free(fo_subframe *sf) {
    free(sf->residual);
    sf->residual = NULL;
}

/* src/coding.rs:384-418  encode_subframe */
void fo_encode_subframe(const fo_config *cfg, const int32_t *samples, int n, int bps, fo_subframe *out) {
    if (cfg->use_constant && fo_is_constant(samples, n)) {
        fo_subframe_reset(out, FO_SF_CONSTANT, samples, n, bps);
        out->bits = 8 + (uint64_t)bps; /* src/component/bitrepr.rs:445-447 */
        return;
    }
    uint64_t verbatim_bits = 8 + (uint64_t)n * bps; /* src/component/datatype.rs:1944-1949 */
    int too_short = n < 64;                        /* src/constant.rs:51 */
    fo_subframe fixed, lpc;
    int have_fixed = 0, have_lpc = 0;
    if (!too_short && cfg->use_fixed) have_fixed = fo_fixed_lpc(cfg, samples, n, bps, verbatim_bits, &fixed);
    uint64_t baseline_bits = have_fixed ? FO_MIN(verbatim_bits, fixed.bits) : verbatim_bits;
    if (!too_short && cfg->use_lpc) {
        fo_estimated_qlpc(cfg, samples, n, bps, &lpc);
        if (lpc.bits < baseline_bits) {
            have_lpc = 1;
        } else {
            fo_subframe_free(&lpc);
        }
    }
    if (have_lpc) {
        if (have_fixed) fo_subframe_free(&fixed);
        if (lpc.bits < verbatim_bits) {
            *out = lpc;
            return;
        }
        fo_subframe_free(&lpc);
    } else if (have_fixed) {
        if (fixed.bits < verbatim_bits) {
            *out = fixed;
            return;
        }
        fo_subframe_free(&fixed);
    }
    fo_subframe_reset(out, FO_SF_VERBATIM, samples, n, bps);
    out->bits = verbatim_bits;
}

/* src/source.rs:262-275  FrameBuf::verify_samples */
static int fo_verify_samples(const int32_t *planar, int channels, int stride, int n, int bps) {
    int32_t max_allowed = (int32_t)((1u << (bps - 1)) - 1u);
    int32_t min_allowed = -(int32_t)(1u << (bps - 1));
    for (int ch = 0; ch < channels; ch++) {
        const int32_t *s = planar + (size_t)ch * stride;
        for (int t = 0; t < n; t++)
            if (s[t] < min_allowed || s[t] > max_allowed) return 1;
    }
    return 0;
}

/* src/coding.rs:421-452 encode_frame_impl, :454-527 try_stereo_coding / recombine_stereo_frame,
 * :530-544 encode_frame, :581-606 encode_fixed_size_frame */
int fo_encode_frame(const fo_config *cfg, const int32_t *planar, int channels, int stride, int n, int bps,
                    int sample_rate, uint32_t frame_number, fo_frame *out) {
    memset(out, 0, sizeof(*out));
    if (frame_number >= (1u << 31)) return 1;
    if (fo_verify_samples(planar, channels, stride, n, bps)) return 1;
    out->channels = channels;
    out->n = n;
    out->bps = bps;
    out->sample_rate = sample_rate;
    out->frame_number = frame_number;
    out->ch_assignment = FO_CH_INDEPENDENT;
    for (int ch = 0; ch < channels; ch++) {
        fo_encode_subframe(cfg, planar + (size_t)ch * stride, n, bps, &out->sub[ch]);
    }
    if (channels == 2) {
        const int32_t *l = planar, *r = planar + stride;
        out->ms_buf = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)n);
        int32_t *m = out->ms_buf, *s = out->ms_buf + n;
        for (int t = 0; t < n; t++) {
            m[t] = (l[t] + r[t]) >> 1;
            s[t] = l[t] - r[t];
        }
        fo_subframe sm, ss;
        fo_encode_subframe(cfg, m, n, bps, &sm);     /* MidSide: offset(0) = 0 */
        fo_encode_subframe(cfg, s, n, bps + 1, &ss); /* MidSide: offset(1) = 1 */
        uint64_t bl = out->sub[0].bits, br = out->sub[1].bits, bm = sm.bits, bs = ss.bits;
        uint64_t min_bits = bl + br;
        int tag = FO_CH_INDEPENDENT;
        if (cfg->use_leftside && bl + bs < min_bits) {
            min_bits = bl + bs;
            tag = FO_CH_LEFT_SIDE;
        }
        if (cfg->use_rightside && br + bs < min_bits) {
            min_bits = br + bs;
            tag = FO_CH_RIGHT_SIDE;
        }
        if (cfg->use_midside && bm + bs < min_bits) {
            min_bits = bm + bs;
            tag = FO_CH_MID_SIDE;
        }
        out->ch_assignment = tag;
        /* ChannelAssignment::select_channels, src/component/datatype.rs:1171-1184 */
        switch (tag) {
        case FO_CH_LEFT_SIDE:
            fo_subframe_free(&out->sub[1]);
            fo_subframe_free(&sm);
            out->sub[1] = ss;
            break;
        case FO_CH_RIGHT_SIDE:
            fo_subframe_free(&out->sub[0]);
            fo_subframe_free(&sm);
            out->sub[0] = ss;
            break;
        case FO_CH_MID_SIDE:
            fo_subframe_free(&out->sub[0]);
            fo_subframe_free(&out->sub[1]);
            out->sub[0] = sm;
            out->sub[1] = ss;
            break;
        default:
            fo_subframe_free(&sm);
            fo_subframe_free(&ss);
            break;
        }
    }
    return 0;
}

void fo_frame_free(fo_frame *f) {
    for (int ch = 0; ch < FO_MAX_CHANNELS; ch++) fo_subframe_free(&f->sub[ch]);
    free(f->ms_buf);
    f->ms_buf = NULL;
}

/* ------------------------------------------------------------------------------------------
 * component/bitrepr.rs + bitsink.rs: MSB-first bit writer
 * ------------------------------------------------------------------------------------------ */

typedef struct {
    uint8_t *buf;
    size_t cap;      /* bytes */
    uint64_t bitpos; /* bits written */
    int overflow;
} fo_sink;

/* write the `nbits` least-significant bits of v, MSB first (src/bitsink.rs write_lsbs) */
static void fo_put(fo_sink *s, uint64_t v, int nbits) {
    for (int i = nbits - 1; i >= 0; i--) {
        uint64_t byte = s->bitpos >> 3;
        if (byte >= s->cap) {
            s->overflow = 1;
            s->bitpos++;
            continue;
        }
        int bit = (int)((v >> i) & 1u);
        int sh = 7 - (int)(s->bitpos & 7);
        if (sh == 7) s->buf[byte] = 0;
        s->buf[byte] = (uint8_t)(s->buf[byte] | (bit << sh));
        s->bitpos++;
    }
}

static void fo_put_zeros(fo_sink *s, uint64_t count) {
    while (count > 0) {
        int k = count > 32 ? 32 : (int)count;
        fo_put(s, 0, k);
        count -= (uint64_t)k;
        if (s->overflow && (s->bitpos >> 3) >= s->cap) {
            s->bitpos += count;
            return;
        }
    }
}

/* src/bitsink.rs:209-217 write_twoc */
static void fo_put_twoc(fo_sink *s, int32_t v, int nbits) {
    uint64_t mask = nbits >= 64 ? ~0ull : ((1ull << nbits) - 1);
    fo_put(s, (uint64_t)(int64_t)v & mask, nbits);
}

/* crc crate: CRC_8_SMBUS (poly 0x07, init 0, no reflection, xorout 0) */
uint8_t fo_crc8(const uint8_t *data, size_t len) {
    uint8_t crc = 0;
    for (size_t i = 0; i < len; i++) {
        crc ^= data[i];
        for (int b = 0; b < 8; b++) crc = (uint8_t)((crc & 0x80) ? ((crc << 1) ^ 0x07) : (crc << 1));
    }
    return crc;
}

/* crc crate: CRC_16_UMTS (poly 0x8005, init 0, no reflection, xorout 0) */
uint16_t fo_crc16(const uint8_t *data, size_t len) {
    static uint16_t table[256];
    static int init = 0;
    if (!init) {
        for (int i = 0; i < 256; i++) {
            uint16_t c = (uint16_t)(i << 8);
            for (int b = 0; b < 8; b++) c = (uint16_t)((c & 0x8000) ? ((c << 1) ^ 0x8005) : (c << 1));
            table[i] = c;
        }
        init = 1;
    }
    uint16_t crc = 0;
    for (size_t i = 0; i < len; i++) crc = (uint16_t)((crc << 8) ^ table[((crc >> 8) ^ data[i]) & 0xFF]);
    return crc;
}

/* src/component/bitrepr.rs:121-159  encode_to_utf8like */
int fo_encode_utf8like(uint64_t val, uint8_t *out) {
    int code_bits = val ? 64 - __builtin_clzll(val) : 0;
    if (code_bits <= 7) {
        out[0] = (uint8_t)val;
        return 1;
    }
    if (code_bits > 36) return -1;
    int trailing = (code_bits - 2) / 5;
    int n = 0;
    static const uint8_t heads[7] = {0x80, 0xC0, 0xE0, 0xF0, 0xF8, 0xFC, 0xFE};
    if (trailing == 6) {
        out[n++] = 0xFE;
    } else {
        out[n++] = (uint8_t)(heads[trailing] | (uint8_t)(val >> (6 * trailing)));
    }
    for (int i = trailing - 1; i >= 0; i--) out[n++] = (uint8_t)(0x80 | ((val >> (6 * i)) & 0x3F));
    return n;
}

/* src/component/datatype.rs:1239-1249,1281-1291 (BlockSizeSpec), :1350-1360 (SampleSizeSpec),
 * :1427-1453,1502-1520 (SampleRateSpec) */
static int fo_block_size_tag(int size, int *extra_bits, uint32_t *extra) {
    *extra_bits = 0;
    *extra = 0;
    if (size == 192) return 1;
    if (size == 576 || size == 1152 || size == 2304 || size == 4608) return 2 + __builtin_ctz((unsigned)(size / 576));
    if (size == 256 || size == 512 || size == 1024 || size == 2048 || size == 4096 || size == 8192 ||
        size == 16384 || size == 32768)
        return 8 + __builtin_ctz((unsigned)(size / 256));
    if (size <= 256) {
        *extra_bits = 8;
        *extra = (uint32_t)(size - 1);
        return 6;
    }
    *extra_bits = 16;
    *extra = (uint32_t)(size - 1);
    return 7;
}

static int fo_sample_size_tag(int bits) {
    switch (bits) {
    case 8: return 1;
    case 12: return 2;
    case 16: return 4;
    case 20: return 5;
    case 24: return 6;
    case 32: return 7;
    default: return 0; /* Unspecified, src/coding.rs:433 */
    }
}

static int fo_sample_rate_tag(uint32_t freq, int *extra_bits, uint32_t *extra) {
    *extra_bits = 0;
    *extra = 0;
    switch (freq) {
    case 88200: return 1;
    case 176400: return 2;
    case 192000: return 3;
    case 8000: return 4;
    case 16000: return 5;
    case 22050: return 6;
    case 24000: return 7;
    case 32000: return 8;
    case 44100: return 9;
    case 48000: return 10;
    case 96000: return 11;
    default: break;
    }
    if (freq % 1000 == 0 && freq / 1000 <= 255) {
        *extra_bits = 8;
        *extra = freq / 1000;
        return 12;
    }
    if (freq % 10 == 0 && freq / 10 <= 65535) {
        *extra_bits = 16;
        *extra = freq / 10;
        return 14;
    }
    if (freq <= 65535) {
        *extra_bits = 16;
        *extra = freq;
        return 13;
    }
    return 0; /* Unspecified, src/coding.rs:434-435 */
}

/* src/component/bitrepr.rs:359-420  FrameHeader::write (+ CRC-8). ch_tag: (channels-1) or 8/9/10. */
int fo_frame_header_bytes(int n, int ch_tag, int bps, int sample_rate, int variable, uint64_t number,
                          uint8_t *out) {
    int bs_extra_bits, sr_extra_bits;
    uint32_t bs_extra, sr_extra;
    int bs_tag = fo_block_size_tag(n, &bs_extra_bits, &bs_extra);
    int sr_tag = fo_sample_rate_tag((uint32_t)sample_rate, &sr_extra_bits, &sr_extra);
    int k = 0;
    out[k++] = 0xFF;
    out[k++] = (uint8_t)(0xF8 + (variable ? 1 : 0));
    out[k++] = (uint8_t)((bs_tag << 4) | sr_tag);
    out[k++] = (uint8_t)((ch_tag << 4) | (fo_sample_size_tag(bps) << 1));
    int u = fo_encode_utf8like(number, out + k);
    if (u < 0) return -1;
    k += u;
    if (bs_extra_bits == 8) out[k++] = (uint8_t)bs_extra;
    if (bs_extra_bits == 16) {
        out[k++] = (uint8_t)(bs_extra >> 8);
        out[k++] = (uint8_t)bs_extra;
    }
    if (sr_extra_bits == 8) out[k++] = (uint8_t)sr_extra;
    if (sr_extra_bits == 16) {
        out[k++] = (uint8_t)(sr_extra >> 8);
        out[k++] = (uint8_t)sr_extra;
    }
    out[k] = fo_crc8(out, (size_t)k);
    return k + 1;
}

static int fo_ch_tag(const fo_frame *f) {
    return f->ch_assignment == FO_CH_INDEPENDENT ? f->channels - 1 : f->ch_assignment;
}

/* src/component/bitrepr.rs:275-287 Frame::count_bits, :361-371 FrameHeader::count_bits */
uint64_t fo_frame_count_bits(const fo_frame *f) {
    uint8_t hdr[16];
    int hb = fo_frame_header_bytes(f->n, fo_ch_tag(f), f->bps, f->sample_rate, 0, f->frame_number, hdr);
    uint64_t header = (uint64_t)(hb - 1) * 8 + 8; /* 40 + utf8 + extras, CRC-8 included in the 40 */
    uint64_t body = 0;
    for (int ch = 0; ch < f->channels; ch++) body += f->sub[ch].bits;
    uint64_t aligned = ((header + body + 7) >> 3) << 3;
    return aligned + 16;
}

/* src/component/bitrepr.rs:550-597  Residual::write */
static void fo_write_residual(fo_sink *s, const fo_subframe *sf) {
    int nparts = 1 << sf->part_order;
    int param_bits = sf->rice2 ? 5 : 4;
    fo_put(s, (uint64_t)((sf->rice2 ? 1 : 0) << 4) | (uint64_t)sf->part_order, 6);
    int part_len = sf->n >> sf->part_order;
    int offset = 0;
    for (int p = 0; p < nparts; p++) {
        int rice_p = sf->rice_params[p];
        fo_put(s, (uint64_t)rice_p, param_bits);
        int start = FO_MAX(sf->order, offset); /* warmup_length == order for fixed and lpc */
        offset += part_len;
        for (int t = start; t < offset; t++) {
            uint32_t u = fo_encode_signbit(sf->residual[t]);
            uint32_t q = u >> rice_p;
            uint32_t r = u & ((1u << rice_p) - 1u);
            fo_put_zeros(s, q);
            fo_put(s, (uint64_t)(r | (1u << rice_p)), rice_p + 1);
        }
    }
}

/* src/component/bitrepr.rs:441-528  Constant/Verbatim/FixedLpc/Lpc ::write */
static void fo_write_subframe(fo_sink *s, const fo_subframe *sf) {
    switch (sf->type) {
    case FO_SF_CONSTANT:
        fo_put(s, 0x00, 8);
        fo_put_twoc(s, sf->samples[0], sf->bps);
        break;
    case FO_SF_VERBATIM:
        fo_put(s, 0x02, 8);
        for (int t = 0; t < sf->n; t++) fo_put_twoc(s, sf->samples[t], sf->bps);
        break;
    case FO_SF_FIXED:
        fo_put(s, (uint64_t)(0x10 | (sf->order << 1)), 8);
        for (int t = 0; t < sf->order; t++) fo_put_twoc(s, sf->samples[t], sf->bps);
        fo_write_residual(s, sf);
        break;
    case FO_SF_LPC:
        fo_put(s, (uint64_t)(0x40 | ((sf->order - 1) << 1)), 8);
        for (int t = 0; t < sf->order; t++) fo_put_twoc(s, sf->samples[t], sf->bps);
        fo_put(s, (uint64_t)(sf->precision - 1), 4);
        fo_put_twoc(s, sf->shift, 5);
        for (int j = 0; j < sf->order; j++) fo_put_twoc(s, sf->qlp[j], sf->precision);
        fo_write_residual(s, sf);
        break;
    default:
        assert(0);
    }
}

/* test hook: BitRepr::write of one subframe into a zeroed buffer; returns bits written */
int64_t fo_subframe_write(const fo_subframe *sf, uint8_t *out, size_t cap) {
    fo_sink s = {out, cap, 0, 0};
    fo_write_subframe(&s, sf);
    return s.overflow ? -1 : (int64_t)s.bitpos;
}

/* src/component/bitrepr.rs:289-320  Frame::write */
int64_t fo_frame_write(const fo_frame *f, uint8_t *out, size_t cap) {
    uint64_t total_bits = fo_frame_count_bits(f);
    if ((total_bits >> 3) > cap) return -1;
    uint8_t hdr[16];
    int hb = fo_frame_header_bytes(f->n, fo_ch_tag(f), f->bps, f->sample_rate, 0, f->frame_number, hdr);
    if (hb < 0) return -1;
    memcpy(out, hdr, (size_t)hb);
    fo_sink s = {out, cap, (uint64_t)hb * 8, 0};
    for (int ch = 0; ch < f->channels; ch++) fo_write_subframe(&s, &f->sub[ch]);
    if (s.bitpos & 7) fo_put(&s, 0, 8 - (int)(s.bitpos & 7)); /* align_to_byte */
    size_t nbytes = (size_t)(s.bitpos >> 3);
    if (s.overflow || nbytes + 2 > cap) return -1;
    assert((uint64_t)(nbytes + 2) * 8 == total_bits); /* count_bits == written bits (bitrepr.rs:96-105) */
    uint16_t crc = fo_crc16(out, nbytes);
    out[nbytes] = (uint8_t)(crc >> 8);
    out[nbytes + 1] = (uint8_t)crc;
    return (int64_t)nbytes + 2;
}

/* ------------------------------------------------------------------------------------------
 * Driver: frames of a stream.  Serial loop = src/coding.rs:662-674; nthreads > 1 mirrors par.rs
 * (src/par.rs:355-449: workers pull frame indices, results re-ordered by frame number).
 * ------------------------------------------------------------------------------------------ */

typedef struct {
    const fo_config *cfg;
    const int32_t *interleaved;
    uint64_t n_samples;
    int channels, bps, sample_rate, block_size;
    uint32_t first_frame_number;
    uint64_t n_frames;
    uint64_t next; /* atomic work counter */
    uint8_t **frame_bytes;
    int64_t *frame_len;
    int error;
} fo_job;

static void fo_encode_one(fo_job *job, uint64_t fi, int32_t *planar) {
    int bs = job->block_size;
    uint64_t start = fi * (uint64_t)bs;
    int n = (int)FO_MIN((uint64_t)bs, job->n_samples - start);
    int ch_n = job->channels;
    /* src/arrayutils.rs:248-264 deinterleave into FrameBuf layout (stride = block_size) */
    const int32_t *src = job->interleaved + start * (uint64_t)ch_n;
    for (int t = 0; t < n; t++)
        for (int ch = 0; ch < ch_n; ch++) planar[(size_t)ch * bs + t] = src[(size_t)t * ch_n + ch];
    fo_frame fr;
    int rc = fo_encode_frame(job->cfg, planar, ch_n, bs, n, job->bps, job->sample_rate,
                             job->first_frame_number + (uint32_t)fi, &fr);
    if (rc) {
        job->frame_len[fi] = -1;
        __atomic_store_n(&job->error, 1, __ATOMIC_RELAXED);
        fo_frame_free(&fr);
        return;
    }
    uint64_t bits = fo_frame_count_bits(&fr);
    size_t cap = (size_t)(bits >> 3);
    uint8_t *buf = (uint8_t *)malloc(cap ? cap : 1);
    int64_t len = fo_frame_write(&fr, buf, cap);
    job->frame_bytes[fi] = buf;
    job->frame_len[fi] = len;
    fo_frame_free(&fr);
}

static void *fo_worker(void *arg) {
    fo_job *job = (fo_job *)arg;
    int32_t *planar = (int32_t *)malloc(sizeof(int32_t) * (size_t)job->block_size * (size_t)job->channels);
    for (;;) {
        uint64_t fi = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
        if (fi >= job->n_frames) break;
        fo_encode_one(job, fi, planar);
    }
    free(planar);
    return NULL;
}

int64_t fo_encode_frames(const fo_config *cfg, const int32_t *interleaved, uint64_t n_samples, int channels,
                         int bps, int sample_rate, int block_size, uint32_t first_frame_number, int nthreads,
                         uint8_t *out, size_t cap, uint32_t *frame_sizes) {
    uint64_t n_frames = (n_samples + (uint64_t)block_size - 1) / (uint64_t)block_size;
    if (n_frames == 0) return 0;
    fo_job job;
    memset(&job, 0, sizeof(job));
    job.cfg = cfg;
    job.interleaved = interleaved;
    job.n_samples = n_samples;
    job.channels = channels;
    job.bps = bps;
    job.sample_rate = sample_rate;
    job.block_size = block_size;
    job.first_frame_number = first_frame_number;
    job.n_frames = n_frames;
    job.frame_bytes = (uint8_t **)calloc(n_frames, sizeof(uint8_t *));
    job.frame_len = (int64_t *)calloc(n_frames, sizeof(int64_t));
    if (nthreads <= 1) {
        fo_worker(&job);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
        for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, fo_worker, &job);
        for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
        free(th);
    }
    int64_t total = 0;
    int64_t rc = 0;
    for (uint64_t i = 0; i < n_frames; i++) {
        if (job.frame_len[i] < 0) rc = -1;
    }
    if (!rc) {
        for (uint64_t i = 0; i < n_frames; i++) {
            if ((size_t)(total + job.frame_len[i]) > cap) {
                rc = -2;
                break;
            }
            memcpy(out + total, job.frame_bytes[i], (size_t)job.frame_len[i]);
            if (frame_sizes) frame_sizes[i] = (uint32_t)job.frame_len[i];
            total += job.frame_len[i];
        }
    }
    for (uint64_t i = 0; i < n_frames; i++) free(job.frame_bytes[i]);
    free(job.frame_bytes);
    free(job.frame_len);
    return rc ? rc : total;
}

/* src/coding.rs:645-695 encode_with_fixed_block_size + src/component/bitrepr.rs:172-267
 * (Stream / MetadataBlock / StreamInfo ::write) + src/component/datatype.rs:514-523. */
int64_t fo_encode_stream(const fo_config *cfg, const int32_t *interleaved, uint64_t n_samples, int channels,
                         int bps, int sample_rate, int block_size, int nthreads, uint8_t *out, size_t cap) {
    if (cap < 42) return -2;
    uint64_t n_frames = (n_samples + (uint64_t)block_size - 1) / (uint64_t)block_size;
    uint32_t *sizes = (uint32_t *)calloc(n_frames ? n_frames : 1, sizeof(uint32_t));
    int64_t body = fo_encode_frames(cfg, interleaved, n_samples, channels, bps, sample_rate, block_size, 0,
                                    nthreads, out + 42, cap - 42, sizes);
    if (body < 0) {
        free(sizes);
        return body;
    }
    uint32_t min_block = 0xFFFF, max_block = 0, min_frame = 0xFFFFFFFFu, max_frame = 0;
    for (uint64_t i = 0; i < n_frames; i++) {
        uint32_t b = (uint32_t)FO_MIN((uint64_t)block_size, n_samples - i * (uint64_t)block_size);
        min_block = FO_MIN(min_block, b);
        max_block = FO_MAX(max_block, b);
        min_frame = FO_MIN(min_frame, sizes[i]);
        max_frame = FO_MAX(max_frame, sizes[i]);
    }
    if (n_frames > 0) min_block = max_block; /* src/coding.rs:681-687 */
    free(sizes);
    uint8_t md5[16];
    fo_md5_of_samples(interleaved, (size_t)(n_samples * (uint64_t)channels), (bps + 7) / 8, md5);
    uint8_t *p = out;
    memcpy(p, "fLaC", 4);
    p[4] = 0x80; /* last-metadata-block flag | type 0 (STREAMINFO) */
    p[5] = 0;
    p[6] = 0;
    p[7] = 34;
    fo_sink s = {out, 42, 64, 0};
    fo_put(&s, min_block, 16);
    fo_put(&s, max_block, 16);
    fo_put(&s, min_frame & 0xFFFFFF, 24);
    fo_put(&s, max_frame & 0xFFFFFF, 24);
    fo_put(&s, (uint64_t)sample_rate, 20);
    fo_put(&s, (uint64_t)(channels - 1), 3);
    fo_put(&s, (uint64_t)(bps - 1), 5);
    fo_put(&s, n_samples & 0xFFFFFFFFFull, 36);
    memcpy(out + 26, md5, 16);
    return 42 + body;
}
