/*
 * tq_oracle.c -- scalar CPU restatement of the term-quantization hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see tq_oracle.h).  Written for clarity, not speed:
 * terms are enumerated level by level exactly in the order the reference's
 * merge loop would pick them, which is a different formulation from the CUDA
 * kernels (cut-level search) so that the two check each other.
 *
 * Parity: pinned against oracle/_ref (the reference kernel body built for the
 * host) and the golden vectors under tests/golden/.
 */
#include "tq_oracle.h"

#include <math.h>
#include <stddef.h>
#include <string.h>

/* ---- quantiser: kernels/tr_cuda_kernel.cu:21-22 ------------------------- */

static int32_t sat_trunc_i32(double v)
{
    /* PTX cvt.rzi.s32.f64: NaN -> 0, saturating */
    if (v != v) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}

static int32_t clip_q(int32_t q, int bits)
{
    /* fminf(float(q), float(2^bits - 1)) then float -> int32 (:21-22) */
    float maxv = (float)(pow(2.0, bits) - 1.0);
    float f = fminf((float)q, maxv);
    return sat_trunc_i32((double)f);
}

int32_t tqo_quantize_f32(float x, float sf, int bits)
{
    float r = fabsf(x) / sf;              /* IEEE fp32 divide (div.rn.f32)  */
    double t = (double)r + 0.5;           /* the 0.5 literal is a double    */
    return clip_q(sat_trunc_i32(t), bits);
}

int32_t tqo_quantize_f64(double x, float sf, int bits)
{
    double r = fabs(x) / (double)sf;      /* scalar_t = double: fp64 divide */
    return clip_q(sat_trunc_i32(r + 0.5), bits);
}

/* ---- term expansion: kernels/tr_cuda_kernel.cu:29-55 -------------------- */

void tqo_terms(uint32_t q, int encoding, uint32_t *pos, uint32_t *neg)
{
    uint32_t P = 0, N = 0;
    if (encoding == TQO_ENC_BINARY) {
        P = q;
    } else {
        /* walk maximal runs of 1-bits [lo..hi] */
        int i = 0;
        while (i < 31) {
            if (!((q >> i) & 1u)) { i++; continue; }
            int lo = i;
            while (i < 31 && ((q >> i) & 1u)) i++;
            int hi = i - 1;
            if (hi == lo && encoding == TQO_ENC_HESE) {
                P |= 1u << lo;                 /* window 010 -> +2^i  (:38-41) */
            } else {
                P |= 1u << (hi + 1);           /* window 011 -> +2^(i+1) (:42-44) */
                N |= 1u << lo;                 /* window 110 -> -2^i  (:49-51) */
            }
        }
    }
    *pos = P;
    *neg = N;
}

/* ---- group truncation: kernels/tr_cuda_kernel.cu:85-123 ----------------- */

static void tr_group(const int32_t *q, const int *sgn, int n, int alpha, int encoding,
                     int32_t *val)
{
    uint32_t P[32], N[32];
    for (int j = 0; j < n; j++) {
        tqo_terms((uint32_t)q[j], encoding, &P[j], &N[j]);
        val[j] = 0;
    }
    int kept = 0;
    /* largest magnitude first; equal magnitudes -> lowest index (:96-103) */
    for (int p = 31; p >= 0 && kept < alpha; p--) {
        for (int j = 0; j < n && kept < alpha; j++) {
            if ((P[j] >> p) & 1u)      { val[j] += sgn[j] * (int32_t)(1u << p); kept++; }
            else if ((N[j] >> p) & 1u) { val[j] -= sgn[j] * (int32_t)(1u << p); kept++; }
        }
    }
}

int tqo_tr(const void *in, void *out, int32_t *codes, int dtype,
           int64_t B, int64_t C, int64_t WH, float sf, int bits, int g, int alpha,
           int encoding, int relu)
{
    if (g < 1 || g > 32 || bits < 1 || bits > 24 || alpha < 0) return -1;
    if (encoding < 0 || encoding > 2) return -1;
    if (dtype != TQO_F32 && dtype != TQO_F64) return -1;
    const float *inf = (const float *)in;
    const double *ind = (const double *)in;
    float *outf = (float *)out;
    double *outd = (double *)out;
    int64_t ngroups = (C + g - 1) / g;

    for (int64_t b = 0; b < B; b++)
    for (int64_t cg = 0; cg < ngroups; cg++)
    for (int64_t wh = 0; wh < WH; wh++) {
        int32_t q[32], val[32];
        int sgn[32];
        int n = (int)((cg * g + g <= C) ? g : (C - cg * g));
        int64_t base = b * C * WH + cg * g * WH + wh;
        for (int j = 0; j < n; j++) {
            int64_t idx = base + (int64_t)j * WH;
            if (dtype == TQO_F32) {
                float x = inf[idx];
                if (relu && x < 0) x = 0.0f;
                q[j] = tqo_quantize_f32(x, sf, bits);
                sgn[j] = x < 0 ? -1 : 1;                    /* :23 */
            } else {
                double x = ind[idx];
                if (relu && x < 0) x = 0.0;
                q[j] = tqo_quantize_f64(x, sf, bits);
                sgn[j] = x < 0 ? -1 : 1;
            }
        }
        tr_group(q, sgn, n, alpha, encoding, val);
        for (int j = 0; j < n; j++) {
            int64_t idx = base + (int64_t)j * WH;
            if (codes) codes[idx] = val[j];
            if (out) {
                /* integer-valued accumulate (:112) then one multiply by sf (:122) */
                if (dtype == TQO_F32) outf[idx] = (float)val[j] * sf;
                else                  outd[idx] = (double)val[j] * (double)sf;
            }
        }
    }
    return 0;
}

/* ---- calibration: tr_layer.py:90-94, 43-54 ------------------------------ */

void tqo_hist_f32(const float *x, int64_t n, float *hist, int nbins, float lo, float hi)
{
    for (int64_t i = 0; i < n; i++) {
        float v = x[i];
        if (!(v >= lo && v <= hi)) continue;
        float t = (v - lo) * (float)nbins;
        int bin = (int)(t / (hi - lo));
        if (bin == nbins) bin -= 1;
        hist[bin] += 1.0f;
    }
}

int tqo_mse_profile(const float *hist, const float *x, int nbins,
                    const float *sfs, int nsf, int bits, int terms, double *errs)
{
    int best = 0;
    float best_err = 0;
    for (int s = 0; s < nsf; s++) {
        double acc = 0.0;
        for (int b = 0; b < nbins; b++) {
            float xh;
            tqo_tr(&x[b], &xh, NULL, TQO_F32, 1, 1, 1, sfs[s], bits, 1, terms,
                   TQO_ENC_HESE, 0);
            float d = x[b] - xh;          /* each elementwise op rounds to fp32 (:50) */
            float d2 = d * d;
            float t = hist[b] * d2;
            acc += (double)t;
        }
        if (errs) errs[s] = acc;
        float e = (float)acc;              /* errs live in a float32 tensor (:53) */
        if (s == 0 || e < best_err) { best = s; best_err = e; }
    }
    return best;
}

/* ---- parameter-bit accounting: tr_layer.py:57-63 ------------------------ */

int64_t tqo_hese_term_count_f32(const float *w, int64_t n, float sf)
{
    int64_t total = 0;
    for (int64_t i = 0; i < n; i++) {
        float r = w[i] / sf;
        int32_t k = sat_trunc_i32((double)r);       /* .int(): toward zero (:60) */
        uint32_t m = (uint32_t)(k < 0 ? -(int64_t)k : k);
        uint32_t P, N;
        tqo_terms(m, TQO_ENC_HESE, &P, &N);
        total += __builtin_popcount(P) + __builtin_popcount(N);
    }
    return total;
}

/* ---- integer contraction on truncated codes ----------------------------- */

void tqo_gemm_i32(const int16_t *a, const int16_t *w, int32_t *acc,
                  int64_t M, int64_t N, int64_t K)
{
    for (int64_t m = 0; m < M; m++)
        for (int64_t n = 0; n < N; n++) {
            int64_t s = 0;
            for (int64_t k = 0; k < K; k++)
                s += (int32_t)a[m * K + k] * (int32_t)w[n * K + k];
            acc[m * N + n] = (int32_t)s;
        }
}

/* ---- fused conv tail: one IEEE fp32 operation per step ------------------- */

/* y[i] = fmaf(x[i], a[i % C], b[i % C]) -- the BatchNorm affine of the fused conv epilogue, a single rounding
 * (csrc/tq_gemm.cu epi_affine).  glibc's fmaf is correctly rounded with or without a hardware FMA. */
void tqo_fma_channels_f32(const float *x, const float *a, const float *b, float *y, int64_t n, int64_t C)
{
    for (int64_t i = 0; i < n; i++) y[i] = fmaf(x[i], a[i % C], b[i % C]);
}
