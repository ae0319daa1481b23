/*
 * tq_oracle.h -- CPU restatement of the term-quantization (TR) hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * timed CPU baseline.  The product path (term_quantization_b200/) never links
 * or imports it and fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  oracle/Makefile also builds oracle/_ref/libtq_ref.so,
 * which is the reference's own kernel body (kernels/tr_cuda_kernel.cu:12-125)
 * compiled for the host from where it lies under /root/reference; the
 * restatement below is checked against it (tests/test_oracle_vs_ref.py, run
 * where /root/reference exists) and against the golden vectors generated from
 * it (tests/golden/, tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it restates.
 */
#ifndef TQ_ORACLE_H
#define TQ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* term encodings (reference ships HESE only) */
#define TQO_ENC_HESE   0   /* kernels/tr_cuda_kernel.cu:29-55 == bit_utils.py:10-44 */
#define TQO_ENC_BINARY 1   /* bit_utils.py:63-73 (plain binary digits)               */
#define TQO_ENC_BOOTH  2   /* verilog/booth_encoder.v:57-78 (radix-2 Booth)          */

/* dtype tags shared with include/tq_b200.h */
#define TQO_F32 0
#define TQO_F64 1

/* kernels/tr_cuda_kernel.cu:21-23 : q = min(int32(|x|/sf + 0.5), 2^bits-1).
 * The division is done in the input type, the +0.5 in double, the conversion
 * truncates and saturates (PTX cvt.rzi.s32.f64), NaN -> 0. */
int32_t tqo_quantize_f32(float x, float sf, int bits);
int32_t tqo_quantize_f64(double x, float sf, int bits);

/* Signed power-of-two term masks of a non-negative integer q:
 * bit p of *pos => +2^p, bit p of *neg => -2^p; pos & neg == 0; pos - neg == q. */
void tqo_terms(uint32_t q, int encoding, uint32_t *pos, uint32_t *neg);

/* kernels/tr_cuda_kernel.cu:58-125, one call over a (B, C, WH) tensor; groups of g
 * consecutive channels at fixed (b, wh) with element stride WH, the alpha largest
 * terms of each group survive (level descending, then index ascending: strict '>'
 * at :99).  A tail group (C % g != 0) holds the C % g real values and the full
 * budget (== the reference on a zero-padded tensor).  dtype TQO_F32/TQO_F64.
 * codes (optional, may be NULL) receives sign * truncated integer per element.
 * relu != 0 clamps negative inputs to +0 before quantisation (fused variant).
 * Returns 0, or -1 on invalid arguments. */
int tqo_tr(const void *in, void *out, int32_t *codes, int dtype,
           int64_t B, int64_t C, int64_t WH, float sf, int bits, int g, int alpha,
           int encoding, int relu);

/* torch.histc(x, nbins, lo, hi) as the CUDA kernel computes it
 * (tr_layer.py:92): bin = (int)((x - lo) * nbins / (hi - lo)) in float,
 * bin == nbins -> nbins-1, values outside [lo, hi] ignored.  hist += counts. */
void tqo_hist_f32(const float *x, int64_t n, float *hist, int nbins, float lo, float hi);

/* tr_layer.py:43-54 : for every scale factor, err = sum_b hist[b]*(x[b]-tr(x[b]))^2
 * with g=1, accumulated in double; returns the index of the first minimum.
 * errs (optional) receives the nsf sums. */
int tqo_mse_profile(const float *hist, const float *x, int nbins,
                    const float *sfs, int nsf, int bits, int terms, double *errs);

/* tr_layer.py:57-63 : sum over elements of len(hese(int(w/sf))) where int()
 * truncates toward zero (tr_layer.py:60) -- the caller multiplies by bit_width. */
int64_t tqo_hese_term_count_f32(const float *w, int64_t n, float sf);

/* Exact integer GEMM on term-truncated codes: acc[m,n] = sum_k a[m,k]*w[n,k]. */
void tqo_gemm_i32(const int16_t *a, const int16_t *w, int32_t *acc,
                  int64_t M, int64_t N, int64_t K);

/* y[i] = fmaf(x[i], a[i % C], b[i % C]): the per-channel affine of the fused conv epilogue, one rounding. */
void tqo_fma_channels_f32(const float *x, const float *a, const float *b, float *y, int64_t n, int64_t C);

#ifdef __cplusplus
}
#endif
#endif
