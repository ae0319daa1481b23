// tq_dw.cu -- the memory-bound layers of the depthwise CNNs (BASELINE.json configs[3], MobileNet-V2) on term codes.
//
// The reference wraps a depthwise conv like any other (cnn_models/__init__.py:52-65 gives it the "effectively
// unquantised" weight setting (16, 1, 16)), so per forward it runs: TR encode of the input (tr_layer.py:96-99, g = 1) ->
// cuDNN fp32 depthwise conv on the dequantised values -> BatchNorm -> ReLU6, and the next wrapped conv encodes the
// result again: five passes over the activation.  Here ONE kernel reads the input's fp16 term codes (written by the
// producing conv's epilogue), accumulates the 9 taps in int32 (|a| <= 2^bits, |w| <= 2^15: |acc| < 2^29, exact),
// applies t = float(acc) * (sf_x * sf_w) (+ bias), fma(t, bn_a, bn_b), ReLU / ReLU6, and writes the fp16 term codes of
// the consumer's quantiser (and / or the fp32 value).  Algorithmic bytes: 2 B read + 2 B written per element.
//
// Mapping: NHWC, one thread = 8 channels (one 128-bit load) x DW_PIX consecutive output pixels of a row; the 3 x
// (DW_PIX-1)*stride+3 input pixels are loaded once and reused across the outputs; weights (int32 [9][C]) and the
// (q, sign) -> code table sit in shared memory.  Consecutive threads take consecutive channel blocks: coalesced.
#include "tq_common.cuh"

namespace tq {

constexpr int DW_PIX = 4;

struct DwParams {
    int N, H, W, C, Ho, Wo, stride;
    float scale;
    const float *bias, *bn_a, *bn_b;
    int relu;                   // 0 none, 1 ReLU, 2 ReLU6
    int write_f32, write_codes;
    float next_sf;
    int next_bits, next_terms, next_fastdiv;
};

// activation shared by the epilogues of this file
__device__ __forceinline__ float apply_act(float t, int relu)
{
    if (relu) t = fmaxf(t, 0.0f);
    if (relu == 2) t = fminf(t, 6.0f);
    return t;
}

template <int STRIDE>
__global__ void __launch_bounds__(256)
depthwise3x3_codes_kernel(const __half *__restrict__ act, const int32_t *__restrict__ wgt, float *__restrict__ out_f32,
                          __half *__restrict__ out_codes, DwParams p)
{
    extern __shared__ __align__(16) uint8_t dw_smem[];
    int32_t *sw = reinterpret_cast<int32_t *>(dw_smem);                        // [9][C]
    __half *lut = reinterpret_cast<__half *>(dw_smem + (size_t)9 * p.C * 4);   // (q, sign) -> code
    for (int i = threadIdx.x; i < 9 * p.C; i += blockDim.x) sw[i] = wgt[i];
    if (p.write_codes) {
        for (uint32_t i = threadIdx.x; i < (2u << p.next_bits); i += blockDim.x) {
            const int code = elem_code(i & ((1u << p.next_bits) - 1u), TQ_ENC_HESE, p.next_terms);
            lut[i] = __int2half_rn((i >> p.next_bits) ? -code : code);
        }
    }
    __syncthreads();
    const Quant nq = make_quant(p.write_codes ? p.next_sf : 1.0f, (float)((1u << p.next_bits) - 1u));
    constexpr int IN_PIX = (DW_PIX - 1) * STRIDE + 3;
    const int c8n = p.C >> 3;
    const int wq_n = (p.Wo + DW_PIX - 1) / DW_PIX;
    const int64_t total = (int64_t)p.N * p.Ho * wq_n * c8n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % c8n);
        const int64_t q = t / c8n;
        const int wq = (int)(q % wq_n), ho = (int)((q / wq_n) % p.Ho), n = (int)(q / ((int64_t)wq_n * p.Ho));
        const int wo0 = wq * DW_PIX;
        const int wi0 = wo0 * STRIDE - 1, hi0 = ho * STRIDE - 1;
        int acc[DW_PIX][8];
#pragma unroll
        for (int j = 0; j < DW_PIX; ++j)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[j][c] = 0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int hi = hi0 + dy;
            if (hi < 0 || hi >= p.H) continue;
            int wrow[3][8];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int4 w0 = *reinterpret_cast<const int4 *>(sw + (dy * 3 + dx) * p.C + c8 * 8);
                const int4 w1 = *reinterpret_cast<const int4 *>(sw + (dy * 3 + dx) * p.C + c8 * 8 + 4);
                wrow[dx][0] = w0.x; wrow[dx][1] = w0.y; wrow[dx][2] = w0.z; wrow[dx][3] = w0.w;
                wrow[dx][4] = w1.x; wrow[dx][5] = w1.y; wrow[dx][6] = w1.z; wrow[dx][7] = w1.w;
            }
            const __half *rowp = act + (((int64_t)n * p.H + hi) * p.W) * p.C + c8 * 8;
#pragma unroll
            for (int px = 0; px < IN_PIX; ++px) {
                const int wi = wi0 + px;
                if (wi < 0 || wi >= p.W) continue;
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(rowp + (int64_t)wi * p.C));
                const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
                int v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = __half2int_rn(__ushort_as_half((unsigned short)(rw[c >> 1] >> (16 * (c & 1)))));
#pragma unroll
                for (int j = 0; j < DW_PIX; ++j) {
                    const int dx = px - j * STRIDE;              // compile-time after unrolling
                    if (dx < 0 || dx > 2) continue;
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[j][c] += v[c] * wrow[dx][c];
                }
            }
        }
        // epilogue: float(acc) * scale (+ bias) -> fma(BN) -> activation -> fp32 and / or term codes
        float ba[8], bb[8], bs[8];
        const int ch = c8 * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            ba[c] = p.bn_a ? __ldg(p.bn_a + ch + c) : 1.0f;
            bb[c] = p.bn_b ? __ldg(p.bn_b + ch + c) : 0.0f;
            bs[c] = p.bias ? __ldg(p.bias + ch + c) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < DW_PIX; ++j) {
            const int wo = wo0 + j;
            if (wo >= p.Wo) break;
            const int64_t o = (((int64_t)n * p.Ho + ho) * p.Wo + wo) * p.C + ch;
            float tv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float x = __fmul_rn(__int2float_rn(acc[j][c]), p.scale);
                if (p.bias) x = __fadd_rn(x, bs[c]);
                if (p.bn_a) x = __fmaf_rn(x, ba[c], bb[c]);
                tv[c] = apply_act(x, p.relu);
            }
            if (p.write_f32) {
                reinterpret_cast<float4 *>(out_f32 + o)[0] = make_float4(tv[0], tv[1], tv[2], tv[3]);
                reinterpret_cast<float4 *>(out_f32 + o)[1] = make_float4(tv[4], tv[5], tv[6], tv[7]);
            }
            if (p.write_codes) {
                uint32_t hc[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t neg = __float_as_uint(tv[c]) >> 31;
                    const uint32_t qi = p.next_fastdiv ? quantize_f32<true>(tv[c], nq) : quantize_f32<false>(tv[c], nq);
                    hc[c] = __half_as_ushort(lut[qi | (neg << p.next_bits)]);
                }
                *reinterpret_cast<uint4 *>(out_codes + o) =
                    make_uint4(hc[0] | (hc[1] << 16), hc[2] | (hc[3] << 16), hc[4] | (hc[5] << 16), hc[6] | (hc[7] << 16));
            }
        }
    }
}

// fp32 NHWC -> fma(x, a[c], b[c]) -> activation -> fp32 and / or fp16 term codes: the tail of an unwrapped first conv
// (cnn_models/__init__.py:34-36) fused with the first wrapped layer's LinearQuantize (tr_layer.py:96-99).
__global__ void __launch_bounds__(256)
bn_act_encode_kernel(const float *__restrict__ x, float *__restrict__ out_f32, __half *__restrict__ out_codes, int64_t n4, int C,
                     DwParams p)
{
    extern __shared__ __align__(16) uint8_t dw_smem[];
    __half *lut = reinterpret_cast<__half *>(dw_smem);
    if (p.write_codes) {
        for (uint32_t i = threadIdx.x; i < (2u << p.next_bits); i += blockDim.x) {
            const int code = elem_code(i & ((1u << p.next_bits) - 1u), TQ_ENC_HESE, p.next_terms);
            lut[i] = __int2half_rn((i >> p.next_bits) ? -code : code);
        }
        __syncthreads();
    }
    const Quant nq = make_quant(p.write_codes ? p.next_sf : 1.0f, (float)((1u << p.next_bits) - 1u));
    const int c4n = C >> 2;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(t % c4n);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + t);
        float tv[4] = {v.x, v.y, v.z, v.w};
        if (p.bn_a) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(p.bn_a) + c4);
            const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bn_b) + c4);
            tv[0] = __fmaf_rn(tv[0], a.x, b.x); tv[1] = __fmaf_rn(tv[1], a.y, b.y);
            tv[2] = __fmaf_rn(tv[2], a.z, b.z); tv[3] = __fmaf_rn(tv[3], a.w, b.w);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) tv[e] = apply_act(tv[e], p.relu);
        if (p.write_f32) reinterpret_cast<float4 *>(out_f32)[t] = make_float4(tv[0], tv[1], tv[2], tv[3]);
        if (p.write_codes) {
            uint32_t hc[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t neg = __float_as_uint(tv[e]) >> 31;
                const uint32_t qi = p.next_fastdiv ? quantize_f32<true>(tv[e], nq) : quantize_f32<false>(tv[e], nq);
                hc[e] = __half_as_ushort(lut[qi | (neg << p.next_bits)]);
            }
            reinterpret_cast<uint2 *>(out_codes)[t] = make_uint2(hc[0] | (hc[1] << 16), hc[2] | (hc[3] << 16));
        }
    }
}

static int fill_quant(DwParams &p, void *out_codes, float next_sf, int next_bits, int next_terms)
{
    p.write_codes = out_codes ? 1 : 0;
    p.next_sf = out_codes ? next_sf : 1.0f;
    p.next_bits = out_codes ? next_bits : 1;
    p.next_terms = next_terms;
    if (out_codes) {
        if (!(next_sf > 0.0f) || !(next_sf < INFINITY)) return fail(TQ_ERR_INVALID, "next_sf must be positive and finite");
        // codes are stored as fp16: |code| <= 2^bits must stay exactly representable
        if (next_bits < 1 || next_bits > 11 || next_terms < 0) return fail(TQ_ERR_UNSUPPORTED, "fused encode supports 1..11 bits");
    }
    p.next_fastdiv = (p.next_sf >= 9.313225746154785e-10f && p.next_sf <= 1073741824.0f) ? 1 : 0;
    return TQ_OK;
}

}  // namespace tq

using namespace tq;

extern "C" int tq_depthwise3x3_codes(const void *act_codes, const int32_t *wgt_codes, float *out_f32, void *out_codes,
                                     const float *bias, const float *bn_a, const float *bn_b, int N, int H, int W, int C,
                                     int stride, float scale, int relu, float next_sf, int next_bits, int next_terms,
                                     void *stream)
{
    if (!act_codes || !wgt_codes || (!out_f32 && !out_codes)) return fail(TQ_ERR_INVALID, "NULL pointer");
    if (N < 1 || H < 1 || W < 1 || C < 8 || C % 8) return fail(TQ_ERR_INVALID, "bad shape (C must be a multiple of 8)");
    if (stride != 1 && stride != 2) return fail(TQ_ERR_UNSUPPORTED, "depthwise 3x3: stride 1 or 2");
    if ((bn_a == nullptr) != (bn_b == nullptr)) return fail(TQ_ERR_INVALID, "bn_a and bn_b go together");
    if (relu < 0 || relu > 2) return fail(TQ_ERR_INVALID, "relu: 0 none, 1 ReLU, 2 ReLU6");
    if ((((uintptr_t)act_codes | (uintptr_t)wgt_codes | (uintptr_t)out_f32 | (uintptr_t)out_codes) & 15u) != 0)
        return fail(TQ_ERR_INVALID, "pointers must be 16-byte aligned");
    DwParams p{};
    p.N = N; p.H = H; p.W = W; p.C = C; p.stride = stride;
    p.Ho = (H + 2 - 3) / stride + 1;
    p.Wo = (W + 2 - 3) / stride + 1;
    p.scale = scale; p.bias = bias; p.bn_a = bn_a; p.bn_b = bn_b; p.relu = relu;
    p.write_f32 = out_f32 ? 1 : 0;
    int rc = fill_quant(p, out_codes, next_sf, next_bits, next_terms);
    if (rc != TQ_OK) return rc;
    const size_t smem = (size_t)9 * C * 4 + (out_codes ? (size_t)(2u << p.next_bits) * sizeof(__half) : 0);
    if (smem > 200 * 1024) return fail(TQ_ERR_UNSUPPORTED, "depthwise conv: %d channels do not fit shared memory", C);
    const int64_t total = (int64_t)N * p.Ho * ((p.Wo + DW_PIX - 1) / DW_PIX) * (C / 8);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    auto kern = stride == 1 ? depthwise3x3_codes_kernel<1> : depthwise3x3_codes_kernel<2>;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(depthwise3x3_codes_kernel)");
    }
    kern<<<(int)blocks, 256, smem, (cudaStream_t)stream>>>((const __half *)act_codes, wgt_codes, out_f32, (__half *)out_codes, p);
    count_launch();
    return check_launch("depthwise3x3_codes_kernel");
}

extern "C" int tq_bn_act_encode(const float *x, const float *bn_a, const float *bn_b, float *out_f32, void *out_codes,
                                int64_t npix, int C, int relu, float next_sf, int next_bits, int next_terms, void *stream)
{
    if (!x || (!out_f32 && !out_codes)) return fail(TQ_ERR_INVALID, "NULL pointer");
    if (npix < 0 || C < 4 || C % 4) return fail(TQ_ERR_INVALID, "bad shape (C must be a multiple of 4)");
    if ((bn_a == nullptr) != (bn_b == nullptr)) return fail(TQ_ERR_INVALID, "bn_a and bn_b go together");
    if (relu < 0 || relu > 2) return fail(TQ_ERR_INVALID, "relu: 0 none, 1 ReLU, 2 ReLU6");
    if ((((uintptr_t)x | (uintptr_t)bn_a | (uintptr_t)bn_b | (uintptr_t)out_f32 | (uintptr_t)out_codes) & 15u) != 0)
        return fail(TQ_ERR_INVALID, "pointers must be 16-byte aligned");
    if (npix == 0) return TQ_OK;
    DwParams p{};
    p.C = C; p.bn_a = bn_a; p.bn_b = bn_b; p.relu = relu; p.write_f32 = out_f32 ? 1 : 0;
    int rc = fill_quant(p, out_codes, next_sf, next_bits, next_terms);
    if (rc != TQ_OK) return rc;
    const int64_t n4 = npix * (C / 4);
    int64_t blocks = (n4 + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    const size_t smem = out_codes ? (size_t)(2u << p.next_bits) * sizeof(__half) : 0;
    bn_act_encode_kernel<<<(int)blocks, 256, smem, (cudaStream_t)stream>>>(x, out_f32, (__half *)out_codes, n4, C, p);
    count_launch();
    return check_launch("bn_act_encode_kernel");
}
