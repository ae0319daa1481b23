// ref_gpu_shim.cu -- device harness around the REFERENCE kernel body ("the kernel to beat").
//
// TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Like ref_shim.cpp this file contains none of the
// reference's code: oracle/Makefile extracts kernels/tr_cuda_kernel.cu:13-126 (hese_encode +
// tr_cuda_kernel, byte-identical) into a temporary file at build time and this harness includes it
// unchanged, with the reference's own constants (kernels/tr_cuda_kernel.cu:9-11: MAX_GROUP_SIZE 32,
// MAX_TERMS 64 -- on the GPU the scan's shifts >= 32 clamp, so no deviation is needed here).
// The launcher below repeats what the reference's ATen launcher does (:142-147): output =
// zeros_like(input), 128 threads per block, ceil(B*C*W*H / 128) blocks -- only the ATen calls that
// no longer compile against torch 2.x are replaced by plain CUDA runtime calls.
// Output: oracle/_ref/libtq_ref_gpu.so (git-ignored).  Used by tests/ (GPU-vs-GPU parity at full
// BASELINE sizes) and tools/microbench.py (the reference kernel's GB/s on the same B200).
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef TQ_REF_BODY
#error "build through oracle/Makefile (TQ_REF_BODY = extracted reference kernel body)"
#endif
#include TQ_REF_BODY

template <typename T>
static int launch(const T *in, T *out, float sf, int bits, int g, int keep, int B, int C, int W, int H, cudaStream_t s)
{
    const long long size = (long long)B * C * W * H;
    if (size <= 0 || size >= (1ll << 31) || g < 1 || g > MAX_GROUP_SIZE || C % g != 0) return -1;   // fenced: SURVEY 8a-3
    const int threads = 128;
    const int blocks = (int)((size + threads - 1) / threads);
    if (cudaMemsetAsync(out, 0, sizeof(T) * (size_t)size, s) != cudaSuccess) return -2;             // at::zeros_like
    tr_cuda_kernel<T><<<blocks, threads, 0, s>>>(in, out, sf, bits, g, keep, B, C, W, H);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int tq_refgpu_tr_f32(const float *in, float *out, float sf, int bits, int g, int keep,
                                int B, int C, int W, int H, void *stream)
{
    return launch<float>(in, out, sf, bits, g, keep, B, C, W, H, (cudaStream_t)stream);
}

extern "C" int tq_refgpu_tr_f64(const double *in, double *out, float sf, int bits, int g, int keep,
                                int B, int C, int W, int H, void *stream)
{
    return launch<double>(in, out, sf, bits, g, keep, B, C, W, H, (cudaStream_t)stream);
}
