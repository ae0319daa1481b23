// umma_stage_probe.cu -- what does the producer / MMA-issuer hand-shake of csrc/tq_gemm.cu cost per stage?
// One CTA of 640 threads like the conv kernel: warp 0 lane 0 = "producer" (waits empty[s], arrives full[s]: no data moves),
// warp 1 lane 0 = MMA issuer (waits full[s], issues MMAS_PER_STAGE tcgen05.mma N=128 K=16 with run-time descriptors, commits
// empty[s]), warps 4-19 optionally spin on an mbarrier like idle epilogue warps.  Prints cycles per MMA.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_stage_probe tools/umma_stage_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

template <int MMAS, int STAGES, bool SPIN, int AROW>
__global__ void __launch_bounds__(640, 1) probe(int stages_total, long long *out, int data)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[8], empty[8], done, idle;
    __shared__ uint32_t slot;
    // operand data: 1.0 everywhere, or (data = 1) pseudo-random integer codes in [-255, 255] like the conv's operands --
    // does the rate depend on what the tensor cores multiply?
    for (int i = threadIdx.x; i < 160 * 1024 / 2; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        ((__half *)smem)[i] = data ? __int2half_rn((int)(h % 511u) - 255) : __float2half(1.0f);
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[i])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&idle)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 96) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < stages_total; ++i) {
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_arrive(&full[s]);
            if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1 && lane == 0) {
        constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint64_t DESC_HI = ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t a_lo0 = ((smem_u32(smem) & 0x3FFFFu) >> 4) | (1u << 16);
        uint32_t a_lo = a_lo0;
        int s = 0; uint32_t ph = 0;
        long long t0 = clock64();
        for (int i = 0; i < stages_total; ++i) {
            mbar_wait(&full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // AROW: the A operand starts AROW 128-byte rows into the swizzled buffer (the halo modes' row-offset views)
            const uint64_t da = DESC_HI | (a_lo + (uint32_t)(AROW * 8)), db = DESC_HI | (a_lo + (20480u >> 4));
#pragma unroll
            for (int k = 0; k < MMAS; ++k)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da + (uint64_t)(2 * (k & 3))), "l"(db + (uint64_t)(2 * (k & 3))), "r"(IDESC), "r"((uint32_t)(i | k)) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
            a_lo += 40960u >> 4;
            if (++s == STAGES) { s = 0; ph ^= 1u; a_lo = a_lo0; }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
        mbar_wait(&done, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
        mbar_arrive(&idle);
    } else if (warp >= 4 && SPIN) {
        mbar_wait(&idle, 0);                                  // idle epilogue warps: spin on an mbarrier until the end
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x >= 64 && threadIdx.x < 96) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// grid = 1: one SM alone.  grid = 148: every SM issues at the same time -- does the chip sustain the single-SM rate?
// (cycles from clock64 of the slowest CTA, and the rate from the wall time of the launch: they differ if the SM clock
// drops under the load)
template <int MMAS, int STAGES, bool SPIN, int AROW = 0>
static void run(int grid = 1, int mult = 1, int data = 0)
{
    long long *d, h[148];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(probe<MMAS, STAGES, SPIN, AROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    const int stages_total = 4096 / MMAS * 4 * mult;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MMAS, STAGES, SPIN, AROW><<<grid, 640, 168 * 1024>>>(stages_total, d, data);
    cudaEventRecord(e0);
    probe<MMAS, STAGES, SPIN, AROW><<<grid, 640, 168 * 1024>>>(stages_total, d, data);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long worst = 0;
    for (int i = 0; i < grid; ++i) worst = h[i] > worst ? h[i] : worst;
    const double mmas = (double)stages_total * MMAS;
    printf("N=128: %3d CTAs, %s, %2d MMAs per stage, %d ring stages, A row offset %2d, idle warps %s: %.1f cycles per MMA, %.0f per stage; "
           "%.1f ns per MMA by wall time = %.0f TFLOP/s  %s\n", grid, data ? "random codes" : "all ones    ", MMAS, STAGES, AROW, SPIN ? "spinning" : "parked ",
           (double)worst / mmas, (double)worst / stages_total, ms * 1e6 / mmas, grid * mmas * 2.0 * 128 * 128 * 16 / (ms * 1e-3) / 1e12,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main()
{
    run<4, 4, false>(); run<4, 4, true>(); run<8, 4, false>(); run<8, 4, true>(); run<16, 4, true>(); run<4, 2, true>();
    run<8, 4, true>(148); run<16, 4, true>(148); run<4, 4, true>(148); run<16, 4, true>(148, 16); run<16, 4, true>(1, 16);
    run<4, 4, true>(1, 1, 1); run<8, 4, true>(1, 1, 1); run<16, 4, true>(1, 1, 1);
    run<4, 4, true>(148, 1, 1); run<8, 4, true>(148, 1, 1); run<16, 4, true>(148, 1, 1); run<16, 4, true>(148, 16, 1);
    run<4, 4, true, 1>(); run<4, 4, true, 3>(); run<4, 4, true, 31>(); run<8, 4, true, 1>(); run<8, 4, true, 31>(); run<16, 4, true, 31>();
    return 0;
}
