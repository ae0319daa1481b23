// umma_issue_probe.cu -- how fast does one SM retire tcgen05.mma kind::f16 (M=128, K=16, SS operands)
// as a function of N and of how many independent accumulators the instruction stream rotates over?
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_issue_probe tools/umma_issue_probe.cu
// Prints cycles per MMA.  (Design aid for csrc/tq_gemm.cu; not part of the library.)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int N, int NACC, int PER_COMMIT>
__global__ void __launch_bounds__(128, 1) probe(int total, long long *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u;   // fp16 ones
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t base = smem_u32(smem);
        long long t0 = clock64();
        const uint64_t da0 = desc_sw128(base), db0 = desc_sw128(base + 16384);
        for (int it = 0; it < total; it += 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {            // 4 stages x 4 k-steps, all offsets compile-time
                const uint64_t da = da0 + (uint64_t)((j / 4) * (32768 >> 4) + 2 * (j & 3));
                const uint64_t db = db0 + (uint64_t)((j / 4) * (32768 >> 4) + 2 * (j & 3));
                const uint32_t d = tmem + (uint32_t)((j % NACC) * N);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(IDESC), "r"((it | (j >= NACC)) ? 1u : 0u) : "memory");
                if ((j + 1) % PER_COMMIT == 0)
                    // a commit nobody waits for (arrives on a scratch barrier), as the per-stage "stage free" signal
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        {   // wait for the final commit
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, int NACC, int PER_COMMIT>
static void run()
{
    const int n_acc = NACC, per_commit = PER_COMMIT, stages = 4;
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe<N, NACC, PER_COMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int total = 2048;
    for (int rep = 0; rep < 2; ++rep) probe<N, NACC, PER_COMMIT><<<1, 128, 198 * 1024>>>(total, d);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) printf("launch: %s\n", cudaGetErrorString(le));
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d accumulators=%d mma/commit=%2d stages=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)  %s\n", N, n_acc,
           per_commit, stages, (double)h[0] / total, (double)h[1] / total, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main()
{
    run<64, 1, 4>(); run<64, 2, 4>(); run<64, 4, 4>(); run<64, 8, 4>();
    run<128, 1, 4>(); run<128, 2, 4>(); run<128, 4, 4>();
    run<256, 1, 4>(); run<256, 2, 4>();
    run<64, 1, 16>(); run<64, 4, 16>(); run<128, 1, 16>(); run<128, 4, 16>(); run<256, 1, 16>();
    run<64, 1, 1>(); run<128, 1, 1>();
    return 0;
}
