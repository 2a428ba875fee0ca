// umma_i8_peak.cu -- measured dense int8 tensor-core peak of the B200 at hand (the roofline
// SURVEY 8(d) / BASELINE.md 3.6 name for the pairwise-identity kernel).  Not part of the
// library.  One persistent CTA per SM (or a CTA pair for cta_group::2) issues back-to-back
// tcgen05.mma kind::i8 instructions on operands that stay in shared memory, accumulating
// into TMEM; nothing is loaded or stored inside the timed region, so the figure is the
// issue-limited peak of the tensor pipe, not a GEMM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_i8_peak tools/umma_i8_peak.cu
//   tools/umma_i8_peak            -> one JSON line per shape
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// K-major, no swizzle (8 x 16-byte core matrices): LBO between the two 16-byte K halves,
// SBO between 8-row groups; descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// D = s32, A = B = u8, K-major both, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc(int M, int N)
{
    return (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int M, int N>
__global__ void __launch_bounds__(128) k_peak(int iters, unsigned long long *cycles)
{
    // A: M rows x 32 bytes, B: N rows x 32 bytes, canonical K-major layout:
    // [row group of 8][k half][row 8][16 bytes] -> SBO = 256, LBO = 128
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    uint8_t *sa = smem, *sb = smem + M * 32;
    for (int i = threadIdx.x; i < (M + N) * 32; i += blockDim.x) smem[i] = (uint8_t)(i & 1);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                         smem_u32(&s_tmem))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint64_t da = umma_desc(smem_u32(sa), 128, 256), db = umma_desc(smem_u32(sb), 128, 256);
        const unsigned long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            // two accumulators (N columns each) alternate, every MMA accumulates
            const uint32_t d = tmem + (uint32_t)((it & 1) * N);
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d),
                "l"(da), "l"(db), "r"(idesc(M, N)), "r"(it > 1 ? 1u : 0u), "r"(0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&bar))
                     : "memory");
        asm volatile(
            "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n"
            "@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar))
            : "memory");
        if (cycles) cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

template <int M, int N>
static void run(int sms, int iters)
{
    const size_t smem = (size_t)(M + N) * 32;
    cudaFuncSetAttribute(k_peak<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned long long *d_cyc;
    cudaMalloc(&d_cyc, sms * sizeof(unsigned long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_peak<M, N><<<sms, 128, smem>>>(1024, nullptr);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_peak<M, N><<<sms, 128, smem>>>(iters, d_cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // sustained: the same launch back to back for about two seconds
    float sus_ms = 0.f;
    int launches = 0;
    cudaEventRecord(e0);
    for (; launches < (int)(2000.f / best) + 1; launches++) k_peak<M, N><<<sms, 128, smem>>>(iters, d_cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&sus_ms, e0, e1);
    unsigned long long cyc0 = 0;
    cudaMemcpy(&cyc0, d_cyc, sizeof cyc0, cudaMemcpyDeviceToHost);
    const double ops = 2.0 * M * N * 32 * (double)iters * sms;
    cudaError_t err = cudaGetLastError();
    printf("{\"probe\": \"tcgen05.mma kind::i8 cta_group::1\", \"M\": %d, \"N\": %d, \"K\": 32, \"ctas\": %d, "
           "\"mma_per_cta\": %d, \"burst_ms\": %.4f, \"burst_tops\": %.1f, \"sustained_tops\": %.1f, "
           "\"sm_cycles_per_mma\": %.1f, \"status\": \"%s\"}\n",
           M, N, sms, iters, best, ops / (best * 1e-3) / 1e12,
           ops * launches / (sus_ms * 1e-3) / 1e12, (double)cyc0 / iters, cudaGetErrorString(err));
    cudaFree(d_cyc);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 1 << 15;
    run<128, 256>(sms, iters);
    run<128, 128>(sms, iters);
    run<128, 64>(sms, iters);   // the shape k_identity2 issues (both-gap counts)
    run<64, 256>(sms, iters);
    return 0;
}
