// pipe_probe2.cu -- how many warps per SM sub-partition does the K1 instruction mix need?
// Register-only loop shaped like identity2's inner body: CH independent chains of
// 5 dependent LOP3 + POPC + add, with the popcount consumed DELAY groups later.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, bool DELAY>
__global__ void __launch_bounds__(256) k(unsigned *out, unsigned seed, int iters)
{
    unsigned a[5], b[CH][5], acc[CH], ep[CH], pc[CH];
#pragma unroll
    for (int i = 0; i < 5; i++) a[i] = seed * (threadIdx.x + i + 1);
#pragma unroll
    for (int c = 0; c < CH; c++) {
        acc[c] = 0; ep[c] = 0; pc[c] = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) b[c][i] = a[i] ^ (0x9e3779b9u * (c + 1 + i));
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            unsigned d;
            asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a[0]), "r"(b[c][0]));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xF6;" : "+r"(d) : "r"(a[1]), "r"(b[c][1]));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xF6;" : "+r"(d) : "r"(a[2]), "r"(b[c][2]));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xF6;" : "+r"(d) : "r"(a[3]), "r"(b[c][3]));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x09;" : "+r"(d) : "r"(a[4]), "r"(b[c][4]));
            if (DELAY) {
                acc[c] += pc[c];
                asm volatile("popc.b32 %0, %1;" : "=r"(pc[c]) : "r"(ep[c]));
                ep[c] = d;
            } else {
                unsigned q;
                asm volatile("popc.b32 %0, %1;" : "=r"(q) : "r"(d));
                acc[c] += q;
            }
        }
#pragma unroll
        for (int i = 0; i < 5; i++) a[i] += it;   // new "A row" every group (FMA/ALU add, cheap)
    }
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += acc[c] + pc[c] + ep[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH, bool DELAY>
static void run(int ctas_per_sm, int threads, unsigned *d)
{
    const int iters = 20000, grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH, DELAY><<<grid, threads>>>(d, 3, 64);
    cudaEventRecord(e0);
    k<CH, DELAY><<<grid, threads>>>(d, 3, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bodies = (double)grid * threads * iters * CH;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("chains=%d delay=%d warps/SMSP=%4.1f : %.2f bodies/clk/SM (LOP3-bound peak 12.4), %.3e pair-col/s\n",
           CH, (int)DELAY, ctas_per_sm * threads / 32 / 4.0, bodies / (ms * 1e-3) / 148 / (clk * 1e3),
           bodies * 32 / (ms * 1e-3));
}

int main()
{
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    for (int w = 1; w <= 8; w *= 2) {          // CTAs of 128 threads per SM: 1,2,4,8 warps/SMSP
        run<4, false>(w, 128, d);
        run<4, true>(w, 128, d);
        run<8, false>(w, 128, d);
        run<8, true>(w, 128, d);
    }
    return 0;
}
