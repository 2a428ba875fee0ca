// pipe_probe.cu -- issue-rate probe of the integer pipes K1 lives on (LOP3 / POPC / IMAD /
// IADD3) on the B200 at hand.  Not part of the library; run by tools/gpu_probe.sh.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned *out, unsigned seed, int iters)
{
    unsigned a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + i + 1); b[i] = a[i] ^ 0x9e3779b9u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) {            // LOP3 only
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
            } else if (MODE == 1) {     // POPC only
                asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
            } else if (MODE == 2) {     // IMAD only (fma pipe)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
            } else if (MODE == 3) {     // 5 LOP3 + 1 POPC + 1 IMAD (K1 v2 mix)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xF6;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xF6;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                unsigned p;
                asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(a[i]));
                asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(b[i]) : "r"(p));
            } else if (MODE == 4) {     // current K1 mix: 6 LOP3 + 2 POPC + 2 IADD
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xF6;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xF6;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                unsigned p, q, g;
                asm volatile("and.b32 %0, %1, %2;" : "=r"(g) : "r"(a[i]), "r"(b[i]));
                asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(a[i]));
                asm volatile("popc.b32 %0, %1;" : "=r"(q) : "r"(g));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(p));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(q));
            } else if (MODE == 5) {     // IADD3 only
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char *name, int ops_per_inner, unsigned *d)
{
    const int iters = 4096, grid = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(d, 3, 64);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, 3, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inner = (double)grid * 256 * iters * 8;           // thread-level inner bodies
    double warp_instr = inner / 32 * ops_per_inner;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %.3e thread-bodies/s  %.1f warp-instr/clk/SM @%d MHz(max)  bodies/clk/SM %.2f\n",
           name, ms, inner / (ms * 1e-3), warp_instr / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000,
           inner / (ms * 1e-3) / 148 / (clk * 1e3));
}

int main()
{
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("LOP3", 1, d);
    run<1>("POPC", 1, d);
    run<2>("IMAD", 1, d);
    run<5>("IADD", 1, d);
    run<3>("5LOP3+POPC+IMAD (v2 mix)", 7, d);
    run<4>("6LOP3+2POPC+2IADD (v1 mix)", 10, d);
    return 0;
}
