// tools/fdiv_check.cu -- exhaustive check of div_small_counts (tcu_internal.cuh) against the
// correctly rounded division, for EVERY pair of counts the packed-counter identity kernel can
// produce: 0 <= h <= d, 0 < d < 65536 (2.1e9 pairs), and a second sweep over the denominators
// up to 2^20 with strided numerators (the bound of the argument, not used by the kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -I pytrimal_b200/csrc \
//        tools/fdiv_check.cu -o /tmp/fdiv_check && /tmp/fdiv_check
#include <cstdio>

#include "tcu_internal.cuh"

__global__ void check_all(unsigned dmax, unsigned hstep, unsigned long long *bad, unsigned long long *seen)
{
    unsigned long long mism = 0, cnt = 0;
    for (unsigned d = 1 + blockIdx.x; d < dmax; d += gridDim.x)
        for (unsigned h = threadIdx.x * hstep; h <= d; h += blockDim.x * hstep) {
            const float a = tcu::div_small_counts((float)h, (float)d);
            const float b = __fdiv_rn((float)h, (float)d);
            mism += __float_as_uint(a) != __float_as_uint(b);
            cnt++;
        }
    atomicAdd(bad, mism);
    atomicAdd(seen, cnt);
}

int main()
{
    unsigned long long *d_v, h_v[2];
    cudaMalloc(&d_v, 16);
    int rc = 0;
    const unsigned sweeps[2][2] = {{65536u, 1u}, {1u << 20, 61u}};
    for (auto &s : sweeps) {
        cudaMemset(d_v, 0, 16);
        check_all<<<148 * 8, 256>>>(s[0], s[1], d_v, d_v + 1);
        if (cudaDeviceSynchronize() != cudaSuccess) return 2;
        cudaMemcpy(h_v, d_v, 16, cudaMemcpyDeviceToHost);
        printf("{\"d_below\": %u, \"h_step\": %u, \"pairs\": %llu, \"mismatches_vs_fdiv_rn\": %llu}\n", s[0],
               s[1], h_v[1], h_v[0]);
        if (s[1] == 1 && h_v[0]) rc = 1;
    }
    return rc;
}
