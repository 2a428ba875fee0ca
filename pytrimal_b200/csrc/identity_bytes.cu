// identity_bytes.cu -- byte-wise identity kernel for alignments with more than 126 distinct
// non-gap byte values (the bit-plane operand of identity2.cu holds 7 code planes).
#include <algorithm>

#include "tcu_internal.cuh"

namespace tcu {

// ---------------------------------------------------------------------------
// The same statistic (template.h:346-437) straight from the raw bytes, one thread per
// pair: any byte values, no packing.  trimAl's own validation admits at most 84 distinct
// symbols (isalpha or ispunct), so this is reached only through the C ABI with arbitrary
// bytes; the GPU tests also use it to tell a packing fault from an arithmetic one.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_identity_bytes(const uint8_t *__restrict__ raw,
                                                        size_t pitch, int ncol,
                                                        const int *__restrict__ kept_rows, int nk,
                                                        const uint8_t *__restrict__ col_drop,
                                                        uint8_t indet, float *__restrict__ out,
                                                        int *__restrict__ hit_out,
                                                        int *__restrict__ dst_out)
{
    const long long npairs = (long long)nk * (nk - 1) / 2;
    for (long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x; pos < npairs;
         pos += (long long)gridDim.x * blockDim.x) {
        // invert pos -> (i, j)
        int i = 0;
        {
            const double m = 2.0 * nk - 1.0;
            i = (int)((m - sqrt(m * m - 8.0 * (double)pos)) * 0.5);
            i = max(0, min(i, nk - 2));
            auto start = [&](int r) { return (long long)r * nk - (long long)r * (r + 1) / 2; };
            while (i > 0 && start(i) > pos) i--;
            while (i + 1 < nk - 1 && start(i + 1) <= pos) i++;
        }
        const long long rs = (long long)i * nk - (long long)i * (i + 1) / 2;
        const int j = i + 1 + (int)(pos - rs);
        const uint8_t *a = raw + (size_t)kept_rows[i] * pitch;
        const uint8_t *b = raw + (size_t)kept_rows[j] * pitch;
        int hit = 0, dst = 0;
        for (int k = 0; k < ncol; k++) {
            if (col_drop[k]) continue;
            const uint8_t x = a[k], y = b[k];
            const bool gx = x == '-' || x == indet, gy = y == '-' || y == indet;
            if (gx && gy) continue;
            dst++;
            hit += x == y;
        }
        out[pos] = dst == 0 ? 0.0f : __fdiv_rn((float)hit, (float)dst);
        if (hit_out) hit_out[pos] = hit;
        if (dst_out) dst_out[pos] = dst;
    }
}

cudaError_t launch_identity_bytes(const uint8_t *raw, size_t pitch, int ncol, const int *kept_rows,
                                  int nk, const uint8_t *col_drop, uint8_t indet, float *out,
                                  int *hit_out, int *dst_out, int num_sms, cudaStream_t stream)
{
    const long long npairs = (long long)nk * (nk - 1) / 2;
    if (npairs <= 0) return cudaSuccess;
    const int blocks = (int)std::min<long long>((npairs + 255) / 256, (long long)num_sms * 16);
    k_identity_bytes<<<blocks, 256, 0, stream>>>(raw, pitch, ncol, kept_rows, nk, col_drop, indet,
                                                 out, hit_out, dst_out);
    return cudaGetLastError();
}

}  // namespace tcu
