// columns.cu -- K3 (gap counts) and K2 (spurious / overlap vector).
//
// K3 replaces simd::calculateGapVectors<V> (vendor/trimal/include/Platform/
// template.h:444-502): per column, the number of '-' bytes over kept rows.
// K2 replaces simd::calculateSpuriousVector<V> (template.h:206-318).
//
// Both are one streaming pass over the n x L byte matrix -> HBM-bound.  The
// column-count kernel reads 16 columns per thread per row (one 128-bit load,
// 512 contiguous bytes per warp), compares the four 32-bit words byte-wise
// and keeps 8-bit partial sums packed in registers that are widened every
// 255 rows -- the same idea as the reference's u8 accumulators
// (template.h:452-487) but flushed on the number of rows actually counted,
// so masked rows cannot make a lane wrap.
#include "tcu_internal.cuh"

namespace tcu {

// 0x01 in every byte lane of x that equals the byte replicated in `pat`.
__device__ __forceinline__ uint32_t byte_eq_ones(uint32_t x, uint32_t pat)
{
    return __vcmpeq4(x, pat) & 0x01010101u;
}

// grid.x covers column groups of 16 (blockDim.x threads each), grid.y strides
// over rows.  TWO selects whether a second symbol is counted in the same pass.
template <bool TWO>
__global__ void __launch_bounds__(128) k_column_counts(const uint8_t *__restrict__ raw, int nseq,
                                                       int ncol, size_t pitch,
                                                       const uint8_t *__restrict__ row_drop,
                                                       uint32_t pat_a, uint32_t pat_b,
                                                       int *__restrict__ count_a,
                                                       int *__restrict__ count_b)
{
    const int group = blockIdx.x * blockDim.x + threadIdx.x;  // 16 columns
    const int col0 = group * 16;
    if (col0 >= ncol) return;

    uint32_t acc_a[4] = {0, 0, 0, 0}, acc_b[4] = {0, 0, 0, 0};  // packed u8 partial sums
    int tot_a[16], tot_b[16];
#pragma unroll
    for (int i = 0; i < 16; i++) tot_a[i] = tot_b[i] = 0;
    int pending = 0;

    auto flush = [&]() {
#pragma unroll
        for (int w = 0; w < 4; w++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                tot_a[w * 4 + b] += (acc_a[w] >> (8 * b)) & 0xFF;
                if (TWO) tot_b[w * 4 + b] += (acc_b[w] >> (8 * b)) & 0xFF;
            }
            acc_a[w] = 0;
            acc_b[w] = 0;
        }
        pending = 0;
    };

    for (int r = blockIdx.y; r < nseq; r += gridDim.y) {
        if (row_drop && row_drop[r]) continue;  // warp-uniform
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(raw + (size_t)r * pitch + col0));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            acc_a[q] += byte_eq_ones(w[q], pat_a);
            if (TWO) acc_b[q] += byte_eq_ones(w[q], pat_b);
        }
        if (++pending == 255) flush();
    }
    flush();

#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (col0 + i < ncol) {
            if (tot_a[i]) atomicAdd(&count_a[col0 + i], tot_a[i]);
            if (TWO && tot_b[i]) atomicAdd(&count_b[col0 + i], tot_b[i]);
        }
    }
}

// count_a / count_b must be zeroed by the caller.  sym_b == sym_a -> single count.
cudaError_t launch_column_counts(const uint8_t *raw, int nseq, int ncol, size_t pitch,
                                 const uint8_t *row_drop, uint8_t sym_a, uint8_t sym_b,
                                 int *count_a, int *count_b, int num_sms, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0) return cudaSuccess;
    const int groups = (ncol + 15) / 16;
    const int gx = (groups + 127) / 128;
    // enough row slices to fill the machine (8 CTAs of 128 threads per SM)
    int gy = max(1, min(nseq, (num_sms * 8 + gx - 1) / gx));
    gy = min(gy, 65535);
    const uint32_t pa = 0x01010101u * sym_a, pb = 0x01010101u * sym_b;
    dim3 grid(gx, gy);
    if (count_b)
        k_column_counts<true><<<grid, 128, 0, stream>>>(raw, nseq, ncol, pitch, row_drop, pa, pb,
                                                        count_a, count_b);
    else
        k_column_counts<false><<<grid, 128, 0, stream>>>(raw, nseq, ncol, pitch, row_drop, pa, pb,
                                                         count_a, count_b);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Spurious vector.  The reference counts, for row i and column k, the rows
// j != i with  byte_i == byte_j  or  (both outside the gap class)
// (template.h:264-271, 280-284).  Summed over j this only depends on the
// column's composition: with cg/cx the number of '-' / indet bytes in the
// column and ng = n - cg - cx,
//     hits(i,k) = ng - 1   if byte_i is a residue
//               = cg - 1   if byte_i == '-'
//               = cx - 1   if byte_i == indet
// an integer identity, so the O(n^2 L) loop collapses to two streaming passes:
// column counts (kernel above) then one pass per row testing
// hits >= ovrlap (template.h:301-305) and the final ratio (:309).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_spurious_flags(int nseq, int ncol,
                                                        const int *__restrict__ cnt_gap,
                                                        const int *__restrict__ cnt_indet,
                                                        uint32_t ovrlap,
                                                        uint8_t *__restrict__ col_flags)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncol) return;
    const int cg = cnt_gap[k], cx = cnt_indet[k];
    const int ng = nseq - cg - cx;
    uint8_t f = 0;
    // a class with zero members is never looked up; guard the unsigned compare
    if (ng >= 1 && (uint32_t)(ng - 1) >= ovrlap) f |= 1;
    if (cg >= 1 && (uint32_t)(cg - 1) >= ovrlap) f |= 2;
    if (cx >= 1 && (uint32_t)(cx - 1) >= ovrlap) f |= 4;
    col_flags[k] = f;
}

// one warp per row of [row_begin, row_end)
__global__ void __launch_bounds__(256) k_spurious_rows(const uint8_t *__restrict__ raw,
                                                       int row_begin, int row_end, int ncol,
                                                       size_t pitch, uint8_t indet,
                                                       const uint8_t *__restrict__ col_flags,
                                                       float *__restrict__ out)
{
    const int warp = row_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= row_end) return;
    const uint8_t *row = raw + (size_t)warp * pitch;
    uint32_t good = 0;
    // 4 columns per lane per step: 128 contiguous bytes per warp
    for (int k = lane * 4; k < ncol; k += 128) {
        const uint32_t v = *reinterpret_cast<const uint32_t *>(row + k);
        const uint32_t f = *reinterpret_cast<const uint32_t *>(col_flags + k);
#pragma unroll
        for (int b = 0; b < 4; b++) {
            if (k + b < ncol) {
                const uint32_t c = (v >> (8 * b)) & 0xFF;
                const uint32_t fl = (f >> (8 * b)) & 0xFF;
                const uint32_t bit = c == '-' ? 2u : (c == indet ? 4u : 1u);
                good += (fl & bit) != 0;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) good += __shfl_xor_sync(0xffffffffu, good, o);
    if (lane == 0) out[warp] = __fdiv_rn((float)good, (float)ncol);
}

// col_flags: scratch of at least roundup(ncol, 4) bytes.  The column counts cover all
// nseq rows; out[row_begin .. row_end) is written (a rank's share of the rows).
cudaError_t launch_spurious_rows(const uint8_t *raw, int nseq, int row_begin, int row_end,
                                 int ncol, size_t pitch, uint8_t indet, const int *cnt_gap,
                                 const int *cnt_indet, uint32_t ovrlap, uint8_t *col_flags,
                                 float *out, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0) return cudaSuccess;
    k_spurious_flags<<<(ncol + 255) / 256, 256, 0, stream>>>(nseq, ncol, cnt_gap, cnt_indet,
                                                             ovrlap, col_flags);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || row_end <= row_begin) return e;
    const long long threads = (long long)(row_end - row_begin) * 32;
    k_spurious_rows<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
        raw, row_begin, row_end, ncol, pitch, indet, col_flags, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Sequence lengths: Alignment::getSequenceLength (source/Alignment/Alignment.cpp:296-298)
// for every row = ncol - number of '-' bytes.  One warp per row, 16 bytes per lane per
// step; the zero padding up to `pitch` never matches '-'.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_lengths(const uint8_t *__restrict__ raw, int nseq,
                                                     int ncol, size_t pitch,
                                                     int *__restrict__ lengths)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nseq) return;
    const uint4 *p = reinterpret_cast<const uint4 *>(raw + (size_t)row * pitch);
    const uint32_t dash = 0x2d2d2d2du;
    int c = 0;
    for (int k = lane; k < (int)(pitch / 16); k += 32) {
        const uint4 v = __ldg(p + k);
        c += __popc(byte_eq_ones(v.x, dash)) + __popc(byte_eq_ones(v.y, dash)) +
             __popc(byte_eq_ones(v.z, dash)) + __popc(byte_eq_ones(v.w, dash));
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) lengths[row] = ncol - c;
}

cudaError_t launch_row_lengths(const uint8_t *raw, int nseq, int ncol, size_t pitch, int *lengths,
                               cudaStream_t stream)
{
    if (nseq == 0) return cudaSuccess;
    const long long threads = (long long)nseq * 32;
    k_row_lengths<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(raw, nseq, ncol, pitch,
                                                                         lengths);
    return cudaGetLastError();
}

}  // namespace tcu
