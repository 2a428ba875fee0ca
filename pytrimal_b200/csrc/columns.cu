// columns.cu -- K3 (gap counts) and K2 (spurious / overlap vector).
//
// K3 replaces simd::calculateGapVectors<V> (vendor/trimal/include/Platform/
// template.h:444-502): per column, the number of '-' bytes over kept rows.
// K2 replaces simd::calculateSpuriousVector<V> (template.h:206-318).
//
// Both are one streaming pass over the n x L byte matrix -> HBM-bound.  The
// column-count kernel reads 16 columns per thread per row (one 128-bit load,
// 512 contiguous bytes per warp, eight rows in flight per thread), compares the
// four 32-bit words byte-wise and keeps 8-bit partial sums packed in registers
// that are widened before they can reach 255 -- the same idea as the reference's
// u8 accumulators (template.h:452-487) but flushed on the number of rows actually
// counted, so masked rows cannot make a lane wrap.
#include "tcu_internal.cuh"

namespace tcu {

// 0x01 in every byte lane of x that equals the byte replicated in `pat`.
__device__ __forceinline__ uint32_t byte_eq_ones(uint32_t x, uint32_t pat)
{
    return __vcmpeq4(x, pat) & 0x01010101u;
}

// The same test with the answer in bit 7 of every byte lane (0x80 = equal), four
// instructions instead of the six of __vcmpeq4 + mask (an emulation on sm_100): a byte of
// x ^ pat is zero iff neither its top bit nor a carry out of its low seven bits + 0x7f is set.
__device__ __forceinline__ uint32_t byte_eq_flags80(uint32_t x, uint32_t pat)
{
    const uint32_t d = x ^ pat;
    const uint32_t t = (d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | d) & 0x80808080u;
}
// four 0x80 flags -> one nibble (bit k = byte k): the product lines flag k up at bit 28 + k,
// all partial products land on distinct bits
__device__ __forceinline__ uint32_t flags80_to_nibble(uint32_t f) { return (f * 0x00204081u) >> 28; }

// One CTA = 32 column groups of 16 columns (512 columns, one warp-wide 512-byte row segment)
// x CC_LANES row lanes; grid.x tiles the columns, grid.y cuts the rows into slices.  Each
// thread streams its rows CC_UNROLL at a time (that many independent 16-byte loads in
// flight), the row lanes are summed through shared memory and every CTA issues ONE atomic
// per column.  TWO selects whether a second symbol is counted in the same pass.
constexpr int CC_LANES = 8;
constexpr int CC_UNROLL = 8;

template <bool TWO>
__global__ void __launch_bounds__(32 * CC_LANES) k_column_counts(
    const uint8_t *__restrict__ raw, int nseq, int ncol, size_t pitch,
    const uint8_t *__restrict__ row_drop, uint32_t pat_a, uint32_t pat_b, int rows_per_slice,
    int *__restrict__ count_a, int *__restrict__ count_b, uint16_t *__restrict__ plane_a,
    uint16_t *__restrict__ plane_b)
{
    __shared__ int s_sum[TWO ? 2 : 1][CC_LANES][32 * 16 + 16];
    const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int col0 = (blockIdx.x * 32 + lane) * 16;
    const bool live = col0 < ncol;  // pitch is a multiple of 128: a live group is readable

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

    const int r_begin = blockIdx.y * rows_per_slice;
    const int r_end = min(nseq, r_begin + rows_per_slice);
    if (live) {
        const uint8_t *base = raw + col0;
        for (int r = r_begin + rl; r < r_end; r += CC_LANES * CC_UNROLL) {
            uint4 v[CC_UNROLL];
            bool use[CC_UNROLL];
#pragma unroll
            for (int u = 0; u < CC_UNROLL; u++) {
                const int rr = r + u * CC_LANES;
                use[u] = rr < r_end && !(row_drop && row_drop[rr]);  // warp-uniform
                // read once: streaming hint, so that the bit planes written below (and read
                // back by the row pass) stay in L2 instead of the bytes
                v[u] = use[u] ? __ldcs(reinterpret_cast<const uint4 *>(base + (size_t)rr * pitch))
                              : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < CC_UNROLL; u++) {
                if (!use[u]) continue;
                const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                uint32_t bits_a = 0, bits_b = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t fa = byte_eq_flags80(w[q], pat_a);
                    acc_a[q] += fa >> 7;
                    if (TWO) {
                        const uint32_t fb = byte_eq_flags80(w[q], pat_b);
                        acc_b[q] += fb >> 7;
                        // nibble q of the row's 16-column plane word (bit k = column 4q + k)
                        bits_a += flags80_to_nibble(fa) << (4 * q);
                        bits_b += flags80_to_nibble(fb) << (4 * q);
                    }
                }
                if (TWO && plane_a) {
                    // one 16-column word per thread: 64 contiguous bytes per warp and plane
                    const size_t o = (size_t)(r + u * CC_LANES) * (pitch >> 4) + (col0 >> 4);
                    plane_a[o] = (uint16_t)bits_a;
                    plane_b[o] = (uint16_t)bits_b;
                }
                pending++;
            }
            if (pending > 255 - CC_UNROLL) flush();
        }
    }
    flush();
#pragma unroll
    for (int i = 0; i < 16; i++) {
        s_sum[0][rl][lane * 16 + i + (lane >> 1)] = tot_a[i];  // skewed: no 16-way bank conflict
        if (TWO) s_sum[TWO ? 1 : 0][rl][lane * 16 + i + (lane >> 1)] = tot_b[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 512; c += 32 * CC_LANES) {
        const int col = blockIdx.x * 512 + c;
        if (col >= ncol) break;
        const int idx = c + (c >> 5);  // (c / 16) >> 1 == c >> 5
        int a = 0, b = 0;
#pragma unroll
        for (int l = 0; l < CC_LANES; l++) {
            a += s_sum[0][l][idx];
            if (TWO) b += s_sum[TWO ? 1 : 0][l][idx];
        }
        if (a) atomicAdd(&count_a[col], a);
        if (TWO && b) atomicAdd(&count_b[col], b);
    }
}

// count_a / count_b must be zeroed by the caller.  count_b == nullptr -> single count.
// plane_a / plane_b (optional, with count_b): one bit per cell "byte == sym", rows of pitch / 8
// bytes, bit k of a row = column k -- what the second pass of the spurious vector reads
// instead of the bytes.
cudaError_t launch_column_counts(const uint8_t *raw, int nseq, int ncol, size_t pitch,
                                 const uint8_t *row_drop, uint8_t sym_a, uint8_t sym_b,
                                 int *count_a, int *count_b, uint16_t *plane_a, uint16_t *plane_b,
                                 int num_sms, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0) return cudaSuccess;
    const int gx = (ncol + 511) / 512;
    // one resident wave: as many CTAs as the device holds at once (2-3 per SM), each with a
    // row slice of at least one unrolled pass of the lanes -- no second, partly filled wave
    int per_sm = 2;
    if (count_b)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_column_counts<true>, 32 * CC_LANES, 0);
    else
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_column_counts<false>, 32 * CC_LANES, 0);
    per_sm = max(1, per_sm);
    int gy = max(1, num_sms * per_sm / gx);
    int rows_per_slice = max(CC_LANES * CC_UNROLL, (nseq + gy - 1) / gy);
    gy = min(65535, (nseq + rows_per_slice - 1) / rows_per_slice);
    rows_per_slice = (nseq + gy - 1) / gy;
    const uint32_t pa = 0x01010101u * sym_a, pb = 0x01010101u * sym_b;
    dim3 grid(gx, gy);
    if (count_b)
        k_column_counts<true><<<grid, 32 * CC_LANES, 0, stream>>>(
            raw, nseq, ncol, pitch, row_drop, pa, pb, rows_per_slice, count_a, count_b, plane_a,
            plane_b);
    else
        k_column_counts<false><<<grid, 32 * CC_LANES, 0, stream>>>(
            raw, nseq, ncol, pitch, row_drop, pa, pb, rows_per_slice, count_a, count_b, nullptr,
            nullptr);
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
// column counts (kernel above, which also leaves the '-' / indet bit planes of
// the rows behind) then one pass per row over those planes testing
// hits >= ovrlap (template.h:301-305) and the final ratio (:309).
// ---------------------------------------------------------------------------
// One warp per 32 columns: which classes reach `ovrlap` there, as three bit words
// (residue / '-' / indet).  flag_words: 3 arrays of pitch / 32 words, fully written.
__global__ void __launch_bounds__(256) k_spurious_flags(int nseq, int ncol, int nwords,
                                                        const int *__restrict__ cnt_gap,
                                                        const int *__restrict__ cnt_indet,
                                                        uint32_t ovrlap,
                                                        uint32_t *__restrict__ flag_words)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nwords) return;
    const int k = w * 32 + lane;
    bool fr = false, fg = false, fx = false;
    if (k < ncol) {
        const int cg = cnt_gap[k], cx = cnt_indet[k];
        const int ng = nseq - cg - cx;
        // a class with zero members is never looked up; guard the unsigned compare
        fr = ng >= 1 && (uint32_t)(ng - 1) >= ovrlap;
        fg = cg >= 1 && (uint32_t)(cg - 1) >= ovrlap;
        fx = cx >= 1 && (uint32_t)(cx - 1) >= ovrlap;
    }
    const uint32_t br = __ballot_sync(0xffffffffu, fr), bg = __ballot_sync(0xffffffffu, fg);
    const uint32_t bx = __ballot_sync(0xffffffffu, fx);
    if (lane == 0) {
        flag_words[w] = br;
        flag_words[nwords + w] = bg;
        flag_words[2 * nwords + w] = bx;
    }
}

// Second pass, from the two bit planes the column-count pass left behind (1/4 of the bytes):
// one warp per row of [row_begin, row_end), grid-stride; per 32 columns
//     popc(~G & ~X & F_residue) + popc(G & F_gap) + popc(X & F_indet)
// columns pass (template.h:301-305), then the ratio (:309).  Padding columns have G = X = 0
// and F_residue = 0.
constexpr int SP_SMEM_WORDS = 1024;  // flag words kept in shared memory (32 768 columns)

template <bool SMEM>
__global__ void __launch_bounds__(256) k_spurious_rows(const uint32_t *__restrict__ plane_gap,
                                                       const uint32_t *__restrict__ plane_indet,
                                                       int row_begin, int row_end, int ncol,
                                                       int nwords,
                                                       const uint32_t *__restrict__ flag_words,
                                                       float *__restrict__ out)
{
    __shared__ uint32_t s_flags[SMEM ? 3 * SP_SMEM_WORDS : 1];
    if (SMEM) {
        for (int k = threadIdx.x; k < 3 * nwords; k += 256) s_flags[k] = flag_words[k];
        __syncthreads();
    }
    const uint32_t *fl = SMEM ? s_flags : flag_words;
    const int lane = threadIdx.x & 31;
    const int warps = gridDim.x * 8;
    for (int row = row_begin + blockIdx.x * 8 + (threadIdx.x >> 5); row < row_end; row += warps) {
        const uint32_t *g = plane_gap + (size_t)row * nwords;
        const uint32_t *x = plane_indet + (size_t)row * nwords;
        uint32_t good = 0;
        for (int k = lane; k < nwords; k += 32) {
            const uint32_t G = __ldg(g + k), X = __ldg(x + k);
            good += __popc(~G & ~X & fl[k]) + __popc(G & fl[nwords + k]) +
                    __popc(X & fl[2 * nwords + k]);
        }
        good = __reduce_add_sync(0xffffffffu, good);
        if (lane == 0) out[row] = __fdiv_rn((float)good, (float)ncol);
    }
}

// flag_words: scratch of 3 * pitch / 32 words.  The column counts cover all nseq rows; the
// planes and out[row_begin .. row_end) cover a rank's share of the rows (plane row r is at
// (r - 0) * pitch / 8 bytes: the planes are indexed by absolute row).
cudaError_t launch_spurious_rows(const uint32_t *plane_gap, const uint32_t *plane_indet, int nseq,
                                 int row_begin, int row_end, int ncol, size_t pitch,
                                 const int *cnt_gap, const int *cnt_indet, uint32_t ovrlap,
                                 uint32_t *flag_words, float *out, int num_sms, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0) return cudaSuccess;
    const int nwords = (int)(pitch >> 5);
    k_spurious_flags<<<(nwords * 32 + 255) / 256, 256, 0, stream>>>(nseq, ncol, nwords, cnt_gap,
                                                                    cnt_indet, ovrlap, flag_words);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || row_end <= row_begin) return e;
    const int rows = row_end - row_begin;
    const int grid = min((rows + 7) / 8, num_sms * 8);
    if (nwords <= SP_SMEM_WORDS)
        k_spurious_rows<true><<<grid, 256, 0, stream>>>(plane_gap, plane_indet, row_begin, row_end,
                                                        ncol, nwords, flag_words, out);
    else
        k_spurious_rows<false><<<grid, 256, 0, stream>>>(plane_gap, plane_indet, row_begin, row_end,
                                                         ncol, nwords, flag_words, out);
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

// The same count restricted to the kept columns (keep01: one byte per column up to `pitch`,
// 1 = kept, 0 = removed or padding): what Cleaner::removeAllGapsSeqsAndCols
// (source/Cleaner.cpp:1338-1370) needs to know about a row -- "all gaps" is count == 0.
__global__ void __launch_bounds__(256) k_row_residues(const uint8_t *__restrict__ raw, int nseq,
                                                      size_t pitch,
                                                      const uint8_t *__restrict__ keep01,
                                                      int *__restrict__ residues)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nseq) return;
    const uint4 *p = reinterpret_cast<const uint4 *>(raw + (size_t)row * pitch);
    const uint4 *k4 = reinterpret_cast<const uint4 *>(keep01);
    const uint32_t dash = 0x2d2d2d2du;
    int c = 0;
    for (int k = lane; k < (int)(pitch / 16); k += 32) {
        const uint4 v = __ldg(p + k), m = __ldg(k4 + k);
        c += __popc(~byte_eq_ones(v.x, dash) & m.x) + __popc(~byte_eq_ones(v.y, dash) & m.y) +
             __popc(~byte_eq_ones(v.z, dash) & m.z) + __popc(~byte_eq_ones(v.w, dash) & m.w);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) residues[row] = c;
}

cudaError_t launch_row_residues(const uint8_t *raw, int nseq, size_t pitch, const uint8_t *keep01,
                                int *residues, cudaStream_t stream)
{
    if (nseq == 0) return cudaSuccess;
    const long long threads = (long long)nseq * 32;
    k_row_residues<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(raw, nseq, pitch, keep01,
                                                                          residues);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Two independent 64-bit hashes of every row (all `pitch` bytes: the padding is zero in
// every row).  Cleaner::removeDuplicates (source/Cleaner.cpp:1489-1509) compares every pair
// of rows with std::string::operator== -- O(n^2) memcmp; equal rows have equal hashes, so
// the host only has to compare rows inside groups of equal hashes (and does compare them:
// the hashes select candidates, they do not decide).  One warp per row; every lane folds
// its 16-byte words in order, the lanes are combined with lane-dependent keys.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__global__ void __launch_bounds__(256) k_row_hashes(const uint8_t *__restrict__ raw, int nseq,
                                                    size_t pitch,
                                                    unsigned long long *__restrict__ hashes)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nseq) return;
    const uint4 *p = reinterpret_cast<const uint4 *>(raw + (size_t)row * pitch);
    unsigned long long a = 0x9e3779b97f4a7c15ull * (lane + 1), b = 0xd6e8feb86659fd93ull * (lane + 1);
    for (int k = lane; k < (int)(pitch / 16); k += 32) {
        const uint4 v = __ldg(p + k);
        const unsigned long long lo = ((unsigned long long)v.y << 32) | v.x;
        const unsigned long long hi = ((unsigned long long)v.w << 32) | v.z;
        a = mix64(a ^ lo) + hi;
        b = mix64(b + hi) ^ (lo * 0x9fb21c651e98df25ull);
    }
    a = mix64(a);
    b = mix64(b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b ^= __shfl_xor_sync(0xffffffffu, b, o) * 0x2545f4914f6cdd1dull + 1;
    }
    // the xor-shuffle reduction of b is not symmetric in its operands: take lane 0's value
    if (lane == 0) {
        hashes[2 * (size_t)row] = a;
        hashes[2 * (size_t)row + 1] = b;
    }
}

cudaError_t launch_row_hashes(const uint8_t *raw, int nseq, size_t pitch, unsigned long long *hashes,
                              cudaStream_t stream)
{
    if (nseq == 0) return cudaSuccess;
    const long long threads = (long long)nseq * 32;
    k_row_hashes<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(raw, nseq, pitch, hashes);
    return cudaGetLastError();
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
