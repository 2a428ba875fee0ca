// identity.cu -- K1: pairwise sequence identity on packed bit-planes.
//
// Replaces simd::calculateSeqIdentity<V> (vendor/trimal/include/Platform/
// template.h:320-442).  Per kept pair i<j:
//     dst = #{kept columns : not (gap_i and gap_j)}      (template.h:422)
//     hit = #{those columns : byte_i == byte_j}          (template.h:423-424)
//     identity = dst ? (float)hit / (float)dst : 0       (template.h:427-434)
// written at the packed upper-triangular offset of (i, j) (template.h:436).
//
// Formulation: the pair matrix is tiled into RB x RB row-block tiles; a
// persistent CTA walks its tiles, streaming the two row-blocks' column chunks
// through a ring of shared-memory stages filled by 1-D bulk async copies
// (cp.async.bulk / TMA unit) and signalled through mbarriers.  Thread 0 issues
// the copies a few chunks ahead; the eight warps each own a 16 x 32 patch of the
// tile, every thread a 4 x 4 register tile of (hit, bothgap) counters.  For a
// 32-column word of one pair the work is NP LOP3 + 1 AND + 2 POPC + 2 IADD:
//     differ  = (a.p0 ^ (b.p0 | b.g)) | (a.p1 ^ b.p1) | ... | (a.pN ^ b.pN)
//     hit    += popc(~differ)          bothgap += popc(a.g & b.g)
// and dst = total_bits - bothgap because masked-out / padding columns are
// stored as gaps in every row.  Everything is integer until the single IEEE
// fp32 division of the epilogue, so results are bit-identical to the
// reference for any input bytes.
#include <algorithm>

#include "tcu_internal.cuh"

namespace tcu {

constexpr int ID_WARPS = 8;
constexpr int ID_THREADS = ID_WARPS * 32;

__host__ __device__ constexpr int identity_stages(int np) { return words_stored(np) <= 6 ? 4 : 3; }
__host__ __device__ constexpr size_t identity_smem_bytes(int np)
{
    return (size_t)identity_stages(np) * 2 * tile_bytes(np) + 2 * identity_stages(np) * sizeof(uint64_t);
}

// Linear index over the upper-triangular (incl. diagonal) block matrix, row
// by row: T(b) = b*nb - b*(b-1)/2 tiles precede block-row b.
__device__ __forceinline__ void tile_to_blocks(long long t, int nb, int &bi, int &bj)
{
    const double m = 2.0 * nb + 1.0;
    int b = (int)((m - sqrt(m * m - 8.0 * (double)t)) * 0.5);
    b = max(0, min(b, nb - 1));
    while (b > 0 && (long long)b * nb - (long long)b * (b - 1) / 2 > t) b--;
    while (b + 1 < nb && (long long)(b + 1) * nb - (long long)(b + 1) * b / 2 <= t) b++;
    bi = b;
    bj = b + (int)(t - ((long long)b * nb - (long long)b * (b - 1) / 2));
}

template <int NP>
struct RowWords {
    uint32_t w[NP + 1];  // w[0] = g, w[1 + p] = plane p
};

template <int NP>
__device__ __forceinline__ void load_row(const uint32_t *tile_kw, int row, RowWords<NP> &o)
{
    constexpr int W = NP + 1;
    constexpr int G1 = group1_words(NP);
    const uint4 v0 = *reinterpret_cast<const uint4 *>(tile_kw + row * 4);
    o.w[0] = v0.x;
    o.w[1] = v0.y;
    o.w[2] = v0.z;
    o.w[3] = v0.w;
    const uint32_t *g1 = tile_kw + RB * 4 + row * G1;
    if (G1 == 1) {
        o.w[4] = g1[0];
    } else if (G1 == 2) {
        const uint2 v1 = *reinterpret_cast<const uint2 *>(g1);
        o.w[4] = v1.x;
        o.w[5] = v1.y;
    } else if (G1 == 4) {
        const uint4 v1 = *reinterpret_cast<const uint4 *>(g1);
        o.w[4] = v1.x;
        o.w[5] = v1.y;
        o.w[6] = v1.z;
        if (W == 8) o.w[W - 1] = v1.w;
    }
}

template <int NP>
__global__ void __launch_bounds__(ID_THREADS, identity_stages(NP) == 4 ? 2 : 1)
    k_identity(const IdentityParams p)
{
    constexpr int WS = words_stored(NP);
    constexpr int TW = tile_words(NP);
    constexpr int TB = tile_bytes(NP);
    constexpr int STAGES = identity_stages(NP);
    constexpr int AHEAD = STAGES >= 4 ? STAGES - 2 : STAGES - 1;  // chunks in flight

    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *tiles = reinterpret_cast<uint32_t *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * 2 * TB);
    uint64_t *empty = full + STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], ID_WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    // ---- producer state: thread 0 runs AHEAD chunks in front of the math ----
    // (no dedicated producer warp: 8 warps x 2 CTAs fill the four schedulers
    // evenly and leave 128 registers per thread)
    const long long ntiles_all = p.tile_end - p.tile_begin;
    const long long my_tiles =
        ntiles_all > blockIdx.x ? (ntiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    long long to_issue = threadIdx.x == 0 ? my_tiles * p.nchunks : 0;
    long long p_tile = p.tile_begin + blockIdx.x;
    int p_chunk = 0, p_stage = 0;
    uint32_t p_phase = 0;
    const uint8_t *p_src_i = nullptr, *p_src_j = nullptr;
    auto issue_one = [&]() {
        if (to_issue == 0) return;
        if (p_chunk == 0) {
            int bi, bj;
            tile_to_blocks(p_tile, p.nb, bi, bj);
            const size_t block_bytes = (size_t)p.nchunks * TB;
            p_src_i = reinterpret_cast<const uint8_t *>(p.planes) + (size_t)bi * block_bytes;
            p_src_j = reinterpret_cast<const uint8_t *>(p.planes) + (size_t)bj * block_bytes;
        }
        mbar_wait(&empty[p_stage], p_phase ^ 1u);
        mbar_arrive_expect_tx(&full[p_stage], 2u * TB);
        uint32_t *dst = tiles + (size_t)p_stage * 2 * TW;
        bulk_copy_g2s(dst, p_src_i + (size_t)p_chunk * TB, TB, &full[p_stage]);
        bulk_copy_g2s(dst + TW, p_src_j + (size_t)p_chunk * TB, TB, &full[p_stage]);
        if (++p_chunk == p.nchunks) {
            p_chunk = 0;
            p_tile += gridDim.x;
        }
        if (++p_stage == STAGES) {
            p_stage = 0;
            p_phase ^= 1u;
        }
        to_issue--;
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < AHEAD; i++) issue_one();
    }

    // -------------------- math: 8 warps as 4 (I) x 2 (J) --------------------
    const int wi = warp >> 1, wj = warp & 1;
    const int li = lane & 3, lj = lane >> 2;
    const int row_i0 = wi * 16 + li;  // + 4*t
    const int row_j0 = wj * 32 + lj;  // + 8*u

    int stage = 0;
    uint32_t phase = 0;
    for (long long t = p.tile_begin + blockIdx.x; t < p.tile_end; t += gridDim.x) {
        int bi, bj;
        tile_to_blocks(t, p.nb, bi, bj);

        uint32_t hit[4][4], both[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) hit[a][b] = both[a][b] = 0;

        for (int c = 0; c < p.nchunks; c++) {
            if (threadIdx.x == 0) issue_one();
            mbar_wait(&full[stage], phase);
            const uint32_t *s_i = tiles + (size_t)stage * 2 * TW;
            const uint32_t *s_j = s_i + TW;
#pragma unroll 2
            for (int kw = 0; kw < KC; kw++) {
                RowWords<NP> A[4], B[4];
#pragma unroll
                for (int a = 0; a < 4; a++) load_row<NP>(s_i + kw * RB * WS, row_i0 + 4 * a, A[a]);
#pragma unroll
                for (int b = 0; b < 4; b++) load_row<NP>(s_j + kw * RB * WS, row_j0 + 8 * b, B[b]);
#pragma unroll
                for (int a = 0; a < 4; a++) {
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        // differ = a.p0 ^ (b.p0 | b.g)
                        uint32_t d = lop3<0x1E>(A[a].w[1], B[b].w[1], B[b].w[0]);
#pragma unroll
                        for (int q = 1; q < NP - 1; q++)  // differ |= a.pq ^ b.pq
                            d = lop3<0xF6>(d, A[a].w[1 + q], B[b].w[1 + q]);
                        // equal = ~(differ | (a.pl ^ b.pl)), last plane
                        const uint32_t e = lop3<0x09>(d, A[a].w[NP], B[b].w[NP]);
                        hit[a][b] += __popc(e);
                        both[a][b] += __popc(A[a].w[0] & B[b].w[0]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
            }
        }

        // ------------------------------ epilogue ----------------------------
        const unsigned long long n = (unsigned long long)p.nk;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int i = bi * RB + row_i0 + 4 * a;
            if (i >= p.nk) continue;
            // offset of pair (i, i+1): i*n - i*(i+1)/2 - i - 1 + (i+1)
            const unsigned long long row_base =
                (unsigned long long)i * n - ((unsigned long long)i * (i + 1)) / 2 - i - 1;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int j = bj * RB + row_j0 + 8 * b;
                if (j >= p.nk || j <= i) continue;
                const unsigned long long pos = row_base + j;
                const int h = (int)hit[a][b];
                const int d = p.total_bits - (int)both[a][b];
                const float v = d == 0 ? 0.0f : __fdiv_rn((float)h, (float)d);
                p.out[pos - p.out_base] = v;
                if (p.hit_out) p.hit_out[pos] = h;
                if (p.dst_out) p.dst_out[pos] = d;
            }
        }
    }
}

cudaError_t launch_identity(int np, const IdentityParams &p, int num_sms, cudaStream_t stream)
{
    const long long ntiles = p.tile_end - p.tile_begin;
    if (ntiles <= 0) return cudaSuccess;
#define TCU_ID_CASE(N)                                                                            \
    case N: {                                                                                     \
        const size_t smem = identity_smem_bytes(N);                                               \
        cudaError_t e = cudaFuncSetAttribute(k_identity<N>,                                       \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                             (int)smem);                                          \
        if (e != cudaSuccess) return e;                                                           \
        const int per_sm = identity_stages(N) == 4 ? 2 : 1;                                       \
        const int grid = (int)std::min<long long>(ntiles, (long long)num_sms * per_sm);                \
        k_identity<N><<<grid, ID_THREADS, smem, stream>>>(p);                                     \
        break;                                                                                    \
    }
    switch (np) {
        TCU_ID_CASE(3)
        TCU_ID_CASE(4)
        TCU_ID_CASE(5)
        TCU_ID_CASE(6)
        TCU_ID_CASE(7)
    default: return cudaErrorInvalidValue;
    }
#undef TCU_ID_CASE
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Debug cross-check: the same statistic straight from the raw bytes, one
// thread per pair.  Slow by design; only reachable through the test entry
// point tcu_debug_identity_bytes so that a packing or pipeline fault can be
// told apart from an arithmetic one on the GPU.
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
                                  int *hit_out, int *dst_out, cudaStream_t stream)
{
    const long long npairs = (long long)nk * (nk - 1) / 2;
    if (npairs <= 0) return cudaSuccess;
    const int blocks = (int)std::min<long long>((npairs + 255) / 256, 148 * 16);
    k_identity_bytes<<<blocks, 256, 0, stream>>>(raw, pitch, ncol, kept_rows, nk, col_drop, indet,
                                                 out, hit_out, dst_out);
    return cudaGetLastError();
}

}  // namespace tcu
