// pack.cu -- K0: turn the raw byte matrix into the packed identity operand.
//
// The reference kernels read std::string rows directly (template.h:352-361);
// there is no packing step to mirror.  Layout: see tcu_internal.cuh.
#include <algorithm>

#include "tcu_internal.cuh"

namespace tcu {

// ---------------------------------------------------------------------------
// Which of the 256 byte values occur in columns [0, ncol)?  One pass over the
// matrix, 16 bytes per thread per row, flags collected in shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_byte_presence(const uint8_t *__restrict__ raw, int nseq,
                                                       int ncol, size_t pitch,
                                                       unsigned int *__restrict__ present256)
{
    __shared__ unsigned int flags[256];
    flags[threadIdx.x] = 0;
    __syncthreads();

    const int groups = (ncol + 15) >> 4;  // 16-byte groups per row
    const long long total = (long long)nseq * groups;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / groups);
        const int g = (int)(idx - (long long)r * groups);
        const uint4 v = *reinterpret_cast<const uint4 *>(raw + (size_t)r * pitch + (size_t)g * 16);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const int valid = min(16, ncol - g * 16);
#pragma unroll
        for (int b = 0; b < 16; b++)
            if (b < valid) flags[(w[b >> 2] >> ((b & 3) * 8)) & 0xFF] = 1;
    }
    __syncthreads();
    if (flags[threadIdx.x]) present256[threadIdx.x] = 1;
}

cudaError_t launch_byte_presence(const uint8_t *raw, int nseq, int ncol, size_t pitch,
                                 unsigned int *present256, int num_sms, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0) return cudaSuccess;
    const long long total = (long long)nseq * ((ncol + 15) >> 4);
    int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms * 8);
    k_byte_presence<<<blocks, 256, 0, stream>>>(raw, nseq, ncol, pitch, present256);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Rows at an arbitrary host stride (as they arrived in one linear copy) -> rows at the
// device pitch (a multiple of 128), padding zeroed.  One warp per row, 4 bytes per lane
// per step on the destination side (pitch and dst are 4-byte aligned; the source is not).
// ---------------------------------------------------------------------------
struct RepitchPeers {
    int n;
    uint8_t *dst[ID2_MAX_PEERS];  // the same matrix on other devices (peer memory)
};

__global__ void __launch_bounds__(256) k_repitch_rows(const uint8_t *__restrict__ src,
                                                      size_t stride, int nseq, int ncol,
                                                      uint8_t *__restrict__ dst, size_t pitch,
                                                      size_t dst_offset, const RepitchPeers peers)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nseq) return;
    const uint8_t *s = src + (size_t)row * stride;
    const size_t at = dst_offset + (size_t)row * pitch;
    uint32_t *d = reinterpret_cast<uint32_t *>(dst + at);
    for (int k = lane * 4; k < (int)pitch; k += 128) {
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; b++)
            if (k + b < ncol) w |= (uint32_t)s[k + b] << (8 * b);
        d[k >> 2] = w;
        for (int q = 0; q < peers.n; q++) reinterpret_cast<uint32_t *>(peers.dst[q] + at)[k >> 2] = w;
    }
}

// rows [0, nseq) of `src` to byte offset dst_offset of `dst` -- and of every matrix in
// `peer_dst` (n_peers <= ID2_MAX_PEERS): a rank's share of the rows reaches all ranks from
// the kernel that lays it out, over NVLink
cudaError_t launch_repitch_rows(const uint8_t *src, size_t stride, int nseq, int ncol, uint8_t *dst,
                                size_t pitch, size_t dst_offset, uint8_t *const *peer_dst,
                                int n_peers, cudaStream_t stream)
{
    if (nseq == 0) return cudaSuccess;
    if (n_peers > ID2_MAX_PEERS) return cudaErrorInvalidValue;
    RepitchPeers peers{};
    for (int q = 0; q < n_peers; q++) peers.dst[peers.n++] = peer_dst[q];
    const long long threads = (long long)nseq * 32;
    k_repitch_rows<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(src, stride, nseq, ncol,
                                                                          dst, pitch, dst_offset, peers);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// 256-bin histogram of the bytes in columns [0, ncol) of every row: everything
// utils::checkAlignmentType (source/utils.cpp:476-545) derives from its scan of the
// alignment is a sum of these bins.  16 bytes per thread per step; runs of equal bytes
// inside the 16 (gap stretches) are merged before they reach the warp-private
// shared-memory bins, which keeps same-address atomics rare.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_byte_histogram(const uint8_t *__restrict__ raw, int nseq,
                                                        int ncol, size_t pitch,
                                                        unsigned long long *__restrict__ hist256)
{
    __shared__ unsigned int bins[8][256];
    for (int k = threadIdx.x; k < 8 * 256; k += 256) (&bins[0][0])[k] = 0;
    __syncthreads();
    unsigned int *mine = bins[threadIdx.x >> 5];

    const int groups = (ncol + 15) >> 4;
    const long long total = (long long)nseq * groups;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / groups);
        const int g = (int)(idx - (long long)r * groups);
        const uint4 v = *reinterpret_cast<const uint4 *>(raw + (size_t)r * pitch + (size_t)g * 16);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const int valid = min(16, ncol - g * 16);
        uint32_t prev = w[0] & 0xFF, run = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            if (b < valid) {
                const uint32_t c = (w[b >> 2] >> ((b & 3) * 8)) & 0xFF;
                if (c == prev) {
                    run++;
                } else {
                    atomicAdd(&mine[prev], run);
                    prev = c;
                    run = 1;
                }
            }
        }
        if (run) atomicAdd(&mine[prev], run);
    }
    __syncthreads();
    unsigned long long sum = 0;
#pragma unroll
    for (int wq = 0; wq < 8; wq++) sum += bins[wq][threadIdx.x];
    if (sum) atomicAdd(&hist256[threadIdx.x], sum);
}

// hist256 must be zeroed by the caller
cudaError_t launch_byte_histogram(const uint8_t *raw, int nseq, int ncol, size_t pitch,
                                  unsigned long long *hist256, int num_sms, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0) return cudaSuccess;
    const long long total = (long long)nseq * ((ncol + 15) >> 4);
    int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms * 8);
    k_byte_histogram<<<blocks, 256, 0, stream>>>(raw, nseq, ncol, pitch, hist256);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// The identity operand (layout: tcu_internal.cuh).  One CTA per (chunk, block): 64 rows x
// 128 columns, one thread per (row, 32-column word).  A thread reads the 32 bytes of its
// word, maps them through the code LUT (bit 7 of an entry = gap class) four at a time, and
// squeezes bit p of four codes into a nibble of plane word p with one multiply: no
// cross-lane operation at all (round 1 formed the words with warp ballots, one byte per lane
// and six ballots per word: 0.145 ms at 50 000 x 1 000; a segmented REDUX needs a uniform
// mask and is serialised per group).  The gap flags leave as the bytes of the UMMA operand.
// ---------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(256) k_pack_planes(const uint8_t *__restrict__ raw, size_t pitch,
                                                      int ncol, const int *__restrict__ kept_rows,
                                                      int nk, const uint8_t *__restrict__ col_drop,
                                                      const uint8_t *__restrict__ lut256,
                                                      int nb2, int nchunks,
                                                      uint32_t *__restrict__ planes,
                                                      uint8_t *__restrict__ gbytes)
{
    constexpr int RP = rest_words(NP);
    constexpr int TW = tile2_words(NP);
    constexpr uint32_t GAP_BITS = (1u << NP) - 2u;  // p0 = 0, rest = 1..1
    constexpr uint32_t GAP_ENTRY = 0x80u | GAP_BITS;
    static_assert(NP <= 7, "bit 7 of a LUT entry is the gap flag");

    __shared__ __align__(16) uint32_t tile[TW];
    __shared__ __align__(16) uint8_t gsm[G_STAGES_PER_CHUNK * G_BLOCK_BYTES];
    __shared__ uint8_t lut[256];
    __shared__ uint32_t s_live[KC2 * 8];  // per 4 columns of the chunk: 0xFF for a usable column

    const int chunk = blockIdx.x;
    const int block = blockIdx.y;

    {
        const uint8_t c = lut256[threadIdx.x];
        lut[threadIdx.x] = c == CODE_GAP ? (uint8_t)GAP_ENTRY : c;
    }
    if (threadIdx.x < KC2 * 8) {
        const int c0 = chunk * (KC2 * 32) + 4 * threadIdx.x;
        const uint32_t d4 = *reinterpret_cast<const uint32_t *>(col_drop + c0);
        uint32_t live = 0;
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (c0 + q < ncol && ((d4 >> (8 * q)) & 0xFFu) == 0) live |= 0xFFu << (8 * q);
        s_live[threadIdx.x] = live;
    }
    if (RP > NP - 1) {  // padding words of the rest group must be defined
        for (int i = threadIdx.x; i < TW; i += 256) tile[i] = 0;
    }

    uint32_t *p0b = tile;
    uint32_t *rest = tile + p0_words();
    uint32_t *p0a = rest + KC2 * RB * RP;

    const int r = threadIdx.x & (RB - 1), kw = threadIdx.x / RB;
    const int ki = block * RB + r;
    const bool row_ok = ki < nk;
    uint32_t rw[8];
    {
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
        if (row_ok) {
            const uint4 *src = reinterpret_cast<const uint4 *>(raw + (size_t)kept_rows[ki] * pitch +
                                                               (size_t)(chunk * KC2 + kw) * 32);
            a = src[0];
            b = src[1];
        }
        rw[0] = a.x, rw[1] = a.y, rw[2] = a.z, rw[3] = a.w;
        rw[4] = b.x, rw[5] = b.y, rw[6] = b.z, rw[7] = b.w;
    }
    __syncthreads();  // lut, s_live, zeroed tile

    const uint32_t rowmask = row_ok ? 0xFFFFFFFFu : 0u;
    uint32_t words[NP], gapw = 0, gapb[8];
#pragma unroll
    for (int p = 0; p < NP; p++) words[p] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t x = rw[i];
        uint32_t codes4 = (uint32_t)lut[x & 0xFFu] | ((uint32_t)lut[(x >> 8) & 0xFFu] << 8) |
                          ((uint32_t)lut[(x >> 16) & 0xFFu] << 16) | ((uint32_t)lut[x >> 24] << 24);
        const uint32_t live = s_live[kw * 8 + i] & rowmask;
        codes4 = (codes4 & live) | (~live & (0x01010101u * GAP_ENTRY));
        // bits 0, 8, 16, 24 -> bits 21..24 (no two partial products share a position)
        auto nibble = [&](uint32_t t) { return ((t * 0x00204081u) >> 21) & 0xFu; };
#pragma unroll
        for (int p = 0; p < NP; p++) words[p] |= nibble((codes4 >> p) & 0x01010101u) << (4 * i);
        gapb[i] = (codes4 >> 7) & 0x01010101u;
        gapw |= nibble(gapb[i]) << (4 * i);
    }
    const int cell = kw * RB + r;
    p0b[cell] = words[0];
    p0a[cell] = words[0] | gapw;
#pragma unroll
    for (int p = 1; p < NP; p++) rest[cell * RP + p - 1] = words[p];
    // gap bytes: stage kw / 2, columns (kw % 2) * 32 .. + 31 of the stage
    {
        uint8_t *g = gsm + (kw >> 1) * G_BLOCK_BYTES + (r >> 3) * 512 + ((kw & 1) * 2) * 128 + (r & 7) * 16;
        *reinterpret_cast<uint4 *>(g) = make_uint4(gapb[0], gapb[1], gapb[2], gapb[3]);
        *reinterpret_cast<uint4 *>(g + 128) = make_uint4(gapb[4], gapb[5], gapb[6], gapb[7]);
    }
    __syncthreads();

    uint4 *dst = reinterpret_cast<uint4 *>(planes + ((size_t)block * nchunks + chunk) * TW);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(tile);
    for (int i = threadIdx.x; i < TW / 4; i += 256) dst[i] = s4[i];
    const uint4 *g4 = reinterpret_cast<const uint4 *>(gsm);
#pragma unroll
    for (int s = 0; s < G_STAGES_PER_CHUNK; s++) {
        uint4 *gd = reinterpret_cast<uint4 *>(
            gbytes + ((size_t)(chunk * G_STAGES_PER_CHUNK + s) * nb2 + block) * G_BLOCK_BYTES);
        for (int i = threadIdx.x; i < G_BLOCK_BYTES / 16; i += 256)
            gd[i] = g4[s * (G_BLOCK_BYTES / 16) + i];
    }
}

cudaError_t launch_pack_planes(const uint8_t *raw, size_t pitch, int ncol, const int *kept_rows,
                                int nk, const uint8_t *col_drop, const uint8_t *lut256, int np,
                                int nb2, int nchunks, uint32_t *planes, uint8_t *gbytes,
                                cudaStream_t stream)
{
    if (nb2 == 0 || nchunks == 0) return cudaSuccess;
    dim3 grid(nchunks, nb2);
#define TCU_PACK_CASE(N)                                                                        \
    case N:                                                                                      \
        k_pack_planes<N><<<grid, 256, 0, stream>>>(raw, pitch, ncol, kept_rows, nk, col_drop,   \
                                                    lut256, nb2, nchunks, planes, gbytes);       \
        break;
    switch (np) {
        TCU_PACK_CASE(3)
        TCU_PACK_CASE(4)
        TCU_PACK_CASE(5)
        TCU_PACK_CASE(6)
        TCU_PACK_CASE(7)
    default: return cudaErrorInvalidValue;
    }
#undef TCU_PACK_CASE
    return cudaGetLastError();
}

}  // namespace tcu
