// identity2.cu -- K1: pairwise sequence identity, integer pipes + tensor cores.
//
// Replaces simd::calculateSeqIdentity<V> (vendor/trimal/include/Platform/
// template.h:320-442).  Per kept pair i<j:
//     dst = #{kept columns : not (gap_i and gap_j)}      (template.h:422)
//     hit = #{those columns : byte_i == byte_j}          (template.h:423-424)
//     identity = dst ? (float)hit / (float)dst : 0       (template.h:427-434)
// written at the packed upper-triangular offset of (i, j) (template.h:436).
//
// The two counts run on different units of the SM, concurrently:
//   hit    integer pipes.  Residues are dense codes stored as NP bit-planes; for
//          one 32-column word of one pair
//              differ = (a.p0A ^ b.p0B) | (a.r1 ^ b.r1) | ... | (a.rN ^ b.rN)
//              hit   += popc(~differ)
//          i.e. NP LOP3 + 1 POPC + 1 add per 32 pair-columns (the gap class is
//          encoded so that it never compares equal, see tcu_internal.cuh).
//   both   = #{columns : gap_i and gap_j} = G Gt, a binary matrix product: one
//          tcgen05.mma kind::i8 (M = 128 I rows, N = 64 J rows, K = 32 columns)
//          per 32 columns of a tile, accumulated in TMEM as int32;
//          dst = total_bits - both because padding/masked columns are gaps in
//          every row.
// Everything is integer until the single IEEE fp32 division of the epilogue, so
// the results are bit-identical to the reference for any input bytes.
//
// One CTA (two per SM) walks 128 x 64 tiles of the pair matrix:
//   warps 0-7  math: 32 x 32 pairs per warp, 8 x 4 per thread, plane chunks of
//              128 columns from a 3-stage ring; the popcount of a pair trails its
//              LOP3 chain by one row-group and the add by another (the in-order warp
//              never waits on the XU pipe); epilogue: TMEM -> registers ->
//              warp-private shared-memory transpose -> divide -> store
//   warp 8     producer: 1-D bulk async copies (TMA unit) of plane chunks and of
//              the 64-column gap-byte stages, mbarrier complete_tx
//   warp 9     MMA issuer: one thread, two UMMAs per gap-byte stage, tcgen05.commit
//              frees the stage / publishes the accumulator (two TMEM buffers)
#include <algorithm>
#include <type_traits>

#include "tcu_internal.cuh"

namespace tcu {

constexpr int ID2_MATH_WARPS = 8;
constexpr int ID2_MATH_THREADS = ID2_MATH_WARPS * 32;
constexpr int ID2_THREADS = ID2_MATH_THREADS + 64;  // + producer warp + MMA warp
constexpr int ID2_PSTAGES = 3;                               // plane-chunk ring
constexpr int ID2_GSTAGES = 2;                               // gap-byte ring
constexpr int ID2_GSTAGE_BYTES = (IB + RB) * GS_COLS;         // 12288
constexpr int ID2_XS = 36;                                   // exchange row stride (words)
constexpr int ID2_XCH_BYTES = ID2_MATH_WARPS * 32 * ID2_XS * 4;  // one 32 x 32 patch per warp
constexpr int ID2_TMEM_COLS = 2 * RB;                         // two int32 accumulators of N = 64
constexpr int ID2_NBARS = 2 * ID2_PSTAGES + 2 * ID2_GSTAGES + 4;

__host__ __device__ constexpr int id2_pstage_bytes(int np) { return 3 * role_bytes(np); }
__host__ __device__ constexpr size_t id2_smem_bytes(int np)
{
    return (size_t)ID2_PSTAGES * id2_pstage_bytes(np) + ID2_GSTAGES * ID2_GSTAGE_BYTES +
           ID2_XCH_BYTES + ID2_NBARS * sizeof(uint64_t) + 16;
}

// ------------------------------ tcgen05 helpers ------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no swizzle: 8 x 16-byte core matrices; LBO between the two 16-byte K
// halves of one MMA, SBO between 8-row groups (cute::UMMA::SmemDescriptor:
// start >> 4 at [0,14), LBO >> 4 at [16,30), SBO >> 4 at [32,46), version 1 at [46,48))
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// cute::UMMA::InstrDescriptor: D = s32 (2 at [4,6)), A = B = u8 (0), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t ID2_IDESC = (2u << 4) | ((uint32_t)(RB >> 3) << 17) | ((uint32_t)(IB >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(ID2_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// tile_to_blocks2(): linear tile index -> (I super-block, J block), see tcu_internal.cuh

// One row's plane words for one 32-column word: p0 (the role's copy of plane 0)
// and the NP-1 "rest" planes.
template <int NP>
struct Row {
    uint32_t p0;
    uint32_t r[NP - 1];
};

template <int NP>
__device__ __forceinline__ void load_row(const uint32_t *p0, const uint32_t *rest, int cell, Row<NP> &o)
{
    constexpr int R = NP - 1;
    constexpr int RP = rest_words(NP);
    const uint32_t *q = rest + cell * RP;
    o.p0 = p0[cell];
    if constexpr (RP == 2) {
        const uint2 v = *reinterpret_cast<const uint2 *>(q);
        o.r[0] = v.x;
        o.r[1] = v.y;
    } else {
        const uint4 v = *reinterpret_cast<const uint4 *>(q);
        o.r[0] = v.x;
        o.r[1] = v.y;
        o.r[2] = v.z;
        if constexpr (R >= 4) o.r[3] = v.w;
        if constexpr (RP == 8) {
            const uint4 u = *reinterpret_cast<const uint4 *>(q + 4);
            o.r[4] = u.x;
            if constexpr (R >= 6) o.r[5] = u.y;
        }
    }
}

// ---- order-pinned forms (asm volatile keeps the relative order through NVVM; ptxas
// still schedules, but starts from this order) ----
__device__ __forceinline__ uint32_t v_xor(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
template <int LUT>
__device__ __forceinline__ uint32_t v_lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}
__device__ __forceinline__ uint32_t v_popc(uint32_t a)
{
    uint32_t d;
    asm volatile("popc.b32 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
// equal columns of one pair in one 32-column word: NP LOP3
template <int NP>
__device__ __forceinline__ uint32_t v_equal_bits(const Row<NP> &a, const Row<NP> &b)
{
    uint32_t d = v_xor(a.p0, b.p0);
#pragma unroll
    for (int q = 0; q < NP - 2; q++) d = v_lop3<0xF6>(d, a.r[q], b.r[q]);
    return v_lop3<0x09>(d, a.r[NP - 2], b.r[NP - 2]);
}

// PACKED: two 16-bit hit counters per register (J rows b and b+2 of a thread share one;
// the upper one is fed by an IMAD with 65536 on the otherwise idle FMA pipe), valid
// while a count cannot reach 65536, i.e. total_bits < 65536.  Frees 16 registers.
template <int NP, bool PACKED>
__global__ void __launch_bounds__(ID2_THREADS, NP <= 5 ? 2 : 1) k_identity2(const Identity2Params p)
{
    constexpr int RP = rest_words(NP);
    constexpr int ROLE_W = role_words(NP);
    constexpr int ROLE_B = role_bytes(NP);
    constexpr int PST_B = id2_pstage_bytes(NP);
    constexpr int TILE_B = tile2_bytes(NP);

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_planes = smem;
    uint8_t *s_g = s_planes + ID2_PSTAGES * PST_B;
    uint32_t *s_xch = reinterpret_cast<uint32_t *>(s_g + ID2_GSTAGES * ID2_GSTAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_xch) + ID2_XCH_BYTES);
    uint64_t *pfull = bars, *pempty = bars + ID2_PSTAGES;
    uint64_t *gfull = pempty + ID2_PSTAGES, *gempty = gfull + ID2_GSTAGES;
    uint64_t *accfull = gempty + ID2_GSTAGES, *accempty = accfull + 2;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + ID2_NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < ID2_PSTAGES; s++) {
            mbar_init(&pfull[s], 1);
            mbar_init(&pempty[s], ID2_MATH_WARPS);
        }
        for (int s = 0; s < ID2_GSTAGES; s++) {
            mbar_init(&gfull[s], 1);
            mbar_init(&gempty[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(&accfull[s], 1);
            mbar_init(&accempty[s], ID2_MATH_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 9) {  // TMEM: two 64-column int32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(s_tmem)),
                     "r"((uint32_t)ID2_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const int ngs = p.nchunks * G_STAGES_PER_CHUNK;

    if (warp == 8) {
        // ------------------------------ producer ------------------------------
        if (lane == 0) {
            int ps = 0, gs = 0;
            uint32_t pph = 0, gph = 0;
            const uint8_t *planes = reinterpret_cast<const uint8_t *>(p.planes);
            for (long long t = p.tile_begin + blockIdx.x; t < p.tile_end; t += gridDim.x) {
                int BI, bj;
                tile_to_blocks2(t, p.nb, p.sb_begin, p.sb_end, BI, bj);
                const uint8_t *srcA0 = planes + (size_t)(2 * BI) * p.nchunks * TILE_B + p0_words() * 4;
                const uint8_t *srcA1 = srcA0 + (size_t)p.nchunks * TILE_B;
                const uint8_t *srcB = planes + (size_t)bj * p.nchunks * TILE_B;
                for (int c = 0; c < p.nchunks; c++) {
#pragma unroll
                    for (int h = 0; h < G_STAGES_PER_CHUNK; h++) {
                        const int s = c * G_STAGES_PER_CHUNK + h;
                        const uint8_t *gsrc = p.gbytes + (size_t)s * p.nb2 * G_BLOCK_BYTES;
                        mbar_wait_parked(&gempty[gs], gph ^ 1u);
                        mbar_arrive_expect_tx(&gfull[gs], ID2_GSTAGE_BYTES);
                        uint8_t *gd = s_g + gs * ID2_GSTAGE_BYTES;
                        bulk_copy_g2s(gd, gsrc + (size_t)(2 * BI) * G_BLOCK_BYTES, 2 * G_BLOCK_BYTES,
                                      &gfull[gs]);
                        bulk_copy_g2s(gd + 2 * G_BLOCK_BYTES, gsrc + (size_t)bj * G_BLOCK_BYTES,
                                      G_BLOCK_BYTES, &gfull[gs]);
                        if (++gs == ID2_GSTAGES) {
                            gs = 0;
                            gph ^= 1u;
                        }
                    }
                    mbar_wait_parked(&pempty[ps], pph ^ 1u);
                    mbar_arrive_expect_tx(&pfull[ps], 3u * ROLE_B);
                    uint8_t *pd = s_planes + ps * PST_B;
                    bulk_copy_g2s(pd, srcA0 + (size_t)c * TILE_B, ROLE_B, &pfull[ps]);
                    bulk_copy_g2s(pd + ROLE_B, srcA1 + (size_t)c * TILE_B, ROLE_B, &pfull[ps]);
                    bulk_copy_g2s(pd + 2 * ROLE_B, srcB + (size_t)c * TILE_B, ROLE_B, &pfull[ps]);
                    if (++ps == ID2_PSTAGES) {
                        ps = 0;
                        pph ^= 1u;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ------------------------------ MMA issuer ----------------------------
        if (lane == 0) {
            int gs = 0;
            uint32_t gph = 0;
            long long k = 0;
            for (long long t = p.tile_begin + blockIdx.x; t < p.tile_end; t += gridDim.x, k++) {
                const int buf = (int)(k & 1);
                const uint32_t use_parity = (uint32_t)((k >> 1) & 1);
                mbar_wait_parked(&accempty[buf], use_parity ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(buf * RB);
                for (int s = 0; s < ngs; s++) {
                    mbar_wait_parked(&gfull[gs], gph);
                    tc_fence_after();
                    const uint32_t a = smem_u32(s_g + gs * ID2_GSTAGE_BYTES);
                    const uint32_t b = a + 2 * G_BLOCK_BYTES;
#pragma unroll
                    for (int kk = 0; kk < GS_COLS / 32; kk++) {
                        umma_i8(d, umma_desc(a + kk * 256, 128, 512), umma_desc(b + kk * 256, 128, 512),
                                (s | kk) != 0 ? 1u : 0u);
                    }
                    umma_commit(&gempty[gs]);
                    if (++gs == ID2_GSTAGES) {
                        gs = 0;
                        gph ^= 1u;
                    }
                }
                umma_commit(&accfull[buf]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------ math warps ----------------------------
        // 4 (I) x 2 (J) warps of 32 x 32 pairs; wi = warp % 4 is also the TMEM lane
        // quadrant this warp may read, so its both-gap counts are exactly its own patch
        const int wi = warp & 3, wj = warp >> 2;
        const int li = lane & 3, lj = lane >> 2;
        const int half = wi >> 1;                     // which block of the super-block
        const int rowA0 = 32 * (wi & 1) + li;         // + 4*a, row inside that block
        const int rowB0 = 32 * wj + lj;               // + 8*b
        const unsigned long long n = (unsigned long long)p.nk;

        // threshold mode: the 64 column entries of the tile whose words sit in buffer kb, stored
        // one tile late (see the epilogue): every warp takes eight of them (equal work for
        // all warps keeps them on the same ring stage), 128 contiguous bytes per warp.  The
        // entry's position was parked next to the words, so that nothing stays live across
        // the counting loop.
        auto flush_bits = [&](int kb) {
            if (lane >= 8) return;
            const int col = 8 * warp + lane;  // column of the tile: patch row col % 32 of the warps (.., col / 32)
            const uint32_t *w = s_xch + (4 * (col >> 5)) * (32 * ID2_XS) + (col & 31) * ID2_XS + 32 + kb;
            const uint32_t at4 = w[2];  // 16-byte entry index, ~0 for a column beyond nk
            if (at4 == 0xFFFFFFFFu) return;
            const uint4 v = make_uint4(w[0], w[32 * ID2_XS], w[2 * 32 * ID2_XS], w[3 * 32 * ID2_XS]);
            const size_t at = (size_t)at4 * 4;
            *reinterpret_cast<uint4 *>(p.bits_out + at) = v;
            for (int q = 0; q < p.n_bits_peer; q++) *reinterpret_cast<uint4 *>(p.bits_peer[q] + at) = v;
        };

        // barrier addresses as plain shared-memory offsets, converted once
        const uint32_t pfull_a = smem_u32(pfull), pempty_a = smem_u32(pempty);
        const uint32_t accfull_a = smem_u32(accfull), accempty_a = smem_u32(accempty);

        int stage = 0;
        uint32_t phase = 0;
        long long k = 0;
        for (long long t = p.tile_begin + blockIdx.x; t < p.tile_end; t += gridDim.x, k++) {
            int BI, bj;
            tile_to_blocks2(t, p.nb, p.sb_begin, p.sb_end, BI, bj);

            constexpr int HB = PACKED ? 2 : 4;
            uint32_t hit[8][HB];
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b < HB; b++) hit[a][b] = 0;
            auto add_hit = [&](int a, int b, uint32_t x) {
                if (!PACKED) hit[a][b] += x;
                else if (b < 2) hit[a][b] += x;
                else hit[a][b - 2] = x * 65536u + hit[a][b - 2];
            };

            uint32_t ep[4] = {0, 0, 0, 0}, pc[4] = {0, 0, 0, 0};
            // per word: hold the four J rows, stream the eight I rows (double-buffered);
            // the popcount of a pair trails its LOP3 chain by one row-group and the add
            // by another one, so neither the XU latency nor its queue stalls the in-order
            // warp (ep / pc are the two pipeline registers per J row).
            for (int c = 0; c < p.nchunks; c++) {
                mbar_wait_a(pfull_a + 8u * (uint32_t)stage, phase);
                const uint32_t *base = reinterpret_cast<const uint32_t *>(s_planes + stage * PST_B);
                const uint32_t *a_rest = base + half * ROLE_W, *a_p0 = a_rest + KC2 * RB * RP;
                const uint32_t *b_p0 = base + 2 * ROLE_W, *b_rest = b_p0 + p0_words();
#pragma unroll 1
                for (int kw = 0; kw < KC2; kw++) {
                    Row<NP> B[4], A, An;
#pragma unroll
                    for (int b = 0; b < 4; b++) load_row<NP>(b_p0, b_rest, kw * RB + rowB0 + 8 * b, B[b]);
                    load_row<NP>(a_p0, a_rest, kw * RB + rowA0, A);
#pragma unroll
                    for (int a = 0; a < 8; a++) {
                        if (a < 7) load_row<NP>(a_p0, a_rest, kw * RB + rowA0 + 4 * (a + 1), An);
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const uint32_t e = v_equal_bits<NP>(A, B[b]);
                            add_hit((a + 6) & 7, b, pc[b]);
                            pc[b] = v_popc(ep[b]);
                            ep[b] = e;
                        }
                        A = An;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(pempty_a + 8u * (uint32_t)stage);
                if (++stage == ID2_PSTAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
#pragma unroll
            for (int b = 0; b < 4; b++) {  // drain the popcount pipeline
                add_hit(6, b, pc[b]);
                add_hit(7, b, __popc(ep[b]));
            }

            // ------------------------------ epilogue --------------------------
            // both-gap counts: TMEM lane = I row, column = J row; the warp's patch is
            // lanes 32*wi..+31, columns 32*wj..+31 of the tile's accumulator.  Lane r
            // receives row r; a warp-private shared-memory patch turns that into the
            // (li, lj) ownership of the hit counters -- no CTA-wide barrier.
            const int buf = (int)(k & 1);
            mbar_wait_a(accfull_a + 8u * (uint32_t)buf, (uint32_t)((k >> 1) & 1));
            tc_fence_after();
            uint32_t both[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * wi) << 16) + (uint32_t)(buf * RB + 32 * wj), both);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(accempty_a + 8u * (uint32_t)buf);
            // not needed for correctness (patches are warp-private): re-aligns the eight
            // warps once per tile so that all of them consume the same ring stage and
            // the other stages stay prefetched
            asm volatile("bar.sync 1, %0;" ::"n"(ID2_MATH_THREADS) : "memory");

            uint32_t *patch = s_xch + warp * (32 * ID2_XS);
            {
                uint32_t *dstrow = patch + lane * ID2_XS;
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                    *reinterpret_cast<uint4 *>(dstrow + q) =
                        make_uint4(both[q], both[q + 1], both[q + 2], both[q + 3]);
            }
            __syncwarp();

            if (p.bits_out != nullptr) {
                if (k > 0) flush_bits((int)((k - 1) & 1));  // the previous tile's entries
                // threshold mode (Cleaner.cpp:1435-1440: a pair joins two sequences iff
                // identity > threshold): only the comparison with the threshold leaves the SM,
                // decided in integers (ThresholdRule).  A thread owns the pairs i = li + 4a,
                // j = lj + 8b of the warp's 32 x 32 patch: its bits of column j go to positions
                // li + 4a of cw[b], the four lanes that share lj OR their words together (two
                // shuffles per column instead of a ballot per pair), and lane (lj, li) keeps
                // the word of patch column c = lj + 8 li (sequence j0 + c against the 32
                // sequences i0 ..).  Patches below the diagonal hold no pair j > i and store
                // zeros (the mirror pass fills them); a patch strictly above it and inside the
                // matrix needs no per-pair bounds.
                const int i0 = BI * IB + 32 * wi, j0 = bj * RB + 32 * wj;
                uint32_t cw[4] = {0u, 0u, 0u, 0u};
                if (j0 + 31 > i0 && p.thr.mode != 0) {
                    const bool interior = i0 + 31 < p.nk && j0 + 31 < p.nk && j0 > i0 + 31;
                    auto fill = [&](auto inside, auto always) {
#pragma unroll
                        for (int a = 0; a < 8; a++) {
                            const int i = i0 + li + 4 * a;
#pragma unroll
                            for (int b = 0; b < 4; b++) {
                                const int j = j0 + lj + 8 * b;
                                bool over = true;
                                if (!decltype(always)::value) {
                                    const uint32_t h = !PACKED ? hit[a][b]
                                                               : (b < 2 ? (hit[a][b] & 0xFFFFu)
                                                                        : (hit[a][b - 2] >> 16));
                                    const uint32_t d = (uint32_t)p.total_bits -
                                                       patch[(li + 4 * a) * ID2_XS + lj + 8 * b];
                                    // fl(h / d) > thr, exactly, without dividing
                                    over = h > (uint32_t)(((unsigned long long)p.thr.mul * d) >> p.thr.shift);
                                }
                                if (!decltype(inside)::value) over = over && i < p.nk && j < p.nk && j > i;
                                cw[b] |= (uint32_t)over << (li + 4 * a);
                            }
                        }
                    };
                    using T = std::true_type;
                    using F = std::false_type;
                    if (p.thr.mode == 2) {
                        if (interior) fill(T{}, F{});
                        else fill(F{}, F{});
                    } else {  // negative threshold: every pair
                        if (interior) fill(T{}, T{});
                        else fill(F{}, T{});
                    }
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        cw[b] |= __shfl_xor_sync(0xffffffffu, cw[b], 1);
                        cw[b] |= __shfl_xor_sync(0xffffffffu, cw[b], 2);
                    }
                }
                const uint32_t colword = li == 0 ? cw[0] : (li == 1 ? cw[1] : (li == 2 ? cw[2] : cw[3]));
                const int c = lj + 8 * li;  // this lane's column of the patch
                // The four words of a column (wi = 0..3) are 16 contiguous bytes of the slab
                // layout, held by four warps.  Each warp parks its word in the padding of its
                // patch (two buffers, by tile parity); one tile later -- the barrier above
                // has then seen every warp finish this epilogue -- whole 16-byte column entries
                // are stored, 1 KB contiguous per tile: to the local matrix and, in a multi-GPU
                // run, straight into every peer's over NVLink.
                patch[c * ID2_XS + 32 + (int)(k & 1)] = colword;
                if (wi == 0)  // (nslab * nk entries: below 2^32 up to 740 000 sequences)
                    patch[c * ID2_XS + 34 + (int)(k & 1)] =
                        j0 + c < p.nk ? (uint32_t)((size_t)BI * p.nk + (size_t)(j0 + c)) : 0xFFFFFFFFu;
                __syncwarp();  // the patch is rewritten by the next tile
                continue;
            }
            // a tile strictly above the diagonal and inside the matrix (all but the rim)
            // needs no per-pair bounds
            const bool interior = BI * IB + IB - 1 < p.nk && bj * RB + RB - 1 < p.nk &&
                                  bj * RB > BI * IB + IB - 1;
            auto store_ratios = [&](auto inside) {
#pragma unroll
                for (int a = 0; a < 8; a++) {
                    const int il = 64 * half + rowA0 + 4 * a;  // row inside the super-block
                    const int i = BI * IB + il;
                    if (!decltype(inside)::value && i >= p.nk) continue;
                    // offset of pair (i, i+1): i*n - i*(i+1)/2 - i - 1 + (i+1)
                    const unsigned long long row_base =
                        (unsigned long long)i * n - ((unsigned long long)i * (i + 1)) / 2 - i - 1;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int jl = rowB0 + 8 * b;
                        const int j = bj * RB + jl;
                        if (!decltype(inside)::value && (j >= p.nk || j <= i)) continue;
                        const unsigned long long pos = row_base + j;
                        const int h = !PACKED ? (int)hit[a][b]
                                              : (b < 2 ? (int)(hit[a][b] & 0xFFFFu) : (int)(hit[a][b - 2] >> 16));
                        const int d = p.total_bits - (int)patch[(li + 4 * a) * ID2_XS + lj + 8 * b];
                        // PACKED: fewer than 65 536 columns, the short exact division applies
                        const float v = d == 0 ? 0.0f
                                               : (PACKED ? div_small_counts((float)h, (float)d)
                                                         : __fdiv_rn((float)h, (float)d));
                        // streaming stores: 4 * P bytes pass through L2 once and must not evict
                        // the operand planes every CTA keeps re-reading (40 MB against 5 GB at C4)
                        __stcs(p.out + (pos - p.out_base), v);
                        if (p.hit_out) __stcs(p.hit_out + pos, h);
                        if (p.dst_out) __stcs(p.dst_out + pos, d);
                    }
                }
            };
            if (interior) store_ratios(std::true_type{});
            else store_ratios(std::false_type{});
            __syncwarp();  // the patch is rewritten by the next tile
        }
        if (p.bits_out != nullptr) {  // the last tile's column entries
            asm volatile("bar.sync 1, %0;" ::"n"(ID2_MATH_THREADS) : "memory");
            if (k > 0) flush_bits((int)((k - 1) & 1));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)ID2_TMEM_COLS)
                     : "memory");
    }
}

cudaError_t launch_identity2(int np, const Identity2Params &p, int num_sms, cudaStream_t stream)
{
    const long long ntiles = p.tile_end - p.tile_begin;
    if (ntiles <= 0) return cudaSuccess;
#define TCU_ID2_CASE_P(N, P)                                                                      \
    {                                                                                             \
        const size_t smem = id2_smem_bytes(N);                                                    \
        cudaError_t e = cudaFuncSetAttribute(k_identity2<N, P>,                                   \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                             (int)smem);                                          \
        if (e != cudaSuccess) return e;                                                           \
        e = cudaFuncSetAttribute(k_identity2<N, P>,                                               \
                                 cudaFuncAttributePreferredSharedMemoryCarveout,                  \
                                 cudaSharedmemCarveoutMaxShared);                                 \
        if (e != cudaSuccess) return e;                                                           \
        /* 228 KB per SM, 1 KB reserved per CTA: two CTAs when both fit */                        \
        const int per_sm = 2 * (smem + 1024) <= 228 * 1024 ? 2 : 1;                               \
        const int grid = (int)std::min<long long>(ntiles, (long long)num_sms * per_sm);           \
        k_identity2<N, P><<<grid, ID2_THREADS, smem, stream>>>(p);                                \
    }
#define TCU_ID2_CASE(N)                                                                           \
    case N:                                                                                       \
        if (p.total_bits < 65536) TCU_ID2_CASE_P(N, true) else TCU_ID2_CASE_P(N, false)           \
        break;
    switch (np) {
        TCU_ID2_CASE(3)
        TCU_ID2_CASE(4)
        TCU_ID2_CASE(5)
        TCU_ID2_CASE(6)
        TCU_ID2_CASE(7)
    default: return cudaErrorInvalidValue;
    }
#undef TCU_ID2_CASE
#undef TCU_ID2_CASE_P
    return cudaGetLastError();
}

}  // namespace tcu
