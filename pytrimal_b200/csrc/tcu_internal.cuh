// tcu_internal.cuh -- shared declarations of libtrimal_cuda (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "trimal_cuda.h"

namespace tcu {

constexpr int RB = 64;  // rows per row-block (J side of a tile, MMA N)

constexpr int MIN_PLANES = 3;
constexpr int MAX_PLANES = 7;
constexpr uint8_t CODE_GAP = 0xFF;  // LUT value for the identity gap class

// similarity codes (per byte, after upper-casing)
constexpr uint8_t SIM_GAP = 0xFF;
constexpr uint8_t SIM_INCORRECT = 0xFE;
constexpr uint8_t SIM_UNDEFINED = 0xFD;
constexpr int SIM_MAX_POS = 28;


// ---------------------------------------------------------------------------
// Layout of the identity operand (identity2.cu: hits on the integer pipes,
// both-gap counts on the tensor cores)
//
// Rows are grouped in blocks of RB = 64; a tile of the pair matrix is one I
// super-block (two consecutive blocks, 128 rows: the MMA's M) against one J
// block (64 rows: N).  Columns are grouped in chunks of KC2 = 4 words (128
// columns).
//
// planes   [block][chunk]{ p0B[kw][row] | rest[kw][row][RP] | p0A[kw][row] }
//          NP code bit-planes per (row, 32-column word); plane 0 is stored
//          twice: p0A = p0 | gap for rows used on the I side, p0B = p0 for the
//          J side.  The gap class carries the code 2^NP-2 (p0 = 0, rest = 1..1)
//          so  (a.p0A ^ b.p0B) | (a.r ^ b.r)...  is 1 wherever either side is a
//          gap: gap|gap differ in plane 0, gap|residue in the rest planes.
//          The I side copies {rest, p0A} (one contiguous run), the J side
//          {p0B, rest}: NP+... words per row reach shared memory, no gap plane.
// gbytes   [kstage][block]{ [rowgroup 8][kchunk 4][row 8][16 bytes] }
//          the gap indicator as one u8 per column, 64 columns per k-stage, in
//          the tcgen05 K-major no-swizzle canonical layout (8x16-byte core
//          matrices; leading-dimension byte offset 128 between the 16-column
//          kchunks, stride byte offset 512 between 8-row groups), so that
//          both(i,j) = sum_k g_i(k) g_j(k) is a kind::i8 UMMA straight from a
//          1-D bulk copy.  Two consecutive blocks are contiguous (the A operand
//          of a super-block is one 8 KB copy).
// Padding rows/columns and masked columns are gaps in every row, so
//          dst = total_bits - both.
// ---------------------------------------------------------------------------
constexpr int IB = 2 * RB;        // rows per I super-block
constexpr int KC2 = 4;            // 32-column words per plane chunk (128 columns)
constexpr int GS_COLS = 64;       // columns per gbytes k-stage
constexpr int G_BLOCK_BYTES = RB * GS_COLS;
constexpr int G_STAGES_PER_CHUNK = KC2 * 32 / GS_COLS;

__host__ __device__ constexpr int rest_words(int np) { return np - 1 <= 2 ? 2 : (np - 1 <= 4 ? 4 : 8); }
__host__ __device__ constexpr int p0_words() { return KC2 * RB; }
__host__ __device__ constexpr int tile2_words(int np) { return KC2 * RB * (2 + rest_words(np)); }
__host__ __device__ constexpr int tile2_bytes(int np) { return tile2_words(np) * 4; }
__host__ __device__ constexpr int role_words(int np) { return KC2 * RB * (1 + rest_words(np)); }
__host__ __device__ constexpr int role_bytes(int np) { return role_words(np) * 4; }

// "fl(h / d) > thr" without the division (identity2.cu, threshold mode).  The correctly
// rounded quotient exceeds thr exactly when the exact one exceeds the midpoint between thr
// and the next float above it (a tie cannot occur: the midpoint is an odd multiple of a power
// of two below 2^-24, and d < 2^24).  With thr = m 2^e that midpoint is (2m + 1) 2^(e-1), so
//     fl(h / d) > thr   <=>   h 2^k > (2m + 1) d   <=>   h > ((2m + 1) d) >> k,   k = 1 - e >= 25
// for 0 <= thr < 1; thr < 0 holds for every pair (also d = 0: the reference's 0 / 0 is 0),
// thr >= 1 or NaN for none.
struct ThresholdRule {
    int mode;       // 0: never, 1: always, 2: h > (mul * d) >> shift
    uint32_t mul;   // 2m + 1 < 2^25
    int shift;      // k, clamped to 63 (mul * d < 2^49)
};
inline ThresholdRule threshold_rule(float thr)
{
    uint32_t b;
    memcpy(&b, &thr, sizeof b);
    const uint32_t mag = b & 0x7FFFFFFFu;
    if (mag > 0x7F800000u) return ThresholdRule{0, 0, 0};            // NaN
    if ((b >> 31) && mag != 0) return ThresholdRule{1, 0, 0};         // negative
    if (mag >= 0x3F800000u) return ThresholdRule{0, 0, 0};            // >= 1, +inf
    const uint32_t E = mag >> 23, F = mag & 0x7FFFFFu;
    const uint32_t m = E ? (F | 0x800000u) : F;
    const int e = E ? (int)E - 150 : -149;
    const int k = 1 - e;
    return ThresholdRule{2, 2 * m + 1, k > 63 ? 63 : k};
}

constexpr int ID2_MAX_PEERS = 15;  // other devices whose bit matrix K1 may write

struct Identity2Params {
    const uint32_t *planes;  // v2 plane tiles
    const uint8_t *gbytes;   // gap indicator bytes, UMMA canonical layout
    float *out;              // identities, element 0 = packed offset out_base
    int *hit_out;            // optional, absolute packed offsets
    int *dst_out;            // optional
    unsigned long long out_base;
    long long tile_begin;    // linear tile range [begin, end), see tiles_before2()
    long long tile_end;
    int sb_begin, sb_end;    // the super-block rows the range covers (tile order, identity2.cu)
    int nb;                  // 64-row blocks holding kept rows
    int nb2;                 // blocks allocated (even)
    int nchunks;             // 128-column chunks
    int nk;                  // kept rows
    int total_bits;          // nchunks * 128
    // threshold mode (bits_out != nullptr): instead of the ratios, identity > thr as one bit per
    // pair in the slab layout below; out / hit_out / dst_out are not touched
    uint32_t *bits_out;
    ThresholdRule thr;
    // the same words stored into the bit matrices of other devices as well (peer memory over
    // NVLink: CUDA IPC mappings or peer access inside one process) -- the band exchange of a
    // multi-GPU clustering happens inside the epilogue, tile by tile, under the computation
    int n_bits_peer;
    uint32_t *bits_peer[ID2_MAX_PEERS];
};

// ---------------------------------------------------------------------------
// Threshold bit matrix ("is identity(i, j) > thr"), symmetric, nk x nk bits, stored in slabs
// of 128 columns:  word (row r, 32-column word w)  at  ((w / 4) * nk + r) * 4 + (w % 4),
// i.e. slab s = [nk rows][4 words] holds every sequence's bits against the sequences
// 128 s .. 128 s + 127.  A row-block band of the pair matrix (what one GPU owns) is a
// contiguous run of slabs, so the bands of several GPUs concatenate without a strided
// copy.  K1 writes, for the pairs i < j it owns, the bits of row j against the earlier
// sequences i ("column words").  The clustering walk reads whole rows in an arbitrary order,
// so k_bits_rows then writes the full symmetric matrix with one contiguous row per sequence
// (brow_*: row r = brow_pitch_words words, bit j of the row = pair (r, j); rows are whole
// 128-byte lines).
// ---------------------------------------------------------------------------
__host__ __device__ inline size_t bits_slab_words(int nk) { return (size_t)nk * 4; }
__host__ __device__ inline size_t bits_word_index(int nk, int row, int w)
{
    return ((size_t)(w >> 2) * (size_t)nk + (size_t)row) * 4 + (size_t)(w & 3);
}
__host__ __device__ inline size_t bits_total_words(int nk)
{
    return (size_t)((nk + 127) / 128) * bits_slab_words(nk);
}
__host__ __device__ inline size_t brow_pitch_words(int nk)
{
    return (size_t)(((nk + 127) / 128 + 7) / 8 * 8) * 4;
}
__host__ __device__ inline size_t brow_total_words(int nk) { return (size_t)nk * brow_pitch_words(nk); }

// tiles in super-block rows < sb: each super-block BI pairs with J blocks 2*BI .. nb-1
__host__ __device__ inline long long tiles_before2(long long sb, long long nb)
{
    return sb * nb - sb * (sb - 1);
}

// linear tile index -> (I super-block, J block).  Super-block BI owns the tiles
// (BI, 2*BI .. nb-1) and tiles_before2(BI, nb) tiles precede it, so a launch over
// the super-blocks [sb_begin, sb_end) is one contiguous index range.  Inside the
// range the tiles are visited in groups of ID2_GROUP super-block rows, J block by
// J block (all rows of the group that own the J block, then the next J block):
// the CTAs running at any moment share a few J blocks and ID2_GROUP I blocks, and
// a J block is fetched from HBM once per group instead of once per super-block
// (the operand set plus the output written between two visits of a row-major
// order exceed one L2 half: 3.9 GB of re-reads at 50 000 x 1 000).
constexpr int ID2_GROUP = 8;

__host__ __device__ inline void tile_to_blocks2(long long t, int nb, int sb_begin, int sb_end,
                                                int &BI, int &bj)
{
    const double m = (double)nb + 1.0;
    int b = (int)((m - sqrt(m * m - 4.0 * (double)t > 0.0 ? m * m - 4.0 * (double)t : 0.0)) * 0.5);
    const int nsb = (nb + 1) >> 1;
    b = b < 0 ? 0 : (b > nsb - 1 ? nsb - 1 : b);
    while (b > 0 && tiles_before2(b, nb) > t) b--;
    while (b + 1 < nsb && tiles_before2(b + 1, nb) <= t) b++;
    // b = the row whose row-major range holds t; its group starts at r0
    const int r0 = sb_begin + (b - sb_begin) / ID2_GROUP * ID2_GROUP;
    const int G = ID2_GROUP < sb_end - r0 ? ID2_GROUP : sb_end - r0;
    int u = (int)(t - tiles_before2(r0, nb));
    const int ramp = G * (G - 1);  // J blocks 2*r0 .. 2*(r0+G-1)-1: row r0+s joins at 2*(r0+s)
    int d, row;
    if (u >= ramp) {
        u -= ramp;
        d = 2 * (G - 1) + u / G;
        row = u % G;
    } else {
        int s = 0;
        while (u >= 2 * (s + 1)) {
            u -= 2 * (s + 1);
            s++;
        }
        d = 2 * s + u / (s + 1);
        row = u % (s + 1);
    }
    BI = r0 + row;
    bj = 2 * r0 + d;
}

// launchers (each enqueues on `stream` and returns the launch status)
cudaError_t launch_byte_presence(const uint8_t *raw, int nseq, int ncol, size_t pitch,
                                 unsigned int *present256, int num_sms, cudaStream_t stream);
cudaError_t launch_repitch_rows(const uint8_t *src, size_t stride, int nseq, int ncol, uint8_t *dst,
                                size_t pitch, size_t dst_offset, uint8_t *const *peer_dst,
                                int n_peers, cudaStream_t stream);
cudaError_t launch_byte_histogram(const uint8_t *raw, int nseq, int ncol, size_t pitch,
                                  unsigned long long *hist256, int num_sms, cudaStream_t stream);
cudaError_t launch_pack_planes(const uint8_t *raw, size_t pitch, int ncol, const int *kept_rows,
                               int nk, const uint8_t *col_drop, const uint8_t *lut256, int np,
                               int nb2, int nchunks, uint32_t *planes, uint8_t *gbytes,
                               cudaStream_t stream);
cudaError_t launch_identity2(int np, const Identity2Params &p, int num_sms, cudaStream_t stream);
cudaError_t launch_identity_bytes(const uint8_t *raw, size_t pitch, int ncol, const int *kept_rows,
                                  int nk, const uint8_t *col_drop, uint8_t indet, float *out,
                                  int *hit_out, int *dst_out, int num_sms, cudaStream_t stream);
cudaError_t launch_column_counts(const uint8_t *raw, int nseq, int ncol, size_t pitch,
                                 const uint8_t *row_drop, uint8_t sym_a, uint8_t sym_b,
                                 int *count_a, int *count_b, uint16_t *plane_a, uint16_t *plane_b,
                                 int num_sms, cudaStream_t stream);
cudaError_t launch_spurious_rows(const uint32_t *plane_gap, const uint32_t *plane_indet, int nseq,
                                 int row_begin, int row_end, int ncol, size_t pitch,
                                 const int *cnt_gap, const int *cnt_indet, uint32_t ovrlap,
                                 uint32_t *flag_words, float *out, int num_sms, cudaStream_t stream);
cudaError_t launch_sim_codes(const uint8_t *raw, int nseq, int ncol, size_t pitch, int npad,
                             const uint8_t *lut256, const uint8_t *col_skip, uint8_t *codesT,
                             unsigned long long *first_error, int num_sms, cudaStream_t stream);
cudaError_t launch_sim_rows(const uint8_t *codesT, int nseq, int npad, int ngroups,
                            uint32_t *skipbits, unsigned long long *nbatches, uint32_t *ngmask,
                            cudaStream_t stream);
cudaError_t launch_similarity(const uint8_t *codesT, int nseq, int npad, int ncol,
                              const float *identities, const float *dist, int npos,
                              const uint8_t *col_skip, const uint32_t *skipbits,
                              const uint32_t *ngmask, const unsigned long long *nbatches,
                              int group_begin, int group_end,
                              float *num, float *den, int num_sms, cudaStream_t stream);

cudaError_t launch_row_lengths(const uint8_t *raw, int nseq, int ncol, size_t pitch, int *lengths,
                               cudaStream_t stream);

cudaError_t launch_row_residues(const uint8_t *raw, int nseq, size_t pitch, const uint8_t *keep01,
                                int *residues, cudaStream_t stream);
cudaError_t launch_row_hashes(const uint8_t *raw, int nseq, size_t pitch, unsigned long long *hashes,
                              cudaStream_t stream);

// consumers of the device-resident identity matrix (clusters.cu)
cudaError_t launch_identity_bits(const float *id, int n, float thr, uint32_t *rows, int row_begin,
                                 int row_end, cudaStream_t stream);
cudaError_t launch_bits_rows(const uint32_t *slab_bits, int n, uint32_t *rows, cudaStream_t stream);
cudaError_t launch_row_stats(const float *id, int n, bool upper_only, float *row_max,
                             float *row_min, float *row_sum, cudaStream_t stream);
size_t greedy_scratch_bytes(int n, int total);
cudaError_t launch_greedy_clusters(const uint32_t *rows, int n, const int *order, int total,
                                   void *scratch, int *clusters, int *count, int num_sms,
                                   cudaStream_t stream);

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA unit, SASS: UBLKCP)
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
// fl(h / d), correctly rounded, for integers 0 <= h <= d and 0 < d < 2^16 (the counts of an
// alignment of fewer than 65 536 columns): reciprocal estimate, quotient estimate, its EXACT
// residual (h - d q is a multiple of ulp(q) below 2^17 of them, so the fused multiply-add
// returns it without rounding) and one correction.  q + r / d is the true quotient; using
// r * rc for r / d is off by less than 2^-22 ulp(q), while a quotient of such integers that
// is not a float is at least ulp(q) / 2^17 away from any rounding boundary -- so the single
// rounding of the last operation is the rounding of the exact quotient.  4 instructions
// against the 12 + range check + slow-path call of the general division.  Checked against
// __fdiv_rn for every such pair: tools/fdiv_check.cu, profiles/r03_fdiv_check.txt.
__device__ __forceinline__ float div_small_counts(float h, float d)
{
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(d));
    const float q = __fmul_rn(h, rc);
    const float r = __fmaf_rn(-d, q, h);
    return __fmaf_rn(r, rc, q);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "TCU_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra TCU_DONE;\n"
        "bra TCU_WAIT;\n"
        "TCU_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// The same two on a 32-bit shared-memory address computed once outside a hot loop (the
// conversion from a generic pointer costs a special-register read each time it is redone).
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "TCU_WAITA:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra TCU_DONEA;\n"
        "bra TCU_WAITA;\n"
        "TCU_DONEA:\n"
        "}\n" ::"r"(bar_addr),
        "r"(parity)
        : "memory");
}
// One non-blocking probe (about 90 cycles until the result is usable): issue it early,
// test the result later, fall back to mbar_wait when it is 0.
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
// Same wait with a suspend-time hint: the waiting thread is parked by the hardware
// until the phase completes (or the hint, in ns, expires) instead of re-polling.
// For the single-thread producer / MMA-issuer warps, whose polling would otherwise
// take issue slots from the math warps of the same SM sub-partition.
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "TCU_WAITP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra TCU_DONEP;\n"
        "bra TCU_WAITP;\n"
        "TCU_DONEP:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                              uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}
#endif

}  // namespace tcu
