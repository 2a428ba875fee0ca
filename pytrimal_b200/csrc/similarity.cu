// similarity.cu -- K4: per-column similarity accumulators.
//
// Replaces simd::calculateSimilarityVectors<V> (vendor/trimal/include/Platform/
// template.h:69-204).  The reference adds, for every column, the terms
//     num += (1 - id[j,k]) * D[a_j][a_k];   den += (1 - id[j,k])
// over all residue pairs j<k in lexicographic order with fp32 accumulators and
// no fused multiply-add (template.h:154-183; AVX2.cpp is built with -mavx2
// only).  fp32 addition is not associative and at n ~ 10^4 the running sum is
// far from the exact value, so the only way to reproduce the
// reference's bits is to replay its order: one lane owns one column and
// performs the same rounded additions in the same sequence
// (__fsub_rn / __fmul_rn / __fadd_rn keep nvcc from contracting them).
// Pairs the reference skips because a row holds a gap (template.h:157-160,
// 170-173) are replayed as "+ 0.0f", which is exact, so the chain is
// branch-free.  The kernel is bound by the dependent-add latency, not by
// bandwidth or FLOPs; no roofline fraction applies.  What CAN be
// parallel -- the terms -- is computed by other warps (see k_similarity2).
#include <algorithm>

#include "tcu_internal.cuh"

namespace tcu {

// ---------------------------------------------------------------------------
// Codes.  byte -> similarity code (template.h:129-150): upper-case, gap/indet ->
// gap, outside 'A'..'Z' -> SIM_INCORRECT, no matrix row -> SIM_UNDEFINED (LUT
// built on the host).  The first offending cell in the reference's scan order
// (columns ascending, rows ascending, skipped columns ignored) is found with a
// 64-bit atomicMin on ((col*nseq + row) << 8 | byte).
//
// Output layout: codesT[group][row][32] -- the 32 columns of one column group
// are contiguous per row, rows padded to a multiple of 32 -- holding 4 * code
// (the byte offset of the code's entry in a table row); the gap class, skipped
// columns and all padding hold SIM2_GAPCODE = 4 * 31.
// ---------------------------------------------------------------------------
constexpr uint32_t SIM2_GAPIDX = 31;
constexpr uint32_t SIM2_GAPCODE = 4 * SIM2_GAPIDX;

__global__ void __launch_bounds__(256) k_sim_codes(const uint8_t *__restrict__ raw, int nseq,
                                                   int ncol, size_t pitch, int npad,
                                                   const uint8_t *__restrict__ lut256,
                                                   const uint8_t *__restrict__ col_skip,
                                                   uint8_t *__restrict__ codesT,
                                                   unsigned long long *__restrict__ first_error)
{
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = lut256[threadIdx.x];
    __syncthreads();
    const int groups = (int)(pitch >> 4);  // 16-column groups
    const long long total = (long long)nseq * groups;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / groups);
        const int g = (int)(idx - (long long)r * groups);
        const uint4 v = *reinterpret_cast<const uint4 *>(raw + (size_t)r * pitch + (size_t)g * 16);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t ow = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int col = g * 16 + q * 4 + b;
                const uint32_t byte = (w[q] >> (8 * b)) & 0xFF;
                uint32_t out = SIM2_GAPCODE;
                if (col < ncol && !col_skip[col]) {
                    const uint32_t code = lut[byte];
                    if (code == SIM_INCORRECT || code == SIM_UNDEFINED) {
                        uint32_t up = (byte >= 'a' && byte <= 'z') ? byte - 32 : byte;
                        atomicMin(first_error, (((unsigned long long)col * nseq + r) << 8) | up);
                    } else if (code != SIM_GAP) {
                        out = code * 4;
                    }
                }
                ow |= out << (8 * b);
            }
            o[q] = ow;
        }
        uint8_t *dst = codesT + ((size_t)(g >> 1) * npad + r) * 32 + (size_t)(g & 1) * 16;
        *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

cudaError_t launch_sim_codes(const uint8_t *raw, int nseq, int ncol, size_t pitch, int npad,
                             const uint8_t *lut256, const uint8_t *col_skip, uint8_t *codesT,
                             unsigned long long *first_error, int num_sms, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0) return cudaSuccess;
    // padding rows (and nothing else survives the kernel below) are gaps
    cudaError_t e = cudaMemsetAsync(codesT, (int)SIM2_GAPCODE, (size_t)(pitch >> 5) * npad * 32, stream);
    if (e != cudaSuccess) return e;
    const long long total = (long long)nseq * (long long)(pitch >> 4);
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms * 8);
    k_sim_codes<<<blocks, 256, 0, stream>>>(raw, nseq, ncol, pitch, npad, lut256, col_skip, codesT,
                                            first_error);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Row prepass.  An outer row j whose 32 columns of the group are all gaps adds
// nothing in any lane (template.h:157-160) and is skipped by every warp of the
// main kernel: skipbits[group][j / 32] bit j % 32.  Row nseq-1 (never an outer
// row) and padding are marked too.  nbatches[group] = number of 32-k batches
// the main kernel exchanges for the group.  ngmask[group][row] = the row's
// non-gap columns of the group as a bit mask (two of them ANDed say in which
// columns a pair counts for the denominator).
// ---------------------------------------------------------------------------
constexpr int SIM2_KB = 32;  // inner rows (k) per batch

__device__ __forceinline__ int sim2_row_batches(int j, int nseq)
{
    return ((nseq - 1) >> 5) - ((j + 1) >> 5) + 1;  // aligned batches covering k = j+1 .. nseq-1
}

// non-gap bytes of a word of codes as a 4-bit mask.  Codes are 4 * index with
// index <= 31, the gap class is index 31: a byte is a gap iff bits 2..6 are all set.
__device__ __forceinline__ uint32_t sim2_nongap4(uint32_t w)
{
    const uint32_t x1 = w & (w >> 1);
    const uint32_t x2 = x1 & (x1 >> 2);
    const uint32_t g = x2 & (w >> 4);                 // bit 2 of each byte: gap
    const uint32_t y = (~g >> 2) & 0x01010101u;       // bit 0 of each byte: non-gap
    return (y * 0x01020408u) >> 24;                   // byte b -> bit b
}

__global__ void __launch_bounds__(256) k_sim_rows(const uint8_t *__restrict__ codesT, int nseq,
                                                  int npad, uint32_t *__restrict__ skipbits,
                                                  unsigned long long *__restrict__ nbatches,
                                                  uint32_t *__restrict__ ngmask)
{
    const int group = blockIdx.y;
    const int nwords = npad >> 5;
    const int word = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (word >= nwords) return;
    const int lane = threadIdx.x & 31;
    const int j = word * 32 + lane;
    const uint4 *cp = reinterpret_cast<const uint4 *>(codesT + ((size_t)group * npad + j) * 32);
    const uint4 a = __ldg(cp), b = __ldg(cp + 1);
    const uint32_t mask = sim2_nongap4(a.x) | sim2_nongap4(a.y) << 4 | sim2_nongap4(a.z) << 8 |
                          sim2_nongap4(a.w) << 12 | sim2_nongap4(b.x) << 16 |
                          sim2_nongap4(b.y) << 20 | sim2_nongap4(b.z) << 24 |
                          sim2_nongap4(b.w) << 28;
    ngmask[(size_t)group * npad + j] = mask;
    const bool skip = mask == 0 || j >= nseq - 1;
    const uint32_t bits = __ballot_sync(0xffffffffu, skip);
    unsigned long long nb = skip ? 0ull : (unsigned long long)sim2_row_batches(j, nseq);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nb += __shfl_xor_sync(0xffffffffu, nb, o);
    if (lane == 0) {
        skipbits[(size_t)group * nwords + word] = bits;
        if (nb) atomicAdd(&nbatches[group], nb);
    }
}

// ---------------------------------------------------------------------------
// Main kernel: one CTA per group of 32 adjacent columns.
//
// The chains  num += w * d;  den += w  of a column must run in the reference's
// order, one rounded fp32 add after the other, but the two chains are
// independent of each other and their TERMS are independent.  The twelve warps of a CTA
// get different jobs by SM sub-partition (warp id % 4):
//   numerator warp,    lane = column.  Nothing but LDS.128 (four consecutive inner
//   denominator warp   rows of the lane's column) and one FADD per inner row: the
//                      dependent add (4 cycles) is the critical path.  Both sit on
//                      one sub-partition, which they share with nobody: each chain
//                      leaves three of four issue slots free, so the two interleave.
//                      The denominator's terms are w where the pair counts in the
//                      column and +0 elsewhere.
//   9 producer warps   (the three other sub-partitions; forming the terms is issue-bound,
//                      ~370 instructions per batch) lane = inner row k of a 32-k batch.
//                      Per batch a lane loads its row's 32 codes (32 contiguous
//                      bytes), id[j,k] (coalesced) and the two non-gap masks, forms
//                      w = 1 - id once, and for each of the 32 columns looks up
//                      D[a_j][a_k] (one PRMT forms the table offset; the table has an
//                      all-zero row/column for gaps), multiplies by w and stores both
//                      terms to its ring slot in shared memory ([column][k], so the
//                      consumers read vectors).  One slot per producer: batch g goes
//                      to slot g % 9.
// Terms of pairs the reference skips are exact +0.
// A single warp can start one shared-memory load every ~4 cycles, which is what
// limited the first version of this kernel (one warp, LDS.64 + two FADDs per row:
// 10.6 cycles per row); hence the vector loads and one chain per warp.
// Slots are handed over with mbarriers (full: one arrival; empty: two).
// ---------------------------------------------------------------------------
constexpr int SIM2_NPROD = 9;
constexpr int SIM2_MAX_SLOTS = 9;
constexpr int SIM2_CS = 36;                          // floats per column of a slot: [column][num 32 | pad 4],
                                                     // 8 lanes x LDS.128 cover the 32 banks (36 % 32 == 4)
constexpr int SIM2_W_OFF = 32 * SIM2_CS;             // then the batch's 32 weights w = 1 - id (one per inner row)
constexpr int SIM2_M_OFF = SIM2_W_OFF + 32;          // and one word per column: in which inner rows the pair counts
constexpr int SIM2_SLOT_WORDS = SIM2_M_OFF + 32;
constexpr int SIM2_THREADS = 384;                    // 3 warps per sub-partition
constexpr int SIM2_TROW = 64;                        // table row stride in floats (256 B: offset = a_j << 8 | 4 a_k)
constexpr int SIM2_TABLE_WORDS = 32 * SIM2_TROW;
constexpr int SIM2_PREFETCH_BATCHES = 6;             // own batches ahead (x SIM2_NPROD in array order)

__host__ __device__ constexpr size_t sim2_smem_bytes(int slots)
{
    return (size_t)slots * SIM2_SLOT_WORDS * 4 + SIM2_TABLE_WORDS * 4 + 2 * SIM2_MAX_SLOTS * 8;
}

struct Sim2Params {
    const uint8_t *codesT;
    const float *identities;
    const float *dist;
    const uint8_t *col_skip;
    const uint32_t *skipbits;
    const uint32_t *ngmask;
    const unsigned long long *nbatches;
    float *num_out, *den_out;
    int nseq, npad, ncol, npos;
    int group_begin;
    int slots;      // multiple of SIM2_NPROD
    uint32_t zero;  // 0 (opaque to the compiler, see sim2_consume)
    int num_sms;
};

// The consumer warps' walk over the ring.  Common scheme: the data of a batch (32 inner
// rows) are loaded into registers well before they are added, because a shared-memory load
// takes 40-100 cycles to return while the producers keep the pipe busy; the loads follow
// the first addition of the block they are issued in -- their address is tied to its
// result through `zero`, a kernel argument that is 0 -- so that (a) that addition waits
// for its own operands only, not for the new loads on the same scoreboard, and (b) the
// loads (one per ~4 cycles from one warp) issue in the shadow of the dependent additions.
// The barrier of the batch after the next is probed at the top of a batch (the answer takes
// ~90 cycles) and looked at after the chain; the last three batches run without probes.
struct Sim2Walk {
    const float *ring;
    uint64_t *full, *empty;
    int slots;
    int s0 = 0, s1 = 1, s2 = 2;   // slots of the current batch and the two after it
    uint32_t p1 = 0, p2 = 0;      // phase parities of s1 and s2
    __device__ __forceinline__ void rotate()
    {
        s0 = s1;
        s1 = s2;
        p1 = p2;
        if (++s2 == slots) {
            s2 = 0;
            p2 ^= 1u;
        }
    }
};

// ---- one chain: acc += term, 32 terms of the lane's column per batch.  `off` selects the
// numerator (0) or denominator (SIM2_KB) terms of the [column][2][k] slot layout.
struct Sim2NumBatch {
    float4 a[8];
};
__device__ __forceinline__ void sim2_num_load(Sim2NumBatch &B, const float *ring, int slot, int lane,
                                              uint32_t dep)
{
    // one LDS.128 = four consecutive inner rows of the lane's column; `lane` is the word
    // offset of the lane's terms inside a slot, dep is always 0
    const float4 *src =
        reinterpret_cast<const float4 *>(ring + (size_t)slot * SIM2_SLOT_WORDS + lane + dep);
#pragma unroll
    for (int q = 0; q < 8; q++) B.a[q] = src[q];
}
template <int K0, int K1>
__device__ __forceinline__ float sim2_num_add(float acc, const Sim2NumBatch &B)
{
#pragma unroll
    for (int k = K0; k < K1; k++) {
        const float4 v = B.a[k >> 2];
        acc = __fadd_rn(acc, (k & 3) == 0 ? v.x : (k & 3) == 1 ? v.y : (k & 3) == 2 ? v.z : v.w);
    }
    return acc;
}

// order-pinned forms (asm volatile keeps the relative order of these statements through the
// compiler; ptxas schedules from that order)
__device__ __forceinline__ float v_fadd(float a, float b)
{
    float d;
    asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float4 v_lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t v_lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// One batch of a chain, refilling the registers in place: as soon as the four terms of an
// LDS.128 have been added, the same four registers are loaded with the corresponding terms of
// the NEXT batch.  A batch therefore lives in 32 registers (not 2 x 32), every load has a whole
// batch (128 cycles of dependent additions) to return, and the loads issue in the shadow of
// the chain without being tied to its results.
//   numerator  : R = the lane's (column's) 32 terms w * D
//   denominator: R = the batch's 32 weights w (the same for every column: broadcast loads, one
//                wavefront each) and `mask` says in which inner rows the lane's column counts;
//                a skipped row is a skipped (predicated-off) addition, i.e. + 0 exactly.
template <bool DEN>
__device__ __forceinline__ float sim2_chain_refill(float acc, Sim2NumBatch &R, uint32_t mask,
                                                   uint32_t next_addr)
{
#pragma unroll
    for (int q = 0; q < 8; q++) {
        if (!DEN || (mask & (1u << (4 * q + 0)))) acc = v_fadd(acc, R.a[q].x);
        if (!DEN || (mask & (1u << (4 * q + 1)))) acc = v_fadd(acc, R.a[q].y);
        if (!DEN || (mask & (1u << (4 * q + 2)))) acc = v_fadd(acc, R.a[q].z);
        if (!DEN || (mask & (1u << (4 * q + 3)))) acc = v_fadd(acc, R.a[q].w);
        R.a[q] = v_lds128(next_addr + 16u * q);
    }
    return acc;
}
template <bool DEN>
__device__ __forceinline__ float sim2_chain(float acc, const Sim2NumBatch &R, uint32_t mask)
{
#pragma unroll
    for (int q = 0; q < 8; q++) {
        if (!DEN || (mask & (1u << (4 * q + 0)))) acc = v_fadd(acc, R.a[q].x);
        if (!DEN || (mask & (1u << (4 * q + 1)))) acc = v_fadd(acc, R.a[q].y);
        if (!DEN || (mask & (1u << (4 * q + 2)))) acc = v_fadd(acc, R.a[q].z);
        if (!DEN || (mask & (1u << (4 * q + 3)))) acc = v_fadd(acc, R.a[q].w);
    }
    return acc;
}


// The denominator chain of one batch as ONE block of PTX: a skipped inner row must cost nothing
// in the dependent chain, so every addition is predicated on its bit of the column's mask, and the
// predicates are formed well ahead of their use -- two rotating sets (4 + 3 predicate registers):
// while the additions of one set issue, the other set already holds the bits of the next inner
// rows, and is reloaded right after its last addition.  (Left to the compiler the tests become
// one R2P per byte of the mask immediately in front of the eight additions that need it, whose
// latency then sits in the chain four times per batch.)  With `refill`, the 32 weights of the
// next batch are loaded in place as in sim2_chain_refill.
__device__ __forceinline__ float sim2_den_chain_refill(float acc, Sim2NumBatch &R, uint32_t mask,
                                                       uint32_t next_addr)
{
    asm volatile(
        "{\n"
        ".reg .pred a0, a1, a2, a3, b0, b1, b2;\n"
        ".reg .b32 t;\n"
        "and.b32 t, %33, 1; setp.ne.u32 a0, t, 0;\n"
        "and.b32 t, %33, 2; setp.ne.u32 a1, t, 0;\n"
        "and.b32 t, %33, 4; setp.ne.u32 a2, t, 0;\n"
        "and.b32 t, %33, 8; setp.ne.u32 a3, t, 0;\n"
        "and.b32 t, %33, 16; setp.ne.u32 b0, t, 0;\n"
        "and.b32 t, %33, 32; setp.ne.u32 b1, t, 0;\n"
        "and.b32 t, %33, 64; setp.ne.u32 b2, t, 0;\n"
        "@a0 add.rn.f32 %0, %0, %1;\n"
        "@a1 add.rn.f32 %0, %0, %2;\n"
        "@a2 add.rn.f32 %0, %0, %3;\n"
        "@a3 add.rn.f32 %0, %0, %4;\n"
        "ld.shared.v4.f32 {%1, %2, %3, %4}, [%34+0];\n"
        "and.b32 t, %33, 128; setp.ne.u32 a0, t, 0;\n"
        "and.b32 t, %33, 256; setp.ne.u32 a1, t, 0;\n"
        "and.b32 t, %33, 512; setp.ne.u32 a2, t, 0;\n"
        "and.b32 t, %33, 1024; setp.ne.u32 a3, t, 0;\n"
        "@b0 add.rn.f32 %0, %0, %5;\n"
        "@b1 add.rn.f32 %0, %0, %6;\n"
        "@b2 add.rn.f32 %0, %0, %7;\n"
        "and.b32 t, %33, 2048; setp.ne.u32 b0, t, 0;\n"
        "and.b32 t, %33, 4096; setp.ne.u32 b1, t, 0;\n"
        "and.b32 t, %33, 8192; setp.ne.u32 b2, t, 0;\n"
        "@a0 add.rn.f32 %0, %0, %8;\n"
        "ld.shared.v4.f32 {%5, %6, %7, %8}, [%34+16];\n"
        "@a1 add.rn.f32 %0, %0, %9;\n"
        "@a2 add.rn.f32 %0, %0, %10;\n"
        "@a3 add.rn.f32 %0, %0, %11;\n"
        "and.b32 t, %33, 16384; setp.ne.u32 a0, t, 0;\n"
        "and.b32 t, %33, 32768; setp.ne.u32 a1, t, 0;\n"
        "and.b32 t, %33, 65536; setp.ne.u32 a2, t, 0;\n"
        "and.b32 t, %33, 131072; setp.ne.u32 a3, t, 0;\n"
        "@b0 add.rn.f32 %0, %0, %12;\n"
        "ld.shared.v4.f32 {%9, %10, %11, %12}, [%34+32];\n"
        "@b1 add.rn.f32 %0, %0, %13;\n"
        "@b2 add.rn.f32 %0, %0, %14;\n"
        "and.b32 t, %33, 262144; setp.ne.u32 b0, t, 0;\n"
        "and.b32 t, %33, 524288; setp.ne.u32 b1, t, 0;\n"
        "and.b32 t, %33, 1048576; setp.ne.u32 b2, t, 0;\n"
        "@a0 add.rn.f32 %0, %0, %15;\n"
        "@a1 add.rn.f32 %0, %0, %16;\n"
        "ld.shared.v4.f32 {%13, %14, %15, %16}, [%34+48];\n"
        "@a2 add.rn.f32 %0, %0, %17;\n"
        "@a3 add.rn.f32 %0, %0, %18;\n"
        "and.b32 t, %33, 2097152; setp.ne.u32 a0, t, 0;\n"
        "and.b32 t, %33, 4194304; setp.ne.u32 a1, t, 0;\n"
        "and.b32 t, %33, 8388608; setp.ne.u32 a2, t, 0;\n"
        "and.b32 t, %33, 16777216; setp.ne.u32 a3, t, 0;\n"
        "@b0 add.rn.f32 %0, %0, %19;\n"
        "@b1 add.rn.f32 %0, %0, %20;\n"
        "ld.shared.v4.f32 {%17, %18, %19, %20}, [%34+64];\n"
        "@b2 add.rn.f32 %0, %0, %21;\n"
        "and.b32 t, %33, 33554432; setp.ne.u32 b0, t, 0;\n"
        "and.b32 t, %33, 67108864; setp.ne.u32 b1, t, 0;\n"
        "and.b32 t, %33, 134217728; setp.ne.u32 b2, t, 0;\n"
        "@a0 add.rn.f32 %0, %0, %22;\n"
        "@a1 add.rn.f32 %0, %0, %23;\n"
        "@a2 add.rn.f32 %0, %0, %24;\n"
        "ld.shared.v4.f32 {%21, %22, %23, %24}, [%34+80];\n"
        "@a3 add.rn.f32 %0, %0, %25;\n"
        "and.b32 t, %33, 268435456; setp.ne.u32 a0, t, 0;\n"
        "and.b32 t, %33, 536870912; setp.ne.u32 a1, t, 0;\n"
        "and.b32 t, %33, 1073741824; setp.ne.u32 a2, t, 0;\n"
        "and.b32 t, %33, 2147483648; setp.ne.u32 a3, t, 0;\n"
        "@b0 add.rn.f32 %0, %0, %26;\n"
        "@b1 add.rn.f32 %0, %0, %27;\n"
        "@b2 add.rn.f32 %0, %0, %28;\n"
        "ld.shared.v4.f32 {%25, %26, %27, %28}, [%34+96];\n"
        "@a0 add.rn.f32 %0, %0, %29;\n"
        "@a1 add.rn.f32 %0, %0, %30;\n"
        "@a2 add.rn.f32 %0, %0, %31;\n"
        "@a3 add.rn.f32 %0, %0, %32;\n"
        "ld.shared.v4.f32 {%29, %30, %31, %32}, [%34+112];\n"
        "}\n"
        : "+f"(acc), "+f"(R.a[0].x), "+f"(R.a[0].y), "+f"(R.a[0].z), "+f"(R.a[0].w), "+f"(R.a[1].x), "+f"(R.a[1].y), "+f"(R.a[1].z), "+f"(R.a[1].w), "+f"(R.a[2].x), "+f"(R.a[2].y), "+f"(R.a[2].z), "+f"(R.a[2].w), "+f"(R.a[3].x), "+f"(R.a[3].y), "+f"(R.a[3].z), "+f"(R.a[3].w), "+f"(R.a[4].x), "+f"(R.a[4].y), "+f"(R.a[4].z), "+f"(R.a[4].w), "+f"(R.a[5].x), "+f"(R.a[5].y), "+f"(R.a[5].z), "+f"(R.a[5].w), "+f"(R.a[6].x), "+f"(R.a[6].y), "+f"(R.a[6].z), "+f"(R.a[6].w), "+f"(R.a[7].x), "+f"(R.a[7].y), "+f"(R.a[7].z), "+f"(R.a[7].w)
        : "r"(mask), "r"(next_addr)
        : "memory");
    return acc;
}
__device__ __forceinline__ float sim2_den_chain(float acc, Sim2NumBatch &R, uint32_t mask)
{
    asm volatile(
        "{\n"
        ".reg .pred a0, a1, a2, a3, b0, b1, b2;\n"
        ".reg .b32 t;\n"
        "and.b32 t, %33, 1; setp.ne.u32 a0, t, 0;\n"
        "and.b32 t, %33, 2; setp.ne.u32 a1, t, 0;\n"
        "and.b32 t, %33, 4; setp.ne.u32 a2, t, 0;\n"
        "and.b32 t, %33, 8; setp.ne.u32 a3, t, 0;\n"
        "and.b32 t, %33, 16; setp.ne.u32 b0, t, 0;\n"
        "and.b32 t, %33, 32; setp.ne.u32 b1, t, 0;\n"
        "and.b32 t, %33, 64; setp.ne.u32 b2, t, 0;\n"
        "@a0 add.rn.f32 %0, %0, %1;\n"
        "@a1 add.rn.f32 %0, %0, %2;\n"
        "@a2 add.rn.f32 %0, %0, %3;\n"
        "@a3 add.rn.f32 %0, %0, %4;\n"
        "and.b32 t, %33, 128; setp.ne.u32 a0, t, 0;\n"
        "and.b32 t, %33, 256; setp.ne.u32 a1, t, 0;\n"
        "and.b32 t, %33, 512; setp.ne.u32 a2, t, 0;\n"
        "and.b32 t, %33, 1024; setp.ne.u32 a3, t, 0;\n"
        "@b0 add.rn.f32 %0, %0, %5;\n"
        "@b1 add.rn.f32 %0, %0, %6;\n"
        "@b2 add.rn.f32 %0, %0, %7;\n"
        "and.b32 t, %33, 2048; setp.ne.u32 b0, t, 0;\n"
        "and.b32 t, %33, 4096; setp.ne.u32 b1, t, 0;\n"
        "and.b32 t, %33, 8192; setp.ne.u32 b2, t, 0;\n"
        "@a0 add.rn.f32 %0, %0, %8;\n"
        "@a1 add.rn.f32 %0, %0, %9;\n"
        "@a2 add.rn.f32 %0, %0, %10;\n"
        "@a3 add.rn.f32 %0, %0, %11;\n"
        "and.b32 t, %33, 16384; setp.ne.u32 a0, t, 0;\n"
        "and.b32 t, %33, 32768; setp.ne.u32 a1, t, 0;\n"
        "and.b32 t, %33, 65536; setp.ne.u32 a2, t, 0;\n"
        "and.b32 t, %33, 131072; setp.ne.u32 a3, t, 0;\n"
        "@b0 add.rn.f32 %0, %0, %12;\n"
        "@b1 add.rn.f32 %0, %0, %13;\n"
        "@b2 add.rn.f32 %0, %0, %14;\n"
        "and.b32 t, %33, 262144; setp.ne.u32 b0, t, 0;\n"
        "and.b32 t, %33, 524288; setp.ne.u32 b1, t, 0;\n"
        "and.b32 t, %33, 1048576; setp.ne.u32 b2, t, 0;\n"
        "@a0 add.rn.f32 %0, %0, %15;\n"
        "@a1 add.rn.f32 %0, %0, %16;\n"
        "@a2 add.rn.f32 %0, %0, %17;\n"
        "@a3 add.rn.f32 %0, %0, %18;\n"
        "and.b32 t, %33, 2097152; setp.ne.u32 a0, t, 0;\n"
        "and.b32 t, %33, 4194304; setp.ne.u32 a1, t, 0;\n"
        "and.b32 t, %33, 8388608; setp.ne.u32 a2, t, 0;\n"
        "and.b32 t, %33, 16777216; setp.ne.u32 a3, t, 0;\n"
        "@b0 add.rn.f32 %0, %0, %19;\n"
        "@b1 add.rn.f32 %0, %0, %20;\n"
        "@b2 add.rn.f32 %0, %0, %21;\n"
        "and.b32 t, %33, 33554432; setp.ne.u32 b0, t, 0;\n"
        "and.b32 t, %33, 67108864; setp.ne.u32 b1, t, 0;\n"
        "and.b32 t, %33, 134217728; setp.ne.u32 b2, t, 0;\n"
        "@a0 add.rn.f32 %0, %0, %22;\n"
        "@a1 add.rn.f32 %0, %0, %23;\n"
        "@a2 add.rn.f32 %0, %0, %24;\n"
        "@a3 add.rn.f32 %0, %0, %25;\n"
        "and.b32 t, %33, 268435456; setp.ne.u32 a0, t, 0;\n"
        "and.b32 t, %33, 536870912; setp.ne.u32 a1, t, 0;\n"
        "and.b32 t, %33, 1073741824; setp.ne.u32 a2, t, 0;\n"
        "and.b32 t, %33, 2147483648; setp.ne.u32 a3, t, 0;\n"
        "@b0 add.rn.f32 %0, %0, %26;\n"
        "@b1 add.rn.f32 %0, %0, %27;\n"
        "@b2 add.rn.f32 %0, %0, %28;\n"
        "@a0 add.rn.f32 %0, %0, %29;\n"
        "@a1 add.rn.f32 %0, %0, %30;\n"
        "@a2 add.rn.f32 %0, %0, %31;\n"
        "@a3 add.rn.f32 %0, %0, %32;\n"
        "}\n"
        : "+f"(acc), "+f"(R.a[0].x), "+f"(R.a[0].y), "+f"(R.a[0].z), "+f"(R.a[0].w), "+f"(R.a[1].x), "+f"(R.a[1].y), "+f"(R.a[1].z), "+f"(R.a[1].w), "+f"(R.a[2].x), "+f"(R.a[2].y), "+f"(R.a[2].z), "+f"(R.a[2].w), "+f"(R.a[3].x), "+f"(R.a[3].y), "+f"(R.a[3].z), "+f"(R.a[3].w), "+f"(R.a[4].x), "+f"(R.a[4].y), "+f"(R.a[4].z), "+f"(R.a[4].w), "+f"(R.a[5].x), "+f"(R.a[5].y), "+f"(R.a[5].z), "+f"(R.a[5].w), "+f"(R.a[6].x), "+f"(R.a[6].y), "+f"(R.a[6].z), "+f"(R.a[6].w), "+f"(R.a[7].x), "+f"(R.a[7].y), "+f"(R.a[7].z), "+f"(R.a[7].w)
        : "r"(mask), "r"(0u));
    return acc;
}

template <bool DEN>
__device__ __forceinline__ float sim2_consume(Sim2Walk W, unsigned long long btot, int lane)
{
    float acc = 0.0f;
    if (btot == 0) return acc;
    // lane id and the lane's byte addresses inside slot 0, made opaque once: the compiler would
    // otherwise re-derive them (S2R / S2UR, ~25 cycles each) in every iteration
    uint32_t lane0, ring0, mask0 = 0;
    {
        const uint32_t base = smem_u32(W.ring);
        const uint32_t a = base + 4u * (uint32_t)(DEN ? SIM2_W_OFF : lane * SIM2_CS);
        const uint32_t m = base + 4u * (uint32_t)(SIM2_M_OFF + lane);
        asm volatile("mov.u32 %0, %1;" : "=r"(ring0) : "r"(a));
        asm volatile("mov.u32 %0, %1;" : "=r"(mask0) : "r"(m));
        asm volatile("mov.u32 %0, %1;" : "=r"(lane0) : "r"((uint32_t)(lane == 0)));
    }
    constexpr uint32_t SLOT_B = SIM2_SLOT_WORDS * 4u;
    Sim2NumBatch R;
    uint32_t mask = 0;
    mbar_wait(&W.full[0], 0);
#pragma unroll
    for (int q = 0; q < 8; q++) R.a[q] = v_lds128(ring0 + 16u * q);
    if (DEN) mask = v_lds32(mask0);
    if (btot > 1) mbar_wait(&W.full[1], 0);
    // invariant at the top of a batch: its terms are in R (and mask), the barrier of the next
    // batch (if any) has been seen complete
    auto iter = [&]() {   // batches b+1 and b+2 exist
        uint64_t *const f2 = &W.full[W.s2], *const e0 = &W.empty[W.s0];
        const uint32_t par2 = W.p2;
        const uint32_t ready2 = mbar_try_wait(f2, par2);
        const uint32_t next_off = (uint32_t)W.s1 * SLOT_B;
        uint32_t next_mask = 0;
        if (DEN) next_mask = v_lds32(mask0 + next_off);
        W.rotate();
        if (DEN) acc = sim2_den_chain_refill(acc, R, mask, ring0 + next_off);
        else acc = sim2_chain_refill<false>(acc, R, 0u, ring0 + next_off);
        mask = next_mask;
        if (!ready2) mbar_wait(f2, par2);
        __syncwarp();
        if (lane0) mbar_arrive(e0);
    };
    // Whole turns of the ring (SIM2_MAX_SLOTS batches, starting at slot 0) with the slot numbers
    // as compile-time constants: no slot arithmetic, no address computation and one loop branch
    // per nine batches in the warp that can least afford extra instructions.  `turn` is the
    // phase parity of the slots of the current turn.
    unsigned long long remaining = btot;
    if (W.slots == SIM2_MAX_SLOTS) {
        uint32_t turn = 0;
        unsigned long long turns = remaining > 2 ? (remaining - 2) / SIM2_MAX_SLOTS : 0;
        remaining -= turns * SIM2_MAX_SLOTS;
        while (turns) {
            const uint32_t chunk = (uint32_t)min(turns, 1ull << 30);
            turns -= chunk;
#pragma unroll 1
            for (uint32_t t = 0; t < chunk; t++) {
#pragma unroll
                for (int i = 0; i < SIM2_MAX_SLOTS; i++) {
                    constexpr int S = SIM2_MAX_SLOTS;
                    const int i1 = (i + 1) % S, i2 = (i + 2) % S;
                    const uint32_t par2 = i + 2 >= S ? turn ^ 1u : turn;
                    const uint32_t ready2 = mbar_try_wait(&W.full[i2], par2);
                    const uint32_t next_off = (uint32_t)i1 * SLOT_B;
                    uint32_t next_mask = 0;
                    if (DEN) next_mask = v_lds32(mask0 + next_off);
                    if (DEN) acc = sim2_den_chain_refill(acc, R, mask, ring0 + next_off);
                    else acc = sim2_chain_refill<false>(acc, R, 0u, ring0 + next_off);
                    mask = next_mask;
                    if (!ready2) mbar_wait(&W.full[i2], par2);
                    __syncwarp();
                    if (lane0) mbar_arrive(&W.empty[i]);
                }
                turn ^= 1u;
            }
        }
        // back to the general walk at slot 0 of the next turn
        W.s0 = 0;
        W.s1 = 1;
        W.s2 = 2;
        W.p1 = W.p2 = turn;
    }
    unsigned long long pairs = remaining > 3 ? (remaining - 2) / 2 : 0;
    remaining -= 2 * pairs;
    while (pairs) {
        const uint32_t chunk = (uint32_t)min(pairs, 1ull << 30);
        pairs -= chunk;
#pragma unroll 1
        for (uint32_t i = 0; i < chunk; i++) {
            iter();
            iter();
        }
    }
    // the last (up to three) batches: refill while a next batch exists, no more probes
    while (remaining) {
        if (remaining > 1) {
            const uint32_t next_off = (uint32_t)W.s1 * SLOT_B;
            uint32_t next_mask = 0;
            if (DEN) next_mask = v_lds32(mask0 + next_off);
            if (DEN) acc = sim2_den_chain_refill(acc, R, mask, ring0 + next_off);
            else acc = sim2_chain_refill<false>(acc, R, 0u, ring0 + next_off);
            mask = next_mask;
        } else {
            if (DEN) acc = sim2_den_chain(acc, R, mask);
            else acc = sim2_chain<false>(acc, R, 0u);
        }
        if (remaining > 2) mbar_wait(&W.full[W.s2], W.p2);
        __syncwarp();
        if (lane0) mbar_arrive(&W.empty[W.s0]);
        W.rotate();
        remaining--;
    }
    return acc;
}

__global__ void __launch_bounds__(SIM2_THREADS, 2) k_similarity2(const Sim2Params p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    float *ring = reinterpret_cast<float *>(smem);
    float *T = ring + (size_t)p.slots * SIM2_SLOT_WORDS;
    uint64_t *full = reinterpret_cast<uint64_t *>(T + SIM2_TABLE_WORDS);
    uint64_t *empty = full + SIM2_MAX_SLOTS;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = p.group_begin + blockIdx.x;
    const int n = p.nseq;

    // table: T[a * 64 + b] = D[a][b]; row / column 31 (gap) and unused codes are zero
    for (int i = threadIdx.x; i < SIM2_TABLE_WORDS; i += SIM2_THREADS) {
        const int a = i / SIM2_TROW, b = i % SIM2_TROW;
        T[i] = (a < p.npos && b < p.npos) ? p.dist[a * p.npos + b] : 0.0f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.slots; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 2);
        }
        fence_mbar_init();
    }
    __syncthreads();

    // When more groups than SMs are launched two CTAs share an SM: put their consumer
    // warps on different sub-partitions (warp id % 4).
    const int cn = (2 * (blockIdx.x / max(p.num_sms, 1))) & 3;  // the consumers' sub-partition:
    const int cd = cn + 4;                                      // numerator warp cn, denominator warp cn + 4
    const int sp = warp & 3;
    const uint8_t *gc = p.codesT + (size_t)group * p.npad * 32;

    if (warp == cn || warp == cd) {
        // ------------------------------ consumers -----------------------------
        Sim2Walk W;
        W.ring = ring;
        W.full = full;
        W.empty = empty;
        W.slots = p.slots;
        const float acc = warp == cn ? sim2_consume<false>(W, p.nbatches[group], lane)
                                     : sim2_consume<true>(W, p.nbatches[group], lane);
        const int col = group * 32 + lane;
        if (col < p.ncol && !p.col_skip[col]) (warp == cn ? p.num_out : p.den_out)[col] = acc;
    } else if (sp != cn) {
        // ------------------------------ producers -----------------------------
        // producer index 0..8: the nine warps of the three sub-partitions without a consumer
        // (the third warp of the consumers' sub-partition exits)
        int pi = 0;
        for (int w = 0; w < warp; w++) pi += (w & 3) != cn;
        constexpr int spp = SIM2_MAX_SLOTS / SIM2_NPROD;  // slots per producer (p.slots == SIM2_MAX_SLOTS)
        const int nwords = p.npad >> 5;
        const uint32_t *skipw = p.skipbits + (size_t)group * nwords;
        const uint32_t *ngm = p.ngmask + (size_t)group * p.npad;
        const unsigned long long nn = (unsigned long long)n;
        const unsigned long long npairs = nn * (nn - 1) / 2;

        // iterator over this producer's batches in consumption order
        int word = -1, j = 0, q = 0, nbj = 0, first = 0, b0m = 0;
        uint32_t bits = 0;
        bool have_row = false;
        auto next = [&](int &oj, int &okb) -> bool {
            for (;;) {
                if (have_row) {
                    if (q < nbj) {
                        oj = j;
                        okb = (first + q) * SIM2_KB;
                        q += SIM2_NPROD;
                        return true;
                    }
                    b0m = (b0m + nbj) % SIM2_NPROD;
                    have_row = false;
                }
                while (bits == 0) {
                    if (++word >= nwords) return false;
                    bits = ~__ldg(skipw + word);
                }
                const int r = __ffs(bits) - 1;
                bits &= bits - 1;
                j = word * 32 + r;
                first = (j + 1) >> 5;
                nbj = sim2_row_batches(j, n);
                q = (pi - b0m + SIM2_NPROD) % SIM2_NPROD;
                have_row = true;
            }
        };
        struct Loaded {
            uint4 c0, c1, j0, j1;
            float id;
            uint32_t mk, mj;  // non-gap columns of row k and of row j (ANDed at use: combining them
                              // here would make the prefetch wait for both loads)
        };
        auto fetch = [&](int fj, int fkb, Loaded &L) {
            const int k = fkb + lane;
            const uint4 *cp = reinterpret_cast<const uint4 *>(gc + (size_t)k * 32);
            const uint4 *jp = reinterpret_cast<const uint4 *>(gc + (size_t)fj * 32);
            L.c0 = __ldg(cp);
            L.c1 = __ldg(cp + 1);
            L.j0 = __ldg(jp);
            L.j1 = __ldg(jp + 1);
            L.mk = __ldg(ngm + k);
            L.mj = __ldg(ngm + fj);
            // identities[(fj, k)], packed upper triangle without diagonal (template.h:158,171,181)
            const unsigned long long rowbase =
                (unsigned long long)fj * nn - ((unsigned long long)fj * (fj + 1)) / 2 - fj - 1;
            const bool valid = k > fj && k < n;
            L.id = valid ? __ldg(p.identities + rowbase + k) : 1.0f;
            // the packed array is consumed front to back: pull the lines this warp will
            // want a few batches from now into L2 (HBM latency >> one batch)
            const unsigned long long ahead = rowbase + k + SIM2_PREFETCH_BATCHES * SIM2_NPROD * SIM2_KB;
            if (valid && ahead < npairs)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.identities + ahead));
        };

        int cj, ckb, nj = 0, nkb = 0;
        Loaded cur, nxt;
        bool have = next(cj, ckb);
        if (have) fetch(cj, ckb, cur);
        int i = 0;  // local batch count
        while (have) {
            const bool have_n = next(nj, nkb);
            if (have_n) fetch(nj, nkb, nxt);

            const int slot = pi + SIM2_NPROD * (i % spp);
            const uint32_t par = (uint32_t)((i / spp) & 1);
            mbar_wait(&empty[slot], par ^ 1u);

            const uint32_t cw8[8] = {cur.c0.x, cur.c0.y, cur.c0.z, cur.c0.w,
                                     cur.c1.x, cur.c1.y, cur.c1.z, cur.c1.w};
            const uint32_t jw8[8] = {cur.j0.x, cur.j0.y, cur.j0.z, cur.j0.w,
                                     cur.j1.x, cur.j1.y, cur.j1.z, cur.j1.w};
            float *dst = ring + (size_t)slot * SIM2_SLOT_WORDS + lane;
            const float w = __fsub_rn(1.0f, cur.id);  // 0 for k <= j and padding
            const uint32_t counted = cur.mk & cur.mj;  // columns in which the pair counts
            const char *Tb = reinterpret_cast<const char *>(T);
            // 16 table loads in flight, then the stores (the compiler must assume that the
            // ring stores alias the table and would otherwise serialise load -> store)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float t[16];
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const int wd = h * 4 + (u >> 2), b = u & 3;
                    // row-j codes (uniform): bytes [c0, 0, c2, 0] and [c1, 0, c3, 0] of code index
                    const uint32_t cjs = (jw8[wd] >> 2) & 0x1F1F1F1Fu;
                    const uint32_t y = (b & 1) ? ((cjs >> 8) & 0x00FF00FFu) : (cjs & 0x00FF00FFu);
                    // table byte offset = a_j * 256 + 4 * a_k: one PRMT
                    const uint32_t sel = 0x5500u | ((b & 2) ? 0x60u : 0x40u) | (uint32_t)b;
                    const uint32_t off = __byte_perm(cw8[wd], y, sel);
                    t[u] = *reinterpret_cast<const float *>(Tb + off);
                }
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const int c = h * 16 + u;
                    dst[c * SIM2_CS] = __fmul_rn(w, t[u]);                              // numerator term
                }
            }
            // denominator: the weight of this inner row once, and -- instead of 32 x 32 float
            // terms -- one word per column saying in which inner rows the pair counts: the
            // rows' masks (lane = inner row, bit = column) transposed across the warp
            {
                uint32_t x = counted;
                uint32_t low = 0x0000FFFFu;
#pragma unroll
                for (int jj = 16; jj >= 1; jj >>= 1) {
                    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, jj);
                    x = (lane & jj) ? ((x & ~low) | ((y & ~low) >> jj)) : ((x & low) | ((y & low) << jj));
                    low ^= low << (jj >> 1);
                }
                float *slot0 = ring + (size_t)slot * SIM2_SLOT_WORDS;
                slot0[SIM2_W_OFF + lane] = w;
                reinterpret_cast<uint32_t *>(slot0)[SIM2_M_OFF + lane] = x;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[slot]);
            i++;
            cur = nxt;
            cj = nj;
            ckb = nkb;
            have = have_n;
        }
    }
}

cudaError_t launch_sim_rows(const uint8_t *codesT, int nseq, int npad, int ngroups,
                            uint32_t *skipbits, unsigned long long *nbatches, uint32_t *ngmask,
                            cudaStream_t stream)
{
    if (nseq == 0 || ngroups == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(nbatches, 0, (size_t)ngroups * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    dim3 grid(((npad >> 5) + 7) / 8, ngroups);
    k_sim_rows<<<grid, 256, 0, stream>>>(codesT, nseq, npad, skipbits, nbatches, ngmask);
    return cudaGetLastError();
}

// Column groups [group_begin, group_end) of 32 columns each.
cudaError_t launch_similarity(const uint8_t *codesT, int nseq, int npad, int ncol,
                              const float *identities, const float *dist, int npos,
                              const uint8_t *col_skip, const uint32_t *skipbits,
                              const uint32_t *ngmask, const unsigned long long *nbatches,
                              int group_begin, int group_end,
                              float *num, float *den, int num_sms, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0 || group_end <= group_begin) return cudaSuccess;
    Sim2Params p{};
    p.codesT = codesT;
    p.identities = identities;
    p.dist = dist;
    p.col_skip = col_skip;
    p.skipbits = skipbits;
    p.ngmask = ngmask;
    p.nbatches = nbatches;
    p.num_out = num;
    p.den_out = den;
    p.nseq = nseq;
    p.npad = npad;
    p.ncol = ncol;
    p.npos = npos;
    p.group_begin = group_begin;
    p.num_sms = num_sms;
    const int ngroups = group_end - group_begin;
    p.slots = SIM2_MAX_SLOTS;  // 52 KB: two CTAs fit an SM when there are more groups than SMs
    const size_t smem = sim2_smem_bytes(p.slots);
    cudaError_t e = cudaFuncSetAttribute(k_similarity2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_similarity2, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    k_similarity2<<<ngroups, SIM2_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tcu
