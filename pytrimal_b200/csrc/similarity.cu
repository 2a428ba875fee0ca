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
// are contiguous per row, rows padded to a multiple of 32 -- holding 8 * code
// (the byte offset of the code's entry in a table row); the gap class, skipped
// columns and all padding hold SIM2_GAP8 = 8 * 31.
// ---------------------------------------------------------------------------
constexpr uint32_t SIM2_GAPIDX = 31;
constexpr uint32_t SIM2_GAP8 = 8 * SIM2_GAPIDX;

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
                uint32_t out = SIM2_GAP8;
                if (col < ncol && !col_skip[col]) {
                    const uint32_t code = lut[byte];
                    if (code == SIM_INCORRECT || code == SIM_UNDEFINED) {
                        uint32_t up = (byte >= 'a' && byte <= 'z') ? byte - 32 : byte;
                        atomicMin(first_error, (((unsigned long long)col * nseq + r) << 8) | up);
                    } else if (code != SIM_GAP) {
                        out = code * 8;
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
                             unsigned long long *first_error, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0) return cudaSuccess;
    // padding rows (and nothing else survives the kernel below) are gaps
    cudaError_t e = cudaMemsetAsync(codesT, (int)SIM2_GAP8, (size_t)(pitch >> 5) * npad * 32, stream);
    if (e != cudaSuccess) return e;
    const long long total = (long long)nseq * (long long)(pitch >> 4);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    k_sim_codes<<<blocks, 256, 0, stream>>>(raw, nseq, ncol, pitch, npad, lut256, col_skip, codesT,
                                            first_error);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Row prepass.  An outer row j whose 32 columns of the group are all gaps adds
// nothing in any lane (template.h:157-160) and is skipped by every warp of the
// main kernel: skipbits[group][j / 32] bit j % 32.  Row nseq-1 (never an outer
// row) and padding are marked too.  nbatches[group] = number of 32-k batches
// the main kernel exchanges for the group.
// ---------------------------------------------------------------------------
constexpr int SIM2_KB = 32;  // inner rows (k) per batch

__device__ __forceinline__ int sim2_row_batches(int j, int nseq)
{
    return ((nseq - 1) >> 5) - ((j + 1) >> 5) + 1;  // aligned batches covering k = j+1 .. nseq-1
}

__global__ void __launch_bounds__(256) k_sim_rows(const uint8_t *__restrict__ codesT, int nseq,
                                                  int npad, uint32_t *__restrict__ skipbits,
                                                  unsigned long long *__restrict__ nbatches)
{
    const int group = blockIdx.y;
    const int nwords = npad >> 5;
    const int word = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (word >= nwords) return;
    const int lane = threadIdx.x & 31;
    const int j = word * 32 + lane;
    const uint4 *cp = reinterpret_cast<const uint4 *>(codesT + ((size_t)group * npad + j) * 32);
    const uint4 a = __ldg(cp), b = __ldg(cp + 1);
    const uint32_t G = SIM2_GAP8 * 0x01010101u;
    const bool allgap = (a.x & a.y & a.z & a.w & b.x & b.y & b.z & b.w) == G &&
                        (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) == G;
    const bool skip = allgap || j >= nseq - 1;
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
// The chain  num += w * d;  den += w  of a column must run in the reference's
// order, one rounded fp32 add after the other, but its TERMS are independent.
// So the CTA splits the work:
//   6 producer warps   lane = inner row k of a 32-k batch.  Per batch a lane
//                      loads its row's 32 codes (32 contiguous bytes) and
//                      id[j,k] (coalesced), forms w = 1 - id once, and for each
//                      of the 32 columns looks up {D[a_j][a_k], pair counted ?
//                      1 : 0} (one 8-byte shared-memory load; the table has an
//                      all-zero row/column for gaps), multiplies by w and
//                      stores {w*d, w*e} to a ring slot in shared memory.
//                      Terms of pairs the reference skips are exact +0.
//   1 consumer warp    lane = column.  Waits for a slot, then does nothing but
//                      LDS.64 + FADD + FADD per inner row: the two dependent
//                      add chains (4-cycle FADD latency each) are the critical
//                      path of the whole kernel, everything else runs beside
//                      them on the other three SM sub-partitions.
// Slots are handed over with mbarriers (full/empty, one arrival each).
// ---------------------------------------------------------------------------
constexpr int SIM2_NPROD = 6;
constexpr int SIM2_MAX_SLOTS = 12;
constexpr int SIM2_RS = 33;                         // float2 per k row of a slot (odd: no bank conflicts)
constexpr int SIM2_SLOT_F2 = SIM2_KB * SIM2_RS;     // float2 per slot
constexpr int SIM2_THREADS = 256;
constexpr int SIM2_TABLE_F2 = 32 * 32;
constexpr int SIM2_PREFETCH_BATCHES = 4;            // own batches ahead (x SIM2_NPROD in array order)

__host__ __device__ constexpr size_t sim2_smem_bytes(int slots)
{
    return (size_t)slots * SIM2_SLOT_F2 * 8 + SIM2_TABLE_F2 * 8 + 2 * SIM2_MAX_SLOTS * 8;
}

struct Sim2Params {
    const uint8_t *codesT;
    const float *identities;
    const float *dist;
    const uint8_t *col_skip;
    const uint32_t *skipbits;
    const unsigned long long *nbatches;
    float *num_out, *den_out;
    int nseq, npad, ncol, npos;
    int group_begin;
    int slots;      // SIM2_NPROD or 2 * SIM2_NPROD
    int num_sms;
};

__global__ void __launch_bounds__(SIM2_THREADS, 2) k_similarity2(const Sim2Params p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    float2 *ring = reinterpret_cast<float2 *>(smem);
    float2 *T = ring + (size_t)p.slots * SIM2_SLOT_F2;
    uint64_t *full = reinterpret_cast<uint64_t *>(T + SIM2_TABLE_F2);
    uint64_t *empty = full + SIM2_MAX_SLOTS;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = p.group_begin + blockIdx.x;
    const int n = p.nseq;

    // table: T[a * 32 + b] = {D[a][b], 1}; row / column 31 (gap) and unused codes are zero
    for (int i = threadIdx.x; i < SIM2_TABLE_F2; i += SIM2_THREADS) {
        const int a = i >> 5, b = i & 31;
        T[i] = (a < p.npos && b < p.npos) ? make_float2(p.dist[a * p.npos + b], 1.0f)
                                          : make_float2(0.0f, 0.0f);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.slots; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    // When more groups than SMs are launched two CTAs share an SM: put their consumer
    // warps on different sub-partitions (warp id % 4).
    const int cw = (blockIdx.x / max(p.num_sms, 1)) & 3;
    const uint8_t *gc = p.codesT + (size_t)group * p.npad * 32;

    if (warp == cw) {
        // ------------------------------ consumer ------------------------------
        const unsigned long long btot = p.nbatches[group];
        float num = 0.0f, den = 0.0f;
        float2 va[16], vb[16];
        auto loadh = [&](float2(&v)[16], int slot, int h) {
            const float2 *src = ring + (size_t)slot * SIM2_SLOT_F2 + h * 16 * SIM2_RS + lane;
#pragma unroll
            for (int u = 0; u < 16; u++) v[u] = src[u * SIM2_RS];
        };
        auto addh = [&](const float2(&v)[16]) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                num = __fadd_rn(num, v[u].x);
                den = __fadd_rn(den, v[u].y);
            }
        };
        int slot = 0;
        uint32_t par = 0;
        if (btot) {
            mbar_wait(&full[0], 0);
            loadh(va, 0, 0);
        }
        for (unsigned long long b = 0; b < btot; b++) {
            int nslot = slot + 1;
            uint32_t npar = par;
            if (nslot == p.slots) {
                nslot = 0;
                npar ^= 1u;
            }
            // probe the next slot now (the probe takes ~90 cycles), look at the answer
            // after the first half of this batch has been added
            const bool more = b + 1 < btot;
            const uint32_t ready = more ? mbar_try_wait(&full[nslot], npar) : 1u;
            loadh(vb, slot, 1);
            addh(va);
            if (more) {
                if (!ready) mbar_wait(&full[nslot], npar);
                loadh(va, nslot, 0);
            }
            addh(vb);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            slot = nslot;
            par = npar;
        }
        const int col = group * 32 + lane;
        if (col < p.ncol && !p.col_skip[col]) {
            p.num_out[col] = num;
            p.den_out[col] = den;
        }
    } else if (warp != cw + 4) {
        // ------------------------------ producers -----------------------------
        // producer index 0..5: the six warps that are neither the consumer nor the
        // (idle) warp sharing its sub-partition
        int pi = 0;
        for (int w = 0; w < warp; w++) pi += (w != cw && w != cw + 4);
        const int spp = p.slots / SIM2_NPROD;  // slots per producer
        const int nwords = p.npad >> 5;
        const uint32_t *skipw = p.skipbits + (size_t)group * nwords;
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
        };
        auto fetch = [&](int fj, int fkb, Loaded &L) {
            const int k = fkb + lane;
            const uint4 *cp = reinterpret_cast<const uint4 *>(gc + (size_t)k * 32);
            const uint4 *jp = reinterpret_cast<const uint4 *>(gc + (size_t)fj * 32);
            L.c0 = __ldg(cp);
            L.c1 = __ldg(cp + 1);
            L.j0 = __ldg(jp);
            L.j1 = __ldg(jp + 1);
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

            const float w = __fsub_rn(1.0f, cur.id);  // 0 for k <= j and padding
            const uint32_t cw8[8] = {cur.c0.x, cur.c0.y, cur.c0.z, cur.c0.w,
                                     cur.c1.x, cur.c1.y, cur.c1.z, cur.c1.w};
            const uint32_t jw8[8] = {cur.j0.x, cur.j0.y, cur.j0.z, cur.j0.w,
                                     cur.j1.x, cur.j1.y, cur.j1.z, cur.j1.w};
            float2 *dst = ring + (size_t)slot * SIM2_SLOT_F2 + lane * SIM2_RS;
            const char *Tb = reinterpret_cast<const char *>(T);
            // 16 table loads in flight, then 16 stores (the compiler must assume that the
            // ring stores alias the table and would otherwise serialise load -> store)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float2 t[16];
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const int wd = h * 4 + (u >> 2), b = u & 3;
                    // row-j codes (uniform): bytes [c0, 0, c2, 0] and [c1, 0, c3, 0] of code index
                    const uint32_t cjs = (jw8[wd] >> 3) & 0x1F1F1F1Fu;
                    const uint32_t y = (b & 1) ? ((cjs >> 8) & 0x00FF00FFu) : (cjs & 0x00FF00FFu);
                    // table byte offset = a_j * 256 + 8 * a_k: one PRMT
                    const uint32_t sel = 0x5500u | ((b & 2) ? 0x60u : 0x40u) | (uint32_t)b;
                    const uint32_t off = __byte_perm(cw8[wd], y, sel);
                    t[u] = *reinterpret_cast<const float2 *>(Tb + off);
                }
#pragma unroll
                for (int u = 0; u < 16; u++)
                    dst[h * 16 + u] = make_float2(__fmul_rn(w, t[u].x), __fmul_rn(w, t[u].y));
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
                            uint32_t *skipbits, unsigned long long *nbatches, cudaStream_t stream)
{
    if (nseq == 0 || ngroups == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(nbatches, 0, (size_t)ngroups * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    dim3 grid(((npad >> 5) + 7) / 8, ngroups);
    k_sim_rows<<<grid, 256, 0, stream>>>(codesT, nseq, npad, skipbits, nbatches);
    return cudaGetLastError();
}

// Column groups [group_begin, group_end) of 32 columns each.
cudaError_t launch_similarity(const uint8_t *codesT, int nseq, int npad, int ncol,
                              const float *identities, const float *dist, int npos,
                              const uint8_t *col_skip, const uint32_t *skipbits,
                              const unsigned long long *nbatches, int group_begin, int group_end,
                              float *num, float *den, int num_sms, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0 || group_end <= group_begin) return cudaSuccess;
    Sim2Params p{};
    p.codesT = codesT;
    p.identities = identities;
    p.dist = dist;
    p.col_skip = col_skip;
    p.skipbits = skipbits;
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
    // two CTAs per SM only when there are more groups than SMs; a deep ring otherwise
    p.slots = SIM2_MAX_SLOTS;
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
