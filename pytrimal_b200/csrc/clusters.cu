// clusters.cu -- K5..K7: consumers of the device-resident identity matrix
// (SURVEY 8f rank 1).
//
// The reference computes the packed nseq x nseq identity matrix and then walks it
// on the host in three places of vendor/trimal/source/Cleaner.cpp:
//   selectMethod              :46-99     per-row max and sequential fp32 sum over all j != i
//   getCutPointClusters       :1026-1156 per-row max / min / sum over j > i, then repeated
//                                        greedy clusterings while bisecting the threshold
//   calculateRepresentativeSeq:1398-1466 one greedy clustering at a given threshold
// These walks are the only reason 4*P bytes (5 GB at 50 000 sequences) would cross PCIe.
// Here they run where the matrix already is:
//
//   K5 k_identity_bits   one streaming pass (HBM-bound, 4*P bytes read): identity > threshold
//                        as a full symmetric nseq x nseq BIT matrix, one row per sequence; only
//                        for a matrix that is already resident as floats -- tcu_representatives
//                        gets the bits from K1's epilogue + k_bits_rows
//   K6 k_row_stats       per-row statistics; each lane owns one row and replays the
//                        reference's fp32 additions in the reference's order (j ascending)
//   K7 k_greedy_clusters the greedy clustering.  Both reference loops are the same rule:
//                        in the given order, a sequence opens a new cluster iff no EARLIER
//                        cluster representative has identity > threshold with it (the
//                        lexicographically-first maximal independent set of the threshold
//                        graph).  One persistent kernel walks the order in blocks of 1024
//                        sequences: "scanner" CTAs (all SMs but one) test the sequences of a
//                        block against the representatives of earlier blocks (one AND over two
//                        bit rows per sequence) and gather the adjacency inside the block and
//                        to the block before; the "resolver" CTA settles a block in one go
//                        (fixed point of the greedy rule, all 1024 sequences at once) while
//                        the scanners already work on the next one.
//                        Result and order of the cluster list are exactly the reference's.
#include "tcu_internal.cuh"

namespace tcu {

// offset of pair (i, i+1) in the packed upper-triangular array (Cleaner.cpp:1431-1434:
// pos(i, j) = n*i - (i+1)(i+2)/2 + j)
__device__ __forceinline__ long long pair_row_base(long long i, long long n)
{
    return n * i - (i + 1) * (i + 2) / 2;  // + j gives the element
}

// ---------------------------------------------------------------------------
// K5: threshold -> symmetric bit matrix, one row per sequence (tcu_internal.cuh), from a
// resident float matrix.  One warp per 32 x 32 block of the upper triangle: 32 coalesced row
// reads (issued in batches of 16 before any is used, so that a warp keeps 2 KB in
// flight); the ballots are the row words, the per-lane accumulated bits the words of the
// transposed block.  (tcu_representatives does not come here: K1 thresholds in its epilogue.)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_identity_bits(const float *__restrict__ id, int n, int W,
                                                       float thr, uint32_t *__restrict__ bits,
                                                       int rb_begin)
{
    const size_t pitch = brow_pitch_words(n);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rb = rb_begin + blockIdx.y;  // 32-row block
    const int jw = blockIdx.x * 8 + warp;  // 32-column word
    if (jw < rb || jw >= W) return;        // below the diagonal: written as a transpose
    const int i0 = rb * 32, j = jw * 32 + lane;
    uint32_t mine = 0, colword = 0;
    long long base = pair_row_base(i0, n);
    const float never = __int_as_float(0x7fc00000);  // NaN > thr is false
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 16) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int i = i0 + r0 + k;
            v[k] = (i < n && j > i && j < n) ? id[base + j] : never;
            base += n - i - 2;  // pair_row_base(i+1) - pair_row_base(i)
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const bool bit = v[k] > thr;
            const uint32_t w = __ballot_sync(0xffffffffu, bit);
            if (lane == r0 + k) mine = w;
            colword |= (uint32_t)bit << (r0 + k);
        }
    }
    if (jw == rb) {
        // diagonal block: bits j > i come from the row word, bits j < i from the column word
        if (i0 + lane < n) bits[(size_t)(i0 + lane) * pitch + jw] = mine | colword;
    } else {
        if (i0 + lane < n) bits[(size_t)(i0 + lane) * pitch + jw] = mine;
        if (j < n) bits[(size_t)j * pitch + rb] = colword;
    }
}

// Rows [row_begin, row_end) of the matrix (multiples of 32, or n): `id` is the address
// packed offset 0 WOULD have, so a rank that holds only its band passes band - band_offset.
// Every word of `bits` that belongs to these rows' pairs is written (both mirror images).
cudaError_t launch_identity_bits(const float *id, int n, float thr, uint32_t *bits, int row_begin,
                                 int row_end, cudaStream_t stream)
{
    if (n <= 0 || row_end <= row_begin) return cudaSuccess;
    const int W = (n + 31) / 32;
    dim3 grid((W + 7) / 8, (row_end - row_begin + 31) / 32);
    k_identity_bits<<<grid, 256, 0, stream>>>(id, n, W, thr, bits, row_begin / 32);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Mirror + relayout pass.  K1's threshold epilogue leaves, in the slab layout, the "column
// words" of the pairs (row j against earlier sequences i < j; zeros where j <= i).  This
// pass writes the full symmetric matrix with one contiguous row per sequence (what the
// clustering walk reads, in an order that has nothing to do with the sequence index): output
// block (R, C) of 128 x 128 bits is slab C's rows of R as they are when R > C, the transpose
// of slab R's rows of C when R < C, and U | U^T on the diagonal.  One CTA of 128 threads per
// row block and EIGHT column blocks: all 16-byte loads of a thread are issued first (2 KB
// contiguous per warp and block), a transpose is four 32 x 32 butterflies across the warp
// written straight into the shared-memory tile, and the tile leaves as whole 128-byte lines
// of the rows.  HBM-bound: reads and writes n^2 / 8 bytes each.
// ---------------------------------------------------------------------------
constexpr int BR_COLS = 8;               // column blocks per CTA
constexpr int BR_STRIDE = BR_COLS + 1;   // uint4 per tile row (+1: spreads the banks)

__global__ void __launch_bounds__(128) k_bits_rows(const uint32_t *__restrict__ slab_bits, int n,
                                                   uint32_t *__restrict__ rows)
{
    __shared__ uint4 s_tile[128 * BR_STRIDE];
    const int R = blockIdx.y, C0 = blockIdx.x * BR_COLS;
    const int lane = threadIdx.x & 31, u = threadIdx.x >> 5;
    const int nslab = (n + 127) >> 7;
    const uint4 *slabs = reinterpret_cast<const uint4 *>(slab_bits);
    const size_t pitch4 = brow_pitch_words(n) / 4;
    const int r = R * 128 + threadIdx.x;  // this thread's row of R
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    uint32_t *s_words = reinterpret_cast<uint32_t *>(s_tile);

    // one load per column block: C <= R the row's own entry (bits against earlier sequences),
    // C > R the entry of row 128 C + thread against the sequences of R (to be transposed)
    uint4 v[BR_COLS];
#pragma unroll
    for (int c = 0; c < BR_COLS; c++) {
        const int C = C0 + c;
        v[c] = zero;
        if (C < nslab) {
            if (C <= R) {
                if (r < n) v[c] = slabs[(size_t)C * n + r];
            } else {
                const int x = C * 128 + threadIdx.x;
                if (x < n) v[c] = slabs[(size_t)R * n + x];
            }
        }
    }
    // transposes (C >= R) into the tile
#pragma unroll
    for (int c = 0; c < BR_COLS; c++) {
        const int C = C0 + c;
        if (C < R || C >= nslab) continue;  // uniform
        const uint32_t w4[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            // 32 x 32 bit transpose across the warp (lane = row): five butterfly stages, each
            // swapping the off-diagonal j x j blocks with the lane j away; afterwards lane k
            // holds sequence 128 R + 32 q + k against rows 128 C + 32 u ..
            uint32_t xw = w4[q];
            uint32_t low = 0x0000FFFFu;  // columns with (column & j) == 0
#pragma unroll
            for (int j = 16; j >= 1; j >>= 1) {
                const uint32_t y = __shfl_xor_sync(0xffffffffu, xw, j);
                xw = (lane & j) ? ((xw & ~low) | ((y & ~low) >> j)) : ((xw & low) | ((y & low) << j));
                low ^= low << (j >> 1);  // 0x0000FFFF -> 0x00FF00FF -> ... -> 0x55555555
            }
            s_words[((32 * q + lane) * BR_STRIDE + c) * 4 + u] = xw;
        }
    }
    __syncthreads();
    // the rows' own entries (C <= R); on the diagonal they join the transposed half
#pragma unroll
    for (int c = 0; c < BR_COLS; c++) {
        const int C = C0 + c;
        if (C > R && C < nslab) continue;  // uniform
        uint4 o = v[c];
        if (C == R) {
            const uint4 t4 = s_tile[threadIdx.x * BR_STRIDE + c];
            o.x |= t4.x;
            o.y |= t4.y;
            o.z |= t4.z;
            o.w |= t4.w;
        }
        s_tile[threadIdx.x * BR_STRIDE + c] = o;
    }
    __syncthreads();
    // 128 rows x 128 bytes: eight threads per row
    uint4 *out = reinterpret_cast<uint4 *>(rows);
#pragma unroll
    for (int pass = 0; pass < 8; pass++) {
        const int row = pass * 16 + (threadIdx.x >> 3), piece = threadIdx.x & 7;
        const int rr = R * 128 + row;
        if (rr < n) out[(size_t)rr * pitch4 + C0 + piece] = s_tile[row * BR_STRIDE + piece];
    }
}

cudaError_t launch_bits_rows(const uint32_t *slab_bits, int n, uint32_t *rows, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const int nsb = (n + 127) / 128;
    dim3 grid((nsb + BR_COLS - 1) / BR_COLS, nsb);
    k_bits_rows<<<grid, 128, 0, stream>>>(slab_bits, n, rows);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K6: per-row statistics.  Lane r of a warp owns row i0 + r and keeps the running
// fp32 sum in the reference's order: j = 0 .. n-1, j != i (FULL, Cleaner.cpp:68-80)
// or j = i+1 .. n-1 (upper only, Cleaner.cpp:1054-1063).  Elements with j < i0 are
// read down the columns (consecutive lanes = consecutive addresses), 16 loads in
// flight per lane; elements to the right of the diagonal block are read row by row,
// coalesced, one 32 x 32 tile ahead into registers, passed through a padded
// shared-memory tile and consumed transposed.  Only n/32 warps exist (one chain per
// row), so the loads in flight per warp, not occupancy, carry the bandwidth.
// ---------------------------------------------------------------------------
template <bool FULL>
__global__ void __launch_bounds__(64) k_row_stats(const float *__restrict__ id, int n,
                                                  float *__restrict__ row_max,
                                                  float *__restrict__ row_min,
                                                  float *__restrict__ row_sum)
{
    __shared__ float tile[2][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = (blockIdx.x * 2 + warp) * 32;
    if (i0 >= n) return;
    const int i = i0 + lane;
    const bool live = i < n;
    const long long N = n;
    float sum = 0.f, mx = 0.f, mn = 1.f;
    auto take = [&](float v) {
        sum = __fadd_rn(sum, v);
        mx = fmaxf(mx, v);
        mn = fminf(mn, v);
    };

    if (FULL && live) {
        long long pos = pair_row_base(0, N) + i;
        for (int j = 0; j < i0; j += 16) {  // i0 is a multiple of 32
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                v[k] = id[pos];
                pos += N - (j + k) - 2;
            }
#pragma unroll
            for (int k = 0; k < 16; k++) take(v[k]);
        }
    }
    if (live) {
        const long long mybase = pair_row_base(i, N);
        for (int jj = 0; jj < 32; jj++) {
            const int j = i0 + jj;
            if (j >= n) break;
            if (j < i) {
                if (FULL) take(id[pair_row_base(j, N) + i]);
            } else if (j > i) {
                take(id[mybase + j]);
            }
        }
    }
    float(*t)[33] = tile[warp];
    float nxt[32];
    const long long base0 = pair_row_base(i0, N);
    auto fetch = [&](int j0) {
        long long base = base0;
        const bool incol = j0 + lane < n;
#pragma unroll
        for (int r = 0; r < 32; r++) {
            nxt[r] = (incol && i0 + r < n) ? id[base + j0 + lane] : 0.f;
            base += N - (i0 + r) - 2;
        }
    };
    if (i0 + 32 < n) fetch(i0 + 32);
    for (int j0 = i0 + 32; j0 < n; j0 += 32) {
#pragma unroll
        for (int r = 0; r < 32; r++) t[r][lane] = nxt[r];
        __syncwarp();
        if (j0 + 32 < n) fetch(j0 + 32);  // in flight while this tile is consumed
        const int cmax = min(32, n - j0);
        if (live) {
            if (cmax == 32) {
#pragma unroll
                for (int c = 0; c < 32; c++) take(t[lane][c]);
            } else {
                for (int c = 0; c < cmax; c++) take(t[lane][c]);
            }
        }
        __syncwarp();
    }
    if (live) {
        if (row_max) row_max[i] = mx;
        if (row_min) row_min[i] = mn;
        if (row_sum) row_sum[i] = sum;
    }
}

cudaError_t launch_row_stats(const float *id, int n, bool upper_only, float *row_max,
                             float *row_min, float *row_sum, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const int grid = (n + 63) / 64;
    if (upper_only)
        k_row_stats<false><<<grid, 64, 0, stream>>>(id, n, row_max, row_min, row_sum);
    else
        k_row_stats<true><<<grid, 64, 0, stream>>>(id, n, row_max, row_min, row_sum);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K7: the greedy walk as ONE persistent kernel (cooperative launch: every CTA is resident).
//
// The order is cut into blocks of 1024 sequences.  CTA 0 is the resolver, the others scan:
//
//   scanner, block b, sequence t (four warps; its bit row is one contiguous read):
//     alive[t]    = no representative of the blocks <= b-2 is adjacent to it
//                   (`rep` bitset; waits until the resolver has finished block b-2)
//     own[w][t]   = adjacency bits to the block's own sequences 32 w .. that precede t
//     prev[w][t]  = adjacency bits to the sequences of block b-1
//                   (word-major: the resolver reads one word of 32 consecutive sequences
//                   without bank conflicts)
//   resolver, block b (thread t = sequence t, its `own` words in registers):
//     a sequence adjacent to a representative chosen in block b-1 (prev & the previous
//     block's result) is out; then the greedy rule is iterated to its fixed point for the
//     whole block at once:
//       a sequence with an adjacent representative is out;
//       one none of whose earlier neighbours is undecided or a representative becomes one.
//     The earliest undecided sequence always decides, so the loop ends (after as many
//     rounds as the longest chain of dependent decisions: a handful on real alignments)
//     with exactly the sequential answer.  New representatives are appended in visiting
//     order and entered into `rep`.
//
// Because the scan of block b needs nothing from block b-1's resolution, it runs while the
// resolver is busy with b-1: the critical path is the resolver alone (round 2 launched a
// scan and a resolve kernel per block, 2 x 49 dependent launches at 50 000 sequences).
// Hand-over through global counters with release / acquire semantics: scanned[b] counts
// the scanner CTAs done with block b, `resolved` the blocks the resolver has finished; the
// per-block scratch (alive / own / prev) is double-buffered by block parity.  A scanner may
// see `rep` while the resolver adds block b-1's representatives to it: harmless, those are
// genuine earlier representatives and the union with the prev check is the same.
// ---------------------------------------------------------------------------
#ifndef MIS_FORCE_GLOBAL_REP
#define MIS_FORCE_GLOBAL_REP 0
#endif
constexpr int MIS_NB = 1024;
constexpr int MIS_SEQ_PER_CTA = 8;  // 1024 threads, four warps per sequence

struct GreedyParams {
    const uint32_t *rows;  // symmetric bit matrix, one row per sequence
    size_t pitch_w;
    int n;
    const int *order;
    int total;
    uint32_t *rep;  // 4 * ceil(n / 128) words, zero on entry
    int *scanned;   // one counter per block, zero on entry
    int *resolved;  // zero on entry
    uint8_t *alive8;     // [2][MIS_NB]
    uint32_t *adj_own;   // [2][32][MIS_NB]
    uint32_t *adj_prev;  // [2][32][MIS_NB]
    int *clusters;
    int *count;
    int rep_in_smem;  // the resolver keeps `rep` in (dynamic) shared memory and stores whole words
};

__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_release_gpu(int *p, int v)
{
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void group_bar(int id)  // the 128 threads of one sequence
{
    asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}

__device__ void greedy_scanner(const GreedyParams &p, int (*s_ord)[MIS_NB], uint32_t (*s_hit)[4])
{
    const int tid = threadIdx.x, lane = tid & 31, part = (tid >> 5) & 3, sub = tid >> 7;
    const int nscan = gridDim.x - 1;
    const int slot0 = (blockIdx.x - 1) * MIS_SEQ_PER_CTA + sub, stride = nscan * MIS_SEQ_PER_CTA;
    const int nblk = (p.total + MIS_NB - 1) / MIS_NB;
    const int nslab = (p.n + 127) >> 7;
    const uint4 *rep4 = reinterpret_cast<const uint4 *>(p.rep);
    for (int b = 0; b < nblk; b++) {
        const int base = b * MIS_NB, cnt = min(MIS_NB, p.total - base), buf = b & 1;
        for (int k = tid; k < cnt; k += 1024) s_ord[buf][k] = p.order[base + k];
        // `rep` must hold the blocks <= b-2, and the resolver must be done reading this
        // parity's scratch (it did so for block b-2)
        if (tid == 0 && b >= 2)
            while (ld_acquire_gpu(p.resolved) < b - 1) {
            }
        __syncthreads();
        uint8_t *alive8 = p.alive8 + buf * MIS_NB;
        uint32_t *own = p.adj_own + (size_t)buf * 32 * MIS_NB;
        uint32_t *prev = p.adj_prev + (size_t)buf * 32 * MIS_NB;
        for (int t = slot0; t < cnt; t += stride) {
            const int s = s_ord[buf][t];
            const uint32_t *row = p.rows + (size_t)s * p.pitch_w;
            const uint4 *row4 = reinterpret_cast<const uint4 *>(row);
            uint32_t acc = 0;
            for (int S0 = lane + 32 * part; S0 < nslab; S0 += 4 * 128) {
                // eight loads in flight per lane (the row comes from HBM: one latency, not four)
                uint4 a[4], r[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int S = S0 + 128 * k;
                    a[k] = r[k] = make_uint4(0u, 0u, 0u, 0u);
                    if (S < nslab) {
                        a[k] = row4[S];           // immutable: may stay in L1 for the gathers
                        r[k] = __ldcg(rep4 + S);  // written by the resolver: L2
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    acc |= (a[k].x & r[k].x) | (a[k].y & r[k].y) | (a[k].z & r[k].z) | (a[k].w & r[k].w);
            }
            const uint32_t any = __any_sync(0xffffffffu, acc != 0) ? 1u : 0u;
            if (lane == 0) s_hit[sub][part] = any;
            group_bar(1 + sub);
            const bool dead = (s_hit[sub][0] | s_hit[sub][1] | s_hit[sub][2] | s_hit[sub][3]) != 0;
            if (part == 0 && lane == 0) alive8[t] = dead ? 0 : 1;
            if (!dead) {
                // word w = 8 part + lane / 4 covers the block sequences 32 w ..; four lanes
                // gather 8 bits each
                const int w = 8 * part + (lane >> 2);
                const int a0 = w * 32 + (lane & 3) * 8;
                uint32_t word = 0, pword = 0;
                if (a0 < t) {
                    const int amax = min(8, t - a0);
                    uint32_t g[8];
#pragma unroll
                    for (int a = 0; a < 8; a++) {
                        const int v = s_ord[buf][min(a0 + a, cnt - 1)];
                        g[a] = a < amax ? (row[v >> 5] >> (v & 31)) & 1u : 0u;
                    }
#pragma unroll
                    for (int a = 0; a < 8; a++) word |= g[a] << ((lane & 3) * 8 + a);
                }
                if (b > 0) {  // the block before is always full
                    uint32_t g[8];
#pragma unroll
                    for (int a = 0; a < 8; a++) {
                        const int v = s_ord[buf ^ 1][a0 + a];
                        g[a] = (row[v >> 5] >> (v & 31)) & 1u;
                    }
#pragma unroll
                    for (int a = 0; a < 8; a++) pword |= g[a] << ((lane & 3) * 8 + a);
                }
                word |= __shfl_xor_sync(0xffffffffu, word, 1);
                word |= __shfl_xor_sync(0xffffffffu, word, 2);
                pword |= __shfl_xor_sync(0xffffffffu, pword, 1);
                pword |= __shfl_xor_sync(0xffffffffu, pword, 2);
                if ((lane & 3) == 0) {
                    own[w * MIS_NB + t] = word;
                    prev[w * MIS_NB + t] = pword;
                }
            }
            group_bar(1 + sub);  // s_hit is rewritten by the group's next sequence
        }
        __syncthreads();
        // release is cumulative: it publishes what the CTA wrote before the barrier
        if (tid == 0) red_add_release_gpu(p.scanned + b, 1);
    }
}

__device__ void greedy_resolver(const GreedyParams &p, uint32_t *s_in, uint32_t *s_und,
                                uint32_t *s_pre, uint32_t *s_inprev, uint32_t *s_rep)
{
    const int t = threadIdx.x, lane = t & 31, g = t >> 5;
    if (p.rep_in_smem) {
        const int nwords = 4 * ((p.n + 127) >> 7);
        for (int i = t; i < nwords; i += 1024) s_rep[i] = 0;
        __syncthreads();
    }
    const int nscan = gridDim.x - 1;
    const int nblk = (p.total + MIS_NB - 1) / MIS_NB;
    int found = 0;
    for (int b = 0; b < nblk; b++) {
        const int base = b * MIS_NB, cnt = min(MIS_NB, p.total - base), buf = b & 1;
        const int my_seq = t < cnt ? p.order[base + t] : 0;
        if (t == 0)
            while (ld_acquire_gpu(p.scanned + b) < nscan) {
            }
        __syncthreads();
        // the scanners' output: L2 only (this SM's L1 may hold the lines of two blocks ago)
        const uint32_t *own = p.adj_own + (size_t)buf * 32 * MIS_NB;
        const uint32_t *prev = p.adj_prev + (size_t)buf * 32 * MIS_NB;
        // all loads of the block are issued at once (one L2 latency, not one per dependency);
        // the words of a sequence that is not alive were not written and are never used
        const uint8_t alive = t < cnt ? __ldcg(p.alive8 + buf * MIS_NB + t) : (uint8_t)0;
        uint32_t a[32];
#pragma unroll
        for (int w = 0; w < 32; w++) a[w] = (w <= g && t < cnt) ? __ldcg(own + w * MIS_NB + t) : 0u;
        uint32_t hitp = 0;
        if (b > 0 && t < cnt) {
#pragma unroll
            for (int w = 0; w < 32; w++) hitp |= __ldcg(prev + w * MIS_NB + t) & s_inprev[w];
        }
        bool undec = alive != 0 && hitp == 0;
        {
            const uint32_t bu = __ballot_sync(0xffffffffu, undec);
            if (lane == 0) {
                s_und[g] = bu;
                s_in[g] = 0;
            }
        }
        __syncthreads();
        for (;;) {
            uint32_t hit = 0, wait = 0;
            const uint4 *in4 = reinterpret_cast<const uint4 *>(s_in);
            const uint4 *und4 = reinterpret_cast<const uint4 *>(s_und);
#pragma unroll
            for (int w4 = 0; w4 < 8; w4++) {
                if (4 * w4 > g) break;  // a[w] = 0 beyond the warp's own word (warp-uniform)
                const uint4 i4 = in4[w4], u4 = und4[w4];
                hit |= (a[4 * w4] & i4.x) | (a[4 * w4 + 1] & i4.y) | (a[4 * w4 + 2] & i4.z) | (a[4 * w4 + 3] & i4.w);
                wait |= (a[4 * w4] & u4.x) | (a[4 * w4 + 1] & u4.y) | (a[4 * w4 + 2] & u4.z) | (a[4 * w4 + 3] & u4.w);
            }
            const bool out = undec && hit != 0;
            const bool in = undec && hit == 0 && wait == 0;
            const uint32_t bi = __ballot_sync(0xffffffffu, in), bo = __ballot_sync(0xffffffffu, out);
            undec = undec && !in && !out;
            __syncthreads();  // every thread has read the masks of this round
            if (lane == 0) {
                s_in[g] |= bi;
                s_und[g] &= ~(bi | bo);
            }
            if (!__syncthreads_or(undec)) break;
        }
        if (g == 0) {  // exclusive prefix of the representatives per group
            const uint32_t c = __popc(s_in[lane]);
            uint32_t x = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            s_pre[lane] = x - c;
        }
        __syncthreads();
        const uint32_t mine = s_in[g];
        if ((mine >> lane) & 1u) {
            if (p.clusters) p.clusters[found + s_pre[g] + __popc(mine & ((1u << lane) - 1u))] = my_seq;
            if (p.rep_in_smem)
                atomicOr(&s_rep[my_seq >> 5], 1u << (my_seq & 31));
            else
                atomicOr(&p.rep[my_seq >> 5], 1u << (my_seq & 31));
        }
        found += (int)(s_pre[31] + __popc(s_in[31]));
        if (lane == 0) s_inprev[g] = mine;
        __syncthreads();  // also: every thread has read s_in / s_pre before the next block resets them
        if (p.rep_in_smem) {
            // the bitset lives in shared memory and goes out as whole lines: ~50 plain stores
            // for the publishing fence to wait for, instead of ~900 scattered atomics
            const int nwords = 4 * ((p.n + 127) >> 7);
            for (int i = t; i < nwords; i += 1024) p.rep[i] = s_rep[i];
            __syncthreads();
        }
        // release is cumulative: it publishes what the CTA wrote to rep / clusters before the barrier
        if (t == 0) st_release_gpu(p.resolved, b + 1);
    }
    if (t == 0) *p.count = found;
}

__global__ void __launch_bounds__(1024, 1) k_greedy_clusters(const GreedyParams p)
{
    __shared__ int s_ord[2][MIS_NB];
    __shared__ uint32_t s_hit[MIS_SEQ_PER_CTA][4];
    __shared__ __align__(16) uint32_t s_masks[4][32];
    extern __shared__ __align__(16) uint32_t s_rep[];  // the resolver's copy of `rep` (rep_in_smem)
    if (blockIdx.x == 0)
        greedy_resolver(p, s_masks[0], s_masks[1], s_masks[2], s_masks[3], s_rep);
    else
        greedy_scanner(p, s_ord, s_hit);
}

// bytes of scratch launch_greedy_clusters needs for n sequences and an order of `total`
size_t greedy_scratch_bytes(int n, int total)
{
    const size_t nslab = ((size_t)n + 127) / 128, nblk = ((size_t)total + MIS_NB - 1) / MIS_NB;
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    return up(nslab * 16) + up((nblk + 1) * sizeof(int)) + up(2 * MIS_NB) +
           2 * up((size_t)2 * 32 * MIS_NB * sizeof(uint32_t));
}

// Greedy clustering over the row-per-sequence bit matrix in the given order (device array of
// `total` indices); the cluster list goes to `clusters` (may be NULL), its length to *count.
cudaError_t launch_greedy_clusters(const uint32_t *rows, int n, const int *order, int total,
                                   void *scratch, int *clusters, int *count, int num_sms,
                                   cudaStream_t stream)
{
    if (total <= 0) return cudaMemsetAsync(count, 0, sizeof(int), stream);
    const size_t nslab = ((size_t)n + 127) / 128, nblk = ((size_t)total + MIS_NB - 1) / MIS_NB;
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    uint8_t *q = (uint8_t *)scratch;
    GreedyParams p{};
    p.rows = rows;
    p.pitch_w = brow_pitch_words(n);
    p.n = n;
    p.order = order;
    p.total = total;
    p.rep = (uint32_t *)q;
    q += up(nslab * 16);
    p.scanned = (int *)q;
    p.resolved = p.scanned + nblk;
    q += up((nblk + 1) * sizeof(int));
    const size_t zeroed = (size_t)(q - (uint8_t *)scratch);
    p.alive8 = q;
    q += up(2 * MIS_NB);
    p.adj_own = (uint32_t *)q;
    q += up((size_t)2 * 32 * MIS_NB * sizeof(uint32_t));
    p.adj_prev = (uint32_t *)q;
    p.clusters = clusters;
    p.count = count;
    cudaError_t e = cudaMemsetAsync(scratch, 0, zeroed, stream);
    if (e != cudaSuccess) return e;
    // every CTA must be resident (the CTAs wait for each other): cooperative launch, at most
    // one CTA of 1024 threads per SM
    const size_t rep_smem = nslab * 16;
    // up to 524 288 sequences; beyond, the resolver updates `rep` in global memory (a build with
    // -DMIS_FORCE_GLOBAL_REP=1 takes that path always: how the test suite was run over it once)
    p.rep_in_smem = !MIS_FORCE_GLOBAL_REP && rep_smem <= 64 * 1024;
    const size_t dyn = p.rep_in_smem ? rep_smem : 0;
    e = cudaFuncSetAttribute(k_greedy_clusters, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_greedy_clusters, 1024, dyn);
    if (e != cudaSuccess) return e;
    if (per_sm < 1 || num_sms < 2) return cudaErrorLaunchOutOfResources;
    const int want = (min(total, MIS_NB) + MIS_SEQ_PER_CTA - 1) / MIS_SEQ_PER_CTA;
    const int nscan = max(1, min(want, num_sms - 1));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(1 + nscan);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = dyn;
    return cudaLaunchKernelEx(&cfg, k_greedy_clusters, p);
}

}  // namespace tcu
