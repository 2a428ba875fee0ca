// clusters.cu -- K5..K8: consumers of the device-resident identity matrix
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
//                        as a full symmetric nseq x nseq BIT matrix (nseq^2/8 bytes); only for
//                        a matrix that is already resident as floats -- tcu_representatives gets
//                        the bits from K1's epilogue + k_bits_symmetrize
//   K6 k_row_stats       per-row statistics; each lane owns one row and replays the
//                        reference's fp32 additions in the reference's order (j ascending)
//   K7 k_mis_scan  +  K8 k_mis_resolve
//                        the greedy clustering.  Both reference loops are the same rule:
//                        in the given order, a sequence opens a new cluster iff no EARLIER
//                        cluster representative has identity > threshold with it (the
//                        lexicographically-first maximal independent set of the threshold
//                        graph).  Processed in blocks of 1024 sequences: K7 tests each
//                        sequence of the block against all representatives of earlier
//                        blocks (one AND over two bit rows per sequence, whole GPU) and
//                        gathers the 1024 x 1024 adjacency inside the block; K8 resolves
//                        the block in one CTA (fixed point of the greedy rule, all 1024
//                        sequences at once).
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
// K5: threshold -> symmetric bit matrix in the slab layout (tcu_internal.cuh), from a
// resident float matrix.  One warp per 32 x 32 block of the upper triangle: 32 coalesced row
// reads (issued in batches of 16 before any is used, so that a warp keeps 2 KB in
// flight); the ballots are the row words, the per-lane accumulated bits the words of the
// transposed block.  (tcu_representatives does not come here: K1 thresholds in its epilogue.)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_identity_bits(const float *__restrict__ id, int n, int W,
                                                       float thr, uint32_t *__restrict__ bits,
                                                       int rb_begin)
{
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
        if (i0 + lane < n) bits[bits_word_index(n, i0 + lane, jw)] = mine | colword;
    } else {
        if (i0 + lane < n) bits[bits_word_index(n, i0 + lane, jw)] = mine;
        if (j < n) bits[bits_word_index(n, j, rb)] = colword;
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
// Mirror pass over a bit matrix of which K1's threshold epilogue wrote the "column words"
// (row j against earlier sequences i < j; zeros where j <= i): every 128 x 128 block (BI, BJ),
// BI < BJ, is transposed into its mirror image and a diagonal block becomes U | U^T.  One CTA
// of 128 threads per block: thread = row, one 16-byte load, four 32 x 32 butterfly transposes,
// the mirrored rows leave as 16-byte stores through shared memory.  HBM-bound: reads and
// writes n^2 / 16 bytes each.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bits_symmetrize(uint32_t *__restrict__ bits, int n)
{
    // block (BI, BJ), BJ >= BI: rows of super-block BJ against the sequences of super-block BI
    const int BI = blockIdx.x, BJ = blockIdx.y;
    if (BJ < BI) return;
    __shared__ uint4 s_out[128];  // [sequence of BI] = its words against the rows of BJ
    const int lane = threadIdx.x & 31, u = threadIdx.x >> 5;
    const int r = BJ * 128 + threadIdx.x;  // this thread's row of BJ (warp u = its 32-row group)
    uint4 *slabs = reinterpret_cast<uint4 *>(bits);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const uint4 own = r < n ? slabs[(size_t)BI * n + r] : zero;  // bits (r, i) for i < r
    const uint32_t w4[4] = {own.x, own.y, own.z, own.w};
    uint32_t *s_words = reinterpret_cast<uint32_t *>(s_out);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        // 32 x 32 bit transpose across the warp (lane = row): five butterfly stages, each
        // swapping the off-diagonal j x j blocks with the lane j away; afterwards lane k
        // holds sequence 128 BI + 32 q + k against rows 128 BJ + 32 u ..
        uint32_t x = w4[q];
        uint32_t low = 0x0000FFFFu;  // columns c with (c & j) == 0
#pragma unroll
        for (int j = 16; j >= 1; j >>= 1) {
            const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
            x = (lane & j) ? ((x & ~low) | ((y & ~low) >> j)) : ((x & low) | ((y & low) << j));
            low ^= low << (j >> 1);  // 0x0000FFFF -> 0x00FF00FF -> 0x0F0F0F0F -> 0x33333333 -> 0x55555555
        }
        s_words[(32 * q + lane) * 4 + u] = x;
    }
    __syncthreads();
    const int c = BI * 128 + threadIdx.x;  // sequence of BI whose mirrored words this thread stores
    if (c >= n) return;
    uint4 t4 = s_out[threadIdx.x];
    if (BI == BJ) {
        t4.x |= own.x;
        t4.y |= own.y;
        t4.z |= own.z;
        t4.w |= own.w;
    }
    slabs[(size_t)BJ * n + c] = t4;
}

cudaError_t launch_bits_symmetrize(uint32_t *bits, int n, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const int nsb = (n + 127) / 128;
    dim3 grid(nsb, nsb);
    k_bits_symmetrize<<<grid, 128, 0, stream>>>(bits, n);
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
// K7: four warps per sequence t of the current block (order[base .. base+cnt)).
//   alive8[t]  = no representative found so far (bitset `rep`) is adjacent to it
//   adj[l][t]  = adjacency bits to the block's sequences 32*l .. 32*l+31 that precede t
//                (word-major, so that K8 reads one word of 32 consecutive sequences
//                without bank conflicts); zero for a sequence that is not alive
// A sequence's bits are one 16-byte entry per slab (tcu_internal.cuh); `rep` is padded to
// whole slabs (4 * nslab words, zero beyond n).
// ---------------------------------------------------------------------------
constexpr int MIS_NB = 1024;

__global__ void __launch_bounds__(256) k_mis_scan(const uint32_t *__restrict__ bits, int n,
                                                  const int *__restrict__ order, int base, int cnt,
                                                  const uint32_t *__restrict__ rep,
                                                  uint8_t *__restrict__ alive8,
                                                  uint32_t *__restrict__ adj)
{
    // four warps per sequence, two sequences per CTA: the scan of a block is a chain of
    // dependent memory latencies, so it is spread over many warps with few loads each
    // (2 048 warps per block of 1 024 sequences)
    __shared__ int s_ord[MIS_NB];
    __shared__ uint32_t s_hit[2][4];
    asm volatile("griddepcontrol.launch_dependents;");  // the resolve kernel may be set up now
    for (int k = threadIdx.x; k < cnt; k += 256) s_ord[k] = order[base + k];
    __syncthreads();
    // programmatic dependent launch: everything above ran under the tail of the previous
    // block's resolve kernel; `rep` is its output
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x & 31, part = (threadIdx.x >> 5) & 3, sub = threadIdx.x >> 7;
    const int t = blockIdx.x * 2 + sub;
    const bool live = t < cnt;
    const int s = live ? s_ord[t] : 0;
    const int nslab = (n + 127) >> 7;
    const uint4 *slabs = reinterpret_cast<const uint4 *>(bits) + s;
    const uint4 *rep4 = reinterpret_cast<const uint4 *>(rep);
    uint32_t acc = 0;
    if (live) {
        for (int S = lane + 32 * part; S < nslab; S += 128) {
            const uint4 a = slabs[(size_t)S * n];
            // L2 only: under programmatic dependent launch this grid was already resident
            // while the previous resolve kernel wrote `rep`; an L1 line left by an earlier
            // block's scan on this SM would be stale
            const uint4 r = __ldcg(rep4 + S);
            acc |= (a.x & r.x) | (a.y & r.y) | (a.z & r.z) | (a.w & r.w);
        }
    }
    const uint32_t any = __any_sync(0xffffffffu, acc != 0) ? 1u : 0u;
    if (lane == 0) s_hit[sub][part] = any;
    __syncthreads();
    if (!live) return;
    const bool dead = (s_hit[sub][0] | s_hit[sub][1] | s_hit[sub][2] | s_hit[sub][3]) != 0;
    if (part == 0 && lane == 0) alive8[t] = dead ? 0 : 1;
    // adjacency to the earlier sequences of the block: word w = 8 part + lane / 4 holds the
    // block's sequences 32 w ..; four lanes gather 8 bits each
    const int w = 8 * part + (lane >> 2);
    const int a0 = w * 32 + (lane & 3) * 8;
    uint32_t word = 0;
    if (!dead && a0 < t) {
        const int amax = min(8, t - a0);
        uint32_t g[8];
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const int u = s_ord[min(a0 + a, cnt - 1)];
            g[a] = a < amax ? (bits[bits_word_index(n, s, u >> 5)] >> (u & 31)) & 1u : 0u;
        }
#pragma unroll
        for (int a = 0; a < 8; a++) word |= g[a] << ((lane & 3) * 8 + a);
    }
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    if ((lane & 3) == 0) adj[w * MIS_NB + t] = word;
}

// K8: one CTA of 1024 threads resolves the block; thread t owns sequence t of the block and
// keeps its adjacency words to the earlier sequences of the block in registers.  The greedy
// rule is iterated to its fixed point for the whole block at once:
//   a sequence with an adjacent representative is out;
//   one none of whose earlier neighbours is undecided or a representative becomes one.
// The earliest undecided sequence always decides, so the loop ends (after as many rounds as
// the longest chain of dependent decisions: a handful on real alignments) with exactly the
// sequential answer.  The new representatives are appended in visiting order.
__global__ void __launch_bounds__(1024) k_mis_resolve(const uint32_t *__restrict__ adj,
                                                      const uint8_t *__restrict__ alive8,
                                                      const int *__restrict__ order, int base,
                                                      int cnt, uint32_t *__restrict__ rep,
                                                      int *__restrict__ clusters,
                                                      int *__restrict__ count)
{
    __shared__ uint32_t s_in[32], s_und[32], s_pre[32];
    const int t = threadIdx.x, lane = t & 31, g = t >> 5;
    asm volatile("griddepcontrol.launch_dependents;");  // the next block's scan may be set up now
    const int my_seq = t < cnt ? order[base + t] : 0;  // does not depend on the scan kernel
    asm volatile("griddepcontrol.wait;" ::: "memory");  // adj / alive8 are its output
    uint32_t a[32];
#pragma unroll
    for (int w = 0; w < 32; w++) a[w] = (w <= g && t < cnt) ? __ldcg(adj + w * MIS_NB + t) : 0u;  // L2 only, as above
    bool undec = t < cnt && __ldcg(alive8 + t) != 0;
    {
        const uint32_t b = __ballot_sync(0xffffffffu, undec);
        if (lane == 0) {
            s_und[g] = b;
            s_in[g] = 0;
        }
    }
    __syncthreads();
    for (;;) {
        uint32_t hit = 0, wait = 0;
#pragma unroll
        for (int w = 0; w < 32; w++) {
            hit |= a[w] & s_in[w];
            wait |= a[w] & s_und[w];
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
    const int c0 = __ldcg(count);
    const uint32_t mine = s_in[g];
    if ((mine >> lane) & 1u) {
        const int v = my_seq;
        if (clusters) clusters[c0 + s_pre[g] + __popc(mine & ((1u << lane) - 1u))] = v;
        atomicOr(&rep[v >> 5], 1u << (v & 31));
    }
    __syncthreads();  // every thread has read *count
    if (t == 0) *count = c0 + (int)(s_pre[31] + __popc(s_in[31]));
}

int mis_block() { return MIS_NB; }

// Greedy clustering over the bit matrix (slab layout, n sequences) in the given order; rep
// (4 * ceil(n / 128) words), alive8 (MIS_NB bytes), adj (MIS_NB * 32 words) and count (1 int)
// are scratch; rep and count must be zero on entry.
cudaError_t launch_greedy_clusters(const uint32_t *bits, int n, const int *order, int total,
                                   uint32_t *rep, uint8_t *alive8, uint32_t *adj, int *clusters,
                                   int *count, cudaStream_t stream)
{
    // the 2 x ceil(total / 1024) launches depend on each other one after the other: each is
    // launched with programmatic stream serialization, so that its launch latency and its
    // prologue overlap the tail of its predecessor (griddepcontrol.wait inside the kernels)
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    for (int base = 0; base < total; base += MIS_NB) {
        const int cnt = min(MIS_NB, total - base);
        cudaLaunchConfig_t cfg = {};
        cfg.stream = stream;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        // the very first launch follows copies / memsets / other kernels of the caller: it keeps
        // the ordinary stream dependency
        cfg.numAttrs = base == 0 ? 0 : 1;
        cfg.gridDim = dim3((cnt + 1) / 2);
        cfg.blockDim = dim3(256);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_mis_scan, bits, n, order, base, cnt,
                                           (const uint32_t *)rep, alive8, adj);
        if (e != cudaSuccess) return e;
        cfg.numAttrs = 1;
        cfg.gridDim = dim3(1);
        cfg.blockDim = dim3(1024);
        e = cudaLaunchKernelEx(&cfg, k_mis_resolve, (const uint32_t *)adj, (const uint8_t *)alive8,
                               order, base, cnt, rep, clusters, count);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

}  // namespace tcu
