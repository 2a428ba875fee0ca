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
//                        as a full symmetric nseq x nseq BIT matrix (nseq^2/8 bytes)
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
//                        the block sequentially in one warp (one vote per live sequence).
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
// K5: threshold -> symmetric bit matrix.  One warp per 32 x 32 block of the upper
// triangle: 32 coalesced row reads (issued in batches of 16 before any is used, so
// that a warp keeps 2 KB in flight); the ballots are the row words, the per-lane
// accumulated bits the words of the transposed block.
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
        if (i0 + lane < n) bits[(size_t)(i0 + lane) * W + jw] = mine | colword;
    } else {
        if (i0 + lane < n) bits[(size_t)(i0 + lane) * W + jw] = mine;
        if (j < n) bits[(size_t)j * W + rb] = colword;
    }
}

// Rows [row_begin, row_end) of the matrix (multiples of 32, or n): `id` is the address
// packed offset 0 WOULD have, so a rank that holds only its band passes band - band_offset.
// Every word of `bits` that belongs to these rows' pairs is written; a caller that covers
// only part of the rows must zero `bits` first.
cudaError_t launch_identity_bits(const float *id, int n, int W, float thr, uint32_t *bits,
                                 int row_begin, int row_end, cudaStream_t stream)
{
    if (n <= 0 || row_end <= row_begin) return cudaSuccess;
    dim3 grid((W + 7) / 8, (row_end - row_begin + 31) / 32);
    k_identity_bits<<<grid, 256, 0, stream>>>(id, n, W, thr, bits, row_begin / 32);
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
// K7: one warp per sequence t of the current block (order[base .. base+cnt)).
//   alive8[t]  = no representative found so far (bitset `rep`) is adjacent to it
//   adj[l][t]  = adjacency bits to the block's sequences 32*l .. 32*l+31 that precede t
//                (word-major, so that K8 reads one word of 32 consecutive sequences
//                without bank conflicts); zero for a sequence that is not alive
// ---------------------------------------------------------------------------
constexpr int MIS_NB = 1024;

__global__ void __launch_bounds__(256) k_mis_scan(const uint32_t *__restrict__ bits, int W,
                                                  const int *__restrict__ order, int base, int cnt,
                                                  const uint32_t *__restrict__ rep,
                                                  uint8_t *__restrict__ alive8,
                                                  uint32_t *__restrict__ adj)
{
    __shared__ int s_ord[MIS_NB];
    for (int k = threadIdx.x; k < cnt; k += 256) s_ord[k] = order[base + k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= cnt) return;
    const uint32_t *row = bits + (size_t)s_ord[t] * W;
    uint32_t acc = 0;
    int w = lane;
    for (; w + 96 < W; w += 128) {  // four independent load pairs per step
        const uint32_t a0 = row[w], a1 = row[w + 32], a2 = row[w + 64], a3 = row[w + 96];
        acc |= (a0 & rep[w]) | (a1 & rep[w + 32]) | (a2 & rep[w + 64]) | (a3 & rep[w + 96]);
    }
    for (; w < W; w += 32) acc |= row[w] & rep[w];
    const bool dead = __any_sync(0xffffffffu, acc != 0);
    if (lane == 0) alive8[t] = dead ? 0 : 1;
    uint32_t word = 0;
    const int a0 = lane * 32;
    if (!dead && a0 < t) {
        const int amax = min(32, t - a0);
        if (amax == 32) {
            uint32_t g[32];
#pragma unroll
            for (int a = 0; a < 32; a++) g[a] = row[s_ord[a0 + a] >> 5];
#pragma unroll
            for (int a = 0; a < 32; a++) word |= ((g[a] >> (s_ord[a0 + a] & 31)) & 1u) << a;
        } else {
            for (int a = 0; a < amax; a++) {
                const int u = s_ord[a0 + a];
                word |= ((row[u >> 5] >> (u & 31)) & 1u) << a;
            }
        }
    }
    adj[lane * MIS_NB + t] = word;
}

// K8: one CTA copies the block's adjacency into shared memory; warp 0 then resolves the
// block 32 sequences at a time, lane l owning sequence 32*g + l of group g:
//   1. killed by a representative of an earlier group of this block?  One AND per
//      earlier group, lanes in parallel (the representatives of group w live in lane w).
//   2. inside the group the greedy rule is iterated to its fixed point with ballots: a
//      sequence with an adjacent representative is out; one whose earlier neighbours
//      are all decided and none of them a representative becomes one.  The lowest
//      undecided lane always decides, so this ends after at most 32 rounds (2-3 in practice)
//      and gives exactly the sequential answer.
//   3. the new representatives are appended in visiting order (prefix popcount).
constexpr int MIS_SMEM = MIS_NB * 32 * 4 + MIS_NB * 4 + MIS_NB;

__global__ void __launch_bounds__(1024) k_mis_resolve(const uint32_t *__restrict__ adj,
                                                      const uint8_t *__restrict__ alive8,
                                                      const int *__restrict__ order, int base,
                                                      int cnt, uint32_t *__restrict__ rep,
                                                      int *__restrict__ clusters,
                                                      int *__restrict__ count)
{
    extern __shared__ uint32_t sm[];
    uint32_t *s_adj = sm;                    // [32][MIS_NB]
    int *s_ord = (int *)(sm + MIS_NB * 32);  // [MIS_NB]
    uint8_t *s_alive = (uint8_t *)(s_ord + MIS_NB);  // [MIS_NB]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ngroups = (cnt + 31) >> 5;
    // word w of sequence t only matters for w <= t/32: copy the lower triangle of groups
    for (int w = warp; w < ngroups; w += 32)
        for (int t = w * 32 + lane; t < cnt; t += 32) s_adj[w * MIS_NB + t] = adj[w * MIS_NB + t];
    if (tid < cnt) {
        s_ord[tid] = order[base + tid];
        s_alive[tid] = alive8[tid];
    }
    __syncthreads();
    if (warp != 0) return;
    int c = *count;
    uint32_t repw = 0;  // lane w: representatives of group w
    for (int g = 0; g < ngroups; g++) {
        const int t = g * 32 + lane;
        bool undec = t < cnt && s_alive[t] != 0;
        uint32_t killed = 0;
        for (int w = 0; w < g; w++)
            killed |= s_adj[w * MIS_NB + min(t, cnt - 1)] & __shfl_sync(0xffffffffu, repw, w);
        undec = undec && killed == 0;
        const uint32_t a = t < cnt ? s_adj[g * MIS_NB + t] : 0;  // bits of earlier lanes only
        uint32_t reps = 0, und = __ballot_sync(0xffffffffu, undec);
        while (und) {
            const bool out = undec && (a & reps) != 0;
            const bool in = undec && !out && (a & und) == 0;
            const uint32_t nin = __ballot_sync(0xffffffffu, in);
            const uint32_t nout = __ballot_sync(0xffffffffu, out);
            reps |= nin;
            und &= ~(nin | nout);
            undec = undec && !in && !out;
        }
        if ((reps >> lane) & 1u) {
            const int v = s_ord[t];
            if (clusters) clusters[c + __popc(reps & ((1u << lane) - 1u))] = v;
            atomicOr(&rep[v >> 5], 1u << (v & 31));
        }
        c += __popc(reps);
        repw = lane == g ? reps : repw;
    }
    if (lane == 0) *count = c;
}

int mis_block() { return MIS_NB; }

// Greedy clustering over the bit matrix in the given order; rep (W words), alive8
// (MIS_NB bytes), adj (MIS_NB*32 words) and count (1 int) are scratch; rep and count must
// be zero on entry.
cudaError_t launch_greedy_clusters(const uint32_t *bits, int W, const int *order, int total,
                                   uint32_t *rep, uint8_t *alive8, uint32_t *adj, int *clusters,
                                   int *count, cudaStream_t stream)
{
    // per device, cheap: set on every call rather than tracking which devices have it
    cudaError_t e = cudaFuncSetAttribute(k_mis_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         MIS_SMEM);
    if (e != cudaSuccess) return e;
    for (int base = 0; base < total; base += MIS_NB) {
        const int cnt = min(MIS_NB, total - base);
        k_mis_scan<<<(cnt + 7) / 8, 256, 0, stream>>>(bits, W, order, base, cnt, rep, alive8, adj);
        k_mis_resolve<<<1, 1024, MIS_SMEM, stream>>>(adj, alive8, order, base, cnt, rep, clusters,
                                                     count);
    }
    return cudaGetLastError();
}

}  // namespace tcu
