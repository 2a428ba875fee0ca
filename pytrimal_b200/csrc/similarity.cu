// similarity.cu -- K4: per-column similarity accumulators.
//
// Replaces simd::calculateSimilarityVectors<V> (vendor/trimal/include/Platform/
// template.h:69-204).  The reference adds, for every column, the terms
//     num += (1 - id[j,k]) * D[a_j][a_k];   den += (1 - id[j,k])
// over all residue pairs j<k in lexicographic order with fp32 accumulators and
// no fused multiply-add (template.h:154-183; AVX2.cpp is built with -mavx2
// only).  fp32 addition is not associative and at n ~ 10^4 the running sum is
// far from the exact value, so the only way to reproduce the
// reference's bits is to replay its order: one thread owns one column and
// performs the same rounded operations in the same sequence
// (__fsub_rn / __fmul_rn / __fadd_rn keep nvcc from contracting them).
// Pairs the reference skips because a row holds a gap (template.h:157-160,
// 170-173) are replayed as "+ 0.0f", which is exact, so the loop is
// branch-free.  The kernel is bound by the dependent-add latency, not by
// bandwidth or FLOPs; no roofline fraction applies.
#include <algorithm>

#include "tcu_internal.cuh"

namespace tcu {

constexpr int DSTRIDE = SIM_MAX_POS + 1;  // 29: odd stride, index npos = "gap" (all zeros)

// ---------------------------------------------------------------------------
// byte -> similarity code (template.h:129-150): upper-case, gap/indet -> SIM_GAP,
// outside 'A'..'Z' -> SIM_INCORRECT, no matrix row -> SIM_UNDEFINED.  The LUT is
// built on the host.  The first offending cell in the reference's scan order
// (columns ascending, rows ascending, skipped columns ignored) is found with a
// 64-bit atomicMin on ((col*nseq + row) << 8 | byte).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sim_codes(const uint8_t *__restrict__ raw, int nseq,
                                                   int ncol, size_t pitch,
                                                   const uint8_t *__restrict__ lut256,
                                                   const uint8_t *__restrict__ col_skip,
                                                   uint8_t *__restrict__ codes,
                                                   unsigned long long *__restrict__ first_error)
{
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = lut256[threadIdx.x];
    __syncthreads();
    const int groups = (int)(pitch >> 4);
    const long long total = (long long)nseq * groups;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / groups);
        const int g = (int)(idx - (long long)r * groups);
        const size_t off = (size_t)r * pitch + (size_t)g * 16;
        const uint4 v = *reinterpret_cast<const uint4 *>(raw + off);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t ow = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int col = g * 16 + q * 4 + b;
                const uint32_t byte = (w[q] >> (8 * b)) & 0xFF;
                uint32_t code = SIM_GAP;
                if (col < ncol) {
                    code = lut[byte];
                    if (code >= SIM_UNDEFINED && code != SIM_GAP && !col_skip[col]) {
                        uint32_t up = (byte >= 'a' && byte <= 'z') ? byte - 32 : byte;
                        atomicMin(first_error,
                                  (((unsigned long long)col * nseq + r) << 8) | up);
                    }
                }
                ow |= code << (8 * b);
            }
            o[q] = ow;
        }
        *reinterpret_cast<uint4 *>(codes + off) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

cudaError_t launch_sim_codes(const uint8_t *raw, int nseq, int ncol, size_t pitch,
                             const uint8_t *lut256, const uint8_t *col_skip, uint8_t *codes,
                             unsigned long long *first_error, cudaStream_t stream)
{
    if (nseq == 0 || ncol == 0) return cudaSuccess;
    const long long total = (long long)nseq * (long long)(pitch >> 4);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    k_sim_codes<<<blocks, 256, 0, stream>>>(raw, nseq, ncol, pitch, lut256, col_skip, codes,
                                            first_error);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// One warp per 32 adjacent columns, one column per lane.  For row j the warp
// reads 32 adjacent codes (one 32-byte sector); id[j,k] is the same address
// for every lane (a broadcast load that stays in L2: the packed identity
// array is read once per warp).  dist lives in shared memory with one extra
// all-zero row/column that gap codes are redirected to.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_similarity(const uint8_t *__restrict__ codes, int nseq,
                                                   int ncol, size_t pitch,
                                                   const float *__restrict__ identities,
                                                   const float *__restrict__ dist, int npos,
                                                   const uint8_t *__restrict__ col_skip,
                                                   float *__restrict__ num_out,
                                                   float *__restrict__ den_out)
{
    __shared__ float D[DSTRIDE * DSTRIDE];
    for (int i = threadIdx.x; i < DSTRIDE * DSTRIDE; i += 32) {
        const int a = i / DSTRIDE, b = i - a * DSTRIDE;
        D[i] = (a < npos && b < npos) ? dist[a * npos + b] : 0.0f;
    }
    __syncwarp();

    const int col = blockIdx.x * 32 + threadIdx.x;
    const bool active = col < ncol && !col_skip[col];
    if (!__any_sync(0xffffffffu, active)) return;
    const int ccol = min(col, ncol - 1);  // inactive lanes read a valid address
    const uint8_t *cp = codes + ccol;

    float num = 0.0f, den = 0.0f;
    const unsigned long long n = (unsigned long long)nseq;
    for (int j = 0; j < nseq - 1; j++) {
        const uint32_t cj = cp[(size_t)j * pitch];
        const bool gj = cj == SIM_GAP;
        if (__all_sync(0xffffffffu, gj || !active)) continue;  // nothing to add in any lane
        const float *drow = D + (gj ? npos : cj) * DSTRIDE;
        // identities[(j,k)] = idrow[k]
        const float *idrow = identities + ((unsigned long long)j * n -
                                           ((unsigned long long)j * (j + 1)) / 2 - j - 1);
        int k = j + 1;
#pragma unroll 1
        for (; k + 4 <= nseq; k += 4) {
            uint32_t ck[4];
            float w[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                ck[u] = cp[(size_t)(k + u) * pitch];
                w[u] = __fsub_rn(1.0f, __ldg(idrow + k + u));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const bool gk = ck[u] == SIM_GAP;
                const float d = drow[gk ? npos : ck[u]];
                const float we = (gj || gk) ? 0.0f : w[u];
                num = __fadd_rn(num, __fmul_rn(we, d));
                den = __fadd_rn(den, we);
            }
        }
        for (; k < nseq; k++) {
            const uint32_t c = cp[(size_t)k * pitch];
            const bool gk = c == SIM_GAP;
            const float w = __fsub_rn(1.0f, __ldg(idrow + k));
            const float d = drow[gk ? npos : c];
            const float we = (gj || gk) ? 0.0f : w;
            num = __fadd_rn(num, __fmul_rn(we, d));
            den = __fadd_rn(den, we);
        }
    }
    if (active) {
        num_out[col] = num;
        den_out[col] = den;
    }
}

cudaError_t launch_similarity(const uint8_t *codes, int nseq, int ncol, size_t pitch,
                              const float *identities, const float *dist, int npos,
                              const uint8_t *col_skip, float *num, float *den, int num_sms,
                              cudaStream_t stream)
{
    (void)num_sms;
    if (nseq == 0 || ncol == 0) return cudaSuccess;
    k_similarity<<<(ncol + 31) / 32, 32, 0, stream>>>(codes, nseq, ncol, pitch, identities, dist,
                                                      npos, col_skip, num, den);
    return cudaGetLastError();
}

}  // namespace tcu
