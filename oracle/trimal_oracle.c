/*
 * oracle/trimal_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar, plain-C restatement of the four per-alignment statistics that
 * trimAl's SIMD platform template computes (the AVX2 path of pytrimal 0.8.5,
 * vendored trimAl 2.0 RC).  It exists so the CUDA kernels can be checked
 * bit-for-bit on the GPU box, where /root/reference does not exist.
 *
 * Who may use it: tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py.  Nothing under pytrimal_b200/ links,
 * imports or executes this file.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks every function
 * here against (a) the known-answer vectors of vendor/trimal/dataset/
 * example.001.AA.clw, (b) fixtures under tests/golden/ that were produced by
 * the real reference code (oracle/_ref, compiled from /root/reference by
 * oracle/Makefile; generating script tests/golden/make_golden.py) and, when
 * oracle/_ref is present, (c) the reference itself on random inputs.
 *
 * Each function cites the reference lines it restates; paths are relative to
 * /root/reference/vendor/trimal/.
 *
 * Alignment representation used here: one contiguous byte matrix, row r at
 * msa + r*stride, `ncol` meaningful bytes per row.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared  (never -ffast-math: the
 * similarity statistic depends on the exact fp32 operation order).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_INCORRECT_SYMBOL 1 /* byte outside 'A'..'Z' after upper-casing */
#define ORC_ERR_UNDEFINED_SYMBOL 2 /* letter with no row in the similarity matrix */
#define ORC_ERR_NOMEM 3

static inline int is_idgap(uint8_t c, uint8_t indet) { return c == '-' || c == indet; }

/* ------------------------------------------------------------------------
 * Gaps.  include/Platform/template.h:444-502 (per-column count of '-' over
 * kept rows; '-' only, the indetermination symbol is NOT a gap here; the
 * column keep-mask is ignored), histogram + maximum at :496-501.
 * The u8 partial-sum flush of the SIMD version (:478-486) is an
 * implementation detail; this returns the true counts (= Statistics/
 * Gaps.cpp:593-610), which is what the SIMD code produces whenever no row
 * is masked out.
 * num_cols_with_gaps has nseq+1 entries and is incremented (the reference
 * constructor zeroes it, Gaps.cpp:49-53); *max_gaps is only raised.
 * ---------------------------------------------------------------------- */
void orc_gaps(const uint8_t *msa, int nseq, int ncol, size_t stride,
              const int *save_seq, int *gaps_in_column,
              int *num_cols_with_gaps, int *max_gaps)
{
    for (int k = 0; k < ncol; k++) gaps_in_column[k] = 0;
    for (int r = 0; r < nseq; r++) {
        if (save_seq && save_seq[r] == -1) continue;
        const uint8_t *row = msa + (size_t)r * stride;
        for (int k = 0; k < ncol; k++)
            gaps_in_column[k] += (row[k] == '-');
    }
    for (int k = 0; k < ncol; k++) {
        if (num_cols_with_gaps) num_cols_with_gaps[gaps_in_column[k]]++;
        if (max_gaps && gaps_in_column[k] > *max_gaps) *max_gaps = gaps_in_column[k];
    }
}

/* The same statistic with the SIMD path's 8-bit partial-sum behaviour
 * reproduced literally (template.h:460-487): the flush test `j % 255 == 0`
 * sits after the `continue` for masked rows, so more than 255 kept rows can
 * pile up between flushes and the u8 lane wraps.  Only used by a test that
 * documents the divergence; the CUDA path returns true counts. */
void orc_gaps_simd_quirk(const uint8_t *msa, int nseq, int ncol, size_t stride,
                         const int *save_seq, int *gaps_in_column)
{
    uint8_t *acc = (uint8_t *)calloc((size_t)ncol + 1, 1);
    for (int k = 0; k < ncol; k++) gaps_in_column[k] = 0;
    for (int r = 0; r < nseq; r++) {
        if (save_seq && save_seq[r] == -1) continue;
        const uint8_t *row = msa + (size_t)r * stride;
        for (int k = 0; k < ncol; k++) acc[k] = (uint8_t)(acc[k] + (row[k] == '-'));
        if (r % 255 == 0) {
            for (int k = 0; k < ncol; k++) { gaps_in_column[k] += acc[k]; acc[k] = 0; }
        }
    }
    for (int k = 0; k < ncol; k++) gaps_in_column[k] += acc[k];
    free(acc);
}

/* ------------------------------------------------------------------------
 * Gaps window.  source/Statistics/Gaps.cpp:93-153 with utils::roundInt
 * (source/utils.cpp:68-72).  Returns 0, or -1 when half_window > ncol/4
 * (ErrorCode::GapWindowTooBig).  For half_window < 1 nothing is written.
 * ---------------------------------------------------------------------- */
int orc_gaps_window(const int *gaps_in_column, int ncol, int half_window,
                    int *gaps_window)
{
    if (half_window > ncol / 4) return -1;
    if (half_window < 1) return 0;
    const int width = 2 * half_window + 1;
    for (int i = 0; i < ncol; i++) {
        int s = 0;
        for (int j = i - half_window; j <= i + half_window; j++) {
            int src = j < 0 ? -j : (j >= ncol ? 2 * ncol - j - 2 : j);
            s += gaps_in_column[src];
        }
        gaps_window[i] = (int)((double)s / width + 0.5);
    }
    return 0;
}

/* ------------------------------------------------------------------------
 * Pairwise identity.  template.h:320-442; the per-column rule is the scalar
 * tail :419-425, the ratio :427-434, output order :347-358,436 (kept pairs
 * i<j, row-major, no diagonal).
 *   gap class   = raw byte '-' or raw byte `indet`   (case-sensitive)
 *   dst         = kept columns where NOT both rows are in the gap class
 *   hit         = those of them where the raw bytes are equal
 *   identity    = dst ? (float)hit / (float)dst : 0
 * hit_out / dst_out (optional) receive the integer counts the reference
 * never exposes.  Returns the number of pairs written.
 * ---------------------------------------------------------------------- */
size_t orc_identity(const uint8_t *msa, int nseq, int ncol, size_t stride,
                    const int *save_seq, const int *save_res, uint8_t indet,
                    float *identities, int *hit_out, int *dst_out)
{
    size_t pos = 0;
    for (int i = 0; i < nseq; i++) {
        if (save_seq && save_seq[i] == -1) continue;
        const uint8_t *a = msa + (size_t)i * stride;
        for (int j = i + 1; j < nseq; j++) {
            if (save_seq && save_seq[j] == -1) continue;
            const uint8_t *b = msa + (size_t)j * stride;
            int hit = 0, dst = 0;
            for (int k = 0; k < ncol; k++) {
                if (save_res && save_res[k] == -1) continue;
                int both_gap = is_idgap(a[k], indet) && is_idgap(b[k], indet);
                if (both_gap) continue;
                dst++;
                hit += (a[k] == b[k]);
            }
            if (identities) identities[pos] = dst == 0 ? 0.0f : (float)hit / (float)dst;
            if (hit_out) hit_out[pos] = hit;
            if (dst_out) dst_out[pos] = dst;
            pos++;
        }
    }
    return pos;
}

/* ------------------------------------------------------------------------
 * Spurious / overlap vector, pairwise form.  template.h:206-318; per-column
 * rule :280-284, threshold :217-218, final ratio :301-309.  All rows and all
 * columns take part (keep-masks are ignored by the reference).
 * hits_out (optional, nseq*ncol uint32) receives the per-(row,column) counts.
 * ---------------------------------------------------------------------- */
void orc_spurious_pairwise(const uint8_t *msa, int nseq, int ncol, size_t stride,
                           uint8_t indet, float overlap, float *spurious,
                           uint32_t *hits_out)
{
    const uint32_t need = (uint32_t)ceil(overlap * (float)(nseq - 1));
    uint32_t *hits = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(ncol > 0 ? ncol : 1));
    for (int i = 0; i < nseq; i++) {
        const uint8_t *a = msa + (size_t)i * stride;
        memset(hits, 0, sizeof(uint32_t) * (size_t)ncol);
        for (int j = 0; j < nseq; j++) {
            if (j == i) continue;
            const uint8_t *b = msa + (size_t)j * stride;
            for (int k = 0; k < ncol; k++) {
                int res_a = !is_idgap(a[k], indet);
                int res_b = !is_idgap(b[k], indet);
                hits[k] += (uint32_t)((res_a && res_b) || a[k] == b[k]);
            }
        }
        uint32_t good = 0;
        for (int k = 0; k < ncol; k++) good += (hits[k] >= need);
        spurious[i] = (float)good / ncol;
        if (hits_out) memcpy(hits_out + (size_t)i * ncol, hits, sizeof(uint32_t) * (size_t)ncol);
    }
    free(hits);
}

/* Exact closed form of the same vector from per-column byte histograms
 * (derivable from template.h:280-284): with ng_k the number of
 * rows whose byte in column k is outside the gap class and cnt_k[c] the
 * number of rows holding byte c there,
 *     hits(i,k) = ng_k - 1            if row i holds a residue in k
 *               = cnt_k[byte] - 1     otherwise.
 * tests/test_oracle_pinned.py proves it equal to orc_spurious_pairwise; it
 * is used as the checker at sizes where the O(n^2 L) form takes too long. */
void orc_spurious_hist(const uint8_t *msa, int nseq, int ncol, size_t stride,
                       uint8_t indet, float overlap, float *spurious)
{
    const uint32_t need = (uint32_t)ceil(overlap * (float)(nseq - 1));
    uint32_t *ng = (uint32_t *)calloc((size_t)ncol + 1, sizeof(uint32_t));
    uint32_t *cg = (uint32_t *)calloc((size_t)ncol + 1, sizeof(uint32_t)); /* '-'   */
    uint32_t *cx = (uint32_t *)calloc((size_t)ncol + 1, sizeof(uint32_t)); /* indet */
    for (int r = 0; r < nseq; r++) {
        const uint8_t *row = msa + (size_t)r * stride;
        for (int k = 0; k < ncol; k++) {
            if (row[k] == '-') cg[k]++;
            else if (row[k] == indet) cx[k]++;
            else ng[k]++;
        }
    }
    for (int i = 0; i < nseq; i++) {
        const uint8_t *row = msa + (size_t)i * stride;
        uint32_t good = 0;
        for (int k = 0; k < ncol; k++) {
            uint32_t h = row[k] == '-' ? cg[k] : (row[k] == indet ? cx[k] : ng[k]);
            good += (h - 1u >= need) && h >= 1u;
        }
        spurious[i] = (float)good / ncol;
    }
    free(ng); free(cg); free(cx);
}

/* ------------------------------------------------------------------------
 * Distance matrix of a similarity matrix.  source/Statistics/
 * similarityMatrix.cpp:273-283 (same loop in defaultAASimMatrix :259-271 of
 * that function): Euclidean distance between columns, accumulated in fp32,
 * sqrt in double then narrowed.  sim and dist are npos*npos row-major.
 * ---------------------------------------------------------------------- */
void orc_distance_matrix(const float *sim, int npos, float *dist)
{
    for (int i = 0; i < npos * npos; i++) dist[i] = 0.0f;
    for (int j = 0; j < npos; j++)
        for (int i = 0; i < npos; i++) {
            if (i == j || dist[i * npos + j] != 0.0f) continue;
            float sum = 0.0f;
            for (int k = 0; k < npos; k++) {
                float d = sim[k * npos + j] - sim[k * npos + i];
                sum += d * d;
            }
            sum = (float)sqrt(sum);
            dist[i * npos + j] = sum;
            dist[j * npos + i] = sum;
        }
}

/* ------------------------------------------------------------------------
 * Column similarity (MDK).  template.h:69-204: gap cut :108,122-125 (note
 * the threshold is 0.8f * number_of_RESIDUES), symbol mapping
 * :129-150, the ordered fp32 accumulation :153-183 and the final transform
 * :186-200.  `identities` is the packed array of orc_identity on the same
 * (unmasked) alignment; it is walked by original row index (:158,171,181).
 * `gaps` is the (windowed) gap vector or NULL for cutByGap=false.
 * vhash has 26 entries ('A'..'Z' -> matrix row or -1).
 * num_out/den_out (optional) receive the raw accumulators per column
 * (0 for skipped columns).  On a symbol error returns the code, and the
 * offending (column,row,byte) through err_col/err_row/err_byte.
 * ---------------------------------------------------------------------- */
int orc_similarity(const uint8_t *msa, int nseq, int ncol, size_t stride,
                   uint8_t indet, const float *identities, const int *gaps,
                   int number_of_residues, const float *dist, int npos,
                   const int *vhash, float *mdk, float *num_out, float *den_out,
                   int *err_col, int *err_row, int *err_byte)
{
    const float gap_threshold = 0.8F * number_of_residues;
    uint8_t *code = (uint8_t *)malloc((size_t)nseq + 1);
    uint8_t *isgap = (uint8_t *)malloc((size_t)nseq + 1);
    if (!code || !isgap) { free(code); free(isgap); return ORC_ERR_NOMEM; }

    for (int c = 0; c < ncol; c++) {
        if (num_out) num_out[c] = 0.0f;
        if (den_out) den_out[c] = 0.0f;
        if (gaps && (float)gaps[c] >= gap_threshold) { mdk[c] = 0.0f; continue; }

        for (int r = 0; r < nseq; r++) {
            uint8_t ch = msa[(size_t)r * stride + c];
            if (ch >= 'a' && ch <= 'z') ch = (uint8_t)(ch & ~0x20);
            if (ch == indet || ch == '-') { isgap[r] = 1; continue; }
            isgap[r] = 0;
            int bad = 0;
            if (ch < 'A' || ch > 'Z') bad = ORC_ERR_INCORRECT_SYMBOL;
            else if (vhash[ch - 'A'] == -1) bad = ORC_ERR_UNDEFINED_SYMBOL;
            if (bad) {
                if (err_col) *err_col = c;
                if (err_row) *err_row = r;
                if (err_byte) *err_byte = ch;
                free(code); free(isgap);
                return bad;
            }
            code[r] = (uint8_t)vhash[ch - 'A'];
        }

        float num = 0.0f, den = 0.0f;
        size_t pos = 0;
        for (int j = 0; j < nseq; j++) {
            if (isgap[j]) { pos += (size_t)(nseq - j - 1); continue; }
            const float *drow = dist + (size_t)code[j] * npos;
            for (int k = j + 1; k < nseq; k++, pos++) {
                if (isgap[k]) continue;
                float w = 1.0F - identities[pos];
                float t = w * drow[code[k]];   /* rounded product ... */
                num = num + t;                  /* ... then rounded sum: no FMA */
                den = den + w;
            }
        }
        if (num_out) num_out[c] = num;
        if (den_out) den_out[c] = den;
        if (den == 0) mdk[c] = 0.0f;
        else {
            float q = num / den;
            mdk[c] = q < 0 ? 1.0f : expf(-q);
        }
    }
    free(code); free(isgap);
    return ORC_OK;
}

/* Final transform alone (template.h:186-200), for checking a device that
 * returns num/den and leaves expf to the host. */
void orc_similarity_finish(const float *num, const float *den, int ncol, float *mdk)
{
    for (int c = 0; c < ncol; c++) {
        if (den[c] == 0) mdk[c] = 0.0f;
        else {
            float q = num[c] / den[c];
            mdk[c] = q < 0 ? 1.0f : expf(-q);
        }
    }
}

/* ------------------------------------------------------------------------
 * Similarity window.  source/Statistics/Similarity.cpp:212-269: fp32 running
 * sum over the mirrored window in ascending j, divided by (float)(2h+1).
 * Returns -1 when half_window > ncol/4 (SimilarityWindowTooBig).
 * ---------------------------------------------------------------------- */
int orc_similarity_window(const float *mdk, int ncol, int half_window, float *mdk_window)
{
    if (half_window > ncol / 4) return -1;
    if (half_window < 1) return 0;
    const int width = 2 * half_window + 1;
    for (int i = 0; i < ncol; i++) {
        float s = 0.F;
        for (int j = i - half_window; j <= i + half_window; j++) {
            int src = j < 0 ? -j : (j >= ncol ? 2 * ncol - j - 2 : j);
            s += mdk[src];
        }
        mdk_window[i] = s / (float)width;
    }
    return 0;
}

/* ========================================================================
 * Consumers of the identity matrix: the three host walks
 * of source/Cleaner.cpp over Identity::identities, for the case the
 * reference's own index arithmetic supports (no masked rows, so
 * numberOfSequences == originalNumberOfSequences == n).
 * ====================================================================== */

/* packed position of pair (i<j): Cleaner.cpp:72-75 / 1105-1108 / 1431-1434 */
static inline size_t orc_pair_pos(size_t n, size_t a, size_t b)
{
    size_t mn = a < b ? a : b, mx = a < b ? b : a;
    size_t sq = (mn + 1) * (mn + 1);
    return n * mn - ((sq + (mn + 1)) / 2) + mx;
}

/* Alignment::getSequenceLength (Alignment/Alignment.cpp:296-298): bytes that are
 * not '-' (only '-'). */
void orc_sequence_lengths(const uint8_t *msa, int nseq, int ncol, size_t stride, int *lengths)
{
    for (int r = 0; r < nseq; r++) {
        int g = 0;
        for (int k = 0; k < ncol; k++) g += msa[(size_t)r * stride + k] == '-';
        lengths[r] = ncol - g;
    }
}

/* utils::quicksort(int **vect, int ini, int fin) (utils.cpp:246-273): sorts
 * (key, index) records by key, pivot = last element held as a float, not
 * stable.  The permutation it leaves decides the clustering order, so it is
 * restated step by step.  key/idx are parallel arrays standing for vect[i][0]
 * and vect[i][1]. */
static void orc_qs(int *key, int *idx, int ini, int fin)
{
    if (ini >= fin || fin < 0) return;
    float div = (float)key[fin];
    int i = ini - 1, j = fin, t;
    for (;;) {
        while ((float)key[++i] < div)
            if (i == fin) break;
        while ((float)key[--j] > div)
            if (j == 0) break;
        if (i < j) {
            t = key[i]; key[i] = key[j]; key[j] = t;
            t = idx[i]; idx[i] = idx[j]; idx[j] = t;
        } else
            break;
    }
    t = key[i]; key[i] = key[fin]; key[fin] = t;
    t = idx[i]; idx[i] = idx[fin]; idx[fin] = t;
    orc_qs(key, idx, ini, i - 1);
    orc_qs(key, idx, i + 1, fin);
}

/* Order in which calculateRepresentativeSeq / getCutPointClusters visit the
 * sequences: seqs[] = (length, index), quicksort ascending, then walked from
 * the END (Cleaner.cpp:1413-1426 / 1078-1089).  order[0] is the first cluster
 * representative. */
int orc_cluster_order(const int *lengths, int nseq, int *order)
{
    int *key = (int *)malloc(sizeof(int) * (nseq ? nseq : 1));
    int *idx = (int *)malloc(sizeof(int) * (nseq ? nseq : 1));
    if (!key || !idx) { free(key); free(idx); return ORC_ERR_NOMEM; }
    for (int i = 0; i < nseq; i++) { key[i] = lengths[i]; idx[i] = i; }
    orc_qs(key, idx, 0, nseq - 1);
    for (int i = 0; i < nseq; i++) order[i] = idx[nseq - 1 - i];
    free(key); free(idx);
    return ORC_OK;
}

/* The greedy walk (Cleaner.cpp:1427-1447, and the inner loop :1100-1118 of
 * getCutPointClusters, which breaks at the first hit instead of looking for
 * the best one -- same set of representatives): order[k] opens a new cluster
 * iff no existing representative has identity > thr with it.  Returns the
 * number of clusters; clusters (optional) lists the representatives in
 * creation order. */
int orc_greedy_clusters(const float *identities, int nseq, const int *order, int count, float thr,
                        int *clusters)
{
    int *cl = clusters ? clusters : (int *)malloc(sizeof(int) * (count ? count : 1));
    int ncl = 0;
    for (int k = 0; k < count; k++) {
        int s = order[k], j;
        for (j = 0; j < ncl; j++)
            if (identities[orc_pair_pos((size_t)nseq, (size_t)s, (size_t)cl[j])] > thr) break;
        if (j == ncl) cl[ncl++] = s;
    }
    if (!clusters) free(cl);
    return ncl;
}

/* Per-row statistics.  upper_only = 0: selectMethod's inner loop
 * (Cleaner.cpp:68-80), all j != i in ascending j; upper_only = 1:
 * getCutPointClusters' (:1054-1063), j > i. */
void orc_identity_row_stats(const float *identities, int nseq, int upper_only, float *row_max,
                            float *row_min, float *row_sum)
{
    for (int i = 0; i < nseq; i++) {
        float mx = 0, mn = 1, avg = 0;
        for (int j = upper_only ? i + 1 : 0; j < nseq; j++) {
            if (j == i) continue;
            float v = identities[orc_pair_pos((size_t)nseq, (size_t)i, (size_t)j)];
            mx = mx < v ? v : mx;
            mn = v < mn ? v : mn;
            avg += v;
        }
        if (row_max) row_max[i] = mx;
        if (row_min) row_min[i] = mn;
        if (row_sum) row_sum[i] = avg;
    }
}

/* Cleaner::selectMethod (Cleaner.cpp:46-99).  Returns 1 for GAPPYOUT, 2 for
 * STRICT (defines.h values are not reproduced; the caller only distinguishes
 * the two); avg_seq / max_seq (optional) receive the two decision values. */
int orc_select_method(const float *identities, int nseq, float *avg_seq, float *max_seq)
{
    float maxSeq = 0, avgSeq = 0;
    for (int i = 0; i < nseq; i++) {
        float mx = 0, avg = 0;
        for (int j = 0; j < nseq; j++) {
            if (i == j) continue;
            float v = identities[orc_pair_pos((size_t)nseq, (size_t)i, (size_t)j)];
            mx = mx < v ? v : mx;
            avg += v;
        }
        avgSeq += avg / (nseq - 1);
        maxSeq += mx;
    }
    avgSeq = avgSeq / nseq;
    maxSeq = maxSeq / nseq;
    if (avg_seq) *avg_seq = avgSeq;
    if (max_seq) *max_seq = maxSeq;
    if (avgSeq >= 0.55) return 1;
    else if (avgSeq <= 0.38) return 2;
    else {
        if (nseq <= 20) return 1;
        if ((maxSeq >= 0.5) && (maxSeq <= 0.65)) return 1;
        return 2;
    }
}

/* Cleaner::getCutPointClusters (Cleaner.cpp:1026-1156): the identity threshold
 * that yields `cluster_number` clusters, found by bisection from the mean
 * identity.  `order` as orc_cluster_order.  iterations (optional) counts the
 * clusterings run. */
float orc_cutpoint_clusters(const float *identities, int nseq, const int *order, int cluster_number,
                            int *iterations)
{
    float max, min, avg, gMax, gMin, startingPoint, prevValue = 0, iter = 0;
    size_t pos = 0;
    int runs = 0;
    if (iterations) *iterations = 0;
    if (cluster_number == nseq) return 1;
    else if (cluster_number == 1) return 0;
    gMax = 0; gMin = 1; startingPoint = 0;
    for (int i = 0; i < nseq; i++) {
        int compared = 0;
        avg = 0; min = 1; max = 0;
        for (int j = i + 1; j < nseq; j++) {
            max = max < identities[pos] ? identities[pos] : max;   /* std::max(max, v) */
            min = identities[pos] < min ? identities[pos] : min;   /* std::min(min, v) */
            avg += identities[pos];
            pos++;
            compared++;
        }
        if (compared > 0) {
            startingPoint += avg / compared;
            gMax = gMax < max ? max : gMax;
            gMin = min < gMin ? min : gMin;
        }
    }
    if (pos > 0) startingPoint /= pos;
    for (;;) {
        int clusterNum = orc_greedy_clusters(identities, nseq, order, nseq, startingPoint, NULL);
        runs++;
        if (clusterNum == cluster_number || iter > 10) break;
        if (clusterNum > cluster_number) {
            gMax = startingPoint;
            startingPoint = (gMax + gMin) / 2;
        } else {
            gMin = startingPoint;
            startingPoint = (gMax + gMin) / 2;
        }
        if (prevValue != clusterNum) {
            iter = 0;
            prevValue = clusterNum;
        } else
            iter++;
    }
    if (iterations) *iterations = runs;
    return startingPoint;
}
