/*
 * trimal_cuda.h -- C ABI of libtrimal_cuda.so, the B200 (sm_100a) implementation
 * of trimAl's per-alignment statistics hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Each entry point replaces one virtual method that trimAl's SIMD platform
 * classes override (reference paths relative to /root/reference/vendor/trimal/):
 *
 *   tcu_gaps        <- statistics::Gaps::CalculateVectors()
 *                      include/Platform/x86/AVX2.h:52-58, source/Platform/x86/AVX2.cpp:133-136,
 *                      kernel include/Platform/template.h:444-502
 *   tcu_identity    <- statistics::Identity::calculateSeqIdentity()
 *                      AVX2.h:67-72, AVX2.cpp:138-141, template.h:320-442
 *   tcu_similarity  <- statistics::Similarity::calculateVectors(bool cutByGap)
 *                      AVX2.h:44-50, AVX2.cpp:128-131, template.h:69-204
 *   tcu_spurious    <- statistics::Overlap::calculateSpuriousVector(float, float*)
 *                      AVX2.h:60-65, AVX2.cpp:143-148, template.h:206-318
 *
 * The host shim that binds them behind a new ComputePlatform::CUDA enumerator
 * is in pytrimal_b200/csrc/shim/ and INTEGRATION.md.
 *
 * Conventions
 *  - Every function returns TCU_OK (0) or a negative tcu_status; the message
 *    of the last failure on the calling thread is in tcu_last_error().
 *  - Host pointers are borrowed for the duration of the call.  Outputs are
 *    caller-owned host buffers (the reference's base classes own them).
 *  - There is no CPU fallback: without a usable CUDA device every compute
 *    call fails with TCU_ERR_NO_DEVICE.
 *  - A handle may be used from one thread at a time; different handles are
 *    independent (own stream, own buffers), so trim() stays re-entrant.
 *  - keep-masks follow trimAl: int arrays where -1 marks a removed
 *    row/column (Alignment::saveSequences / saveResidues); NULL = keep all.
 */
#ifndef TRIMAL_CUDA_H
#define TRIMAL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tcu_status {
    TCU_OK = 0,
    TCU_ERR_NO_DEVICE = -1,        /* no CUDA driver / no sm_100 device                */
    TCU_ERR_OOM = -2,              /* host or device allocation failed                 */
    TCU_ERR_CUDA = -3,             /* any other CUDA runtime failure                   */
    TCU_ERR_INVALID = -4,          /* bad argument                                     */
    TCU_ERR_INCORRECT_SYMBOL = -5, /* similarity: byte outside 'A'..'Z' (template.h:135-138) */
    TCU_ERR_UNDEFINED_SYMBOL = -6, /* similarity: letter without matrix row (:140-144)  */
    TCU_ERR_STATE = -7,            /* call order (e.g. similarity before identity)     */
    TCU_ERR_NCCL = -8              /* NCCL missing or a collective failed (*_all calls) */
} tcu_status;

typedef struct tcu_msa tcu_msa; /* opaque: one alignment resident on one GPU (or replicated
                                  over the devices of the configured set, see below) */

/* Number of usable sm_100 devices (0 when there is no driver or no GPU). */
int tcu_device_count(void);

/*
 * Devices used by handles created with device = TCU_DEVICE_AUTO.  The reference has one compute
 * platform per process and no device notion (statistics::Manager::platform,
 * include/Statistics/Manager.h:51-68); a single-process caller such as pytrimal's
 * platform="cuda" reaches several GPUs by naming them here -- or in the environment variable
 * TRIMAL_CUDA_DEVICES ("all" or e.g. "0,1,2,3"), read on first use.  count = 0 goes back to
 * the environment / device 0.  With more than one device an AUTO handle is replicated: every
 * device uploads 1/N of the rows over its own PCIe link and the shards are exchanged over
 * NVLink; tcu_identity / tcu_representatives split the pair matrix into row-block bands,
 * tcu_gaps / tcu_spurious split the rows; every other call runs on the first device.  Results
 * are bit-identical to the single-device ones.  Alignments below 4 MB stay on one device.
 */
#define TCU_DEVICE_AUTO (-1)
int tcu_set_devices(const int *devices, int count);
int tcu_get_devices(int *devices, int max); /* returns the size of the set */
int tcu_msa_device_count(const tcu_msa *msa); /* devices this handle is replicated on */

/* Message of the last error raised on this thread ("" if none). */
const char *tcu_last_error(void);

/* Library version string. */
const char *tcu_version(void);

/*
 * Upload an alignment.  rows: nseq pointers to ncol bytes each (the
 * std::string::data() of Alignment::sequences).  The bytes are staged through
 * pinned memory and copied to `device`; nothing is retained on the host.
 */
int tcu_msa_create(const char *const *rows, int nseq, int ncol, int device, tcu_msa **out);

/* Same for one contiguous matrix: row r at data + r*stride. */
int tcu_msa_create_strided(const uint8_t *data, int nseq, int ncol, size_t stride, int device,
                           tcu_msa **out);

void tcu_msa_destroy(tcu_msa *msa);

/*
 * The library keeps the large device buffers of destroyed handles (at most 24 per
 * device: identity matrix, threshold bit matrix, packed planes, raw rows) and its pinned
 * staging buffers for reuse by later handles, so that a create / compute / destroy
 * cycle does not go through cudaMalloc / cudaFree; this returns them to the driver.
 */
void tcu_release_cached_memory(void);

int tcu_msa_nseq(const tcu_msa *msa);
int tcu_msa_ncol(const tcu_msa *msa);

/*
 * Gap counts per column over kept rows ('-' only; the column mask is ignored,
 * template.h:460-477).  gaps_in_column: ncol ints.  Optional (may be NULL):
 * num_cols_with_gaps (nseq+1 ints, INCREMENTED like template.h:496-499) and
 * max_gaps (raised, never lowered, :499-500).
 */
int tcu_gaps(tcu_msa *msa, const int *save_seq, int *gaps_in_column, int *num_cols_with_gaps,
             int *max_gaps);

/*
 * Pairwise identity over kept rows i<j and kept columns (template.h:346-437).
 * identities: kept_pairs floats in the reference's packed order.  hit_out /
 * dst_out (optional, kept_pairs ints each) receive the integer counts.
 * identities may be NULL when keep_on_device is set and only the device copy
 * (for a following tcu_similarity) is wanted.
 * indet: 'X' for amino-acid alignments else 'N' (template.h:331).
 * Any byte values are accepted, like the reference's raw byte compares: up to 126 distinct
 * non-gap values go through the packed bit-plane kernel, more (unreachable through trimAl's
 * own symbol validation) through a byte-wise kernel with the same results.
 */
int tcu_identity(tcu_msa *msa, const int *save_seq, const int *save_res, uint8_t indet,
                 float *identities, int *hit_out, int *dst_out, int keep_on_device);

/*
 * Column similarity accumulators (template.h:120-183) for the unmasked
 * alignment.  Uses the device-resident identities of the last tcu_identity
 * (keep_on_device=1, no masks); if `identities` is non-NULL it is uploaded
 * and used instead.
 *   dist  : npos*npos floats, row-major distance matrix (similarityMatrix::distMat)
 *   vhash : 26 ints, letter -> matrix row or -1 (similarityMatrix::vhash)
 *   gaps  : ncol ints (windowed gap counts) or NULL for cutByGap=false
 *   gap_threshold : 0.8f * numberOfResidues, computed by the caller (template.h:108)
 * Outputs (ncol floats each): num, den -- the fp32 accumulators in the
 * reference's exact summation order; mdk (optional) -- the final statistic
 * (template.h:186-200; expf evaluated on the host with the C library).
 * On TCU_ERR_INCORRECT_SYMBOL / TCU_ERR_UNDEFINED_SYMBOL, err_col / err_row /
 * err_byte (optional) identify the first offender in the reference's scan
 * order and the outputs are unspecified.
 */
int tcu_similarity(tcu_msa *msa, uint8_t indet, const float *dist, int npos, const int *vhash,
                   const int *gaps, float gap_threshold, const float *identities, float *num,
                   float *den, float *mdk, int *err_col, int *err_row, int *err_byte);

/*
 * Spurious / overlap vector (template.h:235-310) over ALL rows and columns.
 * ovrlap = (uint32_t)ceil(overlap * (float)(nseq - 1)), computed by the caller
 * (template.h:217-218).  spurious: nseq floats.
 */
int tcu_spurious(tcu_msa *msa, uint8_t indet, uint32_t ovrlap, float *spurious);

/*
 * Row-band variant for drivers that shard the pair matrix across GPUs: only
 * the rows of row-blocks [block_begin, block_end) (tcu_identity_band_rows() =
 * 128 kept rows per block; block_end < 0 = to the end) are computed and copied
 * to `identities`, a HOST slice whose element 0 is packed offset
 * tcu_identity_row_offset(kept, 128*block_begin).  Bands of different ranks
 * concatenate to the full array.
 */
int tcu_identity_band(tcu_msa *msa, const int *save_seq, const int *save_res, uint8_t indet,
                      int block_begin, int block_end, float *identities);

/* ---- device-resident variants (benchmarks, multi-GPU drivers) -------------
 * Same kernels, but results stay in device memory owned by the caller and no
 * host synchronisation is done beyond what the arguments require.            */

/* Rows per row-block of the band interface (the kernel's tile height, 128). */
int tcu_identity_band_rows(void);

/* Number of row-blocks the kept rows are tiled into for tcu_identity_device / _band. */
int tcu_identity_row_blocks(int kept_rows);

/* Work (pair-matrix tiles) in row-blocks < block: lets a driver cut bands of equal work. */
long long tcu_identity_tiles_before(int kept_rows, int block);

/* The order in which a launch over row-blocks [block_begin, block_end) visits its tiles
 * (tile indices tcu_identity_tiles_before(begin) .. tiles_before(end) - 1): row-block and
 * 64-row column block of one tile.  Groups of 8 row-blocks share a column block before the
 * next one is touched (L2 reuse of the operand); exposed so that the order -- a bijection
 * onto {(I, j) : begin <= I < end, 2 I <= j < ceil(kept_rows / 64)} -- can be checked on the
 * host.  No reference counterpart (template.h:395-441 walks pairs row by row). */
int tcu_identity_tile(int kept_rows, int block_begin, int block_end, long long tile,
                      int *row_block, int *col_block64);

/* Packed-array offset of the first pair whose first row is i (kept-index space). */
size_t tcu_identity_row_offset(int kept_rows, int i);

/*
 * Prepare the packed bit-planes for (save_seq, save_res, indet) on the device.
 * Must precede tcu_identity_device; reusable across calls while the masks do
 * not change.  kept_rows_out (optional) receives the number of kept rows.
 */
int tcu_identity_prepare(tcu_msa *msa, const int *save_seq, const int *save_res, uint8_t indet,
                         int *kept_rows_out);

/*
 * Compute the rows of the packed identity array that belong to row-blocks
 * [block_begin, block_end) (tcu_identity_band_rows() kept rows per block) into
 * d_out, a DEVICE buffer whose element 0 corresponds to packed offset
 * tcu_identity_row_offset(kept, 128*block_begin).  Asynchronous on the
 * handle's stream; use tcu_msa_sync() to wait.
 */
int tcu_identity_device(tcu_msa *msa, int block_begin, int block_end, float *d_out);

/* Wait for all work queued on the handle's stream. */
int tcu_msa_sync(tcu_msa *msa);

/* The handle's CUDA stream (cudaStream_t) and device ordinal. */
void *tcu_msa_stream(tcu_msa *msa);
int tcu_msa_device(const tcu_msa *msa);

/*
 * Device-side durations (CUDA events on the handle's stream) of the kernels
 * launched by the most recent call on this handle, in milliseconds.
 */
typedef struct tcu_timings {
    float h2d_ms;        /* host -> device copies                          */
    float pack_ms;       /* bit-plane / code packing kernels               */
    float kernel_ms;     /* the statistic's main kernel(s)                 */
    float d2h_ms;        /* device -> host copies                          */
    int kernel_launches; /* number of kernels launched by the call         */
    float comm_ms;       /* NCCL collectives of the *_all calls            */
} tcu_timings;

int tcu_msa_timings(const tcu_msa *msa, tcu_timings *out);

/* ---- consumers of the device-resident identity matrix ----
 * The three host walks over Identity::identities in vendor/trimal/source/Cleaner.cpp
 * -- selectMethod (:46-99), getCutPointClusters (:1026-1156) and
 * calculateRepresentativeSeq (:1398-1466) -- done where the matrix already is, so
 * that its 4*P bytes never cross PCIe.  All need a preceding
 * tcu_identity(..., keep_on_device=1) with every row kept (the reference's own
 * index arithmetic in these walks assumes that too); otherwise TCU_ERR_STATE.  */

/* 1 when such a matrix is resident on the handle's device. */
int tcu_identity_resident(const tcu_msa *msa);

/* Copy the resident matrix (nseq*(nseq-1)/2 floats, reference's packed order) to the host. */
int tcu_identity_download(tcu_msa *msa, float *identities);

/*
 * Per-row statistics of the matrix read as symmetric, in the reference's fp32
 * order (j ascending): upper_only = 0 -> over all j != i (selectMethod's mx / avg,
 * Cleaner.cpp:68-80); upper_only = 1 -> over j > i (getCutPointClusters' max / min /
 * avg, :1054-1063).  Neutral start values as the reference: max 0, min 1, sum 0.
 * Each output is nseq floats; any may be NULL.
 */
int tcu_identity_row_stats(tcu_msa *msa, int upper_only, float *row_max, float *row_min,
                           float *row_sum);

/*
 * Greedy clustering shared by calculateRepresentativeSeq (Cleaner.cpp:1427-1447) and
 * the search loop of getCutPointClusters (:1100-1118): walking order[0..count) (row
 * indices, distinct), a sequence becomes the representative of a new cluster iff no
 * earlier representative has identity > threshold with it.  clusters (optional,
 * count ints) receives the representatives in creation order, *n_clusters their
 * number.
 */
int tcu_identity_clusters(tcu_msa *msa, const int *order, int count, float threshold,
                          int *clusters, int *n_clusters);

/* ---- alignment-wide scans of the host layer ------------ */

/*
 * Number of occurrences of every byte value over all rows and columns (256 counts).
 * It is all that utils::checkAlignmentType (source/utils.cpp:476-545; reached through
 * Alignment::getAlignmentType on every trim()) needs from its O(nseq*ncol) scan with
 * seven std::string::find calls per byte: each of its counters is a sum of bins.
 */
int tcu_byte_histogram(tcu_msa *msa, unsigned long long *hist256);

/* Alignment::getSequenceLength (source/Alignment/Alignment.cpp:296-298) of every row:
 * ncol minus the number of '-' bytes.  lengths: nseq ints. */
int tcu_sequence_lengths(tcu_msa *msa, int *lengths);

/*
 * Per row, the number of bytes other than '-' over the kept columns (save_res; NULL = all
 * columns = tcu_sequence_lengths).  Together with tcu_gaps(save_seq) -- a column holds only
 * gaps iff its count equals the number of kept rows -- this is everything
 * Cleaner::removeAllGapsSeqsAndCols (source/Cleaner.cpp:1331-1396) scans the alignment for.
 * residues: nseq ints (rows removed by the caller's row mask are simply ignored by it).
 */
int tcu_row_residues(tcu_msa *msa, const int *save_res, int *residues);

/*
 * Two 64-bit hashes of every row over all columns: hashes[2 i], hashes[2 i + 1] (2 * nseq
 * values).  Equal rows have equal hashes; Cleaner::removeDuplicates
 * (source/Cleaner.cpp:1489-1509: O(nseq^2) string compares) then only has to compare rows
 * inside groups of equal hashes.  The hashes select candidates, the caller's byte compare
 * decides.
 */
int tcu_row_hashes(tcu_msa *msa, unsigned long long *hashes);

/*
 * Host only (no device needed).  The order in which both clustering walks visit the
 * sequences: (length, index) records sorted with the reference's own non-stable
 * quicksort (utils.cpp:246-273) and walked from the end (Cleaner.cpp:1413-1426 /
 * 1078-1089); order[0] is the first representative.  lengths, order: nseq ints.
 */
int tcu_cluster_order(const int *lengths, int nseq, int *order);

/*
 * Host only (no device needed).  tcu_representatives never forms the identity ratios: its
 * identity kernel decides "float(hit) / float(dst) > threshold" (Cleaner.cpp:1435-1440 on
 * the values of template.h:427-434) for every pair in integers, as
 *     mode 0: never   mode 1: always   mode 2: hit > (mul * dst) >> shift   (64-bit product)
 * which is exact for every float threshold and 0 <= hit <= dst < 2^24.  This returns the
 * rule for a threshold so that it can be checked against the division on any machine.
 */
void tcu_threshold_rule(float threshold, int *mode, unsigned *mul, int *shift);

/*
 * Cleaner::calculateRepresentativeSeq (Cleaner.cpp:1398-1466) in one call for an
 * alignment with every row kept: sequence lengths, visiting order, and the greedy
 * clustering at `threshold` (maximumIdent) over the identities of the kept columns
 * (save_res).  The walk only compares identities with the threshold (:1435-1440), so the
 * identity kernel emits one bit per pair instead of the ratios: no float matrix is left on
 * the device (tcu_identity_resident() is 0 afterwards) and the call is not limited by the
 * 4*P bytes of the packed array.  clusters: up to nseq ints (may be NULL).
 */
int tcu_representatives(tcu_msa *msa, const int *save_res, uint8_t indet, float threshold,
                        int *clusters, int *n_clusters);

/* ---- several GPUs, one process per GPU --------------------------
 * The reference has no multi-device path; this is the partition north_star
 * asks for.  Every rank uploads the same alignment to its own GPU
 * (tcu_msa_create), joins a communicator, and calls the same *_all function with
 * the same arguments; each computes its share and the shares are exchanged with
 * NCCL over NVLink, so every rank returns the complete result, bit-identical to
 * the single-GPU call.  NCCL is looked up at run time (libnccl.so.2); without it
 * these calls fail with TCU_ERR_NCCL and nothing else is affected.
 *
 *   identity   : row bands of the pair matrix with equal tile counts
 *                (tcu_shard_blocks), all-gather of the bands -> full matrix on
 *                every GPU (kept on the device for tcu_similarity_all)
 *   similarity : 32-column groups (tcu_shard_range), all-gather of num / den
 *   gaps       : row ranges, all-reduce (integer sum) of the column counts
 *   spurious   : row ranges; all-reduce of the column composition counts, then
 *                all-gather of the per-row ratios
 */
typedef struct tcu_comm tcu_comm;
#define TCU_COMM_ID_BYTES 128

/*
 * Environment read by the library (all optional):
 *   TRIMAL_CUDA_DEVICES          device set of TCU_DEVICE_AUTO handles ("all", "0,1,2,3")
 *   TRIMAL_CUDA_MULTI_MIN_BYTES  alignments smaller than this stay on one device (default 4 MB)
 *   TRIMAL_CUDA_NO_PEER=1        ranks never map each other's memory: every exchange of the
 *                                *_all calls goes through NCCL transfers (must be set on all ranks)
 *   TRIMAL_CUDA_NVTX=1           NVTX ranges around the entry points
 *   TCU_TRACE=1                  host-clock phase timings on stderr
 */

/* Rank 0 creates the 128-byte rendezvous id (ncclGetUniqueId); the caller hands it to the
 * other ranks by any means it has (torch.distributed, MPI, a file). */
int tcu_comm_id(void *id);
int tcu_comm_create(const void *id, int rank, int world, int device, tcu_comm **out);
void tcu_comm_destroy(tcu_comm *comm);
int tcu_comm_rank(const tcu_comm *comm);
int tcu_comm_world(const tcu_comm *comm);

/* tcu_msa_create_strided for a group of ranks that all hold the same alignment on the host:
 * each rank uploads its share of the rows (tcu_shard_range) to the communicator's device and
 * the shares are all-gathered over NVLink.  The handle is an ordinary single-device one. */
int tcu_msa_create_all(tcu_comm *comm, const uint8_t *data, int nseq, int ncol, size_t stride,
                       tcu_msa **out);

/* Host-only partition helpers (no device needed). */
int tcu_shard_range(int total, int granule, int rank, int world, int *begin, int *end);
int tcu_shard_blocks(int kept_rows, int rank, int world, int *block_begin, int *block_end);

/* identities (optional): kept_pairs floats on the host, the complete packed array. */
int tcu_identity_all(tcu_msa *msa, tcu_comm *comm, const int *save_seq, const int *save_res,
                     uint8_t indet, float *identities);
/* Needs a preceding tcu_identity_all without masks (the device-resident matrix). */
int tcu_similarity_all(tcu_msa *msa, tcu_comm *comm, uint8_t indet, const float *dist, int npos,
                       const int *vhash, const int *gaps, float gap_threshold, float *num,
                       float *den, float *mdk, int *err_col, int *err_row, int *err_byte);
int tcu_gaps_all(tcu_msa *msa, tcu_comm *comm, const int *save_seq, int *gaps_in_column,
                 int *num_cols_with_gaps, int *max_gaps);
int tcu_spurious_all(tcu_msa *msa, tcu_comm *comm, uint8_t indet, uint32_t ovrlap,
                     float *spurious);
/* tcu_representatives with the pair matrix split into row bands across the ranks: every
 * rank's identity kernel thresholds its band and stores the bits into ALL ranks' matrices over
 * peer memory (NVLink; NCCL transfers when the ranks cannot map each other's memory); the
 * clustering itself is sequential and runs on every rank. */
int tcu_representatives_all(tcu_msa *msa, tcu_comm *comm, const int *save_res, uint8_t indet,
                            float threshold, int *clusters, int *n_clusters);

#ifdef __cplusplus
}
#endif
#endif /* TRIMAL_CUDA_H */
