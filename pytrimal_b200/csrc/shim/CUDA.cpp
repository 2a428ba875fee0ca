// Platform/CUDA/CUDA.cpp -- see CUDA.h.  Mirrors source/Platform/x86/AVX2.cpp:
// 126-148 (four thin overrides), with the SIMD template calls replaced by the
// C ABI of libtrimal_cuda.  Error convention: never throw; report
// through debug.report(...) and return false / leave zero-filled outputs.
#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "Alignment/Alignment.h"
#include "InternalBenchmarker.h"
#include "Platform/CUDA/CUDA.h"
#include "Statistics/Manager.h"
#include "Statistics/similarityMatrix.h"
#include "defines.h"
#include "reportsystem.h"
#include "residueValues.h"
#include "trimal_cuda.h"
#include "utils.h"

namespace statistics {

namespace {

std::mutex g_ctx_mutex;
// rows pointer -> live upload.  weak_ptr: the upload dies with the last
// statistic object that uses it.
std::map<const void *, std::weak_ptr<CUDAContext>> g_contexts;

void report_failure(const char *what)
{
  std::string msg = std::string(what) + ": " + tcu_last_error();
  debug.report(ErrorCode::SomethingWentWrong_reportToDeveloper, new std::string[1]{msg});
}

char indet_of(const Alignment *alig)
{
  return (alig->getAlignmentType() & SequenceTypes::AA) ? 'X' : 'N';  // template.h:99,221,331
}

}  // namespace

int cudaPlatformDeviceCount() { return tcu_device_count(); }

CUDAContext::~CUDAContext()
{
  {
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    auto it = g_contexts.find(rows_key);
    if (it != g_contexts.end() && it->second.expired()) g_contexts.erase(it);
  }
  tcu_msa_destroy(handle);  // waits for the device: outside the table's lock
}

IdentityShare::IdentityShare() : id([] {
  static std::atomic<uint64_t> next{1};
  return next.fetch_add(1);
}())
{
}

std::shared_ptr<CUDAContext> CUDAContext::acquire(Alignment *alig)
{
  const void *key = alig->sequences;
  // this alignment's manager already owns the upload of these rows
  if (alig->Statistics->cudaContext) {
    auto sp = std::static_pointer_cast<CUDAContext>(alig->Statistics->cudaContext);
    if (sp->rows_key == key) return sp;
    alig->Statistics->cudaContext.reset();
  }
  {
    // an alignment sharing the same rows (the source of a copy, or a copy) does
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    auto it = g_contexts.find(key);
    if (it != g_contexts.end())
      if (auto sp = it->second.lock()) {
        alig->Statistics->cudaContext = sp;
        return sp;
      }
  }
  // (inside Alignment::fillMatrices the original* counters are not set yet)
  const int n = alig->originalNumberOfSequences > 0 ? alig->originalNumberOfSequences
                                                    : alig->numberOfSequences;
  const int L = alig->originalNumberOfResidues > 0 ? alig->originalNumberOfResidues
                                                   : alig->numberOfResidues;
  std::vector<const char *> rows((size_t)n);
  for (int i = 0; i < n; i++) rows[i] = alig->sequences[i].data();
  tcu_msa *h = nullptr;
  // TCU_DEVICE_AUTO: the device set of tcu_set_devices / TRIMAL_CUDA_DEVICES (device 0 by
  // default; several GPUs of the box when the user names them)
  if (tcu_msa_create(rows.data(), n, L, TCU_DEVICE_AUTO, &h) != TCU_OK) {
    report_failure("CUDA platform: alignment upload failed");
    return nullptr;
  }
  auto sp = std::make_shared<CUDAContext>();
  sp->handle = h;
  sp->rows_key = key;
  alig->Statistics->cudaContext = sp;
  std::lock_guard<std::mutex> lk(g_ctx_mutex);
  g_contexts[key] = sp;
  return sp;
}

// ---------------------------------------------------------------------------
// Alignment type.  The reference classifies every byte with seven string searches
// (utils.cpp:487-512); all it keeps are six counters, i.e. sums over byte values, so
// one 256-bin histogram from the device and the same membership tests per byte VALUE
// give the same counters, the same early NotDefined and the same warnings.
// ---------------------------------------------------------------------------
bool cudaAlignmentType(const Alignment *calig, int *type)
{
  StartTiming("bool cudaAlignmentType(const Alignment *, int *) ");
  Alignment *alig = const_cast<Alignment *>(calig);
  const int n = alig->originalNumberOfSequences, L = alig->originalNumberOfResidues;
  if (alig->sequences == nullptr || n <= 0 || L <= 0 || alig->numberOfSequences != n) return false;
  for (int i = 0; i < n; i++)
    if (alig->sequences[i].size() != (size_t)L) return false;
  std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig);
  if (!ctx) return false;
  unsigned long long hist[256];
  {
    std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
    if (tcu_byte_histogram(ctx->handle, hist) != TCU_OK) {
      report_failure("CUDA platform: byte histogram failed");
      return false;
    }
  }
  static const std::string dna = "ACGT", rna = "ACGU", gapSymbols = "-?.";  // utils.cpp:477-480
  size_t rnaCount = 0, dnaCount = 0, degNtCount = 0, aaCount = 0, degAaCount = 0, altAaCount = 0;
  for (int b = 0; b < 256; b++) {
    const size_t k = hist[b];
    if (k == 0) continue;
    char c = (char)b;
    if (utils::checkPattern(gapSymbols, c)) continue;
    c = utils::toUpper(c);
    const bool isRNA = utils::checkPattern(rna, c), isDNA = utils::checkPattern(dna, c);
    const bool isDegNN = utils::checkPattern(degenerateNucleotideResidues, c);
    const bool isAA = utils::checkPattern(aminoAcidResidues, c);
    const bool isDegAA = utils::checkPattern(ambiguousAA, c);
    const bool isAltAA = utils::checkPattern(alternativeAminoAcidResidues, c);
    if (!(isRNA || isDNA || isDegNN || isAA || isDegAA || isAltAA)) {
      *type = SequenceTypes::NotDefined;  // utils.cpp:501-503: first unknown symbol ends the scan
      return true;
    }
    if (isRNA) rnaCount += k;
    if (isDNA) dnaCount += k;
    if (isDegNN && !isDNA && !isRNA) degNtCount += k;
    if (isAA) aaCount += k;
    if (isDegAA) degAaCount += k;
    if (isAltAA) altAaCount += k;
  }
  // decision and warnings of utils.cpp:514-545
  dnaCount += degNtCount;
  rnaCount += degNtCount;
  aaCount += degAaCount + altAaCount;
  if (aaCount > dnaCount && aaCount > rnaCount) {
    if (altAaCount > 0) debug.report(WarningCode::AlternativeAminoAcids);
    *type = degAaCount > 0 ? SequenceTypes::AA | SequenceTypes::DEG : SequenceTypes::AA;
    return true;
  }
  const bool asDNA = dnaCount >= aaCount && dnaCount >= rnaCount;
  const char *self = asDNA ? "DNA" : "RNA", *other = asDNA ? "RNA" : "DNA";
  if (aaCount == (asDNA ? dnaCount : rnaCount))
    debug.report(WarningCode::IndeterminateAlignmentType, new std::string[3]{self, "AA", self});
  if (dnaCount == rnaCount)
    debug.report(WarningCode::IndeterminateAlignmentType, new std::string[3]{self, other, self});
  const int base = asDNA ? SequenceTypes::DNA : SequenceTypes::RNA;
  if (degNtCount > 0) {
    debug.report(WarningCode::DegenerateNucleotides);
    *type = base | SequenceTypes::DEG;
  } else
    *type = base;
  return true;
}

// ---------------------------------------------------------------------------
void CUDAGaps::CalculateVectors()
{
  StartTiming("void CUDAGaps::CalculateVectors() ");
  const int L = alig->originalNumberOfResidues;
  // valid output even on failure: Cleaner keeps running after a void override
  memset(gapsInColumn, 0, sizeof(int) * L);
  std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig);
  if (!ctx) return;
  std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
  if (tcu_gaps(ctx->handle, alig->saveSequences, gapsInColumn, numColumnsWithGaps, &maxGaps) !=
      TCU_OK)
    report_failure("CUDA platform: gap statistic failed");
}

// Order of a similarity matrix.  numPositions is private in vanilla trimAl and public in
// pytrimal's build (patches/similarityMatrix.h.patch): read it where it is accessible
// (access failure in a decltype is a substitution failure), else report -1.
template <class M>
static auto matrix_order(M *m, int) -> decltype(m->numPositions)
{
  return m->numPositions;
}
template <class M>
static int matrix_order(M *, long)
{
  return -1;
}

// Set while the caller only needs the matrix on the device (the cuda* walks below and
// CUDASimilarity): CUDAIdentity::calculateSeqIdentity then skips the host array.
static thread_local int t_device_only = 0;
struct DeviceOnlyScope {
  DeviceOnlyScope() { t_device_only++; }
  ~DeviceOnlyScope() { t_device_only--; }
};

static bool all_rows_kept(const Alignment *alig)
{
  const int n = alig->originalNumberOfSequences;
  if (alig->numberOfSequences != n) return false;
  for (int i = 0; i < n; i++)
    if (alig->saveSequences[i] == -1) return false;
  return true;
}

CUDAIdentity::CUDAIdentity(Alignment *parent)
    : Identity(parent), share(std::make_shared<IdentityShare>())
{
}

CUDAIdentity::CUDAIdentity(Alignment *parent, Identity *parentIdentity)
    : Identity(parent, parentIdentity)
{
  if (auto *mold = dynamic_cast<CUDAIdentity *>(parentIdentity)) {
    share = mold->share;
  } else {
    share = std::make_shared<IdentityShare>();
    share->host = identities;
  }
}

CUDAIdentity::~CUDAIdentity()
{
  // the base destructor frees `identities` of the LAST object of the group
  // (Identity.cpp:110-118); hand it the array a sibling may have materialized
  if (identities == nullptr && share) {
    std::lock_guard<std::mutex> lk(share->mutex);
    identities = share->host;
  }
}

bool CUDAIdentity::allocateHost()
{
  const int n = alig->originalNumberOfSequences;
  // same allocation as template.h:327-328 (size computed in fp32); owned and
  // delete[]d by the base class (Identity.cpp:110-118)
  const size_t size = ((float)n * n + n) / 2;
  identities = new (std::nothrow) float[size ? size : 1];
  if (identities == nullptr) {
    report_failure("CUDA platform: identity matrix allocation failed");
    return false;
  }
  return true;
}

// the four-override contract: host array filled, device copy kept when it can feed
// the similarity statistic and the Cleaner walks
void CUDAIdentity::computeToHost()
{
  const int n = alig->originalNumberOfSequences;
  const size_t size = ((float)n * n + n) / 2;
  if (!allocateHost()) return;
  std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig);
  int kept = 0;
  for (int i = 0; i < n; i++) kept += alig->saveSequences[i] != -1;
  const int keep_on_device = kept == n;
  const char indet = indet_of(alig);  // before the lock: type detection may use the handle
  std::unique_lock<std::recursive_mutex> lk;
  if (ctx) lk = std::unique_lock<std::recursive_mutex>(ctx->mutex);
  if (!ctx || tcu_identity(ctx->handle, alig->saveSequences, alig->saveResidues, indet,
                           identities, nullptr, nullptr, keep_on_device) != TCU_OK) {
    if (ctx) {
      report_failure("CUDA platform: identity statistic failed");
      ctx->ident_owner = 0;
    }
    memset(identities, 0, sizeof(float) * size);
  } else {
    ctx->ident_owner = keep_on_device ? share->id : 0;
  }
  std::lock_guard<std::mutex> sl(share->mutex);
  share->host = identities;
}

void CUDAIdentity::calculateSeqIdentity()
{
  StartTiming("void CUDAIdentity::calculateSeqIdentity() ");
  if (t_device_only > 0 && computeOnDevice()) return;
  if (identities != nullptr) return;
  computeToHost();
}

bool CUDAIdentity::computeOnDevice()
{
  if (!all_rows_kept(alig)) return false;
  std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig);
  if (!ctx) return false;
  const char indet = indet_of(alig);  // before the lock: type detection may use the handle
  std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
  if (ctx->ident_owner == share->id && tcu_identity_resident(ctx->handle)) return true;
  {
    // a host copy without a device copy: let the reference code walk the host array
    std::lock_guard<std::mutex> sl(share->mutex);
    if (identities != nullptr || share->host != nullptr) return false;
  }
  if (tcu_identity(ctx->handle, alig->saveSequences, alig->saveResidues, indet, nullptr,
                   nullptr, nullptr, /*keep_on_device=*/1) != TCU_OK) {
    report_failure("CUDA platform: identity statistic failed");
    ctx->ident_owner = 0;
    return false;
  }
  ctx->ident_owner = share->id;
  return true;
}

void CUDAIdentity::materialize()
{
  if (identities != nullptr) return;
  {
    std::lock_guard<std::mutex> sl(share->mutex);
    if (share->host != nullptr) {
      identities = share->host;
      return;
    }
  }
  if (std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig)) {
    std::unique_lock<std::recursive_mutex> lk(ctx->mutex);
    if (ctx->ident_owner == share->id && tcu_identity_resident(ctx->handle)) {
      if (!allocateHost()) return;
      if (tcu_identity_download(ctx->handle, identities) == TCU_OK) {
        std::lock_guard<std::mutex> sl(share->mutex);
        share->host = identities;
        return;
      }
      report_failure("CUDA platform: identity matrix download failed");
      delete[] identities;
      identities = nullptr;
    }
  }
  computeToHost();
}

void cudaMaterializeIdentity(Identity *identity)
{
  if (t_device_only > 0) return;
  if (auto *ci = dynamic_cast<CUDAIdentity *>(identity)) ci->materialize();
}

// ---------------------------------------------------------------------------
// Cleaner's walks over the identity matrix, on the device
// ---------------------------------------------------------------------------
namespace {

// the CUDAIdentity of `alig` with its matrix resident on the device (and the upload it is
// resident on), or nullptr
CUDAIdentity *device_identity(Alignment *alig, std::shared_ptr<CUDAContext> &ctx)
{
  if (!all_rows_kept(alig)) return nullptr;
  {
    DeviceOnlyScope scope;
    if (!alig->Statistics->calculateSeqIdentity()) return nullptr;
  }
  auto *ci = dynamic_cast<CUDAIdentity *>(alig->Statistics->identity);
  if (ci == nullptr || !ci->computeOnDevice()) return nullptr;
  ctx = CUDAContext::acquire(alig);
  return ctx ? ci : nullptr;
}

// visiting order of the clustering walks (Cleaner.cpp:1413-1426 / 1078-1089)
bool cluster_order(CUDAContext &ctx, int n, std::vector<int> &order)
{
  std::vector<int> lengths((size_t)n);
  order.resize((size_t)n);
  std::lock_guard<std::recursive_mutex> lk(ctx.mutex);
  return tcu_sequence_lengths(ctx.handle, lengths.data()) == TCU_OK &&
         tcu_cluster_order(lengths.data(), n, order.data()) == TCU_OK;
}

}  // namespace

bool cudaSelectMethod(Alignment *alig, int *method)
{
  StartTiming("bool cudaSelectMethod(Alignment *, int *) ");
  std::shared_ptr<CUDAContext> ctx;
  CUDAIdentity *ci = device_identity(alig, ctx);
  if (ci == nullptr) return false;
  const int n = alig->numberOfSequences;
  std::vector<float> rowMax((size_t)n), rowSum((size_t)n);
  {
    std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
    if (tcu_identity_row_stats(ctx->handle, /*upper_only=*/0, rowMax.data(), nullptr,
                               rowSum.data()) != TCU_OK) {
      report_failure("CUDA platform: identity row statistics failed");
      return false;
    }
  }
  // Cleaner.cpp:80-85: the per-row values enter two running fp32 sums in row order
  float avgSeq = 0, maxSeq = 0;
  for (int i = 0; i < n; i++) {
    avgSeq += rowSum[i] / (n - 1);
    maxSeq += rowMax[i];
  }
  avgSeq = avgSeq / n;
  maxSeq = maxSeq / n;
  // decision table of Cleaner.cpp:89-98
  const bool gappy = avgSeq >= 0.55 ||
                     (avgSeq > 0.38 && (n <= 20 || (maxSeq >= 0.5 && maxSeq <= 0.65)));
  *method = gappy ? GAPPYOUT : STRICT;
  return true;
}

bool cudaCutPointClusters(Alignment *alig, int clusterNumber, float *cut)
{
  StartTiming("bool cudaCutPointClusters(Alignment *, int, float *) ");
  std::shared_ptr<CUDAContext> ctx;
  CUDAIdentity *ci = device_identity(alig, ctx);
  if (ci == nullptr) return false;
  const int n = alig->numberOfSequences;
  std::vector<float> rowMax((size_t)n), rowMin((size_t)n), rowSum((size_t)n);
  {
    std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
    if (tcu_identity_row_stats(ctx->handle, /*upper_only=*/1, rowMax.data(), rowMin.data(),
                               rowSum.data()) != TCU_OK) {
      report_failure("CUDA platform: identity row statistics failed");
      return false;
    }
  }
  // Cleaner.cpp:1049-1071
  float gMax = 0, gMin = 1, startingPoint = 0;
  for (int i = 0; i < n; i++) {
    const int compared = n - 1 - i;
    if (compared > 0) {
      startingPoint += rowSum[i] / compared;
      gMax = std::max(gMax, rowMax[i]);
      gMin = std::min(gMin, rowMin[i]);
    }
  }
  const size_t pairs = (size_t)n * (size_t)(n - 1) / 2;
  if (pairs > 0) startingPoint /= pairs;

  std::vector<int> order;
  if (!cluster_order(*ctx, n, order)) {
    report_failure("CUDA platform: clustering order failed");
    return false;
  }
  // the bisection of Cleaner.cpp:1098-1147; each probe is one clustering on the device
  float prevValue = 0, iter = 0;
  for (;;) {
    int clusterNum = 0;
    {
      std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
      if (tcu_identity_clusters(ctx->handle, order.data(), n, startingPoint, nullptr,
                                &clusterNum) != TCU_OK) {
        report_failure("CUDA platform: clustering failed");
        return false;
      }
    }
    if (clusterNum == clusterNumber || iter > 10) break;
    if (clusterNum > clusterNumber) gMax = startingPoint;
    else gMin = startingPoint;
    startingPoint = (gMax + gMin) / 2;
    if (prevValue != clusterNum) {
      iter = 0;
      prevValue = clusterNum;
    } else
      iter++;
  }
  *cut = startingPoint;
  return true;
}

int *cudaRepresentativeSeq(Alignment *alig, float maximumIdent)
{
  StartTiming("int *cudaRepresentativeSeq(Alignment *, float) ");
  if (!all_rows_kept(alig)) return nullptr;
  const int n = alig->originalNumberOfSequences;
  std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig);
  if (!ctx) return nullptr;
  // Cleaner.cpp:1435-1440: a hit needs identity > maximumIdent AND > max (0 at first)
  const float threshold = maximumIdent < 0 ? 0 : maximumIdent;
  const char indet = indet_of(alig);  // before the lock: type detection may use the handle
  std::vector<int> reps((size_t)n);
  int count = 0;
  {
    std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
    auto *cid = dynamic_cast<CUDAIdentity *>(alig->Statistics->identity);
    if (cid != nullptr && ctx->ident_owner == cid->share->id && tcu_identity_resident(ctx->handle)) {
      // this alignment's matrix is already on the device (an earlier statistic or walk left
      // it there): threshold and walk it
      std::vector<int> order;
      if (!cluster_order(*ctx, n, order) ||
          tcu_identity_clusters(ctx->handle, order.data(), n, threshold, reps.data(), &count) != TCU_OK) {
        report_failure("CUDA platform: clustering failed");
        return nullptr;
      }
    } else {
      // the walk only compares identities with the threshold: one call, the identity kernel
      // emits one bit per pair and no float matrix exists anywhere (and none is left behind:
      // the operand is repacked, whatever was resident is gone)
      ctx->ident_owner = 0;
      if (tcu_representatives(ctx->handle, alig->saveResidues, (uint8_t)indet, threshold, reps.data(),
                              &count) != TCU_OK) {
        report_failure("CUDA platform: clustering failed");
        return nullptr;
      }
    }
  }
  int *repres = new (std::nothrow) int[count + 1];  // freed by the caller (Cleaner.cpp:1201)
  if (repres == nullptr) return nullptr;
  repres[0] = count;
  for (int i = 0; i < count; i++) repres[i + 1] = reps[i];
  return repres;
}

// ---------------------------------------------------------------------------
// Post-trim scans and symbol validation (rank 2)
// ---------------------------------------------------------------------------
bool cudaRemoveAllGaps(Alignment *alig, bool seqs, bool cols, bool keepSequences)
{
  StartTiming("bool cudaRemoveAllGaps(Alignment *, bool, bool, bool) ");
  const int n = alig->originalNumberOfSequences, L = alig->originalNumberOfResidues;
  if (alig->sequences == nullptr || n <= 0 || L <= 0) return false;
  std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig);
  if (!ctx) return false;
  // A sequence that holds only gaps over the kept columns adds only gaps to every kept
  // column, so dropping it (first loop of the reference) cannot change which columns hold
  // only gaps (second loop): both questions are answered from the masks as they are now.
  std::vector<int> rowResidues, colGaps;
  int keptRows = 0;
  for (int i = 0; i < n; i++) keptRows += alig->saveSequences[i] != -1;
  {
    std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
    if (seqs) {
      rowResidues.resize((size_t)n);
      if (tcu_row_residues(ctx->handle, alig->saveResidues, rowResidues.data()) != TCU_OK) {
        report_failure("CUDA platform: row scan failed");
        return false;
      }
    }
    if (cols) {
      colGaps.resize((size_t)L);
      if (tcu_gaps(ctx->handle, alig->saveSequences, colGaps.data(), nullptr, nullptr) != TCU_OK) {
        report_failure("CUDA platform: column scan failed");
        return false;
      }
    }
  }
  if (seqs) {  // Cleaner.cpp:1338-1370
    int counter = 0;
    for (int i = 0; i < n; i++) {
      if (alig->saveSequences[i] == -1) continue;
      if (rowResidues[i] == 0) {
        if (keepSequences) {
          debug.report(WarningCode::KeepingOnlyGapsSequence, new std::string[1]{alig->seqsName[i]});
          counter++;
        } else {
          debug.report(WarningCode::RemovingOnlyGapsSequence, new std::string[1]{alig->seqsName[i]});
          alig->saveSequences[i] = -1;
        }
      } else
        counter++;
    }
    alig->numberOfSequences = counter;
  }
  if (cols) {  // Cleaner.cpp:1372-1395; the counts were taken over the rows kept at entry
    int counter = 0;
    for (int j = 0; j < L; j++) {
      if (alig->saveResidues[j] == -1) continue;
      if (colGaps[j] == keptRows)
        alig->saveResidues[j] = -1;
      else
        counter++;
    }
    alig->numberOfResidues = counter;
  }
  return true;
}

bool cudaDuplicatePartners(Alignment *alig, std::vector<int> &partner)
{
  StartTiming("bool cudaDuplicatePartners(Alignment *, std::vector<int> &) ");
  const int n = alig->originalNumberOfSequences, L = alig->originalNumberOfResidues;
  if (alig->sequences == nullptr || n <= 1 || L <= 0) return false;
  for (int i = 0; i < n; i++)
    if (alig->sequences[i].size() != (size_t)L) return false;
  std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig);
  if (!ctx) return false;
  std::vector<unsigned long long> h((size_t)2 * n);
  {
    std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
    if (tcu_row_hashes(ctx->handle, h.data()) != TCU_OK) {
      report_failure("CUDA platform: row hashes failed");
      return false;
    }
  }
  // rows ordered by (hash, index): equal rows are neighbours inside a run of equal hashes
  std::vector<int> idx((size_t)n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) {
    if (h[2 * (size_t)a] != h[2 * (size_t)b]) return h[2 * (size_t)a] < h[2 * (size_t)b];
    if (h[2 * (size_t)a + 1] != h[2 * (size_t)b + 1]) return h[2 * (size_t)a + 1] < h[2 * (size_t)b + 1];
    return a < b;
  });
  // Cleaner.cpp:1493-1508: row i goes iff some later row x equals it, and the report names
  // the FIRST such x.  Inside a run (indices ascending) that is the first later member that
  // compares equal byte for byte -- the hashes only chose who is compared.  The caller's loop
  // (the reference's own) marks and reports.
  partner.assign((size_t)n, -1);
  for (size_t a = 0; a < (size_t)n;) {
    size_t b = a + 1;
    while (b < (size_t)n && h[2 * (size_t)idx[b]] == h[2 * (size_t)idx[a]] &&
           h[2 * (size_t)idx[b] + 1] == h[2 * (size_t)idx[a] + 1])
      b++;
    for (size_t u = a; u + 1 < b; u++)
      for (size_t v = u + 1; v < b; v++)
        if (alig->sequences[idx[u]] == alig->sequences[idx[v]]) {
          partner[idx[u]] = idx[v];
          break;
        }
    a = b;
  }
  return true;
}

bool cudaValidateSymbols(Alignment *alig, bool *valid)
{
  StartTiming("bool cudaValidateSymbols(Alignment *, bool *) ");
  static const bool enabled = [] {
    const char *e = getenv("TRIMAL_CUDA_INGEST");
    return e != nullptr && *e != 0 && *e != '0';
  }();
  if (!enabled) return false;
  const int n = alig->numberOfSequences;
  if (alig->sequences == nullptr || n <= 0) return false;
  const size_t L = alig->sequences[0].size();
  if (L == 0 || (size_t)n * L < ((size_t)1 << 20)) return false;  // the scan is cheaper than an upload
  for (int i = 0; i < n; i++)
    if (alig->sequences[i].size() != L) return false;  // ragged: the reference reports it
  if (tcu_device_count() < 1) return false;
  const int savedResidues = alig->numberOfResidues;
  alig->numberOfResidues = (int)L;  // what acquire() uploads (not set yet for file inputs)
  std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig);
  alig->numberOfResidues = savedResidues;
  if (!ctx) return false;
  unsigned long long hist[256];
  {
    std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
    if (tcu_byte_histogram(ctx->handle, hist) != TCU_OK) {
      report_failure("CUDA platform: byte histogram failed");
      return false;
    }
  }
  *valid = true;
  for (int b = 0; b < 256; b++) {
    const char c = (char)b;
    if (hist[b] != 0 && (!isalpha(c)) && (!ispunct(c))) *valid = false;  // Alignment.cpp:660
  }
  return true;
}

bool CUDAOverlap::calculateSpuriousVector(float overlap, float *spuriousVector)
{
  StartTiming("bool CUDAOverlap::calculateSpuriousVector(float, float *) ");
  if (spuriousVector == nullptr) return false;  // template.h:210-211
  const uint32_t ovrlap =
      uint32_t(ceil(overlap * float(alig->originalNumberOfSequences - 1)));  // template.h:217-218
  std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig);
  if (!ctx) return false;
  const char indet = indet_of(alig);  // before the lock: type detection may use the handle
  std::lock_guard<std::recursive_mutex> lk(ctx->mutex);
  if (tcu_spurious(ctx->handle, indet, ovrlap, spuriousVector) != TCU_OK) {
    report_failure("CUDA platform: overlap statistic failed");
    return false;
  }
  return true;
}

bool CUDASimilarity::calculateVectors(bool cutByGap)
{
  StartTiming("bool CUDASimilarity::calculateVectors(bool cutByGap) ");
  if (simMatrix == nullptr) return false;  // template.h:73-74

  // identities first, through the manager so the object is cached (template.h:79); the
  // kernel reads them on the device, so no host copy is asked for here
  {
    DeviceOnlyScope scope;
    alig->Statistics->calculateSeqIdentity();
  }

  int *gaps = nullptr;
  if (cutByGap) {  // template.h:87-91
    if (alig->Statistics->gaps == nullptr) alig->Statistics->calculateGapStats();
    gaps = alig->Statistics->gaps->getGapsWindow();
  }
  const float gapThreshold = 0.8F * alig->numberOfResidues;  // template.h:108 (sic: residues)

  // public accessors only (the raw members are private in vanilla trimAl).  pytrimal's
  // SimilarityMatrix.__init__ (src/pytrimal/_trimal.pyx:1973-1981) fills vhash for the
  // letters of its alphabet only and leaves the other entries of the fresh `new int[]`
  // uninitialised, so an index is trusted only inside [0, order): everything else is
  // "no row" (UndefinedSymbol if the letter occurs, template.h:140-144).
  const int order = matrix_order(simMatrix, 0);
  const int limit = order >= 0 ? order : 28;  // 28 = pytrimal's alphabet limit (_trimal.pyx:1969)
  int vhash[26], npos = 0;
  for (int c = 0; c < 26; c++) {
    const int v = simMatrix->getLetterIndex((char)('A' + c));
    vhash[c] = (v >= 0 && v < limit) ? v : -1;
    if (vhash[c] + 1 > npos) npos = vhash[c] + 1;
  }
  if (order >= 0) npos = order;
  if (npos < 1) return false;
  const float **distMat = simMatrix->getDistanceMatrix();
  std::vector<float> dist((size_t)npos * npos);
  for (int i = 0; i < npos; i++)
    for (int j = 0; j < npos; j++) dist[(size_t)i * npos + j] = distMat[i][j];

  const int L = alig->originalNumberOfResidues;
  std::vector<float> num((size_t)L), den((size_t)L);
  int err_col = -1, err_row = -1, err_byte = 0;
  std::shared_ptr<CUDAContext> ctx = CUDAContext::acquire(alig);
  if (!ctx) return false;
  const char indet = indet_of(alig);  // before the lock: type detection may use the handle
  std::unique_lock<std::recursive_mutex> lk(ctx->mutex);
  // a CUDAIdentity leaves its result on the device; identities computed by any
  // other platform (or whose device copy was replaced) are uploaded from the host
  auto *cid = dynamic_cast<CUDAIdentity *>(alig->Statistics->identity);
  const bool on_device = cid != nullptr && ctx->ident_owner == cid->share->id &&
                         tcu_identity_resident(ctx->handle);
  const float *identities = nullptr;
  if (!on_device) {
    lk.unlock();
    if (cid != nullptr) cid->materialize();
    identities = alig->Statistics->identity->identities;
    lk.lock();
  }
  int rc = tcu_similarity(ctx->handle, indet, dist.data(), npos, vhash, gaps,
                          gapThreshold, on_device ? nullptr : identities, num.data(), den.data(),
                          MDK, &err_col, &err_row, &err_byte);
  // an uploaded matrix replaces whatever was resident
  if (!on_device) ctx->ident_owner = cid != nullptr ? cid->share->id : 0;
  if (rc == TCU_ERR_INCORRECT_SYMBOL) {  // template.h:135-138
    debug.report(ErrorCode::IncorrectSymbol, new std::string[1]{std::string(1, (char)err_byte)});
    return false;
  }
  if (rc == TCU_ERR_UNDEFINED_SYMBOL) {  // template.h:140-144
    debug.report(ErrorCode::UndefinedSymbol, new std::string[1]{std::string(1, (char)err_byte)});
    return false;
  }
  if (rc != TCU_OK) {
    report_failure("CUDA platform: similarity statistic failed");
    return false;
  }
  return true;
}

}  // namespace statistics
