// Platform/CUDA/CUDA.cpp -- see CUDA.h.  Mirrors source/Platform/x86/AVX2.cpp:
// 126-148 (four thin overrides), with the SIMD template calls replaced by the
// C ABI of libtrimal_cuda.  Error convention: never throw; report
// through debug.report(...) and return false / leave zero-filled outputs.
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "Alignment/Alignment.h"
#include "InternalBenchmarker.h"
#include "Platform/CUDA/CUDA.h"
#include "Statistics/Manager.h"
#include "Statistics/similarityMatrix.h"
#include "defines.h"
#include "reportsystem.h"
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
  std::lock_guard<std::mutex> lk(g_ctx_mutex);
  auto it = g_contexts.find(rows_key);
  if (it != g_contexts.end() && it->second.expired()) g_contexts.erase(it);
  tcu_msa_destroy(handle);
}

std::shared_ptr<CUDAContext> CUDAContext::acquire(Alignment *alig)
{
  const void *key = alig->sequences;
  {
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    auto it = g_contexts.find(key);
    if (it != g_contexts.end())
      if (auto sp = it->second.lock()) return sp;
  }
  const int n = alig->originalNumberOfSequences, L = alig->originalNumberOfResidues;
  std::vector<const char *> rows((size_t)n);
  for (int i = 0; i < n; i++) rows[i] = alig->sequences[i].data();
  tcu_msa *h = nullptr;
  if (tcu_msa_create(rows.data(), n, L, /*device=*/0, &h) != TCU_OK) {
    report_failure("CUDA platform: alignment upload failed");
    return nullptr;
  }
  auto sp = std::make_shared<CUDAContext>();
  sp->handle = h;
  sp->rows_key = key;
  std::lock_guard<std::mutex> lk(g_ctx_mutex);
  g_contexts[key] = sp;
  return sp;
}

// ---------------------------------------------------------------------------
void CUDAGaps::CalculateVectors()
{
  StartTiming("void CUDAGaps::CalculateVectors() ");
  const int L = alig->originalNumberOfResidues;
  // valid output even on failure: Cleaner keeps running after a void override
  memset(gapsInColumn, 0, sizeof(int) * L);
  if (!ctx) ctx = CUDAContext::acquire(alig);
  if (!ctx) return;
  std::lock_guard<std::mutex> lk(ctx->mutex);
  if (tcu_gaps(ctx->handle, alig->saveSequences, gapsInColumn, numColumnsWithGaps, &maxGaps) !=
      TCU_OK)
    report_failure("CUDA platform: gap statistic failed");
}

void CUDAIdentity::calculateSeqIdentity()
{
  StartTiming("void CUDAIdentity::calculateSeqIdentity() ");
  const int n = alig->originalNumberOfSequences;
  // same allocation as template.h:327-328 (size computed in fp32); owned and
  // delete[]d by the base class (Identity.cpp:110-118)
  const size_t size = ((float)n * n + n) / 2;
  identities = new (std::nothrow) float[size ? size : 1];
  if (identities == nullptr) {
    report_failure("CUDA platform: identity matrix allocation failed");
    return;
  }
  if (!ctx) ctx = CUDAContext::acquire(alig);
  int kept = 0;
  for (int i = 0; i < n; i++) kept += alig->saveSequences[i] != -1;
  // keep the device copy when it can feed the similarity statistic
  const int keep_on_device = kept == n;
  std::unique_lock<std::mutex> lk;
  if (ctx) lk = std::unique_lock<std::mutex>(ctx->mutex);
  if (!ctx || tcu_identity(ctx->handle, alig->saveSequences, alig->saveResidues, indet_of(alig),
                           identities, nullptr, nullptr, keep_on_device) != TCU_OK) {
    if (ctx) report_failure("CUDA platform: identity statistic failed");
    memset(identities, 0, sizeof(float) * size);
  }
}

bool CUDAOverlap::calculateSpuriousVector(float overlap, float *spuriousVector)
{
  StartTiming("bool CUDAOverlap::calculateSpuriousVector(float, float *) ");
  if (spuriousVector == nullptr) return false;  // template.h:210-211
  const uint32_t ovrlap =
      uint32_t(ceil(overlap * float(alig->originalNumberOfSequences - 1)));  // template.h:217-218
  if (!ctx) ctx = CUDAContext::acquire(alig);
  if (!ctx) return false;
  std::lock_guard<std::mutex> lk(ctx->mutex);
  if (tcu_spurious(ctx->handle, indet_of(alig), ovrlap, spuriousVector) != TCU_OK) {
    report_failure("CUDA platform: overlap statistic failed");
    return false;
  }
  return true;
}

bool CUDASimilarity::calculateVectors(bool cutByGap)
{
  StartTiming("bool CUDASimilarity::calculateVectors(bool cutByGap) ");
  if (simMatrix == nullptr) return false;  // template.h:73-74

  // identities first, through the manager so the object is cached (template.h:79)
  alig->Statistics->calculateSeqIdentity();
  const float *identities = alig->Statistics->identity->identities;

  int *gaps = nullptr;
  if (cutByGap) {  // template.h:87-91
    if (alig->Statistics->gaps == nullptr) alig->Statistics->calculateGapStats();
    gaps = alig->Statistics->gaps->getGapsWindow();
  }
  const float gapThreshold = 0.8F * alig->numberOfResidues;  // template.h:108 (sic: residues)

  // public accessors only (the raw members are private in vanilla trimAl)
  int vhash[26], npos = 0;
  for (int c = 0; c < 26; c++) {
    vhash[c] = simMatrix->getLetterIndex((char)('A' + c));
    if (vhash[c] + 1 > npos) npos = vhash[c] + 1;
  }
  const float **distMat = simMatrix->getDistanceMatrix();
  std::vector<float> dist((size_t)npos * npos);
  for (int i = 0; i < npos; i++)
    for (int j = 0; j < npos; j++) dist[(size_t)i * npos + j] = distMat[i][j];

  const int L = alig->originalNumberOfResidues;
  std::vector<float> num((size_t)L), den((size_t)L);
  int err_col = -1, err_row = -1, err_byte = 0;
  if (!ctx) ctx = CUDAContext::acquire(alig);
  if (!ctx) return false;
  std::lock_guard<std::mutex> lk(ctx->mutex);
  // a CUDAIdentity leaves its result on the device; identities computed by any
  // other platform are uploaded by passing the host pointer
  const bool on_device = dynamic_cast<CUDAIdentity *>(alig->Statistics->identity) != nullptr;
  int rc = tcu_similarity(ctx->handle, indet_of(alig), dist.data(), npos, vhash, gaps,
                          gapThreshold, on_device ? nullptr : identities, num.data(), den.data(),
                          MDK, &err_col, &err_row, &err_byte);
  if (rc == TCU_ERR_STATE)  // device copy not there (masked rows): fall back to an upload
    rc = tcu_similarity(ctx->handle, indet_of(alig), dist.data(), npos, vhash, gaps,
                        gapThreshold, identities, num.data(), den.data(), MDK, &err_col, &err_row,
                        &err_byte);
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
