// Platform/CUDA/CUDA.h -- host shim that plugs libtrimal_cuda into trimAl's
// compute-platform switch, next to Platform/x86/AVX2.h (same four subclasses,
// same two constructors and one override each; vendor/trimal/include/Platform/
// x86/AVX2.h:43-70).  Compiled as part of trimAl (it needs trimAl's headers),
// not as part of libtrimal_cuda.so; the only thing it calls is the C ABI in
// include/trimal_cuda.h.  See INTEGRATION.md for the Manager patches.
#ifndef TRIMAL_PLATFORM_CUDA_H
#define TRIMAL_PLATFORM_CUDA_H

#include <cstdint>
#include <memory>
#include <mutex>
#include <vector>

#include "Statistics/Gaps.h"
#include "Statistics/Identity.h"
#include "Statistics/Overlap.h"
#include "Statistics/Similarity.h"

struct tcu_msa;

namespace statistics {

// One uploaded alignment.  It is owned by the statistics::Manager of every Alignment that
// used it (Manager::cudaContext, patches/Manager.h.patch) and found by the others that share
// the same rows (Alignment copies share `sequences` read-only through SeqRef) through a
// table of weak references -- so the device memory lives exactly as long as an alignment
// that computed on it, and the copies a trim() returns (fresh mold copies that never
// compute) pin nothing.
class CUDAContext {
public:
  // Returns the handle for `alig`'s rows, uploading on first use; nullptr
  // (after debug.report) when the upload fails.
  static std::shared_ptr<CUDAContext> acquire(Alignment *alig);
  ~CUDAContext();
  tcu_msa *handle = nullptr;
  const void *rows_key = nullptr;  // Alignment::sequences pointer this upload belongs to
  // a tcu_msa handle serves one thread at a time; recursive because a statistic may ask for
  // the alignment type (another user of the handle) while it holds the lock
  std::recursive_mutex mutex;
  // Which group of CUDAIdentity objects (IdentityShare::id) the identity matrix resident on
  // the device belongs to; 0 = none.  Alignment copies share the upload but may carry
  // different identity objects (different column masks).  Ids are never reused, so a group
  // that died cannot be mistaken for a new one at the same address.
  uint64_t ident_owner = 0;
};

// State shared by a CUDAIdentity and the copies the mold constructor makes of it (the
// base class shares `identities` and `refCounter` the same way, Identity.cpp:50-58).
struct IdentityShare {
  IdentityShare();
  const uint64_t id;  // process-wide unique, never 0
  std::mutex mutex;
  float *host = nullptr;  // the packed matrix on the host once something asked for it
};

class CUDASimilarity : public Similarity {
public:
  CUDASimilarity(Alignment *parentAlignment) : Similarity(parentAlignment) {}
  CUDASimilarity(Alignment *parentAlignment, Similarity *parentSimilarity)
      : Similarity(parentAlignment, parentSimilarity) {}
  bool calculateVectors(bool cutByGap) override;
};

class CUDAGaps : public Gaps {
public:
  CUDAGaps(Alignment *parentAlignment) : Gaps(parentAlignment) {}
  CUDAGaps(Alignment *parentAlignment, Gaps *parentGaps) : Gaps(parentAlignment, parentGaps) {}
  void CalculateVectors() override;
};

class CUDAOverlap : public Overlap {
public:
  CUDAOverlap(Alignment *parent) : Overlap(parent) {}
  CUDAOverlap(Alignment *parent, Overlap *parentOverlap) : Overlap(parent, parentOverlap) {}
  bool calculateSpuriousVector(float overlap, float *spuriousVector) override;
};

// The identity matrix may live on the device only: Cleaner's three walks over it
// (selectMethod, getCutPointClusters, calculateRepresentativeSeq -- the cuda* functions
// below, called from patches/Cleaner.cpp.patch) run there, so its 4*P bytes cross PCIe
// only when host code really reads Identity::identities; every such reader goes through
// Manager::calculateSeqIdentity, which the Manager patch makes call
// cudaMaterializeIdentity.
class CUDAIdentity : public Identity {
public:
  CUDAIdentity(Alignment *parent);
  CUDAIdentity(Alignment *parent, Identity *parentIdentity);
  ~CUDAIdentity() override;
  void calculateSeqIdentity() override;
  // Matrix of this object's group resident on the device (computing it there when
  // neither a device nor a host copy exists).  false: not applicable (masked rows, a
  // host-only copy) or failed (after debug.report).
  bool computeOnDevice();
  // Make `identities` a valid host array (download, or compute as a last resort).
  void materialize();
  std::shared_ptr<IdentityShare> share;

private:
  bool allocateHost();
  void computeToHost();
};

// Cleaner's walks over the identity matrix on the device.  Each
// returns false / nullptr when it does not apply (masked rows, no device copy); the
// caller then runs the reference code.
bool cudaSelectMethod(Alignment *alig, int *method);                         // Cleaner.cpp:46-99
bool cudaCutPointClusters(Alignment *alig, int clusterNumber, float *cut);   // Cleaner.cpp:1026-1156
int *cudaRepresentativeSeq(Alignment *alig, float maximumIdent);             // Cleaner.cpp:1398-1466
// Called by the patched Manager::calculateSeqIdentity when the object already exists.
void cudaMaterializeIdentity(Identity *identity);

// utils::checkAlignmentType (utils.cpp:476-545) from a byte histogram computed on the
// device; called from patches/Alignment.cpp.patch.  false = not
// applicable (trimmed alignment, ragged rows) or failed: the caller runs the reference scan.
bool cudaAlignmentType(const Alignment *alig, int *type);

// Post-trim scans of the host layer on the device; called from
// patches/Cleaner.cpp.patch, each returns false when it does not apply or failed (the caller
// then runs the reference loop).
//   cudaRemoveAllGaps   Cleaner::removeAllGapsSeqsAndCols (Cleaner.cpp:1331-1396): same mask
//                       updates, same warnings, same counters
//   cudaDuplicatePartners  the search of Cleaner::removeDuplicates (Cleaner.cpp:1489-1509):
//                       partner[i] = the first later row equal to row i, or -1; row hashes on
//                       the device, byte compares only inside groups of equal hashes.  What is
//                       done with a duplicate stays in the reference's loop (pytrimal's build
//                       also decrements numberOfSequences there, vanilla trimAl does not).
bool cudaRemoveAllGaps(Alignment *alig, bool seqs, bool cols, bool keepSequences);
bool cudaDuplicatePartners(Alignment *alig, std::vector<int> &partner);

// Alignment::fillMatrices' symbol validation (Alignment.cpp:657-664) from the device byte
// histogram: true = decided, *valid says whether every byte is isalpha or
// ispunct; false = not applicable (small or ragged alignment, no device, TRIMAL_CUDA_INGEST
// not set).  The upload it makes is the one the statistics of a later trim() use.
bool cudaValidateSymbols(Alignment *alig, bool *valid);

// Number of usable devices; 0 makes the Cython layer refuse platform="cuda".
int cudaPlatformDeviceCount();

}  // namespace statistics

#endif
