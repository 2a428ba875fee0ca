// Platform/CUDA/CUDA.h -- host shim that plugs libtrimal_cuda into trimAl's
// compute-platform switch, next to Platform/x86/AVX2.h (same four subclasses,
// same two constructors and one override each; vendor/trimal/include/Platform/
// x86/AVX2.h:43-70).  Compiled as part of trimAl (it needs trimAl's headers),
// not as part of libtrimal_cuda.so; the only thing it calls is the C ABI in
// include/trimal_cuda.h.  See INTEGRATION.md for the Manager patches.
#ifndef TRIMAL_PLATFORM_CUDA_H
#define TRIMAL_PLATFORM_CUDA_H

#include <memory>
#include <mutex>

#include "Statistics/Gaps.h"
#include "Statistics/Identity.h"
#include "Statistics/Overlap.h"
#include "Statistics/Similarity.h"

struct tcu_msa;

namespace statistics {

// One uploaded alignment, shared by the four statistics of an Alignment and by
// the copies the mold constructors make (rows are shared read-only between
// copies through Alignment::SeqRef, so the device copy can be too).
class CUDAContext {
public:
  // Returns the handle for `alig`'s rows, uploading on first use; nullptr
  // (after debug.report) when the upload fails.
  static std::shared_ptr<CUDAContext> acquire(Alignment *alig);
  ~CUDAContext();
  tcu_msa *handle = nullptr;
  const void *rows_key = nullptr;  // Alignment::sequences pointer this upload belongs to
  std::mutex mutex;                // a tcu_msa handle serves one thread at a time
};

class CUDASimilarity : public Similarity {
public:
  CUDASimilarity(Alignment *parentAlignment) : Similarity(parentAlignment) {}
  CUDASimilarity(Alignment *parentAlignment, Similarity *parentSimilarity)
      : Similarity(parentAlignment, parentSimilarity) {}
  bool calculateVectors(bool cutByGap) override;
  std::shared_ptr<CUDAContext> ctx;  // keeps the upload alive between calls
};

class CUDAGaps : public Gaps {
public:
  CUDAGaps(Alignment *parentAlignment) : Gaps(parentAlignment) {}
  CUDAGaps(Alignment *parentAlignment, Gaps *parentGaps) : Gaps(parentAlignment, parentGaps) {}
  void CalculateVectors() override;
  std::shared_ptr<CUDAContext> ctx;
};

class CUDAOverlap : public Overlap {
public:
  CUDAOverlap(Alignment *parent) : Overlap(parent) {}
  CUDAOverlap(Alignment *parent, Overlap *parentOverlap) : Overlap(parent, parentOverlap) {}
  bool calculateSpuriousVector(float overlap, float *spuriousVector) override;
  std::shared_ptr<CUDAContext> ctx;
};

class CUDAIdentity : public Identity {
public:
  CUDAIdentity(Alignment *parent) : Identity(parent) {}
  CUDAIdentity(Alignment *parent, Identity *parentIdentity) : Identity(parent, parentIdentity) {}
  void calculateSeqIdentity() override;
  std::shared_ptr<CUDAContext> ctx;
};

// Number of usable devices; 0 makes the Cython layer refuse platform="cuda".
int cudaPlatformDeviceCount();

}  // namespace statistics

#endif
