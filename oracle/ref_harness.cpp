/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A small C-ABI driver around the UNMODIFIED reference: vendored trimAl's
 * Alignment / statistics::Manager / Cleaner classes and its AVX2 / SSE2 /
 * generic statistic kernels, compiled from where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libtrimal_ref.so.
 * No reference source is copied into this repository; this file only calls
 * the reference's public (and, via -fno-access-control, a few private)
 * members the way pytrimal's Cython layer does:
 *   - Alignment construction      : src/pytrimal/_trimal.pyx:828-853
 *   - trimmer -> manager fields   : _trimal.pyx:1478-1497, 1651-1659, 1766-1769, 1859-1862
 *   - trim()                      : _trimal.pyx:1291-1365
 *   - manager dispatch            : vendor/trimal/source/trimalManager.cpp:1972-2255
 * (trimAlManager itself is not linked: it drags in the CMake-generated
 * format-handler registry; the three dispatch functions it contributes are
 * restated in ref_trim below.)
 *
 * Used by tests/ (when the .so is present), tests/golden/make_golden.py and
 * bench.py's cpu_baseline / --impl reference legs.  Never by pytrimal_b200/.
 */
#include <cstdint>
#include <cstring>
#include <string>

#include "Alignment/Alignment.h"
#include "Cleaner.h"
#include "FormatHandling/FormatManager.h"
#include "Statistics/Gaps.h"
#include "Statistics/Identity.h"
#include "Statistics/Manager.h"
#include "Statistics/Overlap.h"
#include "Statistics/Similarity.h"
#include "Statistics/similarityMatrix.h"
#include "defines.h"
#include "reportsystem.h"

/* Consistency.cpp references two FormatManager members that live in the
 * format-handler registry we do not build.  They are never reached from the
 * statistics path; abort loudly if that ever changes. */
namespace FormatHandling {
Alignment *FormatManager::loadAlignment(const std::string &) { abort(); }
std::string FormatManager::getFileFormatName(const std::string &) { abort(); }
}  // namespace FormatHandling

namespace {

struct RefAlignment {
    Alignment *ali = nullptr;
    statistics::similarityMatrix *smx = nullptr;
};

statistics::ComputePlatform to_platform(int p)
{
    switch (p) {
    case 1: return statistics::ComputePlatform::SSE2;
    case 2: return statistics::ComputePlatform::AVX2;
#ifdef HAVE_CUDA  /* only in the integration build (integration/Makefile) */
    case 3: return statistics::ComputePlatform::CUDA;
#endif
    default: return statistics::ComputePlatform::NONE;
    }
}

/* trimAlManager::create_or_use_similarity_matrix, default-matrix branch
 * (trimalManager.cpp:1989-2002) + pytrimal's NotDefined fallback
 * (_trimal.pyx:1346-1352). */
void attach_default_matrix(RefAlignment *h)
{
    if (h->smx) return;
    h->smx = new statistics::similarityMatrix();
    int t = h->ali->getAlignmentType();
    if (t == SequenceTypes::AA || t == (SequenceTypes::AA | SequenceTypes::DEG) ||
        t == SequenceTypes::NotDefined)
        h->smx->defaultAASimMatrix();
    else if (t == SequenceTypes::DNA || t == SequenceTypes::RNA)
        h->smx->defaultNTSimMatrix();
    else
        h->smx->defaultNTDegeneratedSimMatrix();
    h->ali->Statistics->setSimilarityMatrix(h->smx);
}

}  // namespace

extern "C" {

/* rows: nseq pointers to ncol bytes each.  datatype 0 = auto-detect. */
void *ref_alignment_new(const char *const *rows, int nseq, int ncol, int datatype)
{
    debug.Level = VerboseLevel::NONE;
    RefAlignment *h = new RefAlignment();
    Alignment *a = new Alignment();
    a->dataType = datatype;
    a->numberOfSequences = nseq;
    a->seqsName = new std::string[nseq];
    a->sequences = new std::string[nseq];
    for (int i = 0; i < nseq; i++) {
        a->seqsName[i] = "s" + std::to_string(i);
        a->sequences[i].assign(rows[i], (size_t)ncol);
    }
    a->numberOfResidues = ncol;
    if (ncol > 0 && !a->fillMatrices(nseq > 1, true)) {
        delete a;
        delete h;
        return nullptr;
    }
    a->originalNumberOfSequences = a->numberOfSequences;
    a->originalNumberOfResidues = a->numberOfResidues;
    h->ali = a;
    return h;
}

void ref_alignment_free(void *hv)
{
    RefAlignment *h = (RefAlignment *)hv;
    if (!h) return;
    delete h->ali;
    /* the matrix is owned by the harness, like pytrimal's stack-local smx */
    delete h->smx;
    delete h;
}

int ref_alignment_type(void *hv) { return ((RefAlignment *)hv)->ali->getAlignmentType(); }

/* Forget the cached type and detect it again (Alignment::getAlignmentType ->
 * utils::checkAlignmentType, utils.cpp:476-545) with the platform set by
 * ref_set_platform: the CUDA platform answers from a device byte histogram. */
int ref_detect_type(void *hv)
{
    Alignment *a = ((RefAlignment *)hv)->ali;
    a->dataType = SequenceTypes::NotDefined;
    return a->getAlignmentType();
}

/* 0 = generic, 1 = SSE2, 2 = AVX2, 3 = CUDA (integration build only).
 * Must be called before any statistic. */
void ref_set_platform(void *hv, int platform)
{
    ((RefAlignment *)hv)->ali->Statistics->platform = to_platform(platform);
}

int ref_get_platform(void *hv) { return (int)((RefAlignment *)hv)->ali->Statistics->platform; }

/* Overwrite the keep-masks (entries -1 = removed), adjusting the live counts
 * the way Cleaner does. */
void ref_set_masks(void *hv, const int *save_seq, const int *save_res)
{
    Alignment *a = ((RefAlignment *)hv)->ali;
    if (save_seq) {
        int kept = 0;
        for (int i = 0; i < a->originalNumberOfSequences; i++) {
            a->saveSequences[i] = save_seq[i];
            kept += save_seq[i] != -1;
        }
        a->numberOfSequences = kept;
    }
    if (save_res) {
        int kept = 0;
        for (int i = 0; i < a->originalNumberOfResidues; i++) {
            a->saveResidues[i] = save_res[i];
            kept += save_res[i] != -1;
        }
        a->numberOfResidues = kept;
    }
}

void ref_set_windows(void *hv, int gap_window, int sim_window)
{
    ((RefAlignment *)hv)->ali->setWindowsSize(gap_window, sim_window);
}

/* Gaps: raw counts, windowed counts (= raw when no window), histogram
 * (nseq+1 ints) and maximum.  Returns 0 on success. */
int ref_gaps(void *hv, int *gaps_in_column, int *gaps_window, int *num_cols_with_gaps,
             int *max_gaps)
{
    Alignment *a = ((RefAlignment *)hv)->ali;
    if (!a->Statistics->calculateGapStats()) return -1;
    statistics::Gaps *g = a->Statistics->gaps;
    const int L = a->originalNumberOfResidues;
    if (gaps_in_column) memcpy(gaps_in_column, g->gapsInColumn, sizeof(int) * L);
    if (gaps_window) memcpy(gaps_window, g->getGapsWindow(), sizeof(int) * L);
    if (num_cols_with_gaps)
        memcpy(num_cols_with_gaps, g->numColumnsWithGaps,
               sizeof(int) * (a->originalNumberOfSequences + 1));
    if (max_gaps) *max_gaps = g->maxGaps;
    return 0;
}

/* Identity: copies the first `count` floats of the packed array. */
int ref_identity(void *hv, float *out, size_t count)
{
    Alignment *a = ((RefAlignment *)hv)->ali;
    if (!a->Statistics->calculateSeqIdentity()) return -1;
    if (out) memcpy(out, a->Statistics->identity->identities, sizeof(float) * count);
    return 0;
}

/* Identity without the copy (timing). */
int ref_identity_nocopy(void *hv)
{
    Alignment *a = ((RefAlignment *)hv)->ali;
    return a->Statistics->calculateSeqIdentity() ? 0 : -1;
}

/* Similarity through Manager::calculateConservationStats (Manager.cpp:61-114):
 * gaps (+window), identity, MDK with the gap cut, similarity window.
 * mdk receives the un-windowed vector, mdk_window the windowed one (or a copy). */
int ref_similarity(void *hv, float *mdk, float *mdk_window)
{
    RefAlignment *h = (RefAlignment *)hv;
    attach_default_matrix(h);
    Alignment *a = h->ali;
    if (!a->Statistics->calculateConservationStats()) return -1;
    statistics::Similarity *s = a->Statistics->similarity;
    const int L = a->originalNumberOfResidues;
    if (mdk) memcpy(mdk, s->MDK, sizeof(float) * L);
    if (mdk_window) memcpy(mdk_window, s->getMdkWindowedVector(), sizeof(float) * L);
    return 0;
}

/* Distance matrix + hash of the default matrix for this alignment's type.
 * dist: 28*28 floats max, vhash: 28 ints.  Returns matrix order. */
int ref_default_matrix(void *hv, float *dist, int *vhash)
{
    RefAlignment *h = (RefAlignment *)hv;
    attach_default_matrix(h);
    const int n = h->smx->numPositions;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) dist[i * n + j] = h->smx->distMat[i][j];
    for (int i = 0; i < 28; i++) vhash[i] = h->smx->vhash[i]; /* TAMABC, similarityMatrix.cpp:34 */
    return n;
}

int ref_spurious(void *hv, float overlap, float *out)
{
    Alignment *a = ((RefAlignment *)hv)->ali;
    return a->Statistics->calculateSpuriousVector(overlap, out) ? 0 : -1;
}

/* The three Cleaner walks over the identity matrix, called on
 * the harness's own alignment with whatever platform ref_set_platform chose:
 *   ref_representatives : Cleaner::calculateRepresentativeSeq (Cleaner.cpp:1398-1466);
 *                         out[0..count) = representatives in creation order
 *   ref_cutpoint        : Cleaner::getCutPointClusters (Cleaner.cpp:1026-1156)
 *   ref_select_method   : Cleaner::selectMethod (Cleaner.cpp:46-99); 1 = GAPPYOUT, 2 = STRICT */
int ref_representatives(void *hv, float max_identity, int *out)
{
    Alignment *a = ((RefAlignment *)hv)->ali;
    int *r = a->Cleaning->calculateRepresentativeSeq(max_identity);
    if (!r) return -1;
    const int count = r[0];
    if (out) memcpy(out, r + 1, sizeof(int) * count);
    delete[] r;
    return count;
}

float ref_cutpoint(void *hv, int clusters)
{
    return ((RefAlignment *)hv)->ali->Cleaning->getCutPointClusters(clusters);
}

int ref_select_method(void *hv)
{
    return ((RefAlignment *)hv)->ali->Cleaning->selectMethod() == GAPPYOUT ? 1 : 2;
}

/* 1 when the identity statistic object holds a host copy of the matrix (the CUDA
 * platform leaves it on the device until something on the host asks for it). */
int ref_identity_on_host(void *hv)
{
    Alignment *a = ((RefAlignment *)hv)->ali;
    return a->Statistics->identity && a->Statistics->identity->identities ? 1 : 0;
}

/*
 * One trim() call.  `method`:
 *   automatic : "strict" "strictplus" "gappyout" "nogaps" "noallgaps"
 *               "automated1" "automated2" "noduplicateseqs"
 *   "manual"        : p = {gapThreshold(already 1-x), similarityThreshold,
 *                          conservationThreshold, window, gapWindow, simWindow}; -1 = unset
 *   "overlap"       : p = {residuesOverlap, sequenceOverlap (percent)}
 *   "representative": p = {clusters, maxIdentity}; -1 = unset
 * The working copy is made first and the platform is then set on the COPY's
 * manager, so the requested platform is the one that runs (the upstream
 * trim() sets it on the source object).
 * Outputs: keep_seq[nseq], keep_res[ncol] = final saveSequences/saveResidues.
 * Returns 0, or -1 when the reference reported an error (nullptr result).
 */
int ref_trim(void *hv, int platform, const char *method, const double *p, int *keep_seq,
             int *keep_res)
{
    RefAlignment *h = (RefAlignment *)hv;
    const std::string m(method);

    bool strict = m == "strict", strictplus = m == "strictplus", gappyout = m == "gappyout";
    bool nogaps = m == "nogaps", noallgaps = m == "noallgaps", automated1 = m == "automated1";
    bool automated2 = m == "automated2", nodup = m == "noduplicateseqs";
    bool automatic = strict || strictplus || gappyout || nogaps || noallgaps || automated1 ||
                     automated2 || nodup;
    float gapThreshold = -1, similarityThreshold = -1, conservationThreshold = -1;
    float residuesOverlap = -1, sequenceOverlap = -1, maxIdentity = -1;
    int windowSize = -1, gapWindow = -1, similarityWindow = -1, clusters = -1;
    if (m == "manual") {
        gapThreshold = (float)p[0];
        similarityThreshold = (float)p[1];
        conservationThreshold = (float)p[2];
        windowSize = (int)p[3];
        gapWindow = (int)p[4];
        similarityWindow = (int)p[5];
    } else if (m == "overlap") {
        residuesOverlap = (float)p[0];
        sequenceOverlap = (float)p[1];
    } else if (m == "representative") {
        clusters = (int)p[0];
        maxIdentity = (float)p[1];
    } else if (!automatic) {
        return -2;
    }

    /* set_window_size (trimalManager.cpp:2241-2255) */
    if (windowSize != -1) gapWindow = similarityWindow = windowSize;
    else {
        if (gapWindow == -1) gapWindow = 0;
        if (similarityWindow == -1) similarityWindow = 0;
    }
    /* similarity matrix (trimalManager.cpp:1972-2011 / _trimal.pyx:1341-1352):
     * attached to the source so the working copy inherits it through the
     * Manager mold constructor (Manager.cpp:446-457). */
    bool want_matrix = strict || strictplus || automated1 || similarityThreshold != -1.0f;
    if (h->ali->getAlignmentType() == SequenceTypes::NotDefined) want_matrix = true;
    if (want_matrix) attach_default_matrix(h);

    Alignment *orig = new Alignment(*h->ali);
    orig->Statistics->platform = to_platform(platform);
    orig->setWindowsSize(gapWindow, similarityWindow);

    Alignment *single = nullptr, *temp = nullptr;
    bool failed = false;

    /* CleanSequences (trimalManager.cpp:2064-2119) */
    bool seq_step = false;
    if (clusters != -1) {
        temp = orig->Cleaning->getClustering(orig->Cleaning->getCutPointClusters(clusters));
        seq_step = true;
    } else if (maxIdentity != -1) {
        temp = orig->Cleaning->getClustering(maxIdentity);
        seq_step = true;
    } else if (residuesOverlap != -1 && sequenceOverlap != -1) {
        temp = orig->Cleaning->cleanSpuriousSeq(residuesOverlap, sequenceOverlap / 100.0F, false);
        seq_step = true;
    } else if (nodup) {
        orig->Cleaning->removeDuplicates();
    }
    if (temp) {
        single = temp->Cleaning->cleanNoAllGaps(false);
        delete temp;
        temp = nullptr;
        if (single) {
            delete single->Statistics->gaps;
            single->Statistics->gaps = nullptr;
            delete single->Statistics->similarity;
            single->Statistics->similarity = nullptr;
        } else
            failed = true;
    } else {
        if (seq_step) failed = true;
        single = orig;
    }

    if (!failed) {
        if (automatic) {
            /* CleanResiduesAuto (trimalManager.cpp:2121-2158) */
            if (automated1) {
                if (single->Cleaning->selectMethod() == GAPPYOUT) gappyout = true;
                else strict = true;
            }
            if (nogaps) temp = single->Cleaning->cleanGaps(0, 0, false);
            else if (noallgaps) temp = single->Cleaning->cleanNoAllGaps(false);
            else if (gappyout) temp = single->Cleaning->clean2ndSlope(false);
            else if (strict) temp = single->Cleaning->cleanCombMethods(false, false);
            else if (strictplus) temp = single->Cleaning->cleanCombMethods(false, true);
            else if (automated2) temp = single->Cleaning->cleanAutomated2(false);
            if (!nodup && temp == nullptr) failed = true;
        } else {
            /* CleanResiduesNonAuto (trimalManager.cpp:2160-2239), without the
             * -selectcols and consistency branches pytrimal never sets. */
            bool col_step = true;
            if (similarityThreshold != -1.0F) {
                if (gapThreshold != -1.0F)
                    temp = single->Cleaning->clean(conservationThreshold, gapThreshold,
                                                   similarityThreshold, false);
                else
                    temp = single->Cleaning->cleanConservation(conservationThreshold,
                                                               similarityThreshold, false);
            } else if (gapThreshold != -1.0F) {
                temp = single->Cleaning->cleanGaps(conservationThreshold, gapThreshold, false);
            } else
                col_step = false;
            if (col_step && temp == nullptr) failed = true;
        }
        if (temp) {
            if (single != orig) delete single;
            single = temp;
            temp = nullptr;
        }
    }

    if (!failed && single) {
        memcpy(keep_seq, single->saveSequences, sizeof(int) * single->originalNumberOfSequences);
        memcpy(keep_res, single->saveResidues, sizeof(int) * single->originalNumberOfResidues);
    }
    if (single && single != orig) delete single;
    delete orig;
    return failed ? -1 : 0;
}

}  /* extern "C" */
