"""Drop-in test: the REFERENCE's own trimAl (Alignment, Cleaner, Manager -- all
trimming logic) built with the CUDA compute platform patched in
(integration/Makefile), driven exactly like the AVX2 oracle.  Trimmed
alignments must be byte-identical to the AVX2 platform's: same kept rows, same
kept columns."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, random_msa

import oracle

pytestmark = pytest.mark.gpu

DROPIN = os.path.join(ROOT, "integration", "_build", "libtrimal_cuda_platform.so")


class DropIn(oracle.Ref):
    PATH = DROPIN
    _lib = None


@pytest.fixture(scope="module")
def dropin(gpu):
    if not os.path.exists(DROPIN):
        pytest.skip("integration/_build not present (built where /root/reference exists)")
    return DropIn


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
IDS = [os.path.basename(f)[:-4] for f in FIXTURES]
AUTO = ["gappyout", "strict", "strictplus", "automated1", "automated2", "nogaps", "noallgaps"]
PYTRIMAL = {
    "cons60.gt90": ("manual", [1 - 0.9, -1, 60, -1, -1, -1]),
    "cons40.gt40": ("manual", [1 - 0.4, -1, 40, -1, -1, -1]),
    "seq80.res80": ("overlap", [0.8, 80]),
    "seq40.res60": ("overlap", [0.6, 40]),
    "clusters5": ("representative", [5, -1]),
    "clusters10": ("representative", [10, -1]),
    "maxidentity75": ("representative", [-1, 0.75]),
    "noduplicateseqs": ("noduplicateseqs", []),
}


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_cuda_platform_reproduces_reference_trims(dropin, path):
    g = np.load(path)
    m = g["matrix"]
    for method in AUTO:
        key = f"trim_{method}_seq"
        if key in g:
            ks, kr = dropin(m, platform=oracle.PLATFORM_CUDA).trim(method)
            assert (ks == g[key]).all(), method
            assert (kr == g[f"trim_{method}_res"]).all(), method
        else:  # the reference reported an error: the CUDA platform must too
            with pytest.raises(ValueError):
                dropin(m, platform=oracle.PLATFORM_CUDA).trim(method)
    for suffix, (method, params) in PYTRIMAL.items():
        key = f"pytrimal_{suffix}_seq"
        if key in g:
            ks, kr = dropin(m, platform=oracle.PLATFORM_CUDA).trim(method, params)
            assert (ks == g[key]).all() and (kr == g[f"pytrimal_{suffix}_res"]).all(), suffix
    if "pytrimal_gt90w3_seq" in g:
        ks, kr = dropin(m, platform=oracle.PLATFORM_CUDA).trim("manual", [1 - 0.9, -1, -1, 3, -1, -1])
        assert (ks == g["pytrimal_gt90w3_seq"]).all() and (kr == g["pytrimal_gt90w3_res"]).all()


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_cuda_platform_statistics_through_manager(dropin, path):
    g = np.load(path)
    m = g["matrix"]
    r = dropin(m, platform=oracle.PLATFORM_CUDA)
    gaps, _, hist, mx = r.gaps()
    assert (gaps == g["gaps"]).all() and (hist == g["gaps_hist"]).all() and mx == int(g["gaps_max"])
    assert (bits(r.identity()) == bits(g["identity"])).all()
    assert (bits(dropin(m, platform=oracle.PLATFORM_CUDA).spurious(0.5)) == bits(g["spurious_50"])).all()
    if "similarity_error" in g:
        with pytest.raises(ValueError):
            r.similarity()
    else:
        assert (bits(r.similarity()[0]) == bits(g["mdk"])).all()


TRIMS = [("strict", []), ("strictplus", []), ("gappyout", []), ("automated1", []),
         ("manual", [1 - 0.9, 0.1, -1, 3, -1, -1]),      # BASELINE config 2
         ("manual", [1 - 0.7, -1, 40, -1, -1, -1]),
         ("overlap", [0.5, 50]), ("overlap", [0.5, 0.5]),  # BASELINE config 5 (both readings)
         ("representative", [-1, 0.8]),                   # BASELINE config 4
         ("representative", [7, -1])]


@pytest.mark.parametrize("shape,seed", [((400, 600), 1), ((1000, 300), 2), ((150, 2000), 3)])
def test_cuda_vs_avx2_platform_synthetic(dropin, shape, seed):
    from pytrimal_b200.synthetic import synthetic_msa
    m = synthetic_msa(shape[0], shape[1], seed)
    for method, params in TRIMS:
        a = dropin(m, platform=oracle.PLATFORM_AVX2).trim(method, params)
        c = dropin(m, platform=oracle.PLATFORM_CUDA).trim(method, params)
        assert (a[0] == c[0]).all() and (a[1] == c[1]).all(), (method, params)


def test_cuda_platform_symbol_error_is_reported(dropin):
    """test_automatic_trimmer.py:74-79: Alignment(["MKKBO","MKKAY"]) + strict -> error."""
    m = np.frombuffer(b"MKKBOMKKAY", np.uint8).reshape(2, 5)
    with pytest.raises(ValueError):
        dropin(m, platform=oracle.PLATFORM_AVX2).trim("strict")
    with pytest.raises(ValueError):
        dropin(m, platform=oracle.PLATFORM_CUDA).trim("strict")


def test_cuda_platform_large_window_is_reported(dropin):
    """test_manual_trimmer.py:49-52: window > L/4 fails (Gaps.cpp:98-101)."""
    g = np.load(os.path.join(GOLDEN, "example.001.AA.npz"))
    with pytest.raises(ValueError):
        dropin(g["matrix"], platform=oracle.PLATFORM_CUDA).trim("manual", [1 - 0.9, -1, -1, 100, -1, -1])


# ---- Cleaner's walks over the identity matrix on the device ----------
@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_cuda_platform_cleaner_walks_match_reference_fixture(dropin, path):
    """Cleaner::calculateRepresentativeSeq / getCutPointClusters / selectMethod of the
    patched reference with platform CUDA, against what the unpatched reference returned."""
    g = np.load(path)
    if "select_method" not in g:
        pytest.skip("fewer than two sequences")
    m = g["matrix"]
    for thr in (0.5, 0.75, 0.9):
        r = dropin(m, platform=oracle.PLATFORM_CUDA)
        assert r.representatives(thr).tolist() == g[f"repr_{int(thr * 100)}"].tolist()
        assert not r.identity_on_host()          # the matrix never crossed PCIe
    for k, want in zip(g["cutpoint_k"], g["cutpoint_thr"]):
        assert bits(dropin(m, platform=oracle.PLATFORM_CUDA).cutpoint(int(k))) == bits(want)
    assert dropin(m, platform=oracle.PLATFORM_CUDA).select_method() == int(g["select_method"])


@pytest.mark.parametrize("shape,seed", [((400, 600), 1), ((1500, 200), 2), ((2300, 128), 3)])
def test_cuda_platform_cleaner_walks_vs_avx2(dropin, shape, seed):
    from pytrimal_b200.synthetic import synthetic_msa
    m = synthetic_msa(shape[0], shape[1], seed)
    for thr in (0.0, 0.35, 0.8, 1.0):
        a = dropin(m, platform=oracle.PLATFORM_AVX2).representatives(thr)
        c = dropin(m, platform=oracle.PLATFORM_CUDA).representatives(thr)
        assert a.tolist() == c.tolist(), thr
    for k in (2, 9, shape[0] // 5, shape[0] - 1):
        a = dropin(m, platform=oracle.PLATFORM_AVX2).cutpoint(k)
        c = dropin(m, platform=oracle.PLATFORM_CUDA).cutpoint(k)
        assert bits(a) == bits(c), k
    assert dropin(m, platform=oracle.PLATFORM_AVX2).select_method() == \
        dropin(m, platform=oracle.PLATFORM_CUDA).select_method()


def test_cuda_platform_identity_materializes_on_demand(dropin):
    """After a device-side walk the host array does not exist; the first host reader
    (through Manager::calculateSeqIdentity) downloads it, bit-identical to AVX2; the
    similarity statistic keeps using the device copy."""
    from pytrimal_b200.synthetic import synthetic_msa
    m = synthetic_msa(500, 300, 11)
    want = dropin(m, platform=oracle.PLATFORM_AVX2).identity()
    r = dropin(m, platform=oracle.PLATFORM_CUDA)
    reps = r.representatives(0.8)
    assert not r.identity_on_host()
    assert r.select_method() == dropin(m, platform=oracle.PLATFORM_AVX2).select_method()
    assert not r.identity_on_host()
    mdk = r.similarity()[0]                      # CUDASimilarity: device copy, no download
    assert not r.identity_on_host()
    assert (bits(mdk) == bits(dropin(m, platform=oracle.PLATFORM_AVX2).similarity()[0])).all()
    got = r.identity()                           # a host reader: materialized now
    assert r.identity_on_host()
    assert (bits(got) == bits(want)).all()
    assert r.representatives(0.8).tolist() == reps.tolist()   # walks still on the device copy


# ---- alignment type from a device byte histogram ----------------------
@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_cuda_platform_alignment_type_fixture(dropin, path):
    g = np.load(path)
    r = dropin(g["matrix"], platform=oracle.PLATFORM_CUDA)
    assert r.detect_type() == int(g["alignment_type"])


def test_cuda_platform_alignment_type_cases(dropin):
    """Every branch of utils.cpp:514-545 plus the early NotDefined, against the AVX2 platform
    (= the reference's own scan)."""
    rng = np.random.default_rng(8)
    cases = {
        "dna": b"ACGT", "rna": b"ACGU", "tie_dna_rna": b"ACG", "deg_dna": b"ACGTRYKM",
        "deg_rna": b"ACGURYKM", "aa": b"ARNDCQEGHILKMFPSTWYV", "aa_deg": b"ARNDCQEGHILKMFPSTWYVBXZ",
        "aa_alt": b"ARNDCQEGHILKMFPSTWYVUO", "lower": b"acgtACGT", "undefined": b"ACGT#",
        "digits": b"ACDE1", "gaps_only": b"-?.",
    }
    for name, pool in cases.items():
        pool = np.frombuffer(pool, np.uint8)
        m = pool[rng.integers(0, len(pool), (37, 101))].copy()
        m[rng.random(m.shape) < 0.2] = ord("-")
        try:
            want = dropin(m, platform=oracle.PLATFORM_AVX2, datatype=0).detect_type()
        except ValueError:      # Alignment::fillMatrices refuses the symbols outright
            assert name in ("undefined", "digits")
            continue
        got = dropin(m, platform=oracle.PLATFORM_CUDA, datatype=0).detect_type()
        assert got == want, name


# ---- removeDuplicates / removeAllGapsSeqsAndCols on the device -------------
def test_cuda_platform_noduplicateseqs_and_all_gap_scans(dropin):
    """Cleaner::removeDuplicates (row hashes on the device) and removeAllGapsSeqsAndCols (row /
    column gap scans on the device) through the patched reference, against the AVX2 platform =
    the reference's own host loops."""
    from pytrimal_b200.synthetic import synthetic_msa
    rng = np.random.default_rng(21)
    m = synthetic_msa(1200, 300, 13)
    for r in rng.choice(1200, 150, replace=False):             # duplicates, some in chains
        m[r] = m[rng.integers(0, 1200)]
    m[5] = m[900]
    m[900] = m[1100]
    ks_a, kr_a = dropin(m, platform=oracle.PLATFORM_AVX2).trim("noduplicateseqs")
    ks_c, kr_c = dropin(m, platform=oracle.PLATFORM_CUDA).trim("noduplicateseqs")
    assert (ks_a == ks_c).all() and (kr_a == kr_c).all()
    assert int((ks_a == -1).sum()) >= 100
    # rows and columns that hold only gaps once other rows / columns are gone
    m2 = synthetic_msa(400, 500, 14)
    m2[:, 100:140] = ord("-")
    m2[::7, :] = ord("-")
    m2[3, 100:140] = ord("A")                                  # one row keeps those columns alive ...
    for method, params in (("noallgaps", []), ("overlap", [0.3, 30]), ("gappyout", []),
                           ("representative", [-1, 0.9])):
        a = dropin(m2, platform=oracle.PLATFORM_AVX2).trim(method, params)
        c = dropin(m2, platform=oracle.PLATFORM_CUDA).trim(method, params)
        assert (a[0] == c[0]).all() and (a[1] == c[1]).all(), method
