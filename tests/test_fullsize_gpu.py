"""BASELINE.json's configurations at FULL size, CUDA path against the reference.

The reference (unmodified trimAl AVX2, oracle/_ref) was run once on the seeded synthetic
alignments of every configuration (tests/golden/make_golden_full.py: 6 to 40 minutes of one
host core each); its results are committed under tests/golden/full/ -- small arrays verbatim,
the packed identity arrays as digests (tests/golden/digest.py).  Here the same inputs are
regenerated and pushed through the C ABI, and -- where the reference's Python package built
with the CUDA platform is present (integration/_build/pkg) -- through pytrimal's own
`trim()` with platform="cuda".  Everything is compared bit for bit.

north_star's target sentence ("the synthetic 50k x 1k RepresentativeTrimmer config produces
byte-identical trimmed alignments to pytrimal's AVX2 backend") is
test_c4_representative_trimmer_full_size below.
"""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, GOLDEN)
from digest import block_sums, sha256_hex  # noqa: E402

pytestmark = pytest.mark.gpu

FULL = os.path.join(GOLDEN, "full")
PKG = os.path.join(ROOT, "integration", "_build", "pkg")
X = ord("X")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def golden(job):
    path = os.path.join(FULL, job + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"tests/golden/full/{job}.npz not generated")
    return np.load(path)


_msa_cache = {}


def msa_of(cfg, g):
    """The seeded alignment of a configuration, checked against the generator's digest so
    that a drift of the random stream cannot pass as a parity failure (or success)."""
    if cfg not in _msa_cache:
        from pytrimal_b200.synthetic import CONFIGS, synthetic_msa
        _msa_cache.clear()                       # one big matrix at a time
        n, L, seed = CONFIGS[cfg]
        _msa_cache[cfg] = synthetic_msa(n, L, seed)
    m = _msa_cache[cfg]
    assert list(m.shape) + [int(g["shape_seed"][2])] == g["shape_seed"].tolist()
    assert sha256_hex(m) == str(g["matrix_sha256"]), "synthetic generator drifted"
    return m


@pytest.fixture(scope="module")
def pytrimal(gpu):
    if not os.path.isdir(os.path.join(PKG, "pytrimal")):
        pytest.skip("integration/_build/pkg not present (built where /root/reference exists)")
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    import pytrimal
    assert pytrimal._trimal._CUDA_RUNTIME_SUPPORT
    return pytrimal


def trimmed_equals_masks(out, m, keep_seq, keep_res):
    rows = np.nonzero(keep_seq != -1)[0]
    cols = np.nonzero(keep_res != -1)[0]
    if list(out.names) != [b"s%d" % i for i in rows]:
        return False
    want = m[np.ix_(rows, cols)] if len(cols) else np.zeros((len(rows), 0), np.uint8)
    got = list(out.sequences)
    return len(got) == len(rows) and all(got[k].encode() == bytes(want[k]) for k in range(len(rows)))


def check_identity_digests(g, ident, n):
    assert ident.size == int(g["identity_count"])
    s, w = block_sums(ident, n, 128)
    assert (s == g["identity_block_sum"]).all() and (w == g["identity_block_wsum"]).all()
    assert sha256_hex(ident) == str(g["identity_sha256"])


# ---- C4: RepresentativeTrimmer(identity_threshold=0.8), 50 000 x 1 000 ---------------------
def test_c4_identity_matrix_and_representatives_full_size(gpu):
    g = golden("C4")
    m = msa_of("C4", g)
    n = m.shape[0]
    with gpu.DeviceAlignment(m) as d:
        reps = d.representatives(0.8, indet=X)          # K1 threshold mode + K7
        assert reps.tolist() == g["representatives_80"].tolist()
        assert (d.gaps()[0] == g["gaps"]).all()
        ident = d.identity(X, keep_on_device=True)      # 5 GB of floats to the host
        check_identity_digests(g, ident, n)
        del ident
        order = gpu.cluster_order(d.sequence_lengths())
        assert d.clusters(order, 0.8).tolist() == g["representatives_80"].tolist()   # float path (K5)
        assert d.select_method()[0] == {1: "gappyout", 2: "strict"}[int(g["select_method"])]
    kept = np.zeros(n, bool)
    kept[reps] = True
    assert (kept == (g["trim_maxidentity80_seq"] != -1)).all()


def test_c4_representative_trimmer_full_size(gpu, pytrimal):
    g = golden("C4")
    m = msa_of("C4", g)
    ali = pytrimal.Alignment([b"s%d" % i for i in range(m.shape[0])], [bytes(r) for r in m])
    out = pytrimal.RepresentativeTrimmer(identity_threshold=0.8, platform="cuda").trim(ali)
    assert trimmed_equals_masks(out, m, g["trim_maxidentity80_seq"], g["trim_maxidentity80_res"])


# ---- C3: AutomaticTrimmer gappyout / strict / strictplus, 10 000 x 5 000 -------------------
def test_c3_statistics_full_size(gpu):
    g = golden("C3")
    m = msa_of("C3", g)
    n, L = m.shape
    smx = gpu.SimilarityMatrix.aa()
    with gpu.DeviceAlignment(m) as d:
        gaps = d.gaps()[0]
        assert (gaps == g["gaps"]).all()
        ident = d.identity(X, keep_on_device=True)
        check_identity_digests(g, ident, n)
        del ident
        mdk, num, den = d.similarity(smx, gaps=gaps, indet=X)
    # n = 10 000 is where the reference's sequential fp32 denominator saturates:
    # only the reference's own summation order reproduces these bits
    assert (bits(mdk) == bits(g["mdk"])).all()
    assert float(den.max()) >= 2.0 ** 24 * 0.99


@pytest.mark.parametrize("method", ["gappyout", "strict", "strictplus", "automated1"])
def test_c3_automatic_trimmer_full_size(gpu, pytrimal, method):
    g = golden("C3" if method == "gappyout" else "C3." + method)
    m = msa_of("C3", g)
    ali = pytrimal.Alignment([b"s%d" % i for i in range(m.shape[0])], [bytes(r) for r in m])
    out = pytrimal.AutomaticTrimmer(method, platform="cuda").trim(ali)
    assert trimmed_equals_masks(out, m, g[f"trim_{method}_seq"], g[f"trim_{method}_res"])


# ---- C2: ManualTrimmer(gap_threshold=.9, similarity_threshold=.1, window=3), 1 000 x 2 000 --
def test_c2_statistics_and_manual_trimmer(gpu, pytrimal):
    g = golden("C2")
    m = msa_of("C2", g)
    n, L = m.shape
    smx = gpu.SimilarityMatrix.aa()
    with gpu.DeviceAlignment(m) as d:
        gaps = d.gaps()[0]
        assert (gaps == g["gaps"]).all()
        ident = d.identity(X, keep_on_device=True)
        assert (bits(ident) == bits(g["identity"])).all()
        mdk, _, _ = d.similarity(smx, gaps=gaps, indet=X)
        assert (bits(mdk) == bits(g["mdk"])).all()
        gw = gpu.gaps_window(gaps, 3)
        assert (gw == g["gaps_w3"]).all()
        mdk3, _, _ = d.similarity(smx, gaps=gw, indet=X)
        assert (bits(gpu.similarity_window(mdk3, 3)) == bits(g["mdk_w3"])).all()
    ali = pytrimal.Alignment([b"s%d" % i for i in range(n)], [bytes(r) for r in m])
    # the literal configuration (it removes every column of this alignment: a degenerate
    # but valid answer) and the same trimmer at thresholds that keep 646 / 1910 columns
    cases = [("literal", dict(gap_threshold=0.9, similarity_threshold=0.1, window=3))]
    for tag in ("gt50_st001", "gt20_st0"):
        gt, st, w = g[f"trim_{tag}_params"].tolist()
        cases.append((tag, dict(gap_threshold=gt, similarity_threshold=st, window=int(w))))
    for tag, kwargs in cases:
        try:
            out = pytrimal.ManualTrimmer(platform="cuda", **kwargs).trim(ali)
        except Exception:
            out = None
        ks, kr = g[f"trim_{tag}_seq"], g[f"trim_{tag}_res"]
        if (kr == -1).all():
            # the reference leaves nothing: an error or an empty alignment, never residues
            assert out is None or len(out.sequences) == 0 or len(out.sequences[0]) == 0, tag
        else:
            assert out is not None and trimmed_equals_masks(out, m, ks, kr), tag
    assert int((g["trim_gt50_st001_res"] != -1).sum()) == 646


# ---- C5: OverlapTrimmer(sequence_overlap, residue_overlap=0.5), 100 000 x 2 000 ------------
def test_c5_spurious_vector_full_size(gpu, port):
    g = golden("C5")
    m = msa_of("C5", g)
    with gpu.DeviceAlignment(m) as d:
        sp = d.spurious(0.5, indet=X)
        assert (d.gaps()[0] == g["gaps"]).all()
    assert (bits(sp) == bits(g["spurious_50"])).all()
    # and the oracle's O(n L) closed form, pinned against the pairwise loop on small inputs
    assert (bits(port.spurious_hist(m, X, 0.5)) == bits(sp)).all()


@pytest.mark.parametrize("seq_overlap,res_overlap,job", [(0.5, 0.5, "C5.seq0.5"), (50, 0.5, "C5.seq50"),
                                                         (75, 0.7, "C5.seq75.res0.7")])
def test_c5_overlap_trimmer_full_size(gpu, pytrimal, seq_overlap, res_overlap, job):
    """The literal configuration (sequence_overlap=0.5 is 0.5 %) and its 50 % reading
    keep every sequence of this alignment; the third case drops about a quarter of them."""
    g = golden(job)
    m = msa_of("C5", g)
    ali = pytrimal.Alignment([b"s%d" % i for i in range(m.shape[0])], [bytes(r) for r in m])
    out = pytrimal.OverlapTrimmer(sequence_overlap=seq_overlap, residue_overlap=res_overlap,
                                  platform="cuda").trim(ali)
    tag = ("%g" % seq_overlap).replace(".", "p")
    if res_overlap != 0.5:
        tag += "_res" + ("%g" % res_overlap).replace(".", "p")
    assert trimmed_equals_masks(out, m, g[f"trim_overlap_seq{tag}_seq"],
                                g[f"trim_overlap_seq{tag}_res"])


# ---- beyond the packed float array: 200 000 sequences ---------------------------------------
def test_representatives_200k_sequences(gpu):
    """tcu_representatives thresholds inside the identity kernel, so nothing of size 4*P exists:
    200 000 sequences (P = 2 x 10^10 pairs; the packed float array alone would be 80 GB, its
    host copy in the reference as much again) cluster with a 5 GB bit matrix.  The input is
    built so that the answer is known: 4 000 families of 50 near-identical members (identity
    ~0.94 inside a family, ~0.05 across), hence exactly one representative per family -- its
    first member in the reference's visiting order."""
    n, L, nfam = 200_000, 256, 4000
    rng = np.random.default_rng(200)
    aa = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", np.uint8)
    anc = aa[rng.integers(0, 20, (nfam, L))]
    fam = rng.permutation(np.repeat(np.arange(nfam), n // nfam))
    m = anc[fam].copy()
    sub = rng.random((n, L)) < 0.03
    m[sub] = aa[rng.integers(0, 20, int(sub.sum()))]
    lead = rng.integers(0, 9, n)
    m[np.arange(L)[None, :] < lead[:, None]] = ord("-")          # ragged starts: lengths differ
    with gpu.DeviceAlignment(m) as d:
        reps = d.representatives(0.8, indet=X)
        assert not d.identity_resident
        order = gpu.cluster_order(d.sequence_lengths())
    seen = np.zeros(nfam, bool)
    want = []
    for s in order:
        f = fam[s]
        if not seen[f]:
            seen[f] = True
            want.append(int(s))
    assert len(want) == nfam and reps.tolist() == want
