"""Pin the plain-C oracle: known answers, golden fixtures produced by the real
reference (tests/golden/make_golden.py) and, when oracle/_ref was built here,
the reference itself on random inputs.  CPU only."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_msa

import oracle


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
IDS = [os.path.basename(f)[:-4] for f in FIXTURES]


def test_fixtures_present():
    assert len(FIXTURES) >= 12
    idx = json.load(open(os.path.join(GOLDEN, "INDEX.json")))
    assert set(idx["inputs"]) == set(IDS)


# ---- SURVEY A.5 known answers for example.001.AA.clw --------------------------
def test_known_answers_example001(port):
    g = np.load(os.path.join(GOLDEN, "example.001.AA.npz"))
    m = g["matrix"]
    assert m.shape == (6, 46)
    gaps = port.gaps(m)[0]
    assert gaps.tolist() == [5, 5, 4, 4, 4, 2, 2, 0, 0, 0, 0, 1, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3,
                             0, 0, 0, 5, 5] + [0] * 18
    ident = port.identity(m, ord("X"))
    assert np.round(ident, 4).tolist() == pytest.approx(
        [.4000, .2609, .4000, .2045, .2439, .3409, .4474, .2439, .3611, .4545, .2955, .3636,
         .2381, .3846, .3171], abs=1e-4)
    assert [i for i in range(46) if g["trim_strictplus_res"][i] != -1] == \
        list(range(14, 18)) + list(range(29, 46))
    assert [i for i in range(46) if g["pytrimal_gt90w3_res"][i] != -1] == \
        list(range(15, 23)) + list(range(31, 46))


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_port_matches_reference_fixture(port, path):
    g = np.load(path)
    m = g["matrix"]
    indet = ord("X") if int(g["alignment_type"]) & 8 else ord("N")
    gaps, hist, mx = port.gaps(m)
    assert (gaps == g["gaps"]).all() and (hist == g["gaps_hist"]).all() and mx == int(g["gaps_max"])
    if "gaps_w3" in g:
        assert (port.gaps_window(gaps, 3) == g["gaps_w3"]).all()
    ident = port.identity(m, indet)
    assert (bits(ident) == bits(g["identity"])).all()
    for ov in (50, 80):
        want = g[f"spurious_{ov}"]
        assert (bits(port.spurious_pairwise(m, indet, ov / 100)) == bits(want)).all()
        assert (bits(port.spurious_hist(m, indet, ov / 100)) == bits(want)).all()
    if "similarity_error" in g:
        with pytest.raises(oracle.SymbolError):
            port.similarity(m, indet, ident, gaps, m.shape[1], g["dist"], g["vhash"])
    else:
        mdk, _, _ = port.similarity(m, indet, ident, gaps, m.shape[1], g["dist"], g["vhash"])
        assert (bits(mdk) == bits(g["mdk"])).all()
        if "mdk_w1" in g:
            assert (bits(port.similarity_window(mdk, 1)) == bits(g["mdk_w1"])).all()


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_port_identity_consumers_match_reference_fixture(port, path):
    """The restated Cleaner walks (calculateRepresentativeSeq, getCutPointClusters,
    selectMethod) against what the reference's own Cleaner returned."""
    g = np.load(path)
    m = g["matrix"]
    n = m.shape[0]
    if "select_method" not in g:
        pytest.skip("fewer than two sequences")
    ident = g["identity"]
    order = port.cluster_order(port.sequence_lengths(m))
    assert sorted(order.tolist()) == list(range(n))
    for thr in (0.5, 0.75, 0.9):
        assert port.greedy_clusters(ident, n, order, thr).tolist() == \
            g[f"repr_{int(thr * 100)}"].tolist()
    for k, want in zip(g["cutpoint_k"], g["cutpoint_thr"]):
        got, _ = port.cutpoint_clusters(ident, n, order, int(k))
        assert bits(got) == bits(want), (int(k), got, want)
    assert port.select_method(ident, n)[0] == int(g["select_method"])


def test_spurious_closed_form_equals_pairwise(port):
    rng = np.random.default_rng(0)
    for n, L in [(2, 3), (7, 50), (40, 33), (120, 64)]:
        m = random_msa(rng, n, L, gap=0.4, indet=0.15)
        for ov in (0.0, 0.3, 0.5, 0.99, 1.0):
            a = port.spurious_pairwise(m, ord("X"), ov)
            b = port.spurious_hist(m, ord("X"), ov)
            assert (bits(a) == bits(b)).all()


def test_gaps_simd_quirk_documented(port):
    """the SIMD u8 accumulator wraps when masked rows delay the flush;
    the restatement (and the CUDA path) return true counts."""
    n = 1200
    m = np.full((n, 40), ord("-"), np.uint8)
    ss = np.arange(n, dtype=np.int32)
    ss[0:n:255] = -1                      # every flush row is masked out
    true = port.gaps(m, ss)[0]
    quirk = port.gaps_simd_quirk(m, ss)
    assert (true == (ss != -1).sum()).all()
    assert (quirk != true).any()
    assert (port.gaps_simd_quirk(m, None) == port.gaps(m)[0]).all()


needs_ref = pytest.mark.skipif(not oracle.ref_available(),
                               reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_port_vs_live_reference_random(port, seed):
    rng = np.random.default_rng(seed)
    n, L = int(rng.integers(2, 120)), int(rng.integers(1, 400))
    m = random_msa(rng, n, L, gap=float(rng.uniform(0.05, 0.6)), lower=0.2 * (seed % 2))
    for platform in (oracle.PLATFORM_AVX2, oracle.PLATFORM_SSE2):
        r = oracle.Ref(m, platform=platform, datatype=8)      # force AA like sequence_type="protein"
        rg, _, rhist, rmx = r.gaps()
        g, hist, mx = port.gaps(m)
        assert (g == rg).all() and (hist == rhist).all() and mx == rmx
        assert (bits(port.identity(m, ord("X"))) == bits(r.identity())).all()
        assert (bits(port.spurious_pairwise(m, ord("X"), 0.5)) == bits(r.spurious(0.5))).all()
        dist, vhash = r.default_matrix()
        mdk, _, _ = port.similarity(m, ord("X"), port.identity(m, ord("X")), g, L, dist, vhash)
        assert (bits(mdk) == bits(r.similarity()[0])).all()


@needs_ref
@pytest.mark.parametrize("seed", range(4))
def test_port_identity_consumers_vs_live_reference(port, seed):
    from pytrimal_b200.synthetic import synthetic_msa
    n, L = [(150, 200), (333, 90), (64, 700), (500, 120)][seed]
    m = synthetic_msa(n, L, 50 + seed)
    if seed == 3:
        m[:, : L // 2] = m[0, : L // 2]        # many equal lengths / high identities
    ident = oracle.Ref(m, datatype=8).identity()
    lengths = port.sequence_lengths(m)
    assert (lengths == (m != ord("-")).sum(1)).all()
    order = port.cluster_order(lengths)
    for thr in (0.0, 0.4, 0.8, 1.0):
        want = oracle.Ref(m, datatype=8).representatives(thr)
        assert port.greedy_clusters(ident, n, order, thr).tolist() == want.tolist()
    for k in (2, 7, n // 4):
        want = oracle.Ref(m, datatype=8).cutpoint(k)
        assert bits(port.cutpoint_clusters(ident, n, order, k)[0]) == bits(want)
    assert port.select_method(ident, n)[0] == oracle.Ref(m, datatype=8).select_method()


@needs_ref
def test_port_vs_live_reference_masks(port):
    rng = np.random.default_rng(42)
    m = random_msa(rng, 90, 300)
    ss = np.arange(90, dtype=np.int32)
    ss[rng.random(90) < 0.3] = -1
    sr = np.arange(300, dtype=np.int32)
    sr[rng.random(300) < 0.3] = -1
    r = oracle.Ref(m, datatype=8)
    r.set_masks(ss, sr)
    assert (bits(r.identity()) == bits(port.identity(m, ord("X"), ss, sr))).all()
    r2 = oracle.Ref(m, datatype=8)
    r2.set_masks(ss, None)
    assert (r2.gaps()[0] == port.gaps(m, ss)[0]).all()    # < 255 rows: no u8 wrap possible


@needs_ref
def test_similarity_gap_cut_is_simd_form(port):
    """AVX2 cuts on 0.8 * numberOfResidues; the generic code on sequences."""
    rng = np.random.default_rng(9)
    m = random_msa(rng, 200, 60, gap=0.3)          # 0.8*60 = 48 gaps cuts; generic would need 160
    r = oracle.Ref(m, platform=oracle.PLATFORM_AVX2, datatype=8)
    mdk = r.similarity()[0]
    g = port.gaps(m)[0]
    assert (g >= 48).any()
    assert (mdk[g >= 48] == 0).all()
    dist, vhash = r.default_matrix()
    mine, _, _ = port.similarity(m, ord("X"), port.identity(m, ord("X")), g, 60, dist, vhash)
    assert (bits(mine) == bits(mdk)).all()
