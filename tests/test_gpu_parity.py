"""GPU parity: the CUDA library (through its C ABI) against the oracle.

Integer counts and identity ratios must be bit-exact; similarity accumulators
are compared bit-for-bit as well (the kernel replays the reference's fp32
order), with the 1e-5 relative bound of BASELINE.json as the hard limit.
"""
import numpy as np
import pytest

from conftest import random_msa

pytestmark = pytest.mark.gpu

X = ord("X")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


SHAPES = [(2, 1), (2, 31), (3, 32), (5, 33), (6, 46), (17, 255), (64, 256), (65, 257),
          (130, 300), (200, 1000), (257, 513)]


@pytest.mark.parametrize("n,L", SHAPES)
def test_gaps(gpu, port, n, L):
    rng = np.random.default_rng(n * 1000 + L)
    m = random_msa(rng, n, L)
    with gpu.DeviceAlignment(m) as d:
        g, hist, mx = d.gaps()
    og, ohist, omx = port.gaps(m)
    assert (g == og).all() and (hist == ohist).all() and mx == omx


def test_gaps_masked_rows_true_counts(gpu, port):
    """Masked rows: true counts (generic path), not the SIMD u8 wrap."""
    rng = np.random.default_rng(5)
    m = random_msa(rng, 700, 100, gap=0.9)
    ss = np.arange(700, dtype=np.int32)
    ss[rng.random(700) < 0.3] = -1
    with gpu.DeviceAlignment(m) as d:
        g, _, _ = d.gaps(save_seq=ss)
    assert (g == port.gaps(m, ss)[0]).all()


def test_gaps_many_rows(gpu, port):
    rng = np.random.default_rng(6)
    m = random_msa(rng, 5000, 333, gap=0.5)
    with gpu.DeviceAlignment(m) as d:
        g, hist, mx = d.gaps()
    og, ohist, omx = port.gaps(m)
    assert (g == og).all() and (hist == ohist).all() and mx == omx


@pytest.mark.parametrize("n,L", SHAPES)
def test_identity_counts_and_ratio(gpu, port, n, L):
    rng = np.random.default_rng(n * 7919 + L)
    m = random_msa(rng, n, L)
    with gpu.DeviceAlignment(m) as d:
        ident, hit, dst = d.identity(X, counts=True)
    oi, oh, od = port.identity(m, X, counts=True)
    assert (hit == oh).all()
    assert (dst == od).all()
    assert (bits(ident) == bits(oi)).all()


@pytest.mark.parametrize("n,L", [(12, 100), (65, 257), (130, 300)])
def test_identity_arbitrary_bytes_bytewise_kernel(gpu, port, n, L):
    """More than 126 distinct non-gap byte values (nothing trimAl's validation admits, but the
    reference's kernels compare raw bytes): the byte-wise kernel behind the same entry points,
    with masks, and the clustering on top of it."""
    rng = np.random.default_rng(n + L)
    m = rng.integers(0, 256, (n, L), dtype=np.uint8)
    m[rng.random((n, L)) < 0.2] = ord("-")
    m[rng.random((n, L)) < 0.05] = X
    m[1:n // 2] = np.where(rng.random((n // 2 - 1, L)) < 0.7, m[0], m[1:n // 2])   # related rows
    assert len(np.unique(m)) > 128
    sr = np.arange(L, dtype=np.int32)
    sr[rng.random(L) < 0.3] = -1
    with gpu.DeviceAlignment(m) as d:
        ident, hit, dst = d.identity(X, counts=True)
        masked = d.identity(X, save_res=sr)
        reps = d.representatives(0.5, indet=X)
    oi, oh, od = port.identity(m, X, counts=True)
    assert (hit == oh).all() and (dst == od).all() and (bits(ident) == bits(oi)).all()
    assert (bits(masked) == bits(port.identity(m, X, None, sr))).all()
    order = port.cluster_order(port.sequence_lengths(m))
    assert reps.tolist() == port.greedy_clusters(oi, n, order, 0.5).tolist()


def test_identity_masks(gpu, port):
    rng = np.random.default_rng(11)
    m = random_msa(rng, 150, 400)
    ss = np.arange(150, dtype=np.int32)
    ss[rng.random(150) < 0.25] = -1
    sr = np.arange(400, dtype=np.int32)
    sr[rng.random(400) < 0.4] = -1
    with gpu.DeviceAlignment(m) as d:
        ident, hit, dst = d.identity(X, save_seq=ss, save_res=sr, counts=True)
    oi, oh, od = port.identity(m, X, ss, sr, counts=True)
    assert (hit == oh).all() and (dst == od).all() and (bits(ident) == bits(oi)).all()


def test_identity_raw_bytes_case_sensitive(gpu, port):
    """Lower case, punctuation and >30 distinct symbols (6-7 planes); 'x' is a
    residue, 'X' a gap."""
    rng = np.random.default_rng(12)
    m = random_msa(rng, 90, 500, lower=0.3, extra=b"BZJUO*.?~!#")
    m[rng.random(m.shape) < 0.02] = ord("x")
    with gpu.DeviceAlignment(m) as d:
        ident, hit, dst = d.identity(X, counts=True)
    oi, oh, od = port.identity(m, X, counts=True)
    assert (hit == oh).all() and (dst == od).all() and (bits(ident) == bits(oi)).all()


@pytest.mark.parametrize("alphabet,indet", [(b"ACGT", ord("N")), (b"AC", ord("N"))])
def test_identity_small_alphabets(gpu, port, alphabet, indet):
    rng = np.random.default_rng(13)
    m = random_msa(rng, 70, 200, alphabet=alphabet, indet=0.0)
    m[rng.random(m.shape) < 0.03] = indet
    with gpu.DeviceAlignment(m) as d:
        ident, hit, dst = d.identity(indet, counts=True)
    oi, oh, od = port.identity(m, indet, counts=True)
    assert (hit == oh).all() and (dst == od).all() and (bits(ident) == bits(oi)).all()


@pytest.mark.parametrize("n,L", [(40, 65536), (130, 66000)])
def test_identity_very_long_rows_unpacked_counters(gpu, port, n, L):
    """>= 65536 columns: the kernel variant with 32-bit hit counters (the packed
    16-bit pairs could overflow); also many k-stages of the both-gap UMMA."""
    rng = np.random.default_rng(L + n)
    m = random_msa(rng, n, L, gap=0.3)
    # long identical gap-free stretches: hit counts beyond 16 bits' half range
    m[: n // 2, : L // 2 + 100] = random_msa(rng, 1, L // 2 + 100, gap=0.0, indet=0.0)[0]
    with gpu.DeviceAlignment(m) as d:
        ident, hit, dst = d.identity(X, counts=True)
    oi, oh, od = port.identity(m, X, counts=True)
    assert hit.max() > 32767
    assert (hit == oh).all() and (dst == od).all() and (bits(ident) == bits(oi)).all()


@pytest.mark.parametrize("n,L", [(127, 129), (128, 128), (129, 127), (191, 64), (193, 65),
                                 (600, 130), (1100, 70), (2500, 140), (2700, 33)])
def test_identity_tile_edges(gpu, port, n, L):
    """Row counts around the 128-row super-block / 64-row block edges, several
    tiles per CTA (both TMEM accumulator buffers, ring wrap-around) and several
    groups of super-block rows in the grouped tile order (full, partial, single)."""
    rng = np.random.default_rng(n * 31 + L)
    m = random_msa(rng, n, L, gap=0.4)
    with gpu.DeviceAlignment(m) as d:
        ident, hit, dst = d.identity(X, counts=True)
    oi, oh, od = port.identity(m, X, counts=True)
    assert (hit == oh).all() and (dst == od).all() and (bits(ident) == bits(oi)).all()


@pytest.mark.parametrize("nsym", [8, 14, 15, 30, 31, 62, 63, 100])
def test_identity_plane_counts(gpu, port, nsym):
    """Alphabets that need 4, 5, 6 and 7 code planes (2^NP - 2 residue codes)."""
    rng = np.random.default_rng(nsym)
    pool = bytes(b for b in range(33, 127) if b not in (ord("-"), X))[:nsym] if nsym <= 92 else \
        bytes(b for b in range(33, 256) if b not in (ord("-"), X))[:nsym]
    m = random_msa(rng, 140, 300, alphabet=pool)
    with gpu.DeviceAlignment(m) as d:
        ident, hit, dst = d.identity(X, counts=True)
    oi, oh, od = port.identity(m, X, counts=True)
    assert (hit == oh).all() and (dst == od).all() and (bits(ident) == bits(oi)).all()


def test_identity_all_gap_pairs(gpu, port):
    """dst == 0 -> identity 0 (template.h:427-428)."""
    m = np.full((5, 70), ord("-"), np.uint8)
    m[0, :10] = ord("A")
    m[3, 5:20] = X
    with gpu.DeviceAlignment(m) as d:
        ident, hit, dst = d.identity(X, counts=True)
    oi, oh, od = port.identity(m, X, counts=True)
    assert (hit == oh).all() and (dst == od).all() and (bits(ident) == bits(oi)).all()
    assert (ident[od == 0] == 0).all()


def test_identity_medium_vs_oracle(gpu, port):
    from pytrimal_b200.synthetic import synthetic_msa
    m = synthetic_msa(700, 1300, 21)
    with gpu.DeviceAlignment(m) as d:
        ident, hit, dst = d.identity(X, counts=True)
    oi, oh, od = port.identity(m, X, counts=True)
    assert (hit == oh).all() and (dst == od).all() and (bits(ident) == bits(oi)).all()


@pytest.mark.parametrize("n,L", SHAPES)
@pytest.mark.parametrize("overlap", [0.0, 0.5, 0.8, 1.0])
def test_spurious(gpu, port, n, L, overlap):
    rng = np.random.default_rng(n * 31 + L)
    m = random_msa(rng, n, L, gap=0.4, indet=0.1)
    with gpu.DeviceAlignment(m) as d:
        s = d.spurious(overlap, indet=X)
    o = port.spurious_pairwise(m, X, overlap)
    assert (bits(s) == bits(o)).all()


@pytest.mark.parametrize("n,L", [(2, 5), (6, 46), (40, 130), (150, 257), (300, 96)])
@pytest.mark.parametrize("cut", [False, True])
def test_similarity(gpu, port, n, L, cut):
    rng = np.random.default_rng(n * 17 + L)
    m = random_msa(rng, n, L, gap=0.3, lower=0.2)
    smx = gpu.SimilarityMatrix.aa()
    og = port.gaps(m)[0]
    oi = port.identity(m, X)
    gaps = og if cut else None
    omdk, onum, oden = port.similarity(m, X, oi, gaps, L, smx.distances, smx.vhash)
    with gpu.DeviceAlignment(m) as d:
        d.identity(X, keep_on_device=True)
        mdk, num, den = d.similarity(smx, gaps=gaps, indet=X)
    assert (bits(num) == bits(onum)).all()
    assert (bits(den) == bits(oden)).all()
    assert (bits(mdk) == bits(omdk)).all()
    np.testing.assert_allclose(mdk, omdk, rtol=1e-5, atol=0)   # BASELINE.json tolerance


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8, 9, 33, 34, 66])
def test_similarity_few_batches(gpu, port, n):
    """One to a handful of 32-row batches per column group (the consumers' peeled last
    batches, odd and even counts), with all-gap rows and an all-gap group in between."""
    rng = np.random.default_rng(900 + n)
    L = 100
    m = random_msa(rng, n, L, gap=0.45)
    m[:, 32:64] = ord("-")                       # a column group without any batch
    m[n // 2, 64:96] = ord("-")                  # a row the third group skips
    smx = gpu.SimilarityMatrix.aa()
    oi = port.identity(m, X)
    omdk, onum, oden = port.similarity(m, X, oi, None, L, smx.distances, smx.vhash)
    with gpu.DeviceAlignment(m) as d:
        d.identity(X, keep_on_device=True)
        mdk, num, den = d.similarity(smx, gaps=None, indet=X)
    assert (bits(num) == bits(onum)).all()
    assert (bits(den) == bits(oden)).all()
    assert (bits(mdk) == bits(omdk)).all()


def test_similarity_gap_cut_uses_residue_count(gpu, port):
    """threshold = 0.8 * numberOfResidues (columns!), SURVEY F4."""
    rng = np.random.default_rng(3)
    n, L = 300, 100                      # 0.8*L = 80 < n: columns with >= 80 gaps are cut
    m = random_msa(rng, n, L, gap=0.3)
    smx = gpu.SimilarityMatrix.aa()
    og = port.gaps(m)[0]
    assert (og >= 80).any() and (og < 80).any()
    oi = port.identity(m, X)
    omdk, onum, oden = port.similarity(m, X, oi, og, L, smx.distances, smx.vhash)
    with gpu.DeviceAlignment(m) as d:
        d.identity(X, keep_on_device=True)
        mdk, num, den = d.similarity(smx, gaps=og, indet=X)
    assert (mdk[og >= 80] == 0).all()
    assert (bits(mdk) == bits(omdk)).all()


def test_similarity_symbol_errors(gpu, port):
    """First offender in column-major scan order; error class as template.h:135-145."""
    import oracle
    rng = np.random.default_rng(4)
    m = random_msa(rng, 20, 60, gap=0.1)
    smx = gpu.SimilarityMatrix.aa()
    m[7, 30] = ord("B")      # undefined in BLOSUM62's 20 letters
    m[3, 41] = ord("*")      # incorrect symbol
    m[15, 30] = ord("O")
    oi = port.identity(m, X)
    with pytest.raises(oracle.SymbolError) as oe:
        port.similarity(m, X, oi, None, 60, smx.distances, smx.vhash)
    with gpu.DeviceAlignment(m) as d:
        d.identity(X, keep_on_device=True)
        with pytest.raises(gpu.SymbolError) as ge:
            d.similarity(smx, indet=X)
    assert (ge.value.col, ge.value.row, ge.value.byte) == (oe.value.col, oe.value.row, oe.value.byte)
    assert ge.value.code == -6 and oe.value.code == 2
    m[2, 5] = ord("?")
    with pytest.raises(oracle.SymbolError) as oe:
        port.similarity(m, X, oi, None, 60, smx.distances, smx.vhash)
    with gpu.DeviceAlignment(m) as d:
        d.identity(X, keep_on_device=True)
        with pytest.raises(gpu.SymbolError) as ge:
            d.similarity(smx, indet=X)
    assert (ge.value.col, ge.value.row, ge.value.byte) == (5, 2, ord("?")) == \
        (oe.value.col, oe.value.row, oe.value.byte)
    assert ge.value.code == -5 and oe.value.code == 1


def test_similarity_requires_identity(gpu):
    m = random_msa(np.random.default_rng(1), 10, 40)
    with gpu.DeviceAlignment(m) as d:
        with pytest.raises(gpu.TrimalCudaError):
            d.similarity(gpu.SimilarityMatrix.aa(), indet=X)


def test_identity_row_band_matches_full(gpu, port):
    """Row-block bands written by tcu_identity_device tile the packed array."""
    import ctypes as C
    import torch
    from pytrimal_b200.synthetic import synthetic_msa
    m = synthetic_msa(600, 700, 5)
    oi = port.identity(m, X)
    lib = gpu.load()
    R = lib.tcu_identity_band_rows()
    with gpu.DeviceAlignment(m) as d:
        nk = d.identity_prepare(X)
        nb = lib.tcu_identity_row_blocks(nk)
        out = torch.full((nk * (nk - 1) // 2,), -1.0, dtype=torch.float32, device="cuda:0")
        for b0, b1 in [(0, 2), (2, 3), (3, nb)]:
            off = lib.tcu_identity_row_offset(nk, R * b0)
            d.identity_device(b0, b1, out.data_ptr() + 4 * off)
        d.sync()
        got = out.cpu().numpy()
    assert (bits(got) == bits(oi)).all()
