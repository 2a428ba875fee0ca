"""GPU parity of the identity-matrix consumers: per-row statistics,
greedy clustering, cluster cut point, selectMethod and sequence lengths, through the C
ABI, against the oracle's restatement of the Cleaner.cpp walks and against the fixtures
produced by the reference's own Cleaner (tests/golden)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_msa

pytestmark = pytest.mark.gpu

X = ord("X")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
IDS = [os.path.basename(f)[:-4] for f in FIXTURES]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def family_msa(n, L, seed):
    from pytrimal_b200.synthetic import synthetic_msa
    return synthetic_msa(n, L, seed)


@pytest.mark.parametrize("n,L", [(2, 5), (3, 40), (31, 64), (32, 64), (33, 100), (64, 65),
                                 (65, 300), (257, 129), (1000, 200), (1025, 64), (2100, 96)])
def test_row_stats_lengths_and_clusters(gpu, port, n, L):
    rng = np.random.default_rng(n * 31 + L)
    m = family_msa(n, L, n + L) if n >= 64 else random_msa(rng, n, L)
    with gpu.DeviceAlignment(m) as d:
        assert not d.identity_resident
        d.identity_on_device(X)
        assert d.identity_resident
        ident = d.identity_download()
        oi = port.identity(m, X)
        assert (bits(ident) == bits(oi)).all()
        lengths = d.sequence_lengths()
        assert (lengths == port.sequence_lengths(m)).all()
        order = gpu.cluster_order(lengths)
        assert order.tolist() == port.cluster_order(lengths).tolist()
        for upper in (False, True):
            mx, mn, sm = d.identity_row_stats(upper_only=upper)
            omx, omn, osm = port.identity_row_stats(oi, n, upper)
            assert (bits(mx) == bits(omx)).all()
            assert (bits(mn) == bits(omn)).all()
            assert (bits(sm) == bits(osm)).all()
        qs = np.quantile(oi, [0.0, 0.1, 0.5, 0.9, 0.99, 1.0]) if len(oi) else [0.5]
        for thr in [0.0, 1.0, 0.8] + [float(q) for q in qs]:
            want = port.greedy_clusters(oi, n, order, thr)
            got = d.clusters(order, thr)
            assert got.tolist() == want.tolist(), thr
            assert d.clusters(order, thr, count_only=True) == len(want)
        # any visiting order, not only the length-sorted one
        perm = rng.permutation(n).astype(np.int32)
        assert d.clusters(perm, float(qs[len(qs) // 2])).tolist() == \
            port.greedy_clusters(oi, n, perm, float(qs[len(qs) // 2])).tolist()
        # a prefix of the order
        half = order[: max(n // 2, 1)]
        assert d.clusters(half, 0.5).tolist() == port.greedy_clusters(oi, n, half, 0.5).tolist()


def test_consumers_need_a_resident_matrix(gpu):
    m = family_msa(100, 80, 3)
    with gpu.DeviceAlignment(m) as d:
        with pytest.raises(gpu.TrimalCudaError):
            d.identity_row_stats()
        ss = np.arange(100, dtype=np.int32)
        ss[3] = -1
        d.identity(X, save_seq=ss, keep_on_device=True)      # masked rows: not usable
        with pytest.raises(gpu.TrimalCudaError):
            d.clusters(np.arange(100, dtype=np.int32), 0.5)
        d.identity_on_device(X)
        with pytest.raises(ValueError):
            d.clusters(np.array([0, 100], np.int32), 0.5)    # index out of range
        assert d.clusters(np.zeros(0, np.int32), 0.5).tolist() == []


@pytest.mark.parametrize("n,L,seed", [(300, 200, 1), (77, 500, 2), (1200, 150, 3), (2500, 100, 4)])
def test_select_method_cutpoint_and_representatives(gpu, port, n, L, seed):
    m = family_msa(n, L, seed)
    oi = port.identity(m, X)
    order = port.cluster_order(port.sequence_lengths(m))
    with gpu.DeviceAlignment(m) as d:
        for thr in (0.3, 0.8, 0.95):
            assert d.representatives(thr, indet=X).tolist() == \
                port.greedy_clusters(oi, n, order, thr).tolist()
        assert not d.identity_resident      # the one-call form thresholds inside K1: no floats
        d.identity_on_device(X)
        for thr in (0.3, 0.8, 0.95):        # same walk over the resident float matrix (K5)
            assert d.clusters(order, thr).tolist() == port.greedy_clusters(oi, n, order, thr).tolist()
        name, avg_seq, max_seq = d.select_method()
        code, oavg, omax = port.select_method(oi, n)
        assert bits(avg_seq) == bits(oavg) and bits(max_seq) == bits(omax)
        assert name == {1: "gappyout", 2: "strict"}[code]
        for k in (2, 5, n // 3, n - 1, 1, n):
            got, runs = d.cutpoint_clusters(k)
            want, oruns = port.cutpoint_clusters(oi, n, order, k)
            assert bits(got) == bits(want) and runs == oruns, k


@pytest.mark.parametrize("n,L", [(1, 7), (2, 5), (3, 40), (31, 64), (33, 100), (63, 33), (64, 65),
                                 (65, 300), (127, 130), (128, 128), (129, 127), (257, 129),
                                 (383, 700), (1000, 200), (1025, 64), (2100, 96), (2700, 33)])
def test_representatives_threshold_epilogue(gpu, port, n, L):
    """tcu_representatives = K1 in threshold mode (one bit per pair, slab layout) + the mirror
    pass + the clustering kernel (K7): every tile-edge shape, thresholds at exact identity values (the > must not
    become >=), negative and >= 1 thresholds, masked columns."""
    rng = np.random.default_rng(n * 131 + L)
    m = family_msa(n, L, n + L) if n >= 64 else random_msa(rng, n, L)
    oi = port.identity(m, X)
    order = port.cluster_order(port.sequence_lengths(m))
    qs = np.quantile(oi, [0.0, 0.25, 0.5, 0.9, 1.0]).tolist() if len(oi) else []
    exact = [float(v) for v in rng.choice(oi, min(3, len(oi)), replace=False)] if len(oi) else []
    with gpu.DeviceAlignment(m) as d:
        for thr in [-1.0, 0.0, 0.8, 1.0, 2.0] + qs + exact:
            got = d.representatives(thr, indet=X)
            assert got.tolist() == port.greedy_clusters(oi, n, order, np.float32(thr)).tolist(), thr
        sr = np.arange(L, dtype=np.int32)
        sr[rng.random(L) < 0.5] = -1
        om = port.identity(m, X, None, sr)
        assert d.representatives(0.5, indet=X, save_res=sr).tolist() == \
            port.greedy_clusters(om, n, order, 0.5).tolist()


def test_representatives_dense_graph_long_chains(gpu, port):
    """Worst case for the fixed-point resolve: a path-like threshold graph in visiting order
    (each sequence within the threshold of its successor only), and a complete graph."""
    n, L = 1500, 256
    rng = np.random.default_rng(5)
    m = np.empty((n, L), np.uint8)
    base = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", np.uint8)[rng.integers(0, 20, L)]
    m[0] = base
    for r in range(1, n):                      # drift: row r differs from row r-1 in 8 columns
        m[r] = m[r - 1]
        cols = rng.choice(L, 8, replace=False)
        m[r, cols] = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", np.uint8)[rng.integers(0, 20, 8)]
    oi = port.identity(m, X)
    order = port.cluster_order(port.sequence_lengths(m))
    with gpu.DeviceAlignment(m) as d:
        for thr in (0.96, 0.9, 0.5, 0.0):
            assert d.representatives(thr, indet=X).tolist() == \
                port.greedy_clusters(oi, n, order, thr).tolist(), thr


def test_representatives_with_column_mask(gpu, port):
    m = family_msa(400, 300, 9)
    sr = np.arange(300, dtype=np.int32)
    sr[np.random.default_rng(1).random(300) < 0.4] = -1
    oi = port.identity(m, X, None, sr)
    order = port.cluster_order(port.sequence_lengths(m))
    with gpu.DeviceAlignment(m) as d:
        assert d.representatives(0.7, indet=X, save_res=sr).tolist() == \
            port.greedy_clusters(oi, 400, order, 0.7).tolist()


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_consumers_match_reference_fixture(gpu, path):
    """Against what the reference's own Cleaner returned (tests/golden/make_golden.py)."""
    g = np.load(path)
    if "select_method" not in g:
        pytest.skip("fewer than two sequences")
    a = gpu.Alignment.from_matrix(g["matrix"])
    n = a.nseq
    with gpu.DeviceAlignment(a) as d:
        for thr in (0.5, 0.75, 0.9):
            assert d.representatives(thr).tolist() == g[f"repr_{int(thr * 100)}"].tolist()
        d.identity_on_device()
        assert d.select_method()[0] == {1: "gappyout", 2: "strict"}[int(g["select_method"])]
        for k, want in zip(g["cutpoint_k"], g["cutpoint_thr"]):
            assert bits(d.cutpoint_clusters(int(k))[0]) == bits(want)


def test_representatives_full_size_properties(gpu):
    """BASELINE configs[3] (50 000 x 1 000, RepresentativeTrimmer 0.8) at full size: the
    result must be a maximal independent set of the threshold graph in visiting order --
    checked on the host from the downloaded matrix with vectorised numpy."""
    from pytrimal_b200.synthetic import CONFIGS
    n, L, seed = CONFIGS["C4"]
    m = family_msa(n, L, seed)
    thr = 0.8
    with gpu.DeviceAlignment(m) as d:
        reps = d.representatives(thr, indet=X)
        d.identity_on_device(X)
        ident = d.identity_download()
        lengths = d.sequence_lengths()
        order = gpu.cluster_order(lengths)
    assert (lengths == (m != ord("-")).sum(1)).all()
    assert len(set(reps.tolist())) == len(reps) and reps[0] == order[0]
    rank = np.empty(n, np.int64)
    rank[order] = np.arange(n)
    assert (np.diff(rank[reps]) > 0).all()              # creation order = visiting order
    is_rep = np.zeros(n, bool)
    is_rep[reps] = True
    rows = np.arange(n, dtype=np.int64)
    row_base = rows * n - (rows + 1) * (rows + 2) // 2  # + j = packed position of (i, j)

    def ident_row(v):
        out = np.zeros(n, np.float32)
        out[v + 1:] = ident[row_base[v] + v + 1: row_base[v] + n]
        out[:v] = ident[row_base[:v] + v]
        return out

    rng = np.random.default_rng(0)
    sample = np.concatenate([reps[:200], rng.choice(n, 1500, replace=False)])
    for v in sample:
        adj = ident_row(int(v)) > np.float32(thr)
        earlier_reps = adj & is_rep & (rank < rank[v])
        # representative <=> no earlier representative within the threshold
        assert is_rep[v] == (not earlier_reps.any()), int(v)


@pytest.mark.parametrize("n,L", [(1, 1), (3, 15), (7, 16), (9, 17), (100, 333), (2000, 1000)])
def test_byte_histogram(gpu, n, L):
    rng = np.random.default_rng(n + L)
    m = random_msa(rng, n, L, lower=0.1, extra=b"?.*BJZUO")
    m[rng.random((n, L)) < 0.01] = 0            # NUL bytes are data, the row padding is not
    m[rng.random((n, L)) < 0.01] = 255
    with gpu.DeviceAlignment(m) as d:
        h = d.byte_histogram()
    assert (h == np.bincount(m.reshape(-1), minlength=256).astype(np.uint64)).all()
    assert int(h.sum()) == n * L


def test_byte_histogram_gap_runs(gpu):
    """Long runs of one byte (the merged-run path) and a single hot symbol."""
    m = np.full((513, 777), ord("-"), np.uint8)
    m[::7, 100:130] = ord("A")
    with gpu.DeviceAlignment(m) as d:
        h = d.byte_histogram()
    assert (h == np.bincount(m.reshape(-1), minlength=256).astype(np.uint64)).all()


@pytest.mark.parametrize("n,L", [(1, 1), (5, 127), (33, 128), (40, 129), (1000, 1000), (64, 4100)])
def test_pinned_strided_upload_equals_staged(gpu, port, n, L):
    """tcu_msa_create_strided from page-locked memory (one linear DMA + device re-pitch)
    must give the same device rows as the staged path: same histogram, lengths, gaps."""
    import ctypes as C
    import torch
    from pytrimal_b200 import _lib
    rng = np.random.default_rng(n * 7 + L)
    for stride in (L, L + 3, L + 64):
        buf = torch.zeros(n * stride + 8, dtype=torch.uint8).pin_memory()
        view = buf.numpy()[: n * stride].reshape(n, stride)
        m = random_msa(rng, n, L)
        view[:, :L] = m
        view[:, L:] = ord("-")                      # bytes between rows must never be read as data
        lib = gpu.load()
        h = C.c_void_p()
        _lib.check(lib.tcu_msa_create_strided(C.c_void_p(buf.data_ptr()), n, L, stride, 0,
                                              C.byref(h)))
        try:
            hist = (C.c_ulonglong * 256)()
            _lib.check(lib.tcu_byte_histogram(h, hist))
            lengths = np.zeros(n, np.int32)
            _lib.check(lib.tcu_sequence_lengths(h, lengths.ctypes.data_as(C.POINTER(C.c_int))))
            gaps = np.zeros(L, np.int32)
            _lib.check(lib.tcu_gaps(h, None, gaps.ctypes.data_as(C.POINTER(C.c_int)), None, None))
        finally:
            lib.tcu_msa_destroy(h)
        assert (np.array(hist[:], np.uint64) ==
                np.bincount(m.reshape(-1), minlength=256).astype(np.uint64)).all()
        assert (lengths == port.sequence_lengths(m)).all()
        assert (gaps == port.gaps(m)[0]).all()


# ---- post-trim scans -------------------------------------------------------
@pytest.mark.parametrize("n,L", [(1, 1), (7, 15), (33, 129), (500, 777), (3000, 260)])
def test_row_residues_and_hashes(gpu, n, L):
    """tcu_row_residues (rows of Cleaner::removeAllGapsSeqsAndCols, Cleaner.cpp:1338-1370) and
    tcu_row_hashes (candidates of removeDuplicates, :1489-1509) against numpy."""
    rng = np.random.default_rng(n * 17 + L)
    m = random_msa(rng, n, L, gap=0.6)
    m[rng.random(n) < 0.1] = ord("-")                         # rows of gaps only
    dup = rng.random(n) < 0.2
    for r in np.nonzero(dup)[0]:
        m[r] = m[rng.integers(0, n)]                          # exact duplicates
    if n > 2:
        m[n - 1] = m[0]
        m[n - 1, L - 1] = ord("A") if m[0, L - 1] != ord("A") else ord("C")   # differs in the last byte only
    sr = np.arange(L, dtype=np.int32)
    sr[rng.random(L) < 0.5] = -1
    with gpu.DeviceAlignment(m) as d:
        assert (d.row_residues() == (m != ord("-")).sum(1)).all()
        assert (d.row_residues(save_res=sr) == (m[:, sr != -1] != ord("-")).sum(1)).all()
        h = d.row_hashes()
    keys = {}
    for i in range(n):
        keys.setdefault((int(h[i, 0]), int(h[i, 1])), []).append(i)
    for group in keys.values():                               # equal hashes <=> equal rows here
        for i in group[1:]:
            assert (m[i] == m[group[0]]).all()
    rows = {}
    for i in range(n):
        rows.setdefault(bytes(m[i]), []).append(i)
    assert len(rows) == len(keys)
