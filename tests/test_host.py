"""Host-side logic and the C-ABI surface, CPU only (no compute calls)."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, random_msa


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "trimal_cuda.h")).read()
    declared = set(re.findall(r"\b(tcu_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in trimal_cuda.h but not exported"
    from pytrimal_b200 import _lib
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"


def test_library_links_no_torch_and_no_oracle():
    import subprocess
    import pytrimal_b200
    out = subprocess.run(["ldd", pytrimal_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "trimal_ref" not in out


def test_product_sources_do_not_reference_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "pytrimal_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
            text = open(path, errors="replace").read()
            assert "import oracle" not in text and "liboracle" not in text and \
                "trimal_ref" not in text, path


def test_no_device_fails_loudly(lib):
    import pytrimal_b200 as pb
    if pb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pb.NoDeviceError):
        pb.DeviceAlignment(np.full((3, 5), 65, np.uint8))
    assert "no CPU fallback" in lib.tcu_last_error().decode() or "CUDA" in lib.tcu_last_error().decode()


def test_row_offset_and_blocks_match_python(lib):
    from pytrimal_b200 import sharding
    for n in [0, 1, 2, 63, 64, 65, 1000, 50000, 100000]:
        assert lib.tcu_identity_row_blocks(n) == sharding.row_blocks(n)
        for b in [0, 1, 2, sharding.row_blocks(n) // 2, sharding.row_blocks(n), sharding.row_blocks(n) + 3]:
            assert lib.tcu_identity_tiles_before(n, b) == sharding.tiles_before(b, n), (n, b)
        for i in [0, 1, 63, 64, n // 2, n - 2, n - 1, n, n + 5]:
            assert lib.tcu_identity_row_offset(n, i) == sharding.row_offset(n, i), (n, i)
    n = 1000
    pos = 0
    for i in range(n - 1):
        assert sharding.row_offset(n, i) == pos
        pos += n - 1 - i
    assert sharding.row_offset(n, n - 1) == n * (n - 1) // 2
    assert lib.tcu_identity_band_rows() == sharding.ROW_BLOCK


def test_identity_tile_order_is_a_bijection(lib):
    """tcu_identity_tile: the grouped order in which a K1 launch over row-blocks [b0, b1) visits
    the pair matrix covers every tile (I, j >= 2 I) of those row-blocks exactly once, and the
    tiles that share a 64-row column block inside a group of 8 row-blocks are consecutive."""
    import ctypes as C
    bi, bj = C.c_int(), C.c_int()

    def order(nk, b0, b1):
        t0, t1 = lib.tcu_identity_tiles_before(nk, b0), lib.tcu_identity_tiles_before(nk, b1)
        out = []
        for t in range(t0, t1):
            assert lib.tcu_identity_tile(nk, b0, b1, t, C.byref(bi), C.byref(bj)) == 0
            out.append((bi.value, bj.value))
        return out

    for nk in [1, 64, 65, 128, 129, 700, 1100, 2500, 5000]:
        nsb, nb = lib.tcu_identity_row_blocks(nk), (nk + 63) // 64
        for b0, b1 in {(0, nsb), (0, max(1, nsb // 2)), (nsb // 2, nsb), (min(3, nsb - 1), nsb)}:
            if b0 >= b1:
                continue
            got = order(nk, b0, b1)
            want = {(I, j) for I in range(b0, b1) for j in range(2 * I, nb)}
            assert len(got) == len(set(got)) and set(got) == want, (nk, b0, b1)
            # inside a group, a column block is finished before the next one starts
            for g0 in range(b0, b1, 8):
                cols = [j for I, j in got if g0 <= I < min(g0 + 8, b1)]
                assert cols == sorted(cols), (nk, b0, b1, g0)
    assert lib.tcu_identity_tile(700, 0, 2, 10**9, C.byref(bi), C.byref(bj)) != 0   # out of range


def test_band_partition_balanced_and_covering():
    from pytrimal_b200.sharding import band_partition, band_slice, row_blocks, tiles_before
    for n in [100, 4097, 50000, 100000]:
        nb = row_blocks(n)
        nj = (n + 63) // 64
        for world in [1, 2, 3, 4, 8]:
            b = band_partition(n, world)
            assert b[0] == 0 and b[-1] == nb and all(x <= y for x, y in zip(b, b[1:]))
            counts = [tiles_before(b[g + 1], n) - tiles_before(b[g], n) for g in range(world)]
            # block B meets the 64-row column blocks 2B .. nj-1
            assert sum(counts) == sum(nj - 2 * B for B in range(nb))
            if nb >= 16 * world:
                assert max(counts) <= 1.05 * sum(counts) / world + nj
            total = 0
            for g in range(world):
                off, cnt = band_slice(n, b, g)
                assert off == total
                total += cnt
            assert total == n * (n - 1) // 2


def test_alignment_type_detection_matches_fixtures():
    import pytrimal_b200 as pb
    for path in glob.glob(os.path.join(GOLDEN, "*.npz")):
        g = np.load(path)
        a = pb.Alignment.from_matrix(g["matrix"])
        assert a.alignment_type == int(g["alignment_type"]), path


def test_default_matrices_match_fixtures():
    import pytrimal_b200 as pb
    for path in glob.glob(os.path.join(GOLDEN, "*.npz")):
        g = np.load(path)
        t = int(g["alignment_type"])
        if t in (8, 24, 0):
            smx = pb.SimilarityMatrix.aa()
        elif t in (2, 4):
            smx = pb.SimilarityMatrix.nt()
        else:
            smx = pb.SimilarityMatrix.nt(degenerated=True)
        assert (smx.distances.view(np.uint32) == g["dist"].view(np.uint32)).all(), path
        assert (smx.vhash == g["vhash"]).all()


def test_windows_match_oracle(port):
    from pytrimal_b200 import gaps_window, similarity_window
    rng = np.random.default_rng(1)
    for L in [8, 46, 200]:
        g = rng.integers(0, 50, L).astype(np.int32)
        mdk = rng.random(L).astype(np.float32)
        for h in [0, 1, 2, L // 4]:
            assert (gaps_window(g, h) == port.gaps_window(g, h)).all()
            assert (similarity_window(mdk, h).view(np.uint32) ==
                    port.similarity_window(mdk, h).view(np.uint32)).all()
        with pytest.raises(ValueError):
            gaps_window(g, L // 4 + 1)      # test_manual_trimmer.py:49-52 (window too large)
        with pytest.raises(ValueError):
            similarity_window(mdk, L // 4 + 1)


def test_alignment_constructor_errors():
    import pytrimal_b200 as pb
    with pytest.raises(ValueError):
        pb.Alignment([b"a"], [b"MKK", b"MKA"])
    with pytest.raises(ValueError):
        pb.Alignment([b"a", b"b"], [b"MKK", b"MK"])
    with pytest.raises(ValueError):
        pb.Alignment([b"a", b"b"], [b"MK1", b"MKA"])     # digit: Alignment.cpp:659-664
    with pytest.raises(ValueError):
        pb.Alignment([b"a"], [b"MKK"], sequence_type="nope")
    a = pb.Alignment([b"a", b"b"], ["MKKBO", "MKKAY"])
    assert a.alignment_type & 8 and chr(a.indet) == "X"


def test_readers(tmp_path):
    from pytrimal_b200 import io as tio
    p = tmp_path / "a.fasta"
    p.write_bytes(b"[junk line]\n>s1 desc\nAC-\nGT\n>s2\nACCGT\n")
    names, seqs = tio.read_alignment(str(p))
    assert names == [b"s1", b"s2"] and seqs == [b"AC-GT", b"ACCGT"]
    q = tmp_path / "a.clw"
    q.write_bytes(b"CLUSTAL W\n\ns1   AC-\ns2   ACC\n     ** \n\ns1   GT\ns2   GT\n")
    names, seqs = tio.read_alignment(str(q))
    assert names == [b"s1", b"s2"] and seqs == [b"AC-GT", b"ACCGT"]


def test_synthetic_generator_is_deterministic():
    from pytrimal_b200.synthetic import synthetic_msa
    a = synthetic_msa(300, 200, 7)
    b = synthetic_msa(300, 200, 7)
    assert (a == b).all() and a.shape == (300, 200)
    assert 0.1 < (a == ord("-")).mean() < 0.5


def test_shard_helpers_cover_and_balance(lib):
    """tcu_shard_range / tcu_shard_blocks (the partition the *_all calls use): contiguous,
    covering, and equal to the Python partition bench.py uses."""
    import pytrimal_b200 as pb
    from pytrimal_b200 import sharding
    for total, gran in [(0, 1), (5, 1), (157, 1), (100000, 1), (1000, 32)]:
        for world in (1, 2, 3, 8):
            cuts = [pb.shard_range(total, gran, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= gran
    for n in (1, 100, 130, 257, 700, 12345, 50000):
        for world in (1, 2, 3, 8):
            bounds = sharding.band_partition(n, world)
            cuts = [pb.shard_blocks(n, r, world) for r in range(world)]
            assert [c[0] for c in cuts] + [cuts[-1][1]] == bounds
    with pytest.raises(ValueError):
        pb.shard_range(10, 1, 3, 2)


def test_comm_calls_fail_loudly_without_device(lib):
    """No GPU here: creating a communicator must fail (NCCL or device), never pretend."""
    import ctypes as C
    import pytrimal_b200 as pb
    if pb.device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.tcu_comm_create(C.create_string_buffer(128), 0, 1, 0, C.byref(h))
    assert rc < 0 and not h.value


def test_cluster_order_matches_oracle(lib, port):
    """tcu_cluster_order is host-only: the reference's non-stable quicksort permutation
    (utils.cpp:246-273) walked from the end, against the oracle's step-by-step restatement."""
    import pytrimal_b200 as pb
    rng = np.random.default_rng(3)
    cases = [np.array([], np.int32), np.array([5], np.int32), np.array([3, 3], np.int32),
             np.arange(200, dtype=np.int32), np.arange(200, dtype=np.int32)[::-1].copy(),
             np.full(300, 7, np.int32), rng.integers(0, 10, 1000).astype(np.int32),
             rng.integers(500, 1000, 50000).astype(np.int32)]
    for lengths in cases:
        got = pb.cluster_order(lengths)
        want = port.cluster_order(lengths)
        assert got.tolist() == want.tolist()
        assert sorted(got.tolist()) == list(range(len(lengths)))
        assert (np.diff(lengths[got]) <= 0).all()     # longest first
    # iterative: a sorted input must not overflow the stack where the recursion is n deep
    big = np.arange(200000, dtype=np.int32)
    assert pb.cluster_order(big)[0] == 199999
