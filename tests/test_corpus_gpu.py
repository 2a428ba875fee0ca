"""The reference's whole bundled corpus (vendor/trimal/dataset, 86 aligned inputs up to
3583 x 7287) through the CUDA path: the four statistics and the Cleaner walks through the C
ABI, and the seven automatic methods of scripts/generate_trimmed_msas.sh through pytrimal's
own trim() with platform="cuda" -- against what the unmodified reference (AVX2 platform)
produced for the same inputs (tests/golden/make_golden_corpus.py -> tests/golden/corpus/).
Arrays are compared through SHA-256 of their bytes, keep-masks bit for bit."""
import json
import lzma
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, GOLDEN)
from digest import sha256_hex  # noqa: E402

pytestmark = pytest.mark.gpu

CORPUS = os.path.join(GOLDEN, "corpus")
PKG = os.path.join(ROOT, "integration", "_build", "pkg")
INDEX = {}
if os.path.exists(os.path.join(CORPUS, "INDEX.json")):
    with open(os.path.join(CORPUS, "INDEX.json")) as _f:
        INDEX = json.load(_f)["inputs"]
STEMS = sorted(INDEX)


def load_matrix(stem):
    e = INDEX[stem]
    with open(os.path.join(CORPUS, stem + ".xz"), "rb") as f:
        m = np.frombuffer(lzma.decompress(f.read()), np.uint8).reshape(e["shape"]).copy()
    assert sha256_hex(m) == e["matrix_sha256"]
    return m


def test_corpus_is_complete():
    if not INDEX:
        pytest.skip("tests/golden/corpus not generated")
    assert len(INDEX) >= 80                                   # 86 aligned inputs upstream
    assert max(e["shape"][0] for e in INDEX.values()) >= 3583  # example.014, pytrimal's own bench input
    for stem in ("example.014.AA.EggNOG.COG0591", "example.028.AA.bctoNOG.ENOG41099PA"):
        assert stem in INDEX                                  # upstream's skip list is NOT skipped


@pytest.mark.parametrize("stem", STEMS)
def test_corpus_statistics(gpu, stem):
    e = INDEX[stem]
    m = load_matrix(stem)
    a = gpu.Alignment.from_matrix(m)
    assert a.alignment_type == e["type"]
    t = e["type"]
    if t in (8, 24, 0):
        smx = gpu.SimilarityMatrix.aa()
    elif t in (2, 4):
        smx = gpu.SimilarityMatrix.nt()
    else:
        smx = gpu.SimilarityMatrix.nt(degenerated=True)
    with gpu.DeviceAlignment(a) as d:
        gaps = d.gaps()[0]
        assert sha256_hex(gaps) == e["gaps_sha256"]
        for ov in (50, 80):
            assert sha256_hex(d.spurious(ov / 100)) == e[f"spurious_{ov}_sha256"]
        if "repr_75_sha256" in e:
            assert sha256_hex(d.representatives(0.75)) == e["repr_75_sha256"]
        ident = d.identity(keep_on_device=True)
        assert ident.size == e["identity_count"] and sha256_hex(ident) == e["identity_sha256"]
        if e["mdk_sha256"] is None:
            with pytest.raises(gpu.SymbolError):
                d.similarity(smx, gaps=gaps)
        else:
            assert sha256_hex(d.similarity(smx, gaps=gaps)[0]) == e["mdk_sha256"]
        if "select_method" in e:
            assert d.select_method()[0] == {1: "gappyout", 2: "strict"}[e["select_method"]]


@pytest.fixture(scope="module")
def pytrimal(gpu):
    if not os.path.isdir(os.path.join(PKG, "pytrimal")):
        pytest.skip("integration/_build/pkg not present (built where /root/reference exists)")
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    import pytrimal
    assert pytrimal._trimal._CUDA_RUNTIME_SUPPORT
    return pytrimal


@pytest.mark.parametrize("stem", STEMS)
def test_corpus_automatic_methods(gpu, pytrimal, stem):
    e = INDEX[stem]
    m = load_matrix(stem)
    n, L = m.shape
    names = [b"s%d" % i for i in range(n)]
    ali = pytrimal.Alignment(names, [bytes(r) for r in m])
    for method, want in e["methods"].items():
        try:
            out = pytrimal.AutomaticTrimmer(method, platform="cuda").trim(ali)
        except Exception:
            out = None
        if want.get("error"):
            assert out is None or len(out.sequences) == 0 or len(out.sequences[0]) == 0, method
            continue
        ks = np.unpackbits(np.frombuffer(bytes.fromhex(want["seq"]), np.uint8))[:n].astype(bool)
        kr = np.unpackbits(np.frombuffer(bytes.fromhex(want["res"]), np.uint8))[:L].astype(bool)
        if not kr.any() or not ks.any():
            assert out is None or len(out.sequences) == 0 or len(out.sequences[0]) == 0, method
            continue
        assert out is not None, method
        rows, cols = np.nonzero(ks)[0], np.nonzero(kr)[0]
        assert list(out.names) == [names[i] for i in rows], method
        got = list(out.sequences)
        sub = m[np.ix_(rows, cols)]
        assert len(got) == len(rows) and all(got[k].encode() == bytes(sub[k]) for k in range(len(rows))), method
