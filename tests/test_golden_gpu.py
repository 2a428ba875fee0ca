"""GPU vs the fixtures produced by the real reference (tests/golden/*.npz)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
IDS = [os.path.basename(f)[:-4] for f in FIXTURES]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_cuda_matches_reference_fixture(gpu, path):
    g = np.load(path)
    a = gpu.Alignment.from_matrix(g["matrix"])
    assert a.alignment_type == int(g["alignment_type"])
    indet = a.indet
    smx = None
    t = a.alignment_type
    if t in (8, 24, 0):
        smx = gpu.SimilarityMatrix.aa()
    elif t in (2, 4):
        smx = gpu.SimilarityMatrix.nt()
    else:
        smx = gpu.SimilarityMatrix.nt(degenerated=True)
    with gpu.DeviceAlignment(a) as d:
        gaps, hist, mx = d.gaps()
        assert (gaps == g["gaps"]).all() and (hist == g["gaps_hist"]).all() and mx == int(g["gaps_max"])
        if "gaps_w3" in g:
            assert (gpu.gaps_window(gaps, 3) == g["gaps_w3"]).all()
        ident = d.identity(keep_on_device=True)
        assert (bits(ident) == bits(g["identity"])).all()
        for ov in (50, 80):
            assert (bits(d.spurious(ov / 100)) == bits(g[f"spurious_{ov}"])).all()
        if "similarity_error" in g:
            with pytest.raises(gpu.SymbolError):
                d.similarity(smx, gaps=gaps)
        else:
            mdk, _, _ = d.similarity(smx, gaps=gaps)
            assert (bits(mdk) == bits(g["mdk"])).all()
            np.testing.assert_allclose(mdk, g["mdk"], rtol=1e-5, atol=0)
            if "mdk_w1" in g:
                assert (bits(gpu.similarity_window(mdk, 1)) == bits(g["mdk_w1"])).all()
