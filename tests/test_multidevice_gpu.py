"""One process, several GPUs (tcu_set_devices / TRIMAL_CUDA_DEVICES + device = TCU_DEVICE_AUTO):
what pytrimal's platform="cuda" uses on a multi-GPU box.  Every statistic of a replicated
handle must be bit-identical to the single-device result (and to the oracle).  Skipped on a
one-GPU box; the host-side partition logic is covered on the CPU in tests/test_host.py."""
import os

import numpy as np
import pytest

from conftest import random_msa

pytestmark = pytest.mark.gpu

X = ord("X")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture()
def multi(gpu):
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    os.environ["TRIMAL_CUDA_MULTI_MIN_BYTES"] = "0"
    gpu.set_devices(list(range(n)))
    yield gpu
    gpu.set_devices(None)
    os.environ.pop("TRIMAL_CUDA_MULTI_MIN_BYTES", None)


@pytest.mark.parametrize("n,L,seed", [(700, 900, 1), (1500, 333, 2), (129, 64, 3), (5, 40, 4),
                                      (3000, 257, 5)])
def test_replicated_handle_equals_single_device(multi, port, n, L, seed):
    from pytrimal_b200.synthetic import synthetic_msa
    rng = np.random.default_rng(seed)
    m = synthetic_msa(n, L, seed) if n >= 64 else random_msa(rng, n, L)
    ss = np.arange(n, dtype=np.int32)
    ss[rng.random(n) < 0.2] = -1
    sr = np.arange(L, dtype=np.int32)
    sr[rng.random(L) < 0.3] = -1
    smx = multi.SimilarityMatrix.aa()
    with multi.DeviceAlignment(m, device=0) as one, multi.DeviceAlignment(m, device="auto") as many:
        assert one.device_count == 1 and many.device_count == multi.device_count()
        for mask in (None, ss):
            a, b = one.gaps(save_seq=mask), many.gaps(save_seq=mask)
            assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2]
        assert (many.byte_histogram() == one.byte_histogram()).all()
        assert (bits(one.spurious(0.5, indet=X)) == bits(many.spurious(0.5, indet=X))).all()
        assert (bits(one.identity(X, save_seq=ss, save_res=sr)) ==
                bits(many.identity(X, save_seq=ss, save_res=sr))).all()
        ia = one.identity(X, keep_on_device=True)
        ib = many.identity(X, keep_on_device=True)
        assert (bits(ia) == bits(ib)).all() and many.identity_resident
        assert (bits(ib) == bits(port.identity(m, X))).all()
        assert (bits(many.identity_download()) == bits(ia)).all()      # gathered on device 0
        for upper in (False, True):
            for x, y in zip(one.identity_row_stats(upper), many.identity_row_stats(upper)):
                assert (bits(x) == bits(y)).all()
        g = many.gaps()[0]
        assert (bits(one.similarity(smx, gaps=g, indet=X)[0]) ==
                bits(many.similarity(smx, gaps=g, indet=X)[0])).all()
        for thr in (0.3, 0.8):
            assert one.representatives(thr, indet=X).tolist() == \
                many.representatives(thr, indet=X).tolist()
        assert one.representatives(0.6, indet=X, save_res=sr).tolist() == \
            many.representatives(0.6, indet=X, save_res=sr).tolist()


def test_environment_variable_and_small_alignments(gpu):
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    gpu.set_devices(list(range(n)))
    try:
        m = np.full((10, 10), ord("A"), np.uint8)
        with gpu.DeviceAlignment(m, device="auto") as d:      # below the 4 MB floor: one device
            assert d.device_count == 1
        assert gpu.get_devices() == list(range(n))
    finally:
        gpu.set_devices(None)
    assert gpu.get_devices()[0] == 0
