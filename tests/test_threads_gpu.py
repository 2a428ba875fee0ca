"""Several host threads, one handle each, on the same device at the same time.

include/trimal_cuda.h's threading contract: a handle is used by one thread at a
time, different handles are independent.  The clustering walk is a persistent kernel whose
CTAs wait for each other (launched cooperatively, so that all of them are resident), the
identity kernel is persistent too, and the buffer / stream / staging pools are shared by all
threads: this runs the whole statistic set from four threads concurrently, repeatedly, and
compares every result with what a single thread gets.  A deadlock shows up as the timeout.
"""
import threading

import numpy as np
import pytest

from pytrimal_b200.synthetic import synthetic_msa

pytestmark = pytest.mark.gpu
X = ord("X")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def results(gpu, m):
    smx = gpu.SimilarityMatrix.aa()
    with gpu.DeviceAlignment(m) as d:
        g = d.gaps()[0]
        out = {"gaps": g, "spurious": d.spurious(0.5, indet=X), "reps": d.representatives(0.8, indet=X),
               "reps60": d.representatives(0.6, indet=X)}
        d.identity_on_device(X)
        out["rowstats"] = np.concatenate(d.identity_row_stats())
        out["mdk"] = d.similarity(smx, gaps=g, indet=X)[0]
    return out


def same(a, b):
    for k in a:
        x, y = a[k], b[k]
        if x.dtype == np.float32:
            if not (bits(x) == bits(y)).all():
                return False
        elif x.tolist() != y.tolist():
            return False
    return True


@pytest.mark.timeout(300)
def test_four_threads_one_handle_each(gpu):
    shapes = [(3000, 300, 1), (2500, 257, 2), (1100, 700, 3), (4100, 130, 4)]
    inputs = [synthetic_msa(n, L, seed) for n, L, seed in shapes]
    want = [results(gpu, m) for m in inputs]
    errors = []

    def worker(k):
        try:
            for _ in range(4):
                if not same(results(gpu, inputs[k]), want[k]):
                    errors.append(f"thread {k}: result differs from the single-threaded run")
                    return
        except Exception as exc:  # noqa: BLE001 -- reported by the main thread
            errors.append(f"thread {k}: {exc!r}")

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(inputs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    assert not any(t.is_alive() for t in threads), "a thread did not finish (deadlock?)"
    assert not errors, errors
