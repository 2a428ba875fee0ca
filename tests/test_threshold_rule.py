"""The integer form of "float32(hit) / float32(dst) > threshold" (tcu_threshold_rule, host only).

tcu_representatives thresholds inside the identity kernel without forming the ratios; the rule
must agree with the reference's comparison (Cleaner.cpp:1435-1440 on the floats of
template.h:427-434) for EVERY pair of counts and EVERY float threshold.  Checked here, without a
GPU, against numpy's correctly rounded float32 division: exhaustively for small denominators, on
random large ones, and at the thresholds where a comparison can flip -- the quotients
themselves and their float neighbours.
"""
import numpy as np
import pytest

import pytrimal_b200 as pb


def rule_bits(thr, h, d):
    mode, mul, shift = pb.threshold_rule(thr)
    if mode == 0:
        return np.zeros(h.shape, bool)
    if mode == 1:
        return np.ones(h.shape, bool)
    q = (np.uint64(mul) * d.astype(np.uint64)) >> np.uint64(shift)
    return h.astype(np.uint64) > q


def reference_bits(thr, h, d):
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.where(d == 0, np.float32(0), h.astype(np.float32) / d.astype(np.float32)).astype(np.float32)
    return v > np.float32(thr)


def all_pairs(dmax):
    d = np.repeat(np.arange(dmax + 1), np.arange(dmax + 1) + 1)
    h = np.concatenate([np.arange(k + 1) for k in range(dmax + 1)])
    return h.astype(np.int64), d.astype(np.int64)


def test_rule_equals_division_exhaustive_small():
    h, d = all_pairs(400)
    with np.errstate(divide="ignore", invalid="ignore"):
        quot = np.unique(np.where(d == 0, 0, h / np.maximum(d, 1)).astype(np.float32))
    rng = np.random.default_rng(1)
    pick = np.concatenate([quot[rng.integers(0, len(quot), 60)], np.float32([0, 0.5, 0.8, 0.25, 1 / 3, 2 / 3, 1])])
    thrs = []
    for t in pick:
        thrs += [t, np.nextafter(t, np.float32(-1)), np.nextafter(t, np.float32(2))]
    thrs += [np.float32(x) for x in (-1.0, -0.0, 0.0, 1e-45, 1e-30, 1e-8, 5.9e-8, 0.999999, 1.0, 1.0000001,
                                      2.0, np.inf, -np.inf, np.nan, -1e-45)]
    for t in thrs:
        assert (rule_bits(t, h, d) == reference_bits(t, h, d)).all(), float(t)


def test_rule_equals_division_large_counts():
    rng = np.random.default_rng(2)
    d = rng.integers(1, 1 << 24, 200_000)
    h = (d * rng.random(len(d))).astype(np.int64)
    h[:1000] = d[:1000]                      # identity 1
    h[1000:2000] = 0
    v = (h.astype(np.float32) / d.astype(np.float32)).astype(np.float32)
    for t in np.concatenate([v[rng.integers(0, len(v), 40)], rng.random(40).astype(np.float32)]):
        for tt in (t, np.nextafter(t, np.float32(-1)), np.nextafter(t, np.float32(2))):
            assert (rule_bits(tt, h, d) == reference_bits(tt, h, d)).all(), float(tt)


@pytest.mark.parametrize("thr", [0.8, 0.5, 0.3, 0.95])
def test_rule_shape(thr):
    mode, mul, shift = pb.threshold_rule(thr)
    assert mode == 2 and mul % 2 == 1 and mul < (1 << 25) and 25 <= shift <= 63
