"""Seeded synthetic protein MSAs for parity tests and benchmarks.

Families of sequences derived from a common root so that identity thresholds
are non-degenerate; gap runs with geometric lengths plus ragged ends; a little
'X'.  Deterministic for a given (nseq, ncol, seed).
"""
from __future__ import annotations

import numpy as np

AA = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", np.uint8)

#: the BASELINE.json shapes
CONFIGS = {
    "C2": (1000, 2000, 2),
    "C3": (10000, 5000, 3),
    "C4": (50000, 1000, 4),
    "C5": (100000, 2000, 5),
}


def _substitute(rng, seq, rate):
    """Replace a `rate` fraction of positions by uniformly random residues."""
    out = np.broadcast_to(seq, rate.shape[:1] + seq.shape[-1:]).copy() if seq.ndim == 1 else seq.copy()
    mask = rng.random(out.shape, dtype=np.float32) < rate[:, None]
    out[mask] = AA[rng.integers(0, 20, int(mask.sum()), dtype=np.uint8)]
    return out


def synthetic_msa(nseq, ncol, seed, lowercase=False, chunk=4096, near_duplicates=0.1) -> np.ndarray:
    """(nseq, ncol) uint8 matrix.  A `near_duplicates` fraction of the rows are
    copies of the row above them with 1-5 % substitutions (same gap pattern), so
    that identity thresholds around 0.8 select something."""
    rng = np.random.default_rng(seed)
    root = AA[rng.integers(0, 20, ncol)]
    nfam = max(1, nseq // 100)
    anc = _substitute(rng, np.tile(root, (nfam, 1)), np.full(nfam, 0.30, np.float32))
    fam = rng.integers(0, nfam, nseq)
    out = np.empty((nseq, ncol), np.uint8)
    pos = np.arange(ncol, dtype=np.int32)[None, :]
    for r0 in range(0, nseq, chunk):
        r1 = min(nseq, r0 + chunk)
        m = r1 - r0
        rows = _substitute(rng, anc[fam[r0:r1]], rng.uniform(0.02, 0.35, m).astype(np.float32))
        # gap runs: start probability 0.02, geometric length (mean 8)
        start = rng.random((m, ncol), dtype=np.float32) < 0.02
        length = rng.geometric(1.0 / 8.0, (m, ncol)).astype(np.int32)
        reach = np.where(start, pos + length, 0)
        np.maximum.accumulate(reach, axis=1, out=reach)
        gap = reach > pos
        # ragged ends
        lead = rng.integers(0, ncol // 10 + 1, m)[:, None]
        trail = rng.integers(0, ncol // 10 + 1, m)[:, None]
        gap |= (pos < lead) | (pos >= ncol - trail)
        rows[rng.random((m, ncol), dtype=np.float32) < 0.002] = ord("X")
        if lowercase:
            low = rng.random((m, ncol), dtype=np.float32) < 0.05
            rows[low] |= 0x20
        rows[gap] = ord("-")
        dups = np.nonzero(rng.random(m) < near_duplicates)[0]
        for r in dups[dups > 0]:
            rows[r] = rows[r - 1]
            res = np.nonzero(rows[r] != ord("-"))[0]
            hit = res[rng.random(res.size) < rng.uniform(0.01, 0.05)]
            rows[r, hit] = AA[rng.integers(0, 20, hit.size)]
        out[r0:r1] = rows
    return out


def config(name):
    n, L, seed = CONFIGS[name]
    return synthetic_msa(n, L, seed)
