"""Digests of large result arrays (TEST INFRASTRUCTURE).

The full-size configurations of BASELINE.json produce results that cannot be
committed (the packed identity array of C4 is 5 GB), so the fixtures hold
digests of what the real reference computed (tests/golden/make_golden_full.py)
and the GPU tests / bench.py recompute the same digests from the CUDA results.

Two kinds:

* ``sha256_hex(a)`` -- SHA-256 of the raw bytes, for whole arrays.
* ``block_sums(identities, nseq, rows)`` -- for the packed upper-triangular
  identity array, per band of ``rows`` consecutive first-rows (128 = the row-block
  of the multi-GPU band partition): the plain sum and a position-weighted sum of
  the fp32 BIT PATTERNS, both modulo 2**64.  Any contiguous band of row-blocks
  (what one rank of an N-GPU run owns) is checked by adding the blocks' entries,
  whatever N is.  The weight of element k (global packed offset) is
  ``(k mod 65521) + 1`` so that a permutation inside a band changes the digest.
"""
from __future__ import annotations

import hashlib

import numpy as np

WEIGHT_MOD = 65521
MASK64 = (1 << 64) - 1


def sha256_hex(a) -> str:
    a = np.ascontiguousarray(a)
    h = hashlib.sha256()
    mv = memoryview(a).cast("B")
    step = 1 << 28
    for o in range(0, len(mv), step):
        h.update(mv[o:o + step])
    return h.hexdigest()


def row_offset(nseq: int, i: int) -> int:
    """Packed offset of pair (i, i+1): rows before i hold nseq-1-r pairs each."""
    i = min(max(i, 0), max(nseq - 1, 0))
    return i * nseq - i * (i + 1) // 2


def range_sums(bits_u32: np.ndarray, first_offset: int):
    """(sum, weighted sum) mod 2**64 of a slice of bit patterns whose element 0 has
    global packed offset `first_offset`."""
    s = 0
    w = 0
    step = 1 << 24
    for o in range(0, bits_u32.size, step):
        v = bits_u32[o:o + step].astype(np.uint64)
        k = np.arange(first_offset + o, first_offset + o + v.size, dtype=np.uint64)
        wt = k % np.uint64(WEIGHT_MOD) + np.uint64(1)
        s = (s + int(v.sum(dtype=np.uint64))) & MASK64
        with np.errstate(over="ignore"):
            w = (w + int((v * wt).sum(dtype=np.uint64))) & MASK64
    return s, w


def block_sums(identities: np.ndarray, nseq: int, rows: int = 128):
    """Per row-block digests of a COMPLETE packed identity array: two uint64 arrays
    (sum, weighted sum) of length ceil(nseq / rows)."""
    bits = np.ascontiguousarray(identities, np.float32).view(np.uint32)
    nb = (nseq + rows - 1) // rows
    s = np.zeros(nb, np.uint64)
    w = np.zeros(nb, np.uint64)
    for b in range(nb):
        lo, hi = row_offset(nseq, b * rows), row_offset(nseq, min((b + 1) * rows, nseq))
        if b == nb - 1:
            hi = bits.size
        a, c = range_sums(bits[lo:hi], lo)
        s[b], w[b] = a, c
    return s, w


def band_expected(s: np.ndarray, w: np.ndarray, block_begin: int, block_end: int):
    """Digest a rank owning row-blocks [block_begin, block_end) must reproduce."""
    a = int(s[block_begin:block_end].sum(dtype=np.uint64)) & MASK64 if block_end > block_begin else 0
    with np.errstate(over="ignore"):
        b = int(w[block_begin:block_end].sum(dtype=np.uint64)) & MASK64 if block_end > block_begin else 0
    return a, b
