"""Partition of the pair matrix across GPUs (one process per GPU).

The packed identity array is row-major over pairs (i, j>i), so a contiguous
range of rows is a contiguous slice of the array.  Rows are grouped in
row-blocks of 128 (the kernel's tile height, tcu_identity_band_rows()); rank g
gets the row-blocks [bounds[g], bounds[g+1]) chosen so that every rank owns
(nearly) the same number of 128x64 tiles, i.e. the same work.  No collective is needed to
compute; assembling the full array is one all-gather of the slices
(SURVEY 8e).
"""
from __future__ import annotations

ROW_BLOCK = 128
J_BLOCK = 64


def row_blocks(kept_rows: int) -> int:
    return (kept_rows + ROW_BLOCK - 1) // ROW_BLOCK


def tiles_before(block: int, kept_rows: int) -> int:
    """128x64 tiles in row-blocks < block: block B pairs with the 64-row column
    blocks 2B .. nj-1 (== tcu_identity_tiles_before)."""
    nj = (kept_rows + J_BLOCK - 1) // J_BLOCK
    b = max(0, min(block, row_blocks(kept_rows)))
    return b * nj - b * (b - 1)


def band_partition(kept_rows: int, world: int):
    """Row-block boundaries, len world+1, monotone, bounds[0]=0, bounds[-1]=row_blocks."""
    nb = row_blocks(kept_rows)
    total = tiles_before(nb, kept_rows)
    bounds = [0]
    for g in range(1, world):
        target = total * g / world
        b = bounds[-1]
        while b < nb and tiles_before(b, kept_rows) < target:
            b += 1
        bounds.append(b)
    bounds.append(nb)
    return bounds


def row_offset(kept_rows: int, row: int) -> int:
    """Packed offset of pair (row, row+1); == total pairs for row >= kept_rows-1."""
    n = max(kept_rows, 0)
    if n < 2:
        return 0
    r = min(max(row, 0), n - 1)
    return r * n - r * (r + 1) // 2


def band_slice(kept_rows: int, bounds, rank: int):
    """(offset, count) of rank's slice of the packed identity array."""
    lo = row_offset(kept_rows, ROW_BLOCK * bounds[rank])
    hi = row_offset(kept_rows, min(ROW_BLOCK * bounds[rank + 1], kept_rows))
    return lo, hi - lo
