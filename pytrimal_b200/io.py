"""Minimal alignment readers (FASTA and Clustal) for the host side.

The reference reads alignments through trimAl's format handlers
(vendor/trimal/source/FormatHandling/*, out of scope here); these two readers
cover the formats the reference's own fixtures use.  Sequences are returned as
raw bytes -- the statistics compare raw bytes, so no case folding
or symbol translation happens here.
"""
from __future__ import annotations

import numpy as np


def _parse_fasta(lines):
    names, seqs, cur = [], [], None
    for line in lines:
        line = line.rstrip(b"\r\n")
        if not line:
            continue
        if line.startswith(b">"):
            names.append(line[1:].split()[0] if line[1:].split() else b"")
            cur = []
            seqs.append(cur)
        elif cur is not None:
            cur.append(line.replace(b" ", b""))
    return names, [b"".join(s) for s in seqs]


def _parse_clustal(lines):
    names, chunks = [], {}
    for line in lines[1:]:
        line = line.rstrip(b"\r\n")
        if not line or line[:1] in (b" ", b"\t"):
            continue
        parts = line.split()
        if len(parts) < 2:
            continue
        name, frag = parts[0], parts[1]
        if name not in chunks:
            chunks[name] = []
            names.append(name)
        chunks[name].append(frag)
    return names, [b"".join(chunks[n]) for n in names]


def read_alignment(path):
    """Return (names: list[bytes], sequences: list[bytes])."""
    with open(path, "rb") as f:
        lines = f.readlines()
    first = next((l for l in lines if l.strip()), b"")
    if first.upper().startswith(b"CLUSTAL") or first.upper().startswith(b"MUSCLE"):
        return _parse_clustal(lines)
    # FASTA; like trimAl's reader, tolerate junk before the first '>' record
    start = next((i for i, l in enumerate(lines) if l.startswith(b">")), None)
    if start is not None:
        return _parse_fasta(lines[start:])
    raise ValueError(f"unsupported alignment format: {path}")


def to_matrix(sequences) -> np.ndarray:
    """Stack equal-length byte sequences into a (nseq, ncol) uint8 matrix."""
    if len(sequences) == 0:
        return np.zeros((0, 0), np.uint8)
    n = len(sequences[0])
    for i, s in enumerate(sequences):
        if len(s) != n:
            raise ValueError(f"Sequence length mismatch in sequence {i}: {len(s)} != {n}")
    return np.frombuffer(b"".join(bytes(s) for s in sequences), np.uint8).reshape(len(sequences), n).copy()
