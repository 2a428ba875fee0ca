"""Host-side alignment container (mirror of ``pytrimal.Alignment``).

Holds the rows as one ``(nseq, ncol)`` uint8 matrix -- the layout the CUDA
library uploads -- plus trimAl's keep-masks.  Mirrors the parts of
src/pytrimal/_trimal.pyx:700-1000 and vendor/trimal Alignment.cpp that the
statistics path reads: names, sequences, ``saveSequences`` / ``saveResidues``
and the alignment type (``utils::checkAlignmentType``, utils.cpp:476-545).
"""
from __future__ import annotations

import os

import numpy as np

from . import io as _io

# SequenceTypes bit tags (vendor/trimal/include/defines.h:75-96)
NOT_DEFINED, DNA, RNA, AA, DEG = 0, 2, 4, 8, 16

_DNA = b"ACGT"
_RNA = b"ACGU"
_DEG_NT = b"ACGTURYKMSWBDHV"          # residueValues.h:40
_AA = b"ARNDCQEGHILKMFPSTWYV"          # residueValues.h:38
_AMBIG_AA = b"BJXZ*"                   # residueValues.h:43
_ALT_AA = b"UO"                        # residueValues.h:47
_GAP_SYMBOLS = b"-?."                  # utils.cpp:480


def detect_type(matrix: np.ndarray) -> int:
    """``utils::checkAlignmentType`` on byte counts instead of a per-character scan."""
    counts = np.bincount(matrix.reshape(-1), minlength=256).astype(np.int64)
    up = counts.copy()
    for c in range(ord("a"), ord("z") + 1):      # utils::toUpper folds a-z only
        up[c - 32] += up[c]
        up[c] = 0
    for g in _GAP_SYMBOLS:
        up[g] = 0
    known = set(_DNA + _RNA + _DEG_NT + _AA + _AMBIG_AA + _ALT_AA)
    for c in np.nonzero(up)[0]:
        if int(c) not in known:
            return NOT_DEFINED

    def total(chars):
        return int(sum(up[c] for c in set(chars)))

    rna, dna = total(_RNA), total(_DNA)
    deg_nt = total(set(_DEG_NT) - set(_DNA) - set(_RNA))
    aa, deg_aa, alt_aa = total(_AA), total(_AMBIG_AA), total(_ALT_AA)
    dna += deg_nt
    rna += deg_nt
    aa += deg_aa + alt_aa
    if aa > dna and aa > rna:
        return (AA | DEG) if deg_aa > 0 else AA
    if dna >= aa and dna >= rna:
        return (DNA | DEG) if deg_nt > 0 else DNA
    return (RNA | DEG) if deg_nt > 0 else RNA


class Alignment:
    """A multiple sequence alignment resident on the host."""

    def __init__(self, names, sequences, sequence_type=None):
        names = [n if isinstance(n, bytes) else str(n).encode() for n in names]
        seqs = [s if isinstance(s, (bytes, bytearray)) else str(s).encode("ascii") for s in sequences]
        if len(names) != len(seqs):
            raise ValueError(f"`Alignment` given {len(names)!r} names but {len(seqs)!r} sequences")
        self.names = names
        self.matrix = _io.to_matrix(seqs)
        # Alignment::fillMatrices (Alignment.cpp:659-664): only letters and punctuation
        if self.matrix.size:
            ok = np.zeros(256, bool)
            for c in range(256):
                ch = chr(c)
                ok[c] = c < 128 and (ch.isalpha() or (ch.isprintable() and not ch.isalnum() and not ch.isspace()))
            bad = ~ok[self.matrix]
            if bad.any():
                r, c = np.argwhere(bad)[0]
                raise ValueError(
                    f"The sequence \"{names[r].decode(errors='replace')}\" has an unknown "
                    f"({int(self.matrix[r, c])}) character")
        types = {None: NOT_DEFINED, "protein": AA, "dna": DNA, "rna": RNA}
        if sequence_type not in types:
            raise ValueError(f"invalid `sequence_type`: {sequence_type!r} "
                             "(expected one of 'protein', 'rna', 'dna' or None)")
        self._type = types[sequence_type]
        n, L = self.matrix.shape
        self.save_sequences = np.arange(n, dtype=np.int32)   # Alignment.cpp:700-713
        self.save_residues = np.arange(L, dtype=np.int32)

    @classmethod
    def load(cls, path, format=None):
        names, seqs = _io.read_alignment(os.fspath(path))
        return cls(names, seqs)

    @classmethod
    def from_matrix(cls, matrix, names=None, sequence_type=None):
        self = cls.__new__(cls)
        self.matrix = np.ascontiguousarray(matrix, np.uint8)
        n, L = self.matrix.shape
        self.names = names or [b"s%d" % i for i in range(n)]
        self._type = {None: NOT_DEFINED, "protein": AA, "dna": DNA, "rna": RNA}[sequence_type]
        self.save_sequences = np.arange(n, dtype=np.int32)
        self.save_residues = np.arange(L, dtype=np.int32)
        return self

    @property
    def nseq(self):
        return self.matrix.shape[0]

    @property
    def ncol(self):
        return self.matrix.shape[1]

    @property
    def alignment_type(self) -> int:
        if self._type == NOT_DEFINED:                      # Alignment.cpp:324-331
            self._type = detect_type(self.matrix)
        return self._type

    @property
    def indet(self) -> int:
        """Indetermination symbol of the statistics (template.h:99,221,331)."""
        return ord("X") if self.alignment_type & AA else ord("N")

    @property
    def sequences(self):
        return [bytes(r) for r in self.matrix]

    def kept(self, save_sequences=None, save_residues=None):
        """Materialise the rows/columns a pair of keep-masks retains."""
        ss = self.save_sequences if save_sequences is None else save_sequences
        sr = self.save_residues if save_residues is None else save_residues
        rows = np.nonzero(np.asarray(ss) != -1)[0]
        cols = np.nonzero(np.asarray(sr) != -1)[0]
        return [self.names[i] for i in rows], [bytes(self.matrix[i, cols]) for i in rows]
