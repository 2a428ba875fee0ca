"""Similarity matrices for the column-similarity statistic.

Mirrors ``pytrimal.SimilarityMatrix`` (src/pytrimal/_trimal.pyx:1867-2060) and
trimAl's ``statistics::similarityMatrix`` (vendor/trimal/source/Statistics/
similarityMatrix.cpp): an alphabet, a square score matrix, the letter hash
``vhash`` and the Euclidean column-distance matrix ``distMat`` that the
statistic actually consumes.  The distance is accumulated in fp32 exactly like
similarityMatrix.cpp:259-271 / _trimal.pyx:1988-1997 so the floats match.
"""
from __future__ import annotations

import numpy as np

AA_ALPHABET = "ARNDCQEGHILKMFPSTWYV"   # residueValues.h:38
NT_ALPHABET = "ACGTU"                  # residueValues.h:36
NT_DEG_ALPHABET = "ACGTURYKMSWBDHV"    # residueValues.h:40

# BLOSUM62 in AA_ALPHABET order (the standard NCBI table; residueValues.h:82-107)
_BLOSUM62 = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4
"""


def _nt_degenerated():
    # residueValues.h:56-76: identity on ACGTU; IUPAC rows spread 1/4 or 1/6
    # over the bases they stand for plus themselves.
    n = len(NT_DEG_ALPHABET)
    m = np.zeros((n, n), np.float32)
    idx = {c: i for i, c in enumerate(NT_DEG_ALPHABET)}
    for c in "ACGTU":
        m[idx[c], idx[c]] = 1.0
    two = {"R": "GA", "Y": "CTU", "K": "GTU", "M": "AC", "S": "GC", "W": "ATU"}
    three = {"B": "GCTU", "D": "GATU", "H": "ACTU", "V": "GCA"}
    for c, bases in two.items():
        for b in bases + c:
            m[idx[c], idx[b]] = np.float32(1 / 4.)
    for c, bases in three.items():
        for b in bases + c:
            m[idx[c], idx[b]] = np.float32(1 / 6.)
    return m


def distance_matrix(sim: np.ndarray) -> np.ndarray:
    """Euclidean distance between matrix columns, fp32 accumulation in k order
    (similarityMatrix.cpp:262-270)."""
    sim = np.asarray(sim, np.float32)
    n = sim.shape[0]
    dist = np.zeros((n, n), np.float32)
    for j in range(n):
        for i in range(n):
            if i == j or dist[i, j] != 0.0:
                continue
            s = np.float32(0.0)
            for k in range(n):
                d = np.float32(sim[k, j] - sim[k, i])
                s = np.float32(s + np.float32(d * d))
            s = np.float32(np.sqrt(np.float64(s)))
            dist[i, j] = s
            dist[j, i] = s
    return dist


class SimilarityMatrix:
    """A similarity matrix for biological sequence characters.

    ``SimilarityMatrix(alphabet, matrix)`` / ``.aa()`` / ``.nt(degenerated=False)``
    follow pytrimal's constructors (_trimal.pyx:1893-1997).
    """

    def __init__(self, alphabet, matrix):
        alphabet = str(alphabet)
        matrix = np.asarray(matrix, np.float32)
        if matrix.ndim != 2 or matrix.shape[0] != matrix.shape[1] or matrix.shape[0] != len(alphabet):
            raise ValueError("`matrix` must be a square matrix of the alphabet's length")
        if len(alphabet) > 28:
            raise ValueError("Cannot use alphabet of more than 28 symbols")
        if not alphabet.isupper():
            raise ValueError("Invalid symbols in alphabet (expected uppercase letters)")
        self.alphabet = alphabet
        self.matrix = matrix.copy()
        self.vhash = np.full(28, -1, np.int32)          # TAMABC, similarityMatrix.cpp:34
        for i, c in enumerate(alphabet):
            if not ("A" <= c <= "Z"):
                raise ValueError(f"Invalid symbol in alphabet: {c!r}")
            self.vhash[ord(c) - ord("A")] = i
        self.distances = distance_matrix(self.matrix)

    def __len__(self):
        return len(self.alphabet)

    @classmethod
    def aa(cls):
        """BLOSUM62 (``similarityMatrix::defaultAASimMatrix``)."""
        return cls(AA_ALPHABET, np.array(_BLOSUM62.split(), np.float32).reshape(20, 20))

    @classmethod
    def nt(cls, degenerated=False):
        """Nucleotide identity matrix (``defaultNTSimMatrix`` / ``defaultNTDegeneratedSimMatrix``)."""
        if degenerated:
            return cls(NT_DEG_ALPHABET, _nt_degenerated())
        return cls(NT_ALPHABET, np.eye(5, dtype=np.float32))

    def similarity(self, a, b):
        ia, ib = self.vhash[ord(a.upper()) - 65], self.vhash[ord(b.upper()) - 65]
        if ia < 0 or ib < 0:
            raise ValueError(f"Invalid symbol: {a if ia < 0 else b!r}")
        return float(self.matrix[ia, ib])
