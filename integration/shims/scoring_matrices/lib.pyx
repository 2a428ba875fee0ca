# cython: language_level=3
"""See scoring_matrices/__init__.py: offline stand-in for scoring_matrices.lib."""
from libc.stdlib cimport calloc, free

# BLOSUM62 over the 20 standard residues, in trimAl's order (the values pytrimal's
# default matrix ends up with after ``shuffle`` to ``aminoAcidResidues``)
_BLOSUM62_ALPHABET = "ARNDCQEGHILKMFPSTWYV"
_BLOSUM62 = [
    [ 4, -1, -2, -2,  0, -1, -1,  0, -2, -1, -1, -1, -1, -2, -1,  1,  0, -3, -2,  0],
    [-1,  5,  0, -2, -3,  1,  0, -2,  0, -3, -2,  2, -1, -3, -2, -1, -1, -3, -2, -3],
    [-2,  0,  6,  1, -3,  0,  0,  0,  1, -3, -3,  0, -2, -3, -2,  1,  0, -4, -2, -3],
    [-2, -2,  1,  6, -3,  0,  2, -1, -1, -3, -4, -1, -3, -3, -1,  0, -1, -4, -3, -3],
    [ 0, -3, -3, -3,  9, -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1],
    [-1,  1,  0,  0, -3,  5,  2, -2,  0, -3, -2,  1,  0, -3, -1,  0, -1, -2, -1, -2],
    [-1,  0,  0,  2, -4,  2,  5, -2,  0, -3, -3,  1, -2, -3, -1,  0, -1, -3, -2, -2],
    [ 0, -2,  0, -1, -3, -2, -2,  6, -2, -4, -4, -2, -3, -3, -2,  0, -2, -2, -3, -3],
    [-2,  0,  1, -1, -3,  0,  0, -2,  8, -3, -3, -1, -2, -1, -2, -1, -2, -2,  2, -3],
    [-1, -3, -3, -3, -1, -3, -3, -4, -3,  4,  2, -3,  1,  0, -3, -2, -1, -3, -1,  3],
    [-1, -2, -3, -4, -1, -2, -3, -4, -3,  2,  4, -2,  2,  0, -3, -2, -1, -2, -1,  1],
    [-1,  2,  0, -1, -3,  1,  1, -2, -1, -3, -2,  5, -1, -3, -1,  0, -1, -3, -2, -2],
    [-1, -1, -2, -3, -1,  0, -2, -3, -2,  1,  2, -1,  5,  0, -2, -1, -1, -1, -1,  1],
    [-2, -3, -3, -3, -2, -3, -3, -3, -1,  0,  0, -3,  0,  6, -4, -2, -2,  1,  3, -1],
    [-1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4,  7, -1, -1, -4, -3, -2],
    [ 1, -1,  1,  0, -1,  0,  0,  0, -1, -2, -2,  0, -1, -2, -1,  4,  1, -3, -2, -2],
    [ 0, -1,  0, -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1,  1,  5, -2, -2,  0],
    [-3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1,  1, -4, -3, -2, 11,  2, -3],
    [-2, -2, -2, -3, -2, -1, -2, -3,  2, -1, -1, -2, -1,  3, -3, -2, -2,  2,  7, -1],
    [ 0, -3, -3, -3, -1, -2, -2, -3, -3,  3,  1, -2,  1, -1, -2, -2,  0, -3, -1,  4],
]


cdef class ScoringMatrix:

    def __cinit__(self):
        self._size = 0
        self._nitems = 0
        self._data = NULL
        self._matrix = NULL
        self.alphabet = ""
        self.name = None

    def __dealloc__(self):
        free(self._data)
        free(self._matrix)

    cdef int _allocate(self, size_t size) except 1:
        cdef size_t i
        free(self._data)
        free(self._matrix)
        self._size = size
        self._nitems = size * size
        self._data = <float*> calloc(max(self._nitems, 1), sizeof(float))
        self._matrix = <float**> calloc(max(size, 1), sizeof(float*))
        if self._data == NULL or self._matrix == NULL:
            raise MemoryError()
        for i in range(size):
            self._matrix[i] = &self._data[i * size]
        return 0

    def __init__(self, object matrix not None, str alphabet = "ARNDCQEGHILKMFPSTWYVBZX*",
                 str name = None):
        cdef size_t i, j
        rows = [list(row) for row in matrix]
        if len(rows) != len(alphabet) or any(len(r) != len(alphabet) for r in rows):
            raise ValueError("matrix must be square and indexed by the alphabet")
        self.alphabet = alphabet
        self.name = name
        self._allocate(len(alphabet))
        for i in range(self._size):
            for j in range(self._size):
                self._matrix[i][j] = rows[i][j]

    @classmethod
    def from_name(cls, str name = "BLOSUM62"):
        if name != "BLOSUM62":
            raise ValueError(
                f"the offline scoring_matrices stand-in only knows BLOSUM62, not {name!r}")
        return ScoringMatrix(_BLOSUM62, alphabet=_BLOSUM62_ALPHABET, name=name)

    def __len__(self):
        return self._size

    def __iter__(self):
        cdef size_t i, j
        for i in range(self._size):
            yield [self._matrix[i][j] for j in range(self._size)]

    def __getitem__(self, index):
        cdef ssize_t i
        if isinstance(index, str):
            i = self.alphabet.index(index)
        else:
            i = index
            if i < 0:
                i += self._size
        if i < 0 or <size_t> i >= self._size:
            raise IndexError(index)
        return [self._matrix[i][j] for j in range(self._size)]

    def __reduce__(self):
        return (type(self), (list(self), self.alphabet, self.name))

    def __eq__(self, other):
        if not isinstance(other, ScoringMatrix):
            return NotImplemented
        return self.alphabet == other.alphabet and list(self) == list(other)

    def shuffle(self, str alphabet):
        """A new matrix with rows and columns reordered (or restricted) to ``alphabet``."""
        idx = [self.alphabet.index(c) for c in alphabet]
        rows = [[self._matrix[i][j] for j in idx] for i in idx]
        return ScoringMatrix(rows, alphabet=alphabet, name=self.name)
