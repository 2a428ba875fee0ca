# cython: language_level=3
cdef class ScoringMatrix:
    cdef readonly str   alphabet
    cdef readonly str   name
    cdef          size_t _size
    cdef          size_t _nitems
    cdef          float*  _data
    cdef          float** _matrix

    cdef int _allocate(self, size_t size) except 1
