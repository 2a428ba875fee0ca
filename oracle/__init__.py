"""CPU checkers for the statistics hot path -- TEST INFRASTRUCTURE ONLY.

Two ctypes front-ends:

* :class:`Port`  -- ``oracle/liboracle.so``, the plain-C restatement in
  ``oracle/trimal_oracle.c`` (always buildable: ``make -C oracle port``).
* :class:`Ref`   -- ``oracle/_ref/libtrimal_ref.so``, the UNMODIFIED reference
  (vendored trimAl incl. its AVX2 kernels) compiled from ``/root/reference`` by
  ``make -C oracle ref`` and driven through ``oracle/ref_harness.cpp``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  Nothing in
``pytrimal_b200/`` does: the product has no CPU fallback.

An alignment is a C-contiguous ``numpy.uint8`` array of shape ``(nseq, ncol)``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(_HERE, "liboracle.so")
REF_PATH = os.path.join(_HERE, "_ref", "libtrimal_ref.so")

ERR_INCORRECT_SYMBOL = 1
ERR_UNDEFINED_SYMBOL = 2

PLATFORM_NONE, PLATFORM_SSE2, PLATFORM_AVX2, PLATFORM_CUDA = 0, 1, 2, 3

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int)
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)


def build(ref: bool | None = None) -> None:
    """Compile the checkers.  ``ref=None`` builds the reference only when its
    sources are present (i.e. in the build container, not on the GPU box)."""
    subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
    if ref is None:
        ref = os.path.isdir("/root/reference/vendor/trimal/source")
    if ref:
        subprocess.run(["make", "-s", "-j8", "-C", _HERE, "ref"], check=True)


def _ptr(a, ty):
    return None if a is None else a.ctypes.data_as(ty)


def _msa(msa) -> np.ndarray:
    msa = np.ascontiguousarray(msa, dtype=np.uint8)
    if msa.ndim != 2:
        raise ValueError("alignment must be a 2-D uint8 array (nseq, ncol)")
    return msa


def _mask(m, n):
    if m is None:
        return None
    m = np.ascontiguousarray(m, dtype=np.int32)
    assert m.shape == (n,)
    return m


def kept_pairs(nseq: int, save_seq=None) -> int:
    k = nseq if save_seq is None else int((np.asarray(save_seq) != -1).sum())
    return k * (k - 1) // 2


class SymbolError(ValueError):
    """Raised by the similarity oracle for a byte the matrix cannot score
    (template.h:135-145).  ``code`` is ERR_INCORRECT_SYMBOL / ERR_UNDEFINED_SYMBOL."""

    def __init__(self, code, col, row, byte):
        super().__init__(f"symbol error {code} at column {col}, row {row}: {chr(byte)!r}")
        self.code, self.col, self.row, self.byte = code, col, row, byte


class Port:
    """The plain-C restatement (``oracle/trimal_oracle.c``)."""

    def __init__(self, path: str = PORT_PATH):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.orc_gaps.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _i32p, _i32p, _i32p, _i32p]
        L.orc_gaps.restype = None
        L.orc_gaps_simd_quirk.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _i32p, _i32p]
        L.orc_gaps_simd_quirk.restype = None
        L.orc_gaps_window.argtypes = [_i32p, C.c_int, C.c_int, _i32p]
        L.orc_gaps_window.restype = C.c_int
        L.orc_identity.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _i32p, _i32p, C.c_uint8,
                                   _f32p, _i32p, _i32p]
        L.orc_identity.restype = C.c_size_t
        L.orc_spurious_pairwise.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_uint8,
                                            C.c_float, _f32p, _u32p]
        L.orc_spurious_pairwise.restype = None
        L.orc_spurious_hist.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_uint8, C.c_float,
                                        _f32p]
        L.orc_spurious_hist.restype = None
        L.orc_distance_matrix.argtypes = [_f32p, C.c_int, _f32p]
        L.orc_distance_matrix.restype = None
        L.orc_similarity.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_uint8, _f32p, _i32p,
                                     C.c_int, _f32p, C.c_int, _i32p, _f32p, _f32p, _f32p, _i32p,
                                     _i32p, _i32p]
        L.orc_similarity.restype = C.c_int
        L.orc_similarity_finish.argtypes = [_f32p, _f32p, C.c_int, _f32p]
        L.orc_similarity_finish.restype = None
        L.orc_similarity_window.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
        L.orc_similarity_window.restype = C.c_int
        L.orc_sequence_lengths.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _i32p]
        L.orc_sequence_lengths.restype = None
        L.orc_cluster_order.argtypes = [_i32p, C.c_int, _i32p]
        L.orc_cluster_order.restype = C.c_int
        L.orc_greedy_clusters.argtypes = [_f32p, C.c_int, _i32p, C.c_int, C.c_float, _i32p]
        L.orc_greedy_clusters.restype = C.c_int
        L.orc_identity_row_stats.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.orc_identity_row_stats.restype = None
        L.orc_select_method.argtypes = [_f32p, C.c_int, _f32p, _f32p]
        L.orc_select_method.restype = C.c_int
        L.orc_cutpoint_clusters.argtypes = [_f32p, C.c_int, _i32p, C.c_int, _i32p]
        L.orc_cutpoint_clusters.restype = C.c_float

    # -- gaps ---------------------------------------------------------------
    def gaps(self, msa, save_seq=None):
        msa = _msa(msa)
        n, L = msa.shape
        ss = _mask(save_seq, n)
        g = np.zeros(L, np.int32)
        hist = np.zeros(n + 1, np.int32)
        mx = C.c_int(0)
        self.lib.orc_gaps(_ptr(msa, _u8p), n, L, msa.strides[0] if n else L, _ptr(ss, _i32p),
                          _ptr(g, _i32p), _ptr(hist, _i32p), C.byref(mx))
        return g, hist, mx.value

    def gaps_simd_quirk(self, msa, save_seq=None):
        msa = _msa(msa)
        n, L = msa.shape
        ss = _mask(save_seq, n)
        g = np.zeros(L, np.int32)
        self.lib.orc_gaps_simd_quirk(_ptr(msa, _u8p), n, L, msa.strides[0], _ptr(ss, _i32p),
                                     _ptr(g, _i32p))
        return g

    def gaps_window(self, gaps, half_window):
        gaps = np.ascontiguousarray(gaps, np.int32)
        out = gaps.copy()
        if self.lib.orc_gaps_window(_ptr(gaps, _i32p), len(gaps), half_window, _ptr(out, _i32p)):
            raise ValueError("gap window too big")
        return out

    # -- identity -----------------------------------------------------------
    def identity(self, msa, indet, save_seq=None, save_res=None, counts=False):
        msa = _msa(msa)
        n, L = msa.shape
        ss, sr = _mask(save_seq, n), _mask(save_res, L)
        npairs = kept_pairs(n, ss)
        ident = np.zeros(npairs, np.float32)
        hit = np.zeros(npairs, np.int32) if counts else None
        dst = np.zeros(npairs, np.int32) if counts else None
        wrote = self.lib.orc_identity(_ptr(msa, _u8p), n, L, msa.strides[0] if n else L,
                                      _ptr(ss, _i32p), _ptr(sr, _i32p), indet, _ptr(ident, _f32p),
                                      _ptr(hit, _i32p), _ptr(dst, _i32p))
        assert wrote == npairs
        return (ident, hit, dst) if counts else ident

    # -- consumers of the identity matrix (Cleaner.cpp walks) ------------------
    def sequence_lengths(self, msa):
        msa = _msa(msa)
        n, L = msa.shape
        out = np.zeros(n, np.int32)
        self.lib.orc_sequence_lengths(_ptr(msa, _u8p), n, L, msa.strides[0] if n else L,
                                      _ptr(out, _i32p))
        return out

    def cluster_order(self, lengths):
        lengths = np.ascontiguousarray(lengths, np.int32)
        out = np.zeros(len(lengths), np.int32)
        if self.lib.orc_cluster_order(_ptr(lengths, _i32p), len(lengths), _ptr(out, _i32p)):
            raise MemoryError
        return out

    def greedy_clusters(self, identities, nseq, order, threshold):
        identities = np.ascontiguousarray(identities, np.float32)
        order = np.ascontiguousarray(order, np.int32)
        out = np.zeros(max(len(order), 1), np.int32)
        k = self.lib.orc_greedy_clusters(_ptr(identities, _f32p), nseq, _ptr(order, _i32p),
                                         len(order), C.c_float(threshold), _ptr(out, _i32p))
        return out[:k].copy()

    def identity_row_stats(self, identities, nseq, upper_only):
        identities = np.ascontiguousarray(identities, np.float32)
        mx, mn, sm = (np.zeros(nseq, np.float32) for _ in range(3))
        self.lib.orc_identity_row_stats(_ptr(identities, _f32p), nseq, int(upper_only),
                                        _ptr(mx, _f32p), _ptr(mn, _f32p), _ptr(sm, _f32p))
        return mx, mn, sm

    def select_method(self, identities, nseq):
        """(1 = gappyout | 2 = strict, avgSeq, maxSeq)"""
        identities = np.ascontiguousarray(identities, np.float32)
        a, m = C.c_float(0), C.c_float(0)
        r = self.lib.orc_select_method(_ptr(identities, _f32p), nseq, C.byref(a), C.byref(m))
        return r, np.float32(a.value), np.float32(m.value)

    def cutpoint_clusters(self, identities, nseq, order, cluster_number):
        """(threshold, number of clusterings run)"""
        identities = np.ascontiguousarray(identities, np.float32)
        order = np.ascontiguousarray(order, np.int32)
        it = C.c_int(0)
        t = self.lib.orc_cutpoint_clusters(_ptr(identities, _f32p), nseq, _ptr(order, _i32p),
                                           int(cluster_number), C.byref(it))
        return np.float32(t), it.value

    # -- spurious -----------------------------------------------------------
    def spurious_pairwise(self, msa, indet, overlap, hits=False):
        msa = _msa(msa)
        n, L = msa.shape
        out = np.zeros(n, np.float32)
        h = np.zeros((n, L), np.uint32) if hits else None
        self.lib.orc_spurious_pairwise(_ptr(msa, _u8p), n, L, msa.strides[0], indet,
                                       C.c_float(overlap), _ptr(out, _f32p), _ptr(h, _u32p))
        return (out, h) if hits else out

    def spurious_hist(self, msa, indet, overlap):
        msa = _msa(msa)
        n, L = msa.shape
        out = np.zeros(n, np.float32)
        self.lib.orc_spurious_hist(_ptr(msa, _u8p), n, L, msa.strides[0], indet,
                                   C.c_float(overlap), _ptr(out, _f32p))
        return out

    # -- similarity ---------------------------------------------------------
    def distance_matrix(self, sim):
        sim = np.ascontiguousarray(sim, np.float32)
        npos = sim.shape[0]
        dist = np.zeros((npos, npos), np.float32)
        self.lib.orc_distance_matrix(_ptr(sim, _f32p), npos, _ptr(dist, _f32p))
        return dist

    def similarity(self, msa, indet, identities, gaps, number_of_residues, dist, vhash):
        """Returns (mdk, num, den).  ``gaps=None`` means cutByGap=False."""
        msa = _msa(msa)
        n, L = msa.shape
        identities = np.ascontiguousarray(identities, np.float32)
        gaps = None if gaps is None else np.ascontiguousarray(gaps, np.int32)
        dist = np.ascontiguousarray(dist, np.float32)
        vhash = np.ascontiguousarray(vhash, np.int32)
        assert vhash.size >= 26 and dist.shape[0] == dist.shape[1]
        mdk, num, den = (np.zeros(L, np.float32) for _ in range(3))
        ec, er, eb = C.c_int(-1), C.c_int(-1), C.c_int(0)
        rc = self.lib.orc_similarity(_ptr(msa, _u8p), n, L, msa.strides[0], indet,
                                     _ptr(identities, _f32p), _ptr(gaps, _i32p),
                                     number_of_residues, _ptr(dist, _f32p), dist.shape[0],
                                     _ptr(vhash, _i32p), _ptr(mdk, _f32p), _ptr(num, _f32p),
                                     _ptr(den, _f32p), C.byref(ec), C.byref(er), C.byref(eb))
        if rc in (ERR_INCORRECT_SYMBOL, ERR_UNDEFINED_SYMBOL):
            raise SymbolError(rc, ec.value, er.value, eb.value)
        if rc:
            raise MemoryError("oracle similarity failed")
        return mdk, num, den

    def similarity_finish(self, num, den):
        num = np.ascontiguousarray(num, np.float32)
        den = np.ascontiguousarray(den, np.float32)
        mdk = np.zeros_like(num)
        self.lib.orc_similarity_finish(_ptr(num, _f32p), _ptr(den, _f32p), len(num),
                                       _ptr(mdk, _f32p))
        return mdk

    def similarity_window(self, mdk, half_window):
        mdk = np.ascontiguousarray(mdk, np.float32)
        out = mdk.copy()
        if self.lib.orc_similarity_window(_ptr(mdk, _f32p), len(mdk), half_window,
                                          _ptr(out, _f32p)):
            raise ValueError("similarity window too big")
        return out


def ref_available() -> bool:
    return os.path.exists(REF_PATH)


class Ref:
    """One reference ``Alignment`` (real trimAl code) built from a byte matrix."""

    _lib = None
    PATH = REF_PATH

    @classmethod
    def lib(cls):
        if cls.__dict__.get("_lib") is None:
            if not os.path.exists(cls.PATH):
                raise RuntimeError(f"{cls.PATH} not built (needs /root/reference)")
            L = C.CDLL(cls.PATH)
            L.ref_alignment_new.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int]
            L.ref_alignment_new.restype = C.c_void_p
            L.ref_alignment_free.argtypes = [C.c_void_p]
            L.ref_alignment_type.argtypes = [C.c_void_p]
            L.ref_set_platform.argtypes = [C.c_void_p, C.c_int]
            L.ref_get_platform.argtypes = [C.c_void_p]
            L.ref_set_masks.argtypes = [C.c_void_p, _i32p, _i32p]
            L.ref_set_windows.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.ref_gaps.argtypes = [C.c_void_p, _i32p, _i32p, _i32p, _i32p]
            L.ref_identity.argtypes = [C.c_void_p, _f32p, C.c_size_t]
            L.ref_identity_nocopy.argtypes = [C.c_void_p]
            L.ref_similarity.argtypes = [C.c_void_p, _f32p, _f32p]
            L.ref_default_matrix.argtypes = [C.c_void_p, _f32p, _i32p]
            L.ref_spurious.argtypes = [C.c_void_p, C.c_float, _f32p]
            L.ref_trim.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_double), _i32p,
                                   _i32p]
            L.ref_representatives.argtypes = [C.c_void_p, C.c_float, _i32p]
            L.ref_cutpoint.argtypes = [C.c_void_p, C.c_int]
            L.ref_cutpoint.restype = C.c_float
            L.ref_select_method.argtypes = [C.c_void_p]
            L.ref_identity_on_host.argtypes = [C.c_void_p]
            L.ref_detect_type.argtypes = [C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, msa, platform=PLATFORM_AVX2, datatype=0):
        self.msa = msa = _msa(msa)
        self.n, self.L = msa.shape
        rows = (C.c_char_p * self.n)()
        base = msa.ctypes.data
        for i in range(self.n):
            rows[i] = C.cast(base + i * msa.strides[0], C.c_char_p)
        self.h = self.lib().ref_alignment_new(rows, self.n, self.L, datatype)
        if not self.h:
            raise ValueError("reference rejected the alignment")
        self.platform = platform
        self.lib().ref_set_platform(self.h, platform)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().ref_alignment_free(self.h)
            self.h = None

    @property
    def alignment_type(self):
        return self.lib().ref_alignment_type(self.h)

    @property
    def indet(self):
        return ord("X") if self.alignment_type & 8 else ord("N")

    def set_masks(self, save_seq=None, save_res=None):
        ss, sr = _mask(save_seq, self.n), _mask(save_res, self.L)
        self.lib().ref_set_masks(self.h, _ptr(ss, _i32p), _ptr(sr, _i32p))
        self._save_seq = ss

    def set_windows(self, gap_window=0, sim_window=0):
        self.lib().ref_set_windows(self.h, gap_window, sim_window)

    def gaps(self):
        g = np.zeros(self.L, np.int32)
        gw = np.zeros(self.L, np.int32)
        hist = np.zeros(self.n + 1, np.int32)
        mx = C.c_int(0)
        if self.lib().ref_gaps(self.h, _ptr(g, _i32p), _ptr(gw, _i32p), _ptr(hist, _i32p),
                               C.byref(mx)):
            raise RuntimeError("reference gap statistic failed")
        return g, gw, hist, mx.value

    def identity(self, copy=True):
        if not copy:
            if self.lib().ref_identity_nocopy(self.h):
                raise RuntimeError("reference identity failed")
            return None
        npairs = kept_pairs(self.n, getattr(self, "_save_seq", None))
        out = np.zeros(npairs, np.float32)
        if self.lib().ref_identity(self.h, _ptr(out, _f32p), npairs):
            raise RuntimeError("reference identity failed")
        return out

    def similarity(self):
        mdk = np.zeros(self.L, np.float32)
        mdkw = np.zeros(self.L, np.float32)
        if self.lib().ref_similarity(self.h, _ptr(mdk, _f32p), _ptr(mdkw, _f32p)):
            raise ValueError("reference similarity statistic failed")
        return mdk, mdkw

    def default_matrix(self):
        dist = np.zeros(28 * 28, np.float32)
        vhash = np.zeros(28, np.int32)
        n = self.lib().ref_default_matrix(self.h, _ptr(dist, _f32p), _ptr(vhash, _i32p))
        return dist[: n * n].reshape(n, n).copy(), vhash

    def spurious(self, overlap):
        out = np.zeros(self.n, np.float32)
        if self.lib().ref_spurious(self.h, C.c_float(overlap), _ptr(out, _f32p)):
            raise RuntimeError("reference spurious vector failed")
        return out

    def representatives(self, max_identity):
        """Cleaner::calculateRepresentativeSeq: representatives in creation order."""
        out = np.zeros(max(self.n, 1), np.int32)
        k = self.lib().ref_representatives(self.h, C.c_float(max_identity), _ptr(out, _i32p))
        if k < 0:
            raise RuntimeError("reference clustering failed")
        return out[:k].copy()

    def cutpoint(self, clusters):
        return np.float32(self.lib().ref_cutpoint(self.h, int(clusters)))

    def select_method(self):
        return self.lib().ref_select_method(self.h)

    def detect_type(self):
        """Alignment::getAlignmentType from scratch with the current platform."""
        return self.lib().ref_detect_type(self.h)

    def identity_on_host(self):
        return bool(self.lib().ref_identity_on_host(self.h))

    def trim(self, method, params=(), platform=None):
        """Returns (keep_seq, keep_res) int32 arrays (-1 = removed)."""
        p = (C.c_double * 8)(*([float(x) for x in params] + [-1.0] * (8 - len(params))))
        ks = np.zeros(self.n, np.int32)
        kr = np.zeros(self.L, np.int32)
        rc = self.lib().ref_trim(self.h, self.platform if platform is None else platform,
                                 method.encode(), p, _ptr(ks, _i32p), _ptr(kr, _i32p))
        if rc:
            raise ValueError(f"reference trim({method}) failed with {rc}")
        return ks, kr
