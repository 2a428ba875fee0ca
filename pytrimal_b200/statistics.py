"""The four per-alignment statistics, computed on a B200 through the C ABI.

Host mirror of ``statistics::Manager`` for the CUDA platform
(vendor/trimal/source/Statistics/Manager.cpp:61-114, 228-370): lazy, cached
gap / identity / similarity / overlap statistics of one alignment, with the
non-kernel parts the reference's base classes keep on the host (gap and
similarity windows, Gaps.cpp:93-153, Similarity.cpp:212-269).

Every number comes from libtrimal_cuda.so; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .alignment import Alignment
from .matrix import SimilarityMatrix

_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(ty)


def _mask(m, n):
    if m is None:
        return None
    m = np.ascontiguousarray(m, np.int32)
    if m.shape != (n,):
        raise ValueError("keep-mask has the wrong length")
    return m


class Communicator:
    """NCCL communicator of the ``*_all`` calls (wraps ``tcu_comm``), one process per GPU.

    The 128-byte rendezvous id is made by rank 0 and handed to the other ranks
    by ``exchange(bytes_or_None) -> bytes``; :meth:`from_torch` uses an initialised
    ``torch.distributed`` process group (any backend) for that and nothing else."""

    ID_BYTES = 128

    def __init__(self, rank, world, device, exchange):
        self.lib = _lib.load()
        ident = None
        if rank == 0:
            buf = C.create_string_buffer(self.ID_BYTES)
            _lib.check(self.lib.tcu_comm_id(buf))
            ident = buf.raw
        ident = exchange(ident)
        if not isinstance(ident, (bytes, bytearray)) or len(ident) != self.ID_BYTES:
            raise ValueError("exchange() must return the 128-byte id of rank 0")
        self._h = C.c_void_p()
        _lib.check(self.lib.tcu_comm_create(C.c_char_p(bytes(ident)), rank, world, device,
                                            C.byref(self._h)))
        self.rank, self.world, self.device = rank, world, device

    @classmethod
    def from_torch(cls, device):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()

        def exchange(ident):
            box = [ident]
            dist.broadcast_object_list(box, src=0, device=torch.device("cpu")
                                       if dist.get_backend() == "gloo" else None)
            return box[0]

        return cls(rank, world, device, exchange)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.tcu_comm_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def shard_range(total, granule, rank, world):
    """``tcu_shard_range``: rank's contiguous share [begin, end) of ``total`` units."""
    a, b = C.c_int(0), C.c_int(0)
    _lib.check(_lib.load().tcu_shard_range(total, granule, rank, world, C.byref(a), C.byref(b)))
    return a.value, b.value


def shard_blocks(kept_rows, rank, world):
    """``tcu_shard_blocks``: rank's band [begin, end) of 128-row blocks of the pair matrix."""
    a, b = C.c_int(0), C.c_int(0)
    _lib.check(_lib.load().tcu_shard_blocks(kept_rows, rank, world, C.byref(a), C.byref(b)))
    return a.value, b.value


DEVICE_AUTO = -1


def set_devices(devices=None):
    """``tcu_set_devices``: the GPUs that ``DeviceAlignment(..., device="auto")`` handles are
    replicated on (one process driving several GPUs); ``None`` goes back to the environment
    variable ``TRIMAL_CUDA_DEVICES`` / device 0."""
    lib = _lib.load()
    if not devices:
        _lib.check(lib.tcu_set_devices(None, 0))
        return
    arr = (C.c_int * len(devices))(*devices)
    _lib.check(lib.tcu_set_devices(arr, len(devices)))


def get_devices():
    lib = _lib.load()
    arr = (C.c_int * 64)()
    k = lib.tcu_get_devices(arr, 64)
    return list(arr[:k])


class DeviceAlignment:
    """One alignment uploaded to one GPU -- or, with ``device="auto"``, replicated over the
    configured device set (wraps a ``tcu_msa`` handle).

    Every statistic takes ``comm=``: with a :class:`Communicator` the ``tcu_*_all``
    entry point is used (each rank computes its share, NCCL exchanges the shares,
    every rank returns the complete result)."""

    def __init__(self, alignment, device=0):
        if device == "auto":
            device = DEVICE_AUTO
        if not isinstance(alignment, Alignment):
            alignment = Alignment.from_matrix(np.asarray(alignment, np.uint8))
        self.alignment = alignment
        self.lib = _lib.load()
        m = np.ascontiguousarray(alignment.matrix)
        self._h = C.c_void_p()
        _lib.check(self.lib.tcu_msa_create_strided(
            m.ctypes.data_as(C.c_void_p), m.shape[0], m.shape[1],
            m.strides[0] if m.shape[0] else max(m.shape[1], 1), device, C.byref(self._h)))
        self.nseq, self.ncol = m.shape
        self.device = device
        self.device_count = self.lib.tcu_msa_device_count(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.tcu_msa_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def timings(self):
        t = _lib.Timings()
        _lib.check(self.lib.tcu_msa_timings(self._h, C.byref(t)))
        return {"h2d_ms": t.h2d_ms, "pack_ms": t.pack_ms, "kernel_ms": t.kernel_ms,
                "d2h_ms": t.d2h_ms, "kernel_launches": t.kernel_launches, "comm_ms": t.comm_ms}

    # -- K3 -------------------------------------------------------------------
    def gaps(self, save_seq=None, comm=None):
        """(gapsInColumn, numColumnsWithGaps, maxGaps) -- template.h:444-502."""
        ss = _mask(save_seq, self.nseq)
        g = np.zeros(self.ncol, np.int32)
        hist = np.zeros(self.nseq + 1, np.int32)
        mx = C.c_int(0)
        if comm is None:
            rc = self.lib.tcu_gaps(self._h, _p(ss, _i32p), _p(g, _i32p), _p(hist, _i32p),
                                   C.byref(mx))
        else:
            rc = self.lib.tcu_gaps_all(self._h, comm._h, _p(ss, _i32p), _p(g, _i32p),
                                       _p(hist, _i32p), C.byref(mx))
        _lib.check(rc)
        return g, hist, mx.value

    # -- K1 -------------------------------------------------------------------
    def identity(self, indet=None, save_seq=None, save_res=None, counts=False,
                 keep_on_device=False, out=None, comm=None):
        """Packed identity array over kept pairs -- template.h:320-442.  With ``comm``
        the matrix always stays on every rank's device as well."""
        indet = self.alignment.indet if indet is None else indet
        ss, sr = _mask(save_seq, self.nseq), _mask(save_res, self.ncol)
        nk = self.nseq if ss is None else int((ss != -1).sum())
        npairs = nk * (nk - 1) // 2
        ident = np.empty(npairs, np.float32) if out is None else out
        assert ident.dtype == np.float32 and ident.size >= npairs and ident.flags.c_contiguous
        if comm is not None:
            if counts:
                raise ValueError("counts are a single-GPU debugging output")
            _lib.check(self.lib.tcu_identity_all(self._h, comm._h, _p(ss, _i32p), _p(sr, _i32p),
                                                 indet, _p(ident, _f32p)))
            return ident[:npairs]
        hit = np.zeros(npairs, np.int32) if counts else None
        dst = np.zeros(npairs, np.int32) if counts else None
        rc = self.lib.tcu_identity(self._h, _p(ss, _i32p), _p(sr, _i32p), indet,
                                   _p(ident, _f32p), _p(hit, _i32p), _p(dst, _i32p),
                                   1 if keep_on_device else 0)
        _lib.check(rc)
        ident = ident[:npairs]
        return (ident, hit, dst) if counts else ident

    # -- K4 -------------------------------------------------------------------
    def similarity(self, matrix: SimilarityMatrix, gaps=None, number_of_residues=None,
                   identities=None, indet=None, comm=None):
        """(mdk, num, den) -- template.h:69-204.  ``gaps=None`` is cutByGap=False.
        Needs the device identities of a previous ``identity(keep_on_device=True)``
        unless ``identities`` is given."""
        indet = self.alignment.indet if indet is None else indet
        nres = self.ncol if number_of_residues is None else number_of_residues
        thr = np.float32(0.8) * np.float32(nres)           # template.h:108
        dist = np.ascontiguousarray(matrix.distances, np.float32)
        vhash = np.ascontiguousarray(matrix.vhash, np.int32)
        g = None if gaps is None else np.ascontiguousarray(gaps, np.int32)
        ids = None if identities is None else np.ascontiguousarray(identities, np.float32)
        num, den, mdk = (np.zeros(self.ncol, np.float32) for _ in range(3))
        ec, er, eb = C.c_int(-1), C.c_int(-1), C.c_int(0)
        if comm is not None:
            if ids is not None:
                raise ValueError("similarity(comm=...) uses the matrix identity(comm=...) left "
                                 "on the device")
            rc = self.lib.tcu_similarity_all(self._h, comm._h, indet, _p(dist, _f32p),
                                             dist.shape[0], _p(vhash, _i32p), _p(g, _i32p),
                                             C.c_float(thr), _p(num, _f32p), _p(den, _f32p),
                                             _p(mdk, _f32p), C.byref(ec), C.byref(er),
                                             C.byref(eb))
        else:
            rc = self.lib.tcu_similarity(self._h, indet, _p(dist, _f32p), dist.shape[0],
                                         _p(vhash, _i32p), _p(g, _i32p), C.c_float(thr),
                                         _p(ids, _f32p), _p(num, _f32p), _p(den, _f32p),
                                         _p(mdk, _f32p), C.byref(ec), C.byref(er), C.byref(eb))
        _lib.check(rc, (ec.value, er.value, eb.value))
        return mdk, num, den

    # -- K2 -------------------------------------------------------------------
    def spurious(self, overlap, indet=None, comm=None):
        """spuriousVector -- template.h:206-318."""
        indet = self.alignment.indet if indet is None else indet
        # template.h:217-218: fp32 product, then ceil
        ovrlap = int(math.ceil(float(np.float32(overlap) * np.float32(self.nseq - 1))))
        out = np.zeros(self.nseq, np.float32)
        if comm is None:
            rc = self.lib.tcu_spurious(self._h, indet, max(ovrlap, 0), _p(out, _f32p))
        else:
            rc = self.lib.tcu_spurious_all(self._h, comm._h, indet, max(ovrlap, 0),
                                           _p(out, _f32p))
        _lib.check(rc)
        return out

    # -- consumers of the device-resident identity matrix (Cleaner.cpp walks) ---
    def identity_on_device(self, indet=None, save_res=None):
        """Identity matrix over all rows, left on the device (nothing crosses PCIe)."""
        indet = self.alignment.indet if indet is None else indet
        sr = _mask(save_res, self.ncol)
        _lib.check(self.lib.tcu_identity(self._h, None, _p(sr, _i32p), indet, None, None, None, 1))

    @property
    def identity_resident(self):
        return bool(self.lib.tcu_identity_resident(self._h))

    def identity_download(self):
        out = np.empty(self.nseq * (self.nseq - 1) // 2, np.float32)
        _lib.check(self.lib.tcu_identity_download(self._h, _p(out, _f32p)))
        return out

    def identity_row_stats(self, upper_only=False):
        """(row_max, row_min, row_sum): Cleaner.cpp:68-80 (all j != i) or :1054-1063 (j > i)."""
        mx, mn, sm = (np.zeros(self.nseq, np.float32) for _ in range(3))
        _lib.check(self.lib.tcu_identity_row_stats(self._h, 1 if upper_only else 0, _p(mx, _f32p),
                                                   _p(mn, _f32p), _p(sm, _f32p)))
        return mx, mn, sm

    def byte_histogram(self):
        """Occurrences of each byte value over the whole matrix (utils.cpp:487-512)."""
        out = (C.c_ulonglong * 256)()
        _lib.check(self.lib.tcu_byte_histogram(self._h, out))
        return np.array(out[:], np.uint64)

    def sequence_lengths(self):
        """Alignment::getSequenceLength of every row (Alignment.cpp:296-298)."""
        out = np.zeros(self.nseq, np.int32)
        _lib.check(self.lib.tcu_sequence_lengths(self._h, _p(out, _i32p)))
        return out

    def row_residues(self, save_res=None):
        """Non-gap bytes of every row over the kept columns (Cleaner.cpp:1338-1370)."""
        sr = _mask(save_res, self.ncol)
        out = np.zeros(self.nseq, np.int32)
        _lib.check(self.lib.tcu_row_residues(self._h, _p(sr, _i32p), _p(out, _i32p)))
        return out

    def row_hashes(self):
        """(nseq, 2) uint64: equal rows have equal hashes (candidates for removeDuplicates)."""
        out = np.zeros((self.nseq, 2), np.uint64)
        _lib.check(self.lib.tcu_row_hashes(self._h, out.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return out

    def clusters(self, order, threshold, count_only=False):
        """Greedy clustering (Cleaner.cpp:1427-1447 / 1100-1118): representatives in
        creation order, or only their number."""
        order = np.ascontiguousarray(order, np.int32)
        out = None if count_only else np.zeros(max(len(order), 1), np.int32)
        k = C.c_int(0)
        _lib.check(self.lib.tcu_identity_clusters(self._h, _p(order, _i32p), len(order),
                                                  C.c_float(threshold), _p(out, _i32p),
                                                  C.byref(k)))
        return k.value if count_only else out[: k.value].copy()

    def representatives(self, threshold, indet=None, save_res=None, comm=None):
        """Cleaner::calculateRepresentativeSeq (Cleaner.cpp:1398-1466) in one library call."""
        indet = self.alignment.indet if indet is None else indet
        sr = _mask(save_res, self.ncol)
        out = np.zeros(max(self.nseq, 1), np.int32)
        k = C.c_int(0)
        if comm is None:
            rc = self.lib.tcu_representatives(self._h, _p(sr, _i32p), indet, C.c_float(threshold),
                                              _p(out, _i32p), C.byref(k))
        else:
            rc = self.lib.tcu_representatives_all(self._h, comm._h, _p(sr, _i32p), indet,
                                                  C.c_float(threshold), _p(out, _i32p), C.byref(k))
        _lib.check(rc)
        return out[: k.value].copy()

    def select_method(self):
        """Cleaner::selectMethod (Cleaner.cpp:46-99) on the resident matrix:
        ("gappyout" | "strict", avgSeq, maxSeq).  The O(n) tail runs here in fp32, in the
        reference's order."""
        n = self.nseq
        mx, _, sm = self.identity_row_stats(upper_only=False)
        avg_seq, max_seq = np.float32(0), np.float32(0)
        nm1 = np.float32(n - 1)
        for i in range(n):
            avg_seq = np.float32(avg_seq + np.float32(sm[i] / nm1))
            max_seq = np.float32(max_seq + mx[i])
        avg_seq = np.float32(avg_seq / np.float32(n))
        max_seq = np.float32(max_seq / np.float32(n))
        if float(avg_seq) >= 0.55:
            r = "gappyout"
        elif float(avg_seq) <= 0.38:
            r = "strict"
        elif n <= 20:
            r = "gappyout"
        else:
            r = "gappyout" if 0.5 <= float(max_seq) <= 0.65 else "strict"
        return r, avg_seq, max_seq

    def cutpoint_clusters(self, cluster_number, order=None):
        """Cleaner::getCutPointClusters (Cleaner.cpp:1026-1156): (threshold, clusterings run)."""
        n = self.nseq
        if cluster_number == n:
            return np.float32(1), 0
        if cluster_number == 1:
            return np.float32(0), 0
        mx, mn, sm = self.identity_row_stats(upper_only=True)
        f = np.float32
        g_max, g_min, start = f(0), f(1), f(0)
        for i in range(n):
            compared = n - 1 - i
            if compared > 0:
                start = f(start + f(sm[i] / f(compared)))
                g_max = max(g_max, mx[i])
                g_min = min(g_min, mn[i])
        pairs = n * (n - 1) // 2
        if pairs > 0:
            start = f(start / f(pairs))
        if order is None:
            order = cluster_order(self.sequence_lengths())
        prev, it, runs = f(0), 0, 0
        while True:
            k = self.clusters(order, start, count_only=True)
            runs += 1
            if k == cluster_number or it > 10:
                break
            if k > cluster_number:
                g_max = start
            else:
                g_min = start
            start = f(f(g_max + g_min) / f(2))
            if prev != f(k):
                it, prev = 0, f(k)
            else:
                it += 1
        return start, runs

    # -- device-resident identity (benchmarks / multi-GPU) ---------------------
    def identity_prepare(self, indet=None, save_seq=None, save_res=None):
        indet = self.alignment.indet if indet is None else indet
        ss, sr = _mask(save_seq, self.nseq), _mask(save_res, self.ncol)
        nk = C.c_int(0)
        _lib.check(self.lib.tcu_identity_prepare(self._h, _p(ss, _i32p), _p(sr, _i32p), indet,
                                                 C.byref(nk)))
        return nk.value

    def identity_device(self, block_begin, block_end, device_ptr):
        _lib.check(self.lib.tcu_identity_device(self._h, block_begin, block_end,
                                                C.c_void_p(device_ptr)))

    def sync(self):
        _lib.check(self.lib.tcu_msa_sync(self._h))

    @property
    def stream(self):
        return self.lib.tcu_msa_stream(self._h)


def cluster_order(lengths):
    """Visiting order of the clustering walks (utils.cpp:246-273 + Cleaner.cpp:1413-1426);
    host only."""
    lengths = np.ascontiguousarray(lengths, np.int32)
    out = np.zeros(len(lengths), np.int32)
    _lib.check(_lib.load().tcu_cluster_order(_p(lengths, _i32p), len(lengths), _p(out, _i32p)))
    return out


def threshold_rule(threshold):
    """(mode, mul, shift): the integer form of ``float32(hit) / float32(dst) > threshold`` the
    identity kernel uses in threshold mode (tcu_threshold_rule); host only."""
    mode, mul, shift = C.c_int(), C.c_uint(), C.c_int()
    _lib.load().tcu_threshold_rule(float(np.float32(threshold)), C.byref(mode), C.byref(mul), C.byref(shift))
    return mode.value, mul.value, shift.value


def gaps_window(gaps_in_column, half_window):
    """``Gaps::applyWindow`` (Gaps.cpp:93-153): mirrored integer mean with
    ``utils::roundInt`` (utils.cpp:68-72).  Host side, O(L*w)."""
    g = np.asarray(gaps_in_column, np.int64)
    L = len(g)
    if half_window > L // 4:
        raise ValueError("gap window too big")            # ErrorCode::GapWindowTooBig
    if half_window < 1:
        return np.asarray(gaps_in_column, np.int32).copy()
    idx = np.arange(-half_window, half_window + 1)[None, :] + np.arange(L)[:, None]
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= L, 2 * L - idx - 2, idx)
    s = g[idx].sum(axis=1)
    return (s.astype(np.float64) / (2 * half_window + 1) + 0.5).astype(np.int32)


def similarity_window(mdk, half_window):
    """``Similarity::applyWindow`` (Similarity.cpp:212-269): fp32 running sum in
    ascending j, divided by (float)(2h+1)."""
    mdk = np.asarray(mdk, np.float32)
    L = len(mdk)
    if half_window > L // 4:
        raise ValueError("similarity window too big")     # ErrorCode::SimilarityWindowTooBig
    if half_window < 1:
        return mdk.copy()
    idx = np.arange(-half_window, half_window + 1)[None, :] + np.arange(L)[:, None]
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= L, 2 * L - idx - 2, idx)
    acc = np.zeros(L, np.float32)
    for j in range(2 * half_window + 1):                   # sequential fp32 adds, same order
        acc = (acc + mdk[idx[:, j]]).astype(np.float32)
    return (acc / np.float32(2 * half_window + 1)).astype(np.float32)
