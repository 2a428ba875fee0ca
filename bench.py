#!/usr/bin/env python
"""bench.py -- throughput of the statistics hot path on 1..8 B200 (one process per GPU).

Default workload (config.workload): BASELINE.json configs[3], the RepresentativeTrimmer
identity matrix of a synthetic 50 000 x 1 000 protein MSA (seed 4): the largest
configuration that is sharded over 1/2/4/8 GPUs and fits one B200.  A "step" is one full
pass of the hot path over the alignment: bit-plane packing (K0) plus the pairwise-identity
kernel (K1) for every pair this rank owns.

  value  : whole-job pair-column comparisons/s (P*L*K / max-over-ranks device time),
           alignment already resident in HBM.
  e2e    : the same metric through the host-buffer C ABI call the CUDA platform makes for
           this configuration's trimmer (RepresentativeTrimmer ->
           Cleaner::calculateRepresentativeSeq -> tcu_msa_create_strided +
           tcu_representatives; at N > 1 tcu_msa_create_all + tcu_representatives_all):
           pinned host rows -> device, pack, identity kernel in threshold mode (one bit per
           pair), mirror pass, sequence lengths, greedy clustering on the device,
           representatives -> host, all inside the timed region.
           e2e_matrix_to_host is the other public path (tcu_identity_band): the packed
           matrix itself copied to pinned host memory (4*P bytes over PCIe).
  N > 1  : the pair matrix is split into contiguous row-block bands of equal work, one band
           per rank (strong scaling: the alignment is fixed).  Every rank's band is checked
           against the digests of the reference's matrix (tests/golden/full/C4.npz), so the
           scaling record also shows bit-identity.

--workload C5 : Overlap::calculateSpuriousVector on 100 000 x 2 000 (BASELINE configs[4]),
                rows sharded over the ranks; HBM roofline.
--workload C3 : Similarity::calculateVectors on 10 000 x 5 000 (BASELINE configs[2]) with the
                identity matrix resident; latency-bound (no roofline applies).

`--impl reference` times the reference's own AVX2 code (oracle/_ref, the unmodified vendored
trimAl compiled by oracle/Makefile) on the box's host CPU.
"""
from __future__ import annotations

import argparse
import ctypes as C
import gc
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "pair-col/s"
METRICS = {
    "C4": "pair-column comparisons/s (pairwise identity)",
    "C5": "ordered pair-column comparisons/s (spurious / overlap vector)",
    "C3": "pair-column comparisons/s (column similarity)",
    "C2": "pair-column comparisons/s (pairwise identity)",
}
OPS_PER_PAIR_COLUMN = 42  # 2*(20 one-hot planes + 1 gap plane) int8 tensor ops


def measured_peaks():
    """HBM / bf16 from the driver-written MEASURED_PEAKS.json; the dense int8 tensor peak
    from this repo's own probe (tools/umma_i8_peak.cu: back-to-back tcgen05.mma kind::i8,
    run on this pool's B200s, result committed under profiles/)."""
    out = {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        out = {"bf16_tflops": float(p["bf16_tflops"]), "hbm_gbs": float(p["hbm_gbs"]),
               "source": "measured (MEASURED_PEAKS.json, burst)"}
    out["int8_tops"], out["int8_source"] = 2.0 * out["bf16_tflops"], "2 x bf16 dense, " + out["source"]
    probe = os.path.join(ROOT, "profiles", "r02_umma_i8_peak.json")
    if os.path.exists(probe):
        with open(probe) as f:
            q = json.load(f)
        out["int8_tops"] = float(q["int8_tops_burst"])
        out["int8_source"] = ("measured on this pool's B200: tcgen05.mma kind::i8 M128 N256 K32 back to "
                              "back (tools/umma_i8_peak.cu, profiles/r02_umma_i8_peak.json)")
    return out


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed summary
    of an `ncu --set full` capture (profiles/ncu_traffic.json); None when not captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f).get(kernel_key)
    return None if t is None else float(t["dram_read_bytes"]) + float(t["dram_write_bytes"])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arms: the reference's own code on the box's host cores
# ---------------------------------------------------------------------------
def cpu_sample(workload, n_full, L, seed, seconds):
    """One bounded pass of the reference's CPU implementation of `workload`'s hot path (the
    AVX2 build in oracle/_ref, else the plain-C port) on the first rows of the seeded
    alignment.  Returns (units processed, seconds, kind, description of the sample)."""
    import oracle
    from pytrimal_b200.synthetic import synthetic_msa
    kind = "reference" if oracle.ref_available() else "port"
    X = ord("X")
    if workload == "C5":        # ordered pairs, ~8e9 /s/core (AVX2)
        rate = 8.0e9 if kind == "reference" else 0.3e9
        ns = int(min(n_full, max(64, math.sqrt(rate * seconds / L))))
    elif workload == "C3":      # scalar similarity ~1e9 /s/core (+ identity at 5e9)
        rate = 0.8e9 if kind == "reference" else 0.5e9
        ns = int(min(n_full, max(64, math.sqrt(2.0 * rate * seconds / L))))
    else:
        rate = 5.0e9 if kind == "reference" else 0.8e9
        ns = int(min(n_full, max(64, math.sqrt(2.0 * rate * seconds / L))))
    m = synthetic_msa(ns, L, seed)
    t0 = time.perf_counter()
    if workload == "C5":
        if kind == "reference":
            oracle.Ref(m, platform=oracle.PLATFORM_AVX2).spurious(0.5)
        else:
            oracle.Port().spurious_pairwise(m, X, 0.5)
        units = ns * (ns - 1) * L
        what = "Overlap::calculateSpuriousVector(0.5), the O(n^2 L) loop of template.h:235-310"
    elif workload == "C3":
        if kind == "reference":
            oracle.Ref(m, platform=oracle.PLATFORM_AVX2).similarity()
        else:
            port = oracle.Port()
            from pytrimal_b200 import SimilarityMatrix
            smx = SimilarityMatrix.aa()
            ident = port.identity(m, X)
            port.similarity(m, X, ident, port.gaps(m)[0], L, smx.distances, smx.vhash)
        units = ns * (ns - 1) // 2 * L
        what = ("Manager::calculateConservationStats = gaps + Identity::calculateSeqIdentity + "
                "Similarity::calculateVectors")
    else:
        if kind == "reference":
            r = oracle.Ref(m, platform=oracle.PLATFORM_AVX2)
            r.representatives(0.8)
            del r
        else:
            port = oracle.Port()
            ident = port.identity(m, X)
            port.greedy_clusters(ident, ns, port.cluster_order(port.sequence_lengths(m)), 0.8)
        units = ns * (ns - 1) // 2 * L
        what = ("Cleaner::calculateRepresentativeSeq(0.8) = Identity::calculateSeqIdentity (AVX2) + "
                "the greedy walk")
    dt = time.perf_counter() - t0
    sample = f"first {ns} of {n_full} rows x {L} cols of the seeded {workload} alignment, one pass of {what}"
    return units, dt, kind, sample


def cpu_baseline(workload, n_full, L, seed):
    units, dt, kind, sample = cpu_sample(workload, n_full, L, seed, 12.0)
    return {"value": units / dt, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
            "sample": sample,
            "note": "trimAl/pytrimal statistics are single-threaded (SURVEY F9); "
                    f"host has {os.cpu_count()} logical cores"}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation, bounded sample per step."""
    if rank != 0:
        return
    from pytrimal_b200.synthetic import CONFIGS
    n, L, seed = CONFIGS[args.workload]
    if args.rows:
        n = args.rows
    steps, warmup = args.steps, args.warmup
    budget = min(8.0, 150.0 / max(1, steps + warmup))        # seconds per step
    kind = sample = None
    for _ in range(warmup):
        cpu_sample(args.workload, n, L, seed, budget)
    units = secs = 0.0
    for _ in range(steps):
        u, dt, kind, sample = cpu_sample(args.workload, n, L, seed, budget)
        units += u
        secs += dt
    value = units / secs
    line = {
        "impl": "reference", "metric": METRICS[args.workload], "value": value, "unit": UNIT,
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * secs / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if args.workload == "C3" else "u8", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, n, L), "sample": sample + " per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "note": "trimAl/pytrimal statistics are single-threaded (SURVEY F9); "
                                 f"host has {os.cpu_count()} logical cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(wl, n, L):
    return {
        "C4": f"C4: pairwise identity {n}x{L} synthetic protein MSA (RepresentativeTrimmer identity matrix)",
        "C2": f"C2: pairwise identity {n}x{L} synthetic protein MSA",
        "C5": f"C5: spurious / overlap vector {n}x{L} synthetic protein MSA (OverlapTrimmer)",
        "C3": f"C3: column similarity {n}x{L} synthetic protein MSA (AutomaticTrimmer strict*), identity resident in HBM",
    }[wl]


# ---------------------------------------------------------------------------
# helpers shared by the GPU workloads
# ---------------------------------------------------------------------------
class Env:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import pytrimal_b200 as pb
        from pytrimal_b200 import _lib
        self.torch, self.dist, self.pb, self._lib = torch, dist, pb, _lib
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if pb.device_count() < 1:
            raise SystemExit("bench.py needs a B200: libtrimal_cuda has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.lib = pb.load()
        self.comm = pb.Communicator.from_torch(self.local_rank) if self.world > 1 else None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def min_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return t.tolist()

    def create(self, host_rows, n, L):
        """The upload a caller of the C ABI does: the whole alignment from pinned memory on one
        GPU; one share per rank + all-gather over NVLink on several."""
        h = C.c_void_p()
        if self.comm is None:
            self._lib.check(self.lib.tcu_msa_create_strided(C.c_void_p(host_rows.data_ptr()), n, L, L,
                                                            self.local_rank, C.byref(h)))
        else:
            self._lib.check(self.lib.tcu_msa_create_all(self.comm._h, C.c_void_p(host_rows.data_ptr()),
                                                        n, L, L, C.byref(h)))
        return h

    def close(self):
        if self.comm is not None:
            self.comm.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_device_steps(env, step, sync, steps, warmup, stream):
    """W warm-up steps, then K steps between two CUDA events on the library's stream, barrier +
    synchronize on both sides; clocks sampled meanwhile on rank 0."""
    torch = env.torch
    for _ in range(warmup):
        step()
    sync()
    env.barrier()
    sampler = ClockSampler(env.local_rank)
    if env.rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    sync()
    env.barrier()
    clock_note = None
    # the timed region may be shorter than a few 100 ms sampling periods: every rank then keeps
    # the same steps running, untimed, for the same number of rounds (the steps of some
    # workloads are collective: all ranks must make the same calls) until rank 0's sampler
    # has seen the clocks under this load
    short = env.max_over_ranks([1.0 if (env.rank == 0 and len(sampler.rows) < 3) else 0.0])[0] > 0
    if short:
        clock_note = ("timed region shorter than the sampling period: clocks sampled over the "
                      "same steps repeated untimed right after it")
        ms_step = max(e0.elapsed_time(e1) / max(steps, 1), 1e-3)
        rounds = int(env.max_over_ranks([min(2000.0, max(1.0, 600.0 / ms_step))])[0])
        for _ in range(rounds):
            step()
        sync()
        env.barrier()
    clocks = sampler.stop() if env.rank == 0 else None
    if clocks is not None and clock_note:
        clocks["note"] = clock_note
    env.barrier()
    return e0.elapsed_time(e1), clocks


def wall_steps(env, fn, steps):
    """Wall clock over `steps` calls of fn() with the interpreter's cyclic GC kept out (as
    timeit does: with torch imported a full collection takes hundreds of ms)."""
    fn()  # warm-up (pinned pool, allocations)
    gc.collect()
    gc.disable()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    env.barrier()
    dt = time.perf_counter() - t0
    gc.enable()
    return env.max_over_ranks([dt])[0]


def band_digest(torch, out, first_offset, count):
    """(sum, position-weighted sum) mod 2**64 of the fp32 bit patterns of out[:count] whose
    element 0 has packed offset `first_offset` -- tests/golden/digest.py on the device."""
    bits = out[:count].view(torch.int32)
    s = 0
    w = 0
    step = 1 << 24
    for o in range(0, count, step):
        v = bits[o:o + step].to(torch.int64)
        k = torch.arange(first_offset + o, first_offset + o + v.numel(), dtype=torch.int64,
                         device=v.device)
        s = (s + int(v.sum().item())) & ((1 << 64) - 1)
        w = (w + int((v * (k % 65521 + 1)).sum().item())) & ((1 << 64) - 1)
    return s, w


# ---------------------------------------------------------------------------
# C4 (default) / C2: pairwise identity
# ---------------------------------------------------------------------------
def run_identity(args, env):
    torch, pb, lib, _lib = env.torch, env.pb, env.lib, env._lib
    from pytrimal_b200.sharding import band_partition
    from pytrimal_b200.synthetic import CONFIGS, synthetic_msa
    rank, world, local_rank = env.rank, env.world, env.local_rank
    n, L, seed = CONFIGS[args.workload]
    m = synthetic_msa(n, L, seed)
    if args.rows:
        m = m[: args.rows].copy()
        n = args.rows
    pairs_total = n * (n - 1) // 2
    X = ord("X")

    host_rows = torch.from_numpy(m).pin_memory()   # e2e uploads come from pinned memory
    host_np = host_rows.numpy()
    band_rows = lib.tcu_identity_band_rows()
    bounds = band_partition(n, world)
    b0, b1 = bounds[rank], bounds[rank + 1]
    off0 = lib.tcu_identity_row_offset(n, band_rows * b0)
    off1 = lib.tcu_identity_row_offset(n, min(band_rows * b1, n))
    my_pairs = off1 - off0

    # ---------------- device-resident steps ---------------------------------
    dev = pb.DeviceAlignment(pb.Alignment.from_matrix(host_np), device=local_rank)
    out = torch.empty(max(my_pairs, 1), dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local_rank))

    def step():
        dev.identity_prepare(X)                       # K0: pack (1 kernel)
        dev.identity_device(b0, b1, out.data_ptr())   # K1 (1 kernel)

    ms_total, clocks = timed_device_steps(env, step, dev.sync, args.steps, args.warmup, stream)
    t = dev.timings   # duration of the dominant kernel alone (last step), events inside the library
    kernel_ms, pack_ms = t["kernel_ms"], t["pack_ms"]
    ms_total_max, kernel_ms_max = env.max_over_ranks([ms_total, kernel_ms])
    value = pairs_total * L * args.steps / (ms_total_max * 1e-3)
    kernel_ms_ranks = [kernel_ms]
    if world > 1:
        box = [None] * world
        env.dist.all_gather_object(box, float(kernel_ms))
        kernel_ms_ranks = box

    # ---------------- bit-identity of this rank's band with the reference ----------------
    verify = {"checked": False}
    gpath = os.path.join(ROOT, "tests", "golden", "full", args.workload + ".npz")
    if os.path.exists(gpath) and not args.rows:
        g = np.load(gpath)
        mask = (1 << 64) - 1
        want_s = int(g["identity_block_sum"][b0:b1].sum(dtype=np.uint64)) & mask if b1 > b0 else 0
        with np.errstate(over="ignore"):
            want_w = int(g["identity_block_wsum"][b0:b1].sum(dtype=np.uint64)) & mask if b1 > b0 else 0
        got_s, got_w = band_digest(torch, out, off0, my_pairs)
        ok = float(got_s == want_s and got_w == want_w)
        all_ok = env.min_over_ranks([ok])[0]
        verify = {"checked": True, "bands_bit_identical_to_reference": bool(all_ok == 1.0),
                  "rank0_band_digest": "%016x:%016x" % (got_s, got_w),
                  "reference": "unmodified trimAl AVX2 (oracle/_ref) on the same seeded alignment, "
                               "per-128-row-block digests in tests/golden/full/%s.npz" % args.workload}

    # ---------------- end-to-end through the host-buffer C ABI ----------------
    reps_host = torch.empty(n, dtype=torch.int32).pin_memory()
    reps_ptr = C.cast(reps_host.data_ptr(), C.POINTER(C.c_int))
    nreps = C.c_int(0)
    e2e_launches = [0]
    e2e_phases = []

    def e2e_step():
        tp0 = time.perf_counter()
        h = env.create(host_rows, n, L)
        tc = _lib.Timings()
        lib.tcu_msa_timings(h, C.byref(tc))
        tp1 = time.perf_counter()
        try:
            if env.comm is None:
                _lib.check(lib.tcu_representatives(h, None, X, C.c_float(0.8), reps_ptr, C.byref(nreps)))
            else:
                _lib.check(lib.tcu_representatives_all(h, env.comm._h, None, X, C.c_float(0.8),
                                                       reps_ptr, C.byref(nreps)))
            tp2 = time.perf_counter()
            t = _lib.Timings()
            lib.tcu_msa_timings(h, C.byref(t))
            e2e_launches[0] = t.kernel_launches
        finally:
            lib.tcu_msa_destroy(h)
        e2e_phases.append({"create_ms": 1e3 * (tp1 - tp0), "create_h2d_ms": tc.h2d_ms,
                           "create_allgather_ms": tc.comm_ms, "call_ms": 1e3 * (tp2 - tp1),
                           "destroy_ms": 1e3 * (time.perf_counter() - tp2),
                           "h2d_ms": t.h2d_ms, "pack_and_mirror_ms": t.pack_ms,
                           "identity_and_clustering_ms": t.kernel_ms,
                           "d2h_ms": t.d2h_ms, "comm_ms": t.comm_ms})

    e2e_s = wall_steps(env, e2e_step, args.e2e_steps)
    e2e_value = pairs_total * L * args.e2e_steps / e2e_s
    e2e_reps = int(nreps.value)
    reps_ok = None
    if verify["checked"] and "representatives_80" in g.files:
        reps_ok = bool(e2e_reps == g["representatives_80"].size and
                       (reps_host.numpy()[:e2e_reps] == g["representatives_80"]).all())

    # (b) the matrix itself to the host (tcu_identity_band, 4*P bytes of D2H)
    host_out = torch.empty(max(my_pairs, 1), dtype=torch.float32).pin_memory()
    out_ptr = C.cast(host_out.data_ptr(), C.POINTER(C.c_float))

    def band_step():
        h = C.c_void_p()
        _lib.check(lib.tcu_msa_create_strided(C.c_void_p(host_rows.data_ptr()), n, L, L, local_rank,
                                              C.byref(h)))
        try:
            _lib.check(lib.tcu_identity_band(h, None, None, X, b0, b1, out_ptr))
        finally:
            lib.tcu_msa_destroy(h)

    band_s = wall_steps(env, band_step, args.e2e_steps)
    band_value = pairs_total * L * args.e2e_steps / band_s
    same = bool(torch.equal(host_out[: min(my_pairs, 1 << 20)], out[: min(my_pairs, 1 << 20)].cpu()))

    # (c) N > 1: the collective forms of the other calls, one timed call each
    collectives = None
    if env.comm is not None:
        collectives = {}
        with pb.DeviceAlignment(pb.Alignment.from_matrix(host_np), device=local_rank) as d2:
            for name, call in (("tcu_gaps_all", lambda: d2.gaps(comm=env.comm)),
                               ("tcu_spurious_all", lambda: d2.spurious(0.5, indet=X, comm=env.comm)),
                               ("tcu_identity_all", lambda: lib.tcu_identity_all(
                                   d2._h, env.comm._h, None, None, X, None))):
                call()
                env.barrier()
                t0 = time.perf_counter()
                call()
                env.barrier()
                dt = env.max_over_ranks([time.perf_counter() - t0])[0]
                tt = d2.timings
                collectives[name] = {"call_ms": 1e3 * dt, "kernel_ms": tt["kernel_ms"],
                                     "comm_ms": tt["comm_ms"]}

    if rank == 0:
        peaks = measured_peaks()
        achieved_tops = OPS_PER_PAIR_COLUMN * (pairs_total / world) * L / (kernel_ms_max * 1e-3) / 1e12
        np_planes = 5
        kernel_name = f"tcu::k_identity2<{np_planes},true>"
        full_size = world == 1 and not args.rows
        line = {
            "metric": METRICS[args.workload], "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {
                "workload": workload_name(args.workload, n, L) + ", K0 pack + K1 identity per step",
                "pairs": pairs_total, "columns": L, "parallelism": f"row-block bands x{world}",
                "l2": "no explicit flush: each step writes %.2f GB of identities per GPU, >> 126 MB L2"
                      % (4.0 * my_pairs / 1e9),
            },
            "clocks": clocks,
            "verify": verify,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n) * int(L),
                    "d2h_bytes_per_step": int(4 * n + 4 * e2e_reps + 4), "steps": args.e2e_steps,
                    "representatives": e2e_reps, "representatives_equal_reference": reps_ok,
                    "gpu_launches_per_step": e2e_launches[0],
                    "last_step_phases_ms": {k: round(v, 3) for k, v in e2e_phases[-1].items()},
                    "step_ms": [round(ph["create_ms"] + ph["call_ms"] + ph["destroy_ms"], 2)
                                for ph in e2e_phases[1:]],
                    "api": ("tcu_msa_create_strided + tcu_representatives" if world == 1 else
                            "tcu_msa_create_all + tcu_representatives_all") +
                           " (pinned host buffers): Cleaner::calculateRepresentativeSeq(0.8) of the "
                           "RepresentativeTrimmer, identity thresholded inside K1, clustering in HBM"},
            "e2e_matrix_to_host": {"value": band_value, "unit": UNIT,
                                   "h2d_bytes_per_step": int(n) * int(L),
                                   "d2h_bytes_per_step": int(4 * my_pairs), "steps": args.e2e_steps,
                                   "matches_device_result": same,
                                   "api": "tcu_msa_create_strided + tcu_identity_band"},
            "gpu_launches": 2 * args.steps,
            "roofline": {
                "bound": "tensor", "achieved": achieved_tops, "peak": peaks["int8_tops"],
                "unit": "TFLOP/s", "frac": achieved_tops / peaks["int8_tops"],
                "traffic": ncu_traffic(f"{kernel_name} {args.workload} {n}x{L}") if full_size else None,
                "kernel": kernel_name, "kernel_ms": kernel_ms_max, "pack_ms": pack_ms,
                "kernel_ms_per_rank": [round(v, 3) for v in kernel_ms_ranks],
                "peak_source": peaks["int8_source"],
                "frac_vs_2x_bf16_dense": achieved_tops / (2.0 * peaks["bf16_tflops"]),
                "note": "algorithmic int8 tensor ops = 42 per pair-column (SURVEY 8d: the one-hot GEMM "
                        "formulation), whatever the kernel does; this kernel counts hits on the "
                        "LOP3/POPC integer pipes (bit planes, 5 LOP3 per 32 pair-columns) and the "
                        "both-gap counts with tcgen05 kind::i8 UMMAs",
            },
        }
        if collectives is not None:
            line["collectives"] = collectives
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, n, L, seed)
        if world == 1 and not args.rows and not args.no_similarity and not args.no_cpu_baseline:
            try:
                line["similarity"] = similarity_extra(pb, CONFIGS, synthetic_msa)
            except Exception as exc:  # never lose the headline line over the secondary figure
                line["similarity"] = {"error": repr(exc)}
            try:
                line["overlap"] = overlap_extra(pb, CONFIGS, synthetic_msa)
            except Exception as exc:
                line["overlap"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    dev.close()


def overlap_extra(pb, CONFIGS, synthetic_msa):
    """BASELINE.json's configs[4] beside the headline: Overlap::calculateSpuriousVector and the gap
    counts over C5, host buffers in and out, checked against the reference's stored vector (see
    --workload C5 for the full line)."""
    n, L, seed = CONFIGS["C5"]
    m = synthetic_msa(n, L, seed)
    X = ord("X")
    peaks = measured_peaks()
    with pb.DeviceAlignment(m) as d:
        best = {"spurious": (1e30, 1e30), "gaps": (1e30, 1e30)}
        for _ in range(3):
            t0 = time.perf_counter()
            sp = d.spurious(0.5, indet=X)
            best["spurious"] = (min(best["spurious"][0], d.timings["kernel_ms"]),
                                min(best["spurious"][1], time.perf_counter() - t0))
            t0 = time.perf_counter()
            g = d.gaps()[0]
            best["gaps"] = (min(best["gaps"][0], d.timings["kernel_ms"]),
                            min(best["gaps"][1], time.perf_counter() - t0))
    out = {"workload": workload_name("C5", n, L)}
    gpath = os.path.join(ROOT, "tests", "golden", "full", "C5.npz")
    if os.path.exists(gpath):
        ref = np.load(gpath)
        out["bit_identical_to_reference"] = bool(
            (sp.view(np.uint32) == ref["spurious_50"].view(np.uint32)).all() and (g == ref["gaps"]).all())
    for name, byts in (("spurious", 2 * n * L + 4 * n), ("gaps", n * L + 4 * L)):
        k, w = best[name]
        out[name] = {"kernel_ms": k, "call_ms_host_buffers": w * 1e3,
                     "roofline": {"bound": "hbm", "achieved": byts / (k * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                  "unit": "GB/s", "frac": byts / (k * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
    out["spurious"]["value"] = n * (n - 1) * L / (best["spurious"][0] * 1e-3)
    out["spurious"]["unit"] = "ordered " + UNIT
    return out


def similarity_extra(pb, CONFIGS, synthetic_msa):
    """The other half of BASELINE.json's metric (identity + similarity), reported beside the
    identity headline: one pass of Similarity::calculateVectors over C3 (see --workload C3)."""
    n, L, seed = CONFIGS["C3"]
    m = synthetic_msa(n, L, seed)
    X = ord("X")
    smx = pb.SimilarityMatrix.aa()
    P = n * (n - 1) // 2
    with pb.DeviceAlignment(m) as d:
        g, _, _ = d.gaps()
        d.identity(X, keep_on_device=True)
        best_k, best_w = 1e30, 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            d.similarity(smx, gaps=g, indet=X)
            best_w = min(best_w, time.perf_counter() - t0)
            best_k = min(best_k, d.timings["kernel_ms"])
    return {"workload": workload_name("C3", n, L),
            "value": P * L / (best_k * 1e-3), "unit": UNIT, "kernel": "tcu::k_similarity2",
            "kernel_ms": best_k, "call_ms_host_buffers": best_w * 1e3,
            "ns_per_chain_step": best_k * 1e6 / P,
            "columns_cut_by_gap_rule": int((g.astype(np.float32) >= np.float32(0.8) * np.float32(L)).sum()),
            "roofline": None,
            "note": "latency-bound: one dependent fp32 add per pair and column in the reference's "
                    "order (floor 4 cycles per step); bit-identical to the reference"}


# ---------------------------------------------------------------------------
# C5: spurious / overlap vector (rows sharded, HBM-bound)
# ---------------------------------------------------------------------------
def run_spurious(args, env):
    torch, pb, lib, _lib = env.torch, env.pb, env.lib, env._lib
    from pytrimal_b200.synthetic import CONFIGS, synthetic_msa
    rank, world, local_rank = env.rank, env.world, env.local_rank
    n, L, seed = CONFIGS[args.workload]
    m = synthetic_msa(n, L, seed)
    if args.rows:
        m = m[: args.rows].copy()
        n = args.rows
    X = ord("X")
    units = n * (n - 1) * L                       # ordered pair-columns
    ovrlap = int(math.ceil(float(np.float32(0.5) * np.float32(n - 1))))
    host_rows = torch.from_numpy(m).pin_memory()
    sp_host = torch.empty(n, dtype=torch.float32).pin_memory()
    sp_ptr = C.cast(sp_host.data_ptr(), C.POINTER(C.c_float))

    dev = pb.DeviceAlignment(pb.Alignment.from_matrix(host_rows.numpy()), device=local_rank)
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local_rank))
    kms = []

    def step():
        if env.comm is None:
            _lib.check(lib.tcu_spurious(dev._h, X, ovrlap, sp_ptr))
        else:
            _lib.check(lib.tcu_spurious_all(dev._h, env.comm._h, X, ovrlap, sp_ptr))
        kms.append(dev.timings["kernel_ms"])

    ms_total, clocks = timed_device_steps(env, step, dev.sync, args.steps, args.warmup, stream)
    kernel_ms = float(np.median(kms[args.warmup:args.warmup + args.steps]))
    ms_total_max, kernel_ms_max = env.max_over_ranks([ms_total, kernel_ms])
    value = units * args.steps / (ms_total_max * 1e-3)

    verify = {"checked": False}
    gpath = os.path.join(ROOT, "tests", "golden", "full", "C5.npz")
    if os.path.exists(gpath) and not args.rows:
        want = np.load(gpath)["spurious_50"]
        verify = {"checked": True,
                  "vector_bit_identical_to_reference": bool(
                      (sp_host.numpy().view(np.uint32) == want.view(np.uint32)).all()),
                  "reference": "unmodified trimAl AVX2 (oracle/_ref), tests/golden/full/C5.npz"}

    def e2e_step():
        h = env.create(host_rows, n, L)
        try:
            if env.comm is None:
                _lib.check(lib.tcu_spurious(h, X, ovrlap, sp_ptr))
            else:
                _lib.check(lib.tcu_spurious_all(h, env.comm._h, X, ovrlap, sp_ptr))
        finally:
            lib.tcu_msa_destroy(h)

    e2e_s = wall_steps(env, e2e_step, args.e2e_steps)
    if rank == 0:
        peaks = measured_peaks()
        # the closed-form (column-histogram) mode is accounted as two passes over the
        # alignment, 2*n*L + 4*n bytes; this rank's share of the rows
        byts = (2 * n * L + 4 * n) / world
        achieved = byts / (kernel_ms_max * 1e-3) / 1e9
        line = {
            "metric": METRICS["C5"], "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload_name("C5", n, L) + ", one tcu_spurious call per step "
                                   "(alignment resident, %d floats to the host)" % n,
                       "parallelism": f"row shards x{world}",
                       "l2": "alignment = %.0f MB per GPU > 126 MB L2" % (n * L / world / 1e6),
                       "note": "exact closed form of the O(n^2 L) loop (SURVEY F7): the rate is the "
                               "reference's work done per second, not compares executed"},
            "clocks": clocks, "verify": verify,
            "e2e": {"value": units * args.e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(n) * int(L), "d2h_bytes_per_step": 4 * int(n),
                    "steps": args.e2e_steps, "ms_per_step": 1e3 * e2e_s / args.e2e_steps,
                    "api": ("tcu_msa_create_strided + tcu_spurious" if world == 1 else
                            "tcu_msa_create_all + tcu_spurious_all") + " (pinned host buffers)"},
            "gpu_launches": 3 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"],
                         "traffic": ncu_traffic(f"spurious C5 {n}x{L}") if world == 1 and not args.rows else None,
                         "kernel": "tcu::k_column_counts<true> + k_spurious_flags + k_spurious_rows",
                         "kernel_ms": kernel_ms_max, "peak_source": peaks["source"],
                         "note": "algorithmic bytes 2*n*L + 4*n (SURVEY 8d, histogram mode = two passes); "
                                 "the kernels move n*L read + n*L/4 written + n*L/4 read (the row pass "
                                 "reads bit planes the column pass left behind)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline("C5", n, L, seed)
        print(json.dumps(line), flush=True)
    dev.close()


# ---------------------------------------------------------------------------
# C3: column similarity (latency-bound)
# ---------------------------------------------------------------------------
def run_similarity(args, env):
    torch, pb, lib, _lib = env.torch, env.pb, env.lib, env._lib
    from pytrimal_b200.synthetic import CONFIGS, synthetic_msa
    rank, world, local_rank = env.rank, env.world, env.local_rank
    n, L, seed = CONFIGS["C3"]
    m = synthetic_msa(n, L, seed)
    if args.rows:
        m = m[: args.rows].copy()
        n = args.rows
    X = ord("X")
    P = n * (n - 1) // 2
    smx = pb.SimilarityMatrix.aa()
    host_rows = torch.from_numpy(m).pin_memory()
    dev = pb.DeviceAlignment(pb.Alignment.from_matrix(host_rows.numpy()), device=local_rank)
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local_rank))
    g, _, _ = dev.gaps(comm=env.comm)
    if env.comm is None:
        dev.identity(X, keep_on_device=True)
    else:
        _lib.check(lib.tcu_identity_all(dev._h, env.comm._h, None, None, X, None))
    kms, res = [], [None]

    def step():
        res[0] = dev.similarity(smx, gaps=g, indet=X, comm=env.comm)
        kms.append(dev.timings["kernel_ms"])

    ms_total, clocks = timed_device_steps(env, step, dev.sync, args.steps, args.warmup, stream)
    kernel_ms = float(np.median(kms[args.warmup:args.warmup + args.steps]))
    ms_total_max, kernel_ms_max = env.max_over_ranks([ms_total, kernel_ms])
    verify = {"checked": False}
    gpath = os.path.join(ROOT, "tests", "golden", "full", "C3.npz")
    if os.path.exists(gpath) and not args.rows:
        want = np.load(gpath)["mdk"]
        verify = {"checked": True, "mdk_bit_identical_to_reference": bool(
            (res[0][0].view(np.uint32) == want.view(np.uint32)).all()),
            "reference": "unmodified trimAl AVX2 (oracle/_ref), tests/golden/full/C3.npz"}

    def e2e_step():
        with pb.DeviceAlignment(pb.Alignment.from_matrix(host_rows.numpy()), device=local_rank) as d:
            gg, _, _ = d.gaps()
            _lib.check(lib.tcu_identity(d._h, None, None, X, None, None, None, 1))
            d.similarity(smx, gaps=gg, indet=X)

    e2e_s = wall_steps(env, e2e_step, max(1, min(args.e2e_steps, 3))) if world == 1 else None
    if rank == 0:
        line = {
            "metric": METRICS["C3"], "value": P * L * args.steps / (ms_total_max * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name("C3", n, L) + ", one tcu_similarity call per step",
                       "parallelism": f"32-column groups x{world}",
                       "columns_cut_by_gap_rule": int((g.astype(np.float32) >= np.float32(0.8) * np.float32(L)).sum()),
                       "l2": "identity matrix = %.0f MB, re-read by every column group" % (4 * P / 1e6)},
            "clocks": clocks, "verify": verify,
            "gpu_launches": 3 * args.steps,
            "roofline": {"bound": "latency", "achieved": kernel_ms_max * 1e6 / max(P, 1), "peak": 4.0 / 1.965,
                         "unit": "ns per chain step", "frac": (4.0 / 1.965) / (kernel_ms_max * 1e6 / max(P, 1)),
                         "traffic": None, "kernel": "tcu::k_similarity2", "kernel_ms": kernel_ms_max,
                         "note": "no HBM/tensor roofline applies (SURVEY 8d): every column replays the "
                                 "reference's sequential fp32 additions, one dependent FADD (4 cycles at "
                                 "1965 MHz) per pair; frac = that floor / achieved"},
        }
        if e2e_s is not None:
            k = max(1, min(args.e2e_steps, 3))
            line["e2e"] = {"value": P * L * k / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(n) * int(L),
                           "d2h_bytes_per_step": 4 * 4 * int(L), "steps": k,
                           "ms_per_step": 1e3 * e2e_s / k,
                           "api": "tcu_msa_create_strided + tcu_gaps + tcu_identity(keep_on_device) + "
                                  "tcu_similarity: what Manager::calculateConservationStats costs with "
                                  "the CUDA platform"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline("C3", n, L, seed)
        print(json.dumps(line), flush=True)
    dev.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--rows", type=int, default=0, help="debug: use only the first ROWS rows")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-similarity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    env = Env(args)
    try:
        if args.workload == "C5":
            run_spurious(args, env)
        elif args.workload == "C3":
            run_similarity(args, env)
        else:
            run_identity(args, env)
    finally:
        env.close()


if __name__ == "__main__":
    main()
