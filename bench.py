#!/usr/bin/env python
"""bench.py -- pair-column comparisons/s of the pairwise-identity hot path.

Workload (config.workload): BASELINE.json configs[3], the RepresentativeTrimmer
identity matrix of a synthetic 50 000 x 1 000 protein MSA (seed 4): the largest
configuration that is sharded over 1/2/4/8 GPUs and fits one B200.  A "step" is
one full pass of the hot path over the alignment: bit-plane packing (K0) plus
the pairwise-identity kernel (K1) for every pair this rank owns.

  value  : whole-job pair-column comparisons/s (P*L*K / max-over-ranks device
           time), alignment already resident in HBM.
  e2e    : the same metric through the host-buffer C ABI call the CUDA platform makes
           for this configuration's trimmer (RepresentativeTrimmer ->
           Cleaner::calculateRepresentativeSeq -> tcu_msa_create_strided +
           tcu_representatives[_all]): pinned host rows -> device, pack, identity
           matrix (kept in HBM), sequence lengths, greedy clustering on the device,
           representatives -> host, all inside the timed region.
           e2e_matrix_to_host is the other public path (tcu_identity_band): the packed
           matrix itself copied to pinned host memory (4*P bytes over PCIe).
  N > 1  : the pair matrix is split into contiguous row-block bands of equal
           pair count, one band per rank, no data-path collective (strong
           scaling: the alignment is fixed, every rank holds a replica).

`--impl reference` times the reference's own AVX2 code (oracle/_ref, the
unmodified vendored trimAl compiled by oracle/Makefile) on the box's host CPU.
"""
from __future__ import annotations

import argparse
import ctypes as C
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

METRIC = "pair-column comparisons/s (pairwise identity)"
UNIT = "pair-col/s"
OPS_PER_PAIR_COLUMN = 42  # 2*(20 one-hot planes + 1 gap plane) int8 tensor ops
# dram__bytes_read.sum + dram__bytes_write.sum of one k_identity2 launch at full C4 size, from
# the ncu --set full capture summarised in profiles/r01k_ncu_identity2_c4.txt (1.62 GB read
# + 5.06 GB written; algorithmic bytes n*L + 4*P = 5.05 GB -- the reads are operand blocks
# that miss L2, once per group of 8 super-block rows)
NCU_TRAFFIC_BYTES = 1.618266e9 + 5.061914e9


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": float(p["bf16_tflops"]), "hbm_gbs": float(p["hbm_gbs"]),
                "source": "measured (MEASURED_PEAKS.json, burst)"}
    return {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


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


def run_reference(args, rank, world):
    """The reference's own CPU implementation (AVX2), bounded sample per step."""
    if rank != 0:
        return
    import oracle
    from pytrimal_b200.synthetic import CONFIGS, synthetic_msa
    n, L, seed = CONFIGS[args.workload]
    if args.rows:
        n = args.rows
    steps, warmup = args.steps, args.warmup
    kind = "reference" if oracle.ref_available() else "port"
    rate = 5.0e9 if kind == "reference" else 0.8e9          # expected pair-col/s per core
    budget = min(8.0, 150.0 / max(1, steps + warmup))        # seconds per step
    ns = int(min(n, max(64, math.sqrt(2.0 * rate * budget / L))))
    m = synthetic_msa(ns, L, seed)
    pairs = ns * (ns - 1) // 2
    port = None if kind == "reference" else oracle.Port()

    def one():
        t0 = time.perf_counter()
        if kind == "reference":
            # Cleaner::calculateRepresentativeSeq: Identity::calculateSeqIdentity (AVX2)
            # + the greedy walk -- what tcu_representatives replaces
            r = oracle.Ref(m, platform=oracle.PLATFORM_AVX2)
            r.representatives(0.8)
            del r
        else:
            ident = port.identity(m, ord("X"))
            port.greedy_clusters(ident, ns, port.cluster_order(port.sequence_lengths(m)), 0.8)
        return time.perf_counter() - t0

    for _ in range(warmup):
        one()
    times = [one() for _ in range(steps)]
    total = sum(times)
    value = pairs * L * steps / total
    sample = f"first {ns} of {n} rows x {L} cols of the seeded {args.workload} alignment per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * total / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: pairwise identity {n}x{L} synthetic protein MSA "
                               f"(RepresentativeTrimmer identity matrix)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "note": "trimAl/pytrimal statistics are single-threaded (SURVEY F9); "
                                 f"host has {os.cpu_count()} logical cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(workload, L, seed, n_full):
    """AVX2 reference (or the port) on a bounded row sample, ~10-20 s of one core."""
    import oracle
    from pytrimal_b200.synthetic import synthetic_msa
    kind = "reference" if oracle.ref_available() else "port"
    rate = 5.0e9 if kind == "reference" else 0.8e9
    ns = int(min(n_full, math.sqrt(2.0 * rate * 12.0 / L)))
    m = synthetic_msa(ns, L, seed)
    t0 = time.perf_counter()
    if kind == "reference":
        r = oracle.Ref(m, platform=oracle.PLATFORM_AVX2)
        r.representatives(0.8)
    else:
        port = oracle.Port()
        ident = port.identity(m, ord("X"))
        port.greedy_clusters(ident, ns, port.cluster_order(port.sequence_lengths(m)), 0.8)
    dt = time.perf_counter() - t0
    return {"value": ns * (ns - 1) // 2 * L / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "seconds": dt,
            "sample": f"first {ns} of {n_full} rows x {L} cols of the seeded {workload} alignment, "
                      "one pass of Cleaner::calculateRepresentativeSeq(0.8) = "
                      "Identity::calculateSeqIdentity (AVX2) + the greedy walk"}


def similarity_line(pb, CONFIGS, synthetic_msa):
    """The other half of BASELINE.json's metric (identity + similarity): one pass of
    Similarity::calculateVectors over the C3 alignment (10 000 x 5 000, BASELINE configs[2])
    through the host-buffer C ABI, identity matrix resident on the device.  Reported beside
    the headline, not folded into `value`: the statistic is bound by the sequential fp32 add
    chain the reference's order mandates, so no roofline fraction applies."""
    import numpy as np
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
    return {"workload": f"C3: per-column similarity {n}x{L} (AutomaticTrimmer strict*), identity resident in HBM",
            "value": P * L / (best_k * 1e-3), "unit": UNIT, "kernel": "tcu::k_similarity2",
            "kernel_ms": best_k, "call_ms_host_buffers": best_w * 1e3,
            "ns_per_chain_step": best_k * 1e6 / P,
            "columns_cut_by_gap_rule": int((g.astype(np.float32) >= np.float32(0.8) * np.float32(L)).sum()),
            "roofline": None,
            "note": "latency-bound: one dependent fp32 add per pair and column in the reference's "
                    "order (floor 4 cycles per step); bit-identical to the reference"}



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
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import pytrimal_b200 as pb
    from pytrimal_b200 import _lib
    from pytrimal_b200.sharding import band_partition
    from pytrimal_b200.synthetic import CONFIGS, synthetic_msa

    if pb.device_count() < 1:
        raise SystemExit("bench.py needs a B200: libtrimal_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n, L, seed = CONFIGS[args.workload]
    m = synthetic_msa(n, L, seed)
    if args.rows:
        m = m[: args.rows].copy()
        n = args.rows
    pairs_total = n * (n - 1) // 2
    X = ord("X")
    lib = pb.load()

    # pinned host copy of the rows (e2e uploads come from pinned memory)
    host_rows = torch.from_numpy(m).pin_memory()
    host_np = host_rows.numpy()

    band_rows = lib.tcu_identity_band_rows()
    bounds = band_partition(n, world)
    b0, b1 = bounds[rank], bounds[rank + 1]
    off0 = lib.tcu_identity_row_offset(n, band_rows * b0)
    off1 = lib.tcu_identity_row_offset(n, min(band_rows * b1, n))
    my_pairs = off1 - off0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident steps ---------------------------------
    dev = pb.DeviceAlignment(pb.Alignment.from_matrix(host_np), device=local_rank)
    out = torch.empty(max(my_pairs, 1), dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local_rank))

    def step():
        dev.identity_prepare(X)                 # K0: pack (1 kernel)
        dev.identity_device(b0, b1, out.data_ptr())   # K1 (1 kernel)

    for _ in range(args.warmup):
        step()
    dev.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_ms = []
    e0.record(stream)
    for _ in range(args.steps):
        step()
        # per-step kernel time from the library's own events on the same stream
    e1.record(stream)
    dev.sync()
    barrier()
    clock_note = None
    if rank == 0 and len(sampler.rows) < 3:
        # the timed region was shorter than a few 100 ms sampling periods (many GPUs):
        # keep the same steps running, untimed, until the sampler has seen the clocks
        # under this load
        clock_note = ("timed region shorter than the sampling period: clocks sampled over the "
                      "same steps repeated untimed right after it")
        t_end = time.perf_counter() + 1.5
        while len(sampler.rows) < 4 and time.perf_counter() < t_end:
            step()
            dev.sync()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None and clock_note:
        clocks["note"] = clock_note
    barrier()
    ms_total = e0.elapsed_time(e1)
    # duration of the dominant kernel alone (last step), CUDA events inside the library
    t = dev.timings
    kernel_ms, pack_ms = t["kernel_ms"], t["pack_ms"]

    tmax = torch.tensor([ms_total, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total_max, kernel_ms_max = tmax.tolist()
    value = pairs_total * L * args.steps / (ms_total_max * 1e-3)

    # ---------------- end-to-end through the host-buffer C ABI ----------------
    # (a) the call the CUDA platform makes for this configuration's trimmer
    comm = None
    if world > 1:
        comm = pb.Communicator.from_torch(local_rank)
    reps_host = torch.empty(n, dtype=torch.int32).pin_memory()
    reps_ptr = C.cast(reps_host.data_ptr(), C.POINTER(C.c_int))
    nreps = C.c_int(0)
    e2e_launches = [0]

    e2e_phases = []

    def e2e_step():
        h = C.c_void_p()
        tp0 = time.perf_counter()
        _lib.check(lib.tcu_msa_create_strided(C.c_void_p(host_rows.data_ptr()), n, L, L, local_rank,
                                              C.byref(h)))
        tp1 = time.perf_counter()
        try:
            if comm is None:
                _lib.check(lib.tcu_representatives(h, None, X, C.c_float(0.8), reps_ptr,
                                                   C.byref(nreps)))
            else:
                _lib.check(lib.tcu_representatives_all(h, comm._h, None, X, C.c_float(0.8),
                                                       reps_ptr, C.byref(nreps)))
            tp2 = time.perf_counter()
            t = _lib.Timings()
            lib.tcu_msa_timings(h, C.byref(t))
            e2e_launches[0] = t.kernel_launches
        finally:
            lib.tcu_msa_destroy(h)
        e2e_phases.append({"create_ms": 1e3 * (tp1 - tp0), "call_ms": 1e3 * (tp2 - tp1),
                           "destroy_ms": 1e3 * (time.perf_counter() - tp2),
                           "h2d_ms": t.h2d_ms, "pack_ms": t.pack_ms, "kernel_ms": t.kernel_ms,
                           "d2h_ms": t.d2h_ms, "comm_ms": t.comm_ms})

    # The interpreter's cyclic GC is kept out of the wall-clock regions (as timeit does):
    # with torch imported a full collection takes hundreds of ms and would land at random
    # inside a 25 ms step.
    import gc
    e2e_step()  # warm-up (pinned pool, allocations)
    gc.collect()
    gc.disable()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    gc.enable()
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = pairs_total * L * args.e2e_steps / te.item()
    e2e_reps = int(nreps.value)
    if comm is not None:
        comm.close()

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

    band_step()
    gc.collect()
    gc.disable()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        band_step()
    barrier()
    band_s = time.perf_counter() - t0
    gc.enable()
    tb = torch.tensor([band_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    band_value = pairs_total * L * args.e2e_steps / tb.item()

    # spot-check the host copy against the device-resident one
    same = bool(torch.equal(host_out[: min(my_pairs, 1 << 20)],
                            out[: min(my_pairs, 1 << 20)].cpu()))

    if rank == 0:
        peaks = measured_peaks()
        peak_tops = 2.0 * peaks["bf16_tflops"]          # int8 dense = 2x bf16 dense
        my_tiles_pairs = my_pairs if world == 1 else None
        # rank 0's launch processes its own band: use the max-over-ranks duration with
        # the per-rank share of the pairs (bands hold equal pair counts)
        achieved_tops = OPS_PER_PAIR_COLUMN * (pairs_total / world) * L / (kernel_ms_max * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {
                "workload": f"{args.workload}: pairwise identity {n}x{L} synthetic protein MSA "
                            "(RepresentativeTrimmer identity matrix), K0 pack + K1 identity per step",
                "pairs": pairs_total, "columns": L, "parallelism": f"row-block bands x{world}",
                "l2": "no explicit flush: each step writes %.2f GB of identities per GPU, >> 126 MB L2"
                      % (4.0 * my_pairs / 1e9),
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n) * int(L),
                    "d2h_bytes_per_step": int(4 * n + 4 * e2e_reps + 4), "steps": args.e2e_steps,
                    "representatives": e2e_reps, "gpu_launches_per_step": e2e_launches[0],
                    "last_step_phases_ms": {k: round(v, 3) for k, v in e2e_phases[-1].items()},
                    "step_ms": [round(ph["create_ms"] + ph["call_ms"] + ph["destroy_ms"], 2)
                                for ph in e2e_phases[1:]],
                    "create_ms_per_step": [round(ph["create_ms"], 2) for ph in e2e_phases[1:]],
                    "api": "tcu_msa_create_strided + tcu_representatives%s (pinned host buffers): "
                           "Cleaner::calculateRepresentativeSeq(0.8) of the RepresentativeTrimmer, "
                           "identity matrix consumed in HBM" % ("_all" if world > 1 else "")},
            "e2e_matrix_to_host": {"value": band_value, "unit": UNIT,
                                   "h2d_bytes_per_step": int(n) * int(L),
                                   "d2h_bytes_per_step": int(4 * my_pairs), "steps": args.e2e_steps,
                                   "matches_device_result": same,
                                   "api": "tcu_msa_create_strided + tcu_identity_band"},
            "gpu_launches": 2 * args.steps,
            "roofline": {
                "bound": "tensor", "achieved": achieved_tops, "peak": peak_tops, "unit": "TFLOP/s",
                "frac": achieved_tops / peak_tops, "traffic": NCU_TRAFFIC_BYTES if world == 1 and not args.rows and args.workload == "C4" else None,
                "kernel": "tcu::k_identity2<5,true>", "kernel_ms": kernel_ms_max, "pack_ms": pack_ms,
                "note": "algorithmic int8 tensor ops = 42 per pair-column (SURVEY 8d); peak = 2 x "
                        "bf16 dense, " + peaks["source"] + "; the kernel counts hits on "
                        "the LOP3/POPC integer pipes (bit-plane formulation, 5 LOP3 per 32 "
                        "pair-columns) and the both-gap counts with tcgen05 kind::i8 UMMAs",
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, L, seed, n)
        if world == 1 and not args.rows and not args.no_similarity and not args.no_cpu_baseline:
            try:
                line["similarity"] = similarity_line(pb, CONFIGS, synthetic_msa)
            except Exception as exc:  # never lose the headline line over the secondary figure
                line["similarity"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)

    dev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
