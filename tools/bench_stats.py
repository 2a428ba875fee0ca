#!/usr/bin/env python
"""Throughput of the other three statistics (gaps K3, spurious K2, similarity K4)
through the host-buffer C ABI, one JSON line per (statistic, workload).

bench.py times the headline identity kernel; this tool gives the SURVEY 8(d)
numbers for the rest: device kernel time (CUDA events inside the library), the
wall time of the C-ABI call with host buffers, and the roofline that applies
(HBM for K2/K3; none for K4, which is bound by the dependent fp32 add chain).

    python tools/bench_stats.py [--only gaps,spurious,similarity] [--workloads C2,C3,C5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="gaps,spurious,similarity")  # + "consumers" (C3/C4)
    ap.add_argument("--workloads", default="C2,C3,C5")
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--rows", type=int, default=0)
    args = ap.parse_args()
    import pytrimal_b200 as pb
    from pytrimal_b200.synthetic import CONFIGS, synthetic_msa

    if pb.device_count() < 1:
        raise SystemExit("needs a B200: libtrimal_cuda has no CPU fallback")
    hbm, src = peaks()
    only = set(args.only.split(","))
    X = ord("X")
    smx = pb.SimilarityMatrix.aa()
    for wl in args.workloads.split(","):
        n, L, seed = CONFIGS[wl]
        if args.rows:
            n = min(n, args.rows)
        m = synthetic_msa(n, L, seed)
        P = n * (n - 1) // 2
        with pb.DeviceAlignment(m) as d:
            if "gaps" in only:
                best_k, best_w = 1e30, 1e30
                for _ in range(args.repeats + 1):
                    t0 = time.perf_counter()
                    g, hist, mx = d.gaps()
                    best_w = min(best_w, time.perf_counter() - t0)
                    best_k = min(best_k, d.timings["kernel_ms"])
                byts = n * L + 4 * L
                print(json.dumps({
                    "stat": "gaps", "workload": f"{wl} {n}x{L}", "kernel_ms": best_k,
                    "call_ms": best_w * 1e3, "cells_per_s": n * L / (best_k * 1e-3),
                    "roofline": {"bound": "hbm", "achieved": byts / (best_k * 1e-3) / 1e9,
                                 "peak": hbm, "unit": "GB/s",
                                 "frac": byts / (best_k * 1e-3) / 1e9 / hbm, "peak_source": src},
                    "max_gaps": int(mx)}), flush=True)
            if "spurious" in only:
                best_k, best_w = 1e30, 1e30
                for _ in range(args.repeats + 1):
                    t0 = time.perf_counter()
                    sp = d.spurious(0.5, indet=X)
                    best_w = min(best_w, time.perf_counter() - t0)
                    best_k = min(best_k, d.timings["kernel_ms"])
                byts = 2 * n * L + 4 * n
                print(json.dumps({
                    "stat": "spurious", "workload": f"{wl} {n}x{L}", "kernel_ms": best_k,
                    "call_ms": best_w * 1e3,
                    "ordered_pair_col_per_s": n * (n - 1) * L / (best_k * 1e-3),
                    "roofline": {"bound": "hbm", "achieved": byts / (best_k * 1e-3) / 1e9,
                                 "peak": hbm, "unit": "GB/s",
                                 "frac": byts / (best_k * 1e-3) / 1e9 / hbm, "peak_source": src,
                                 "note": "closed-form column-histogram mode: 2 passes over n*L bytes"},
                    "mean": float(sp.mean())}), flush=True)
            if "consumers" in only and wl != "C5":
                # the Cleaner.cpp walks over the resident identity matrix
                d.identity_on_device(X)
                lengths = d.sequence_lengths()
                t_len = d.timings["kernel_ms"]
                t0 = time.perf_counter()
                order = pb.cluster_order(lengths)
                t_order = (time.perf_counter() - t0) * 1e3
                rec = {"stat": "consumers", "workload": f"{wl} {n}x{L}",
                       "lengths_kernel_ms": t_len, "cluster_order_host_ms": t_order}
                for thr in (0.8, 0.5):
                    best = None
                    for _ in range(args.repeats):
                        t0 = time.perf_counter()
                        k = d.clusters(order, thr, count_only=True)
                        w = (time.perf_counter() - t0) * 1e3
                        t = d.timings
                        if best is None or w < best["call_ms"]:
                            best = {"clusters": k, "call_ms": w, "bits_kernel_ms": t["pack_ms"],
                                    "greedy_ms": t["kernel_ms"], "launches": t["kernel_launches"]}
                    best["bits_roofline"] = {
                        "bound": "hbm", "achieved": (4 * P + n * n / 8) / (best["bits_kernel_ms"] * 1e-3) / 1e9,
                        "peak": hbm, "unit": "GB/s",
                        "frac": (4 * P + n * n / 8) / (best["bits_kernel_ms"] * 1e-3) / 1e9 / hbm}
                    rec[f"clusters_thr{thr}"] = best
                for upper in (False, True):
                    best = 1e30
                    for _ in range(args.repeats):
                        d.identity_row_stats(upper_only=upper)
                        best = min(best, d.timings["kernel_ms"])
                    reads = (2 if not upper else 1) * 4 * P
                    rec["row_stats_upper_ms" if upper else "row_stats_full_ms"] = best
                    rec["row_stats_upper_gbs" if upper else "row_stats_full_gbs"] = reads / (best * 1e-3) / 1e9
                t0 = time.perf_counter()
                name, avg_seq, max_seq = d.select_method()
                rec["select_method"] = {"result": name, "call_ms": (time.perf_counter() - t0) * 1e3}
                t0 = time.perf_counter()
                thr_k, runs = d.cutpoint_clusters(max(2, n // 100), order=order)
                rec["cutpoint_clusters"] = {"clusters": max(2, n // 100), "threshold": float(thr_k),
                                            "clusterings": runs,
                                            "call_ms": (time.perf_counter() - t0) * 1e3}
                # whole Cleaner::calculateRepresentativeSeq through one C-ABI call, host buffers
                best = None
                for _ in range(args.repeats):
                    with pb.DeviceAlignment(m) as d2:
                        t0 = time.perf_counter()
                        reps = d2.representatives(0.8, indet=X)
                        w = (time.perf_counter() - t0) * 1e3
                        if best is None or w < best["call_ms"]:
                            best = {"call_ms": w, "representatives": len(reps), **d2.timings}
                rec["tcu_representatives_thr0.8"] = best
                rec["pair_col_per_s_through_representatives"] = P * L / (best["call_ms"] * 1e-3)
                print(json.dumps(rec), flush=True)
            if "similarity" in only and wl != "C5":
                g, _, _ = d.gaps()
                t0 = time.perf_counter()
                d.identity(X, keep_on_device=True)
                id_s = time.perf_counter() - t0
                best_k, best_w = 1e30, 1e30
                for _ in range(args.repeats):
                    t0 = time.perf_counter()
                    mdk, num, den = d.similarity(smx, gaps=g, indet=X)
                    best_w = min(best_w, time.perf_counter() - t0)
                    best_k = min(best_k, d.timings["kernel_ms"])
                cut = int((g.astype(np.float32) >= np.float32(0.8) * np.float32(L)).sum())
                print(json.dumps({
                    "stat": "similarity", "workload": f"{wl} {n}x{L}", "kernel_ms": best_k,
                    "call_ms": best_w * 1e3, "identity_call_ms": id_s * 1e3,
                    "pair_col_per_s": P * L / (best_k * 1e-3),
                    "chain_steps_per_column": P, "ns_per_chain_step": best_k * 1e6 / max(P, 1),
                    "columns_cut_by_gap_rule": cut,
                    "roofline": None,
                    "note": "bound by the sequential fp32 add chain the reference's order mandates "
                            "(SURVEY F3/8d): report chain-step latency, no roofline fraction",
                    "mdk_mean": float(mdk.mean())}), flush=True)


if __name__ == "__main__":
    main()
