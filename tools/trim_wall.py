#!/usr/bin/env python
"""trim() wall time through the reference's own Python API (BASELINE.json's second metric):
pytrimal built with the CUDA platform (integration/build_pytrimal.py), platform="cuda"
against platform="avx2" on the same synthetic alignment, one JSON line per configuration.

    python tools/trim_wall.py [--configs C2,C3,C4,C5] [--avx2-rows 4000]

The AVX2 run is limited to the first --avx2-rows rows where the full size would take
minutes on one core (the statistics are single-threaded); its pair-column
rate is what scales.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "integration", "_build", "pkg"))

TRIMMERS = {
    "C2": ("ManualTrimmer", dict(gap_threshold=0.9, similarity_threshold=0.1, window=3)),
    "C3": ("AutomaticTrimmer", dict(method="strictplus")),
    "C4": ("RepresentativeTrimmer", dict(identity_threshold=0.8)),
    "C5": ("OverlapTrimmer", dict(sequence_overlap=0.5, residue_overlap=0.5)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C2,C4,C5")
    ap.add_argument("--avx2-rows", type=int, default=4000)
    ap.add_argument("--repeats", type=int, default=2)
    args = ap.parse_args()
    import pytrimal
    from pytrimal_b200.synthetic import CONFIGS, synthetic_msa

    for cfg in args.configs.split(","):
        n, L, seed = CONFIGS[cfg]
        m = synthetic_msa(n, L, seed)
        names = [b"s%d" % i for i in range(n)]
        t0 = time.perf_counter()
        ali = pytrimal.Alignment(names, [bytes(r) for r in m])
        build_s = time.perf_counter() - t0
        cls, kwargs = TRIMMERS[cfg]
        trimmer = getattr(pytrimal, cls)(platform="cuda", **kwargs)
        times = []
        for _ in range(args.repeats + 1):
            t0 = time.perf_counter()
            out = trimmer.trim(ali)
            times.append(time.perf_counter() - t0)
        rec = {"config": cfg, "shape": [n, L], "trimmer": cls, "kwargs": kwargs,
               "alignment_build_s": build_s, "cuda_trim_s_first": times[0],
               "cuda_trim_s_best": min(times[1:]), "kept_sequences": len(out.sequences),
               "kept_columns": len(out.sequences[0]) if len(out.sequences) else 0}
        # C3 also runs the scalar similarity loop on the CPU (1e9 pair-col/s): fewer rows there
        rows = min(n, args.avx2_rows if cfg != "C3" else min(args.avx2_rows, 1500))
        sub = pytrimal.Alignment(names[:rows], [bytes(r) for r in m[:rows]])
        cpu = getattr(pytrimal, cls)(platform="avx2", **kwargs)
        cpu.trim(sub)                       # platform applies from the 2nd call
        t0 = time.perf_counter()
        ref = cpu.trim(sub)
        rec["avx2_trim_s"] = time.perf_counter() - t0
        rec["avx2_rows"] = rows
        gpu_sub = getattr(pytrimal, cls)(platform="cuda", **kwargs).trim(sub)
        rec["identical_to_avx2_on_subsample"] = (
            list(gpu_sub.names) == list(ref.names) and list(gpu_sub.sequences) == list(ref.sequences))
        t0 = time.perf_counter()
        getattr(pytrimal, cls)(platform="cuda", **kwargs).trim(sub)
        rec["cuda_trim_s_subsample"] = time.perf_counter() - t0
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
