#!/usr/bin/env python
"""trim() wall time through the reference's own Python API (BASELINE.json's second metric):
pytrimal built with the CUDA platform (integration/build_pytrimal.py), platform="cuda", on the
full-size seeded alignments of BASELINE's configurations, one JSON line per configuration.

    python tools/trim_wall.py [--configs C2,C3,C4,C5] [--devices all|0,1,..] [--ingest]

Every result is compared with what the unmodified reference (AVX2 platform) produced for the
same input at FULL size (tests/golden/full/*.npz, tests/golden/make_golden_full.py; the
reference needs 6-40 minutes of one core per configuration, so it is not rerun here -- its
own wall time is in the fixture).  --devices sets TRIMAL_CUDA_DEVICES (one process, several
GPUs); --ingest sets TRIMAL_CUDA_INGEST=1 (symbol validation of Alignment() on the device, the
upload it makes is reused by trim()).
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

CASES = {
    "C2": ("ManualTrimmer", dict(gap_threshold=0.5, similarity_threshold=0.001, window=3), "C2", "gt50_st001"),
    "C2literal": ("ManualTrimmer", dict(gap_threshold=0.9, similarity_threshold=0.1, window=3), "C2", "literal"),
    "C3": ("AutomaticTrimmer", dict(method="strictplus"), "C3.strictplus", "strictplus"),
    "C3strict": ("AutomaticTrimmer", dict(method="strict"), "C3.strict", "strict"),
    "C3gappyout": ("AutomaticTrimmer", dict(method="gappyout"), "C3", "gappyout"),
    "C4": ("RepresentativeTrimmer", dict(identity_threshold=0.8), "C4", "maxidentity80"),
    "C5": ("OverlapTrimmer", dict(sequence_overlap=0.5, residue_overlap=0.5), "C5.seq0.5", "overlap_seq0p5"),
    "C5seq50": ("OverlapTrimmer", dict(sequence_overlap=50, residue_overlap=0.5), "C5.seq50", "overlap_seq50"),
    "C5seq75": ("OverlapTrimmer", dict(sequence_overlap=75, residue_overlap=0.7), "C5.seq75.res0.7", "overlap_seq75_res0p7"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C2,C3,C4,C5")
    ap.add_argument("--devices", default="")
    ap.add_argument("--ingest", action="store_true")
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    if args.devices:
        os.environ["TRIMAL_CUDA_DEVICES"] = args.devices
    if args.ingest:
        os.environ["TRIMAL_CUDA_INGEST"] = "1"
    import numpy as np
    import pytrimal
    import pytrimal_b200 as pb
    from pytrimal_b200.synthetic import CONFIGS, synthetic_msa

    cache = {}
    for case in args.configs.split(","):
        cls, kwargs, job, tag = CASES[case]
        cfg = case[:2]
        n, L, seed = CONFIGS[cfg]
        if cfg not in cache:
            cache.clear()
            cache[cfg] = synthetic_msa(n, L, seed)
        m = cache[cfg]
        names = [b"s%d" % i for i in range(n)]
        seqs = [bytes(r) for r in m]
        build = []
        for _ in range(2):
            t0 = time.perf_counter()
            ali = pytrimal.Alignment(names, seqs)
            build.append(time.perf_counter() - t0)
        trimmer = getattr(pytrimal, cls)(platform="cuda", **kwargs)
        times, out = [], None
        for _ in range(args.repeats + 1):
            t0 = time.perf_counter()
            try:
                out = trimmer.trim(ali)
            except Exception as exc:      # the reference raises when nothing is left
                out = exc
            times.append(time.perf_counter() - t0)
        # a NEW alignment object each time, as a caller that trims many alignments does: build,
        # trim, drop (the upload belongs to the alignment; its device buffers go back to the
        # library's pool when the alignment dies and are reused by the next one)
        del ali
        fresh = []
        for _ in range(args.repeats):
            t0 = time.perf_counter()
            a2 = pytrimal.Alignment(names, seqs)
            t1 = time.perf_counter()
            try:
                getattr(pytrimal, cls)(platform="cuda", **kwargs).trim(a2)
            except Exception:
                pass
            t2 = time.perf_counter()
            del a2
            fresh.append((t2 - t0, t1 - t0, t2 - t1))
        rec = {"config": case, "shape": [n, L], "trimmer": cls, "kwargs": kwargs,
               "devices": pb.get_devices(), "ingest_on_device": bool(args.ingest),
               "alignment_build_s": min(build), "cuda_trim_s_first": times[0],
               "cuda_trim_s_best": min(times[1:]),
               "fresh_alignment_plus_trim_s": min(fresh)[0], "fresh_build_s": min(fresh)[1],
               "fresh_trim_s": min(fresh)[2]}
        if isinstance(out, Exception):
            rec["result"] = "error: " + type(out).__name__
            kept = (0, 0)
        else:
            kept = (len(out.sequences), len(out.sequences[0]) if len(out.sequences) else 0)
            rec["kept_sequences"], rec["kept_columns"] = kept
        gpath = os.path.join(ROOT, "tests", "golden", "full", job + ".npz")
        if os.path.exists(gpath):
            g = np.load(gpath)
            ks, kr = g[f"trim_{tag}_seq"], g[f"trim_{tag}_res"]
            rows, cols = np.nonzero(ks != -1)[0], np.nonzero(kr != -1)[0]
            rec["reference_avx2_seconds_full_size_job"] = float(g["reference_seconds"])
            if len(cols) == 0 or len(rows) == 0:
                rec["identical_to_reference_full_size"] = kept[0] == 0 or kept[1] == 0
            elif isinstance(out, Exception):
                rec["identical_to_reference_full_size"] = False
            else:
                sub = m[np.ix_(rows, cols)]
                got = list(out.sequences)
                rec["identical_to_reference_full_size"] = bool(
                    list(out.names) == [names[i] for i in rows] and len(got) == len(rows) and
                    all(got[k].encode() == bytes(sub[k]) for k in range(len(rows))))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
