#!/usr/bin/env python
"""Full-size fixtures from the REAL reference (oracle/_ref, AVX2 platform).

    python tests/golden/make_golden_full.py JOB [JOB ...]      # one process per job is fine
    python tests/golden/make_golden_full.py index              # merge -> full/INDEX.json

BASELINE.json's configurations at their full sizes (synthetic, seeded:
pytrimal_b200.synthetic.CONFIGS).  The reference needs minutes to most of an
hour per job on one core (it is single-threaded), so this runs once in the build
container and the GPU tests compare the CUDA results with what is stored here:
small result arrays verbatim (keep-masks, representatives, MDK, spurious vector),
the packed identity arrays as digests (tests/golden/digest.py).

Jobs
  C2            ManualTrimmer(gap_threshold=.9, similarity_threshold=.1, window=3) literal, and a
                non-degenerate variant of it, on 1 000 x 2 000; gaps / identity / MDK
  C3            gaps, identity digests, MDK (un-windowed), gappyout keep-masks on 10 000 x 5 000
  C3.strict  C3.strictplus  C3.automated1      keep-masks of those AutomaticTrimmer methods
  C4            RepresentativeTrimmer(identity_threshold=.8) on 50 000 x 1 000: representatives in
                creation order, keep-masks, identity digests, selectMethod
  C5            Overlap::calculateSpuriousVector(0.5) on 100 000 x 2 000 (the vector itself)
  C5.seq0.5  C5.seq50    OverlapTrimmer(sequence_overlap=0.5 | 50, residue_overlap=.5) keep-masks
  C5.seq75.res0.7        OverlapTrimmer(sequence_overlap=75, residue_overlap=.7): a selective variant
                         (drops about a quarter of the sequences)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from digest import block_sums, sha256_hex  # noqa: E402
from pytrimal_b200.synthetic import CONFIGS, synthetic_msa  # noqa: E402

OUT = os.path.join(HERE, "full")


def msa_of(cfg):
    n, L, seed = CONFIGS[cfg]
    return synthetic_msa(n, L, seed)


def identity_digests(out, ident, n):
    s, w = block_sums(ident, n, 128)
    out["identity_block_sum"], out["identity_block_wsum"] = s, w
    out["identity_sha256"] = np.array(sha256_hex(ident))
    out["identity_count"] = np.int64(ident.size)


def job_C2(out):
    m = msa_of("C2")
    n, L = m.shape
    r = oracle.Ref(m)
    out["gaps"] = r.gaps()[0]
    ident = r.identity()
    out["identity"] = ident                      # 2 MB: stored verbatim
    identity_digests(out, ident, n)
    out["mdk"] = r.similarity()[0]
    r3 = oracle.Ref(m)
    r3.set_windows(3, 3)
    g3 = r3.gaps()[1]
    out["gaps_w3"] = g3
    out["mdk_w3"] = r3.similarity()[1]
    # the literal configuration: _gap_threshold = 1 - 0.9 (src/pytrimal/_trimal.pyx:1589)
    ks, kr = oracle.Ref(m).trim("manual", [1 - 0.9, 0.1, -1, 3, -1, -1])
    out["trim_literal_seq"], out["trim_literal_res"] = ks, kr
    # the same trimmer with thresholds that keep a non-trivial part of this alignment
    for tag, gt, st in (("gt30_st01", 0.3, 0.1), ("gt50_st001", 0.5, 0.001), ("gt20_st0", 0.2, 0.0)):
        ks, kr = oracle.Ref(m).trim("manual", [1 - gt, st, -1, 3, -1, -1])
        out[f"trim_{tag}_seq"], out[f"trim_{tag}_res"] = ks, kr
        out[f"trim_{tag}_params"] = np.array([gt, st, 3], np.float64)


def job_C3(out):
    m = msa_of("C3")
    n, L = m.shape
    r = oracle.Ref(m)
    out["gaps"] = r.gaps()[0]
    ident = r.identity()
    identity_digests(out, ident, n)
    del ident
    out["mdk"] = r.similarity()[0]
    ks, kr = r.trim("gappyout")
    out["trim_gappyout_seq"], out["trim_gappyout_res"] = ks, kr


def job_C3_method(method):
    def run(out):
        m = msa_of("C3")
        ks, kr = oracle.Ref(m).trim(method)
        out[f"trim_{method}_seq"], out[f"trim_{method}_res"] = ks, kr
    return run


def job_C4(out):
    m = msa_of("C4")
    n, L = m.shape
    r = oracle.Ref(m)
    out["representatives_80"] = r.representatives(0.8)   # computes the identity matrix (AVX2)
    ident = r.identity()
    identity_digests(out, ident, n)
    del ident
    out["select_method"] = np.int32(r.select_method())
    out["gaps"] = r.gaps()[0]
    ks, kr = r.trim("representative", [-1, 0.8])          # shares the matrix through the mold
    out["trim_maxidentity80_seq"], out["trim_maxidentity80_res"] = ks, kr


def job_C5(out):
    m = msa_of("C5")
    out["spurious_50"] = oracle.Ref(m).spurious(0.5)
    out["gaps"] = oracle.Ref(m).gaps()[0]


def job_C5_trim(seq_overlap, res_overlap=0.5):
    def run(out):
        m = msa_of("C5")
        ks, kr = oracle.Ref(m).trim("overlap", [res_overlap, seq_overlap])
        tag = ("%g" % seq_overlap).replace(".", "p")
        if res_overlap != 0.5:
            tag += "_res" + ("%g" % res_overlap).replace(".", "p")
        out[f"trim_overlap_seq{tag}_seq"], out[f"trim_overlap_seq{tag}_res"] = ks, kr
        out["params"] = np.array([seq_overlap, res_overlap], np.float64)
    return run


JOBS = {
    "C2": ("C2", job_C2),
    "C3": ("C3", job_C3),
    "C3.strict": ("C3", job_C3_method("strict")),
    "C3.strictplus": ("C3", job_C3_method("strictplus")),
    "C3.automated1": ("C3", job_C3_method("automated1")),
    "C4": ("C4", job_C4),
    "C5": ("C5", job_C5),
    "C5.seq0.5": ("C5", job_C5_trim(0.5)),
    "C5.seq50": ("C5", job_C5_trim(50.0)),
    # the literal configuration keeps every sequence of this alignment; a selective variant
    "C5.seq75.res0.7": ("C5", job_C5_trim(75.0, 0.7)),
}


def run_job(name):
    cfg, fn = JOBS[name]
    assert oracle.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    os.makedirs(OUT, exist_ok=True)
    t0 = time.time()
    out = {}
    fn(out)
    n, L, seed = CONFIGS[cfg]
    out["config"] = np.array(cfg)
    out["shape_seed"] = np.array([n, L, seed], np.int64)
    out["matrix_sha256"] = np.array(sha256_hex(msa_of(cfg)))
    out["reference_seconds"] = np.float64(time.time() - t0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "done in %.0f s" % (time.time() - t0), flush=True)


def make_index():
    index = {}
    for name in sorted(JOBS):
        path = os.path.join(OUT, name + ".npz")
        if not os.path.exists(path):
            continue
        g = np.load(path)
        e = {"config": str(g["config"]), "shape_seed": g["shape_seed"].tolist(),
             "matrix_sha256": str(g["matrix_sha256"]),
             "reference_seconds": round(float(g["reference_seconds"]), 1), "arrays": {}}
        for k in g.files:
            a = g[k]
            if k.startswith("trim_") and k.endswith(("_seq", "_res")):
                e["arrays"][k] = {"kept": int((a != -1).sum()), "of": int(a.size)}
            elif k == "identity_sha256":
                e["arrays"][k] = str(a)
            elif k.startswith("representatives"):
                e["arrays"][k] = {"count": int(a.size)}
        index[name] = e
    with open(os.path.join(OUT, "INDEX.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden_full.py",
                   "reference": "pytrimal 0.8.5 / vendored trimAl 2.0 RC, AVX2 platform, oracle/_ref, "
                                "one host core", "jobs": index}, f, indent=1)
    print(json.dumps(index, indent=1))


if __name__ == "__main__":
    for a in sys.argv[1:]:
        if a == "index":
            make_index()
        else:
            run_job(a)
