#!/usr/bin/env python
"""The reference's whole bundled corpus as fixtures (tier ii).

    python tests/golden/make_golden_corpus.py            # needs /root/reference + oracle/_ref

Every aligned FASTA input of vendor/trimal/dataset/ (86 of them; the unaligned ones cannot be
trimmed) is run through the REAL reference (unmodified trimAl, AVX2 platform, oracle/_ref):
the seven automatic methods of scripts/generate_trimmed_msas.sh, the four statistics and the
Cleaner walks.  Stored per input under tests/golden/corpus/:

  <stem>.xz      the input matrix (raw bytes, LZMA) -- /root/reference does not exist on the
                 GPU box
  INDEX.json     shape, alignment type, keep-masks of every method (hex bit strings), SHA-256
                 of the gap counts / packed identities / MDK / spurious vectors, the
                 representatives at 0.75, selectMethod, and how each keep-mask compared with
                 the reference's committed expected output (dataset/trimmed_msas/<method>/)

The upstream comparison script skips examples 014/028/032/041 for the similarity methods
(scripts/compare_trimmed_msas.sh:9: the AVX2 build differs there from the goldens, SURVEY
F4); they are kept here because the parity target is the AVX2 platform itself -- the
"expected" column records the mismatch instead of hiding the input.
"""
import json
import lzma
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from digest import sha256_hex  # noqa: E402
from pytrimal_b200 import io as tio  # noqa: E402

DATASET = "/root/reference/vendor/trimal/dataset"
OUT = os.path.join(HERE, "corpus")
AUTO = ["gappyout", "strict", "strictplus", "automated1", "automated2", "nogaps", "noallgaps"]


def mask_hex(mask):
    return np.packbits(mask != -1).tobytes().hex()


def expected_status(method, stem, names, m, ks, kr):
    path = os.path.join(DATASET, "trimmed_msas", method, stem + ".fasta")
    if not os.path.exists(path):
        return "missing"
    if os.path.getsize(path) == 0:
        return "empty"
    en, es = tio.read_alignment(path)
    rows, cols = np.nonzero(ks != -1)[0], np.nonzero(kr != -1)[0]
    gn = [names[i].split()[0] for i in rows]
    gs = [bytes(m[i, cols]) for i in rows]
    return "ok" if [x.split()[0] for x in en] == gn and es == gs else "differs"


def main():
    assert oracle.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    os.makedirs(OUT, exist_ok=True)
    index = {}
    files = sorted(f for f in os.listdir(DATASET)
                   if f.startswith("example.") and f.endswith((".fasta", ".fa")))
    for f in files:
        t0 = time.time()
        names, seqs = tio.read_alignment(os.path.join(DATASET, f))
        if len({len(s) for s in seqs}) != 1:
            continue                                   # not aligned: nothing to trim
        stem = os.path.splitext(f)[0]
        m = tio.to_matrix(seqs)
        n, L = m.shape
        try:
            r = oracle.Ref(m)
        except ValueError:
            continue                                   # the reference rejects the input
        e = {"file": f, "shape": [int(n), int(L)], "type": int(r.alignment_type),
             "indet": chr(r.indet), "matrix_sha256": sha256_hex(m)}
        e["gaps_sha256"] = sha256_hex(r.gaps()[0])
        ident = r.identity()
        e["identity_sha256"], e["identity_count"] = sha256_hex(ident), int(ident.size)
        try:
            e["mdk_sha256"] = sha256_hex(r.similarity()[0])
        except ValueError:
            e["mdk_sha256"] = None                     # IncorrectSymbol / UndefinedSymbol
        for ov in (0.5, 0.8):
            e[f"spurious_{int(ov * 100)}_sha256"] = sha256_hex(oracle.Ref(m).spurious(ov))
        if n >= 2:
            e["repr_75_sha256"] = sha256_hex(oracle.Ref(m).representatives(0.75))
            e["select_method"] = int(oracle.Ref(m).select_method())
        e["methods"] = {}
        for method in AUTO:
            try:
                ks, kr = oracle.Ref(m).trim(method)
            except ValueError:
                e["methods"][method] = {"error": True}
                continue
            e["methods"][method] = {"seq": mask_hex(ks), "res": mask_hex(kr),
                                    "kept": [int((ks != -1).sum()), int((kr != -1).sum())],
                                    "expected": expected_status(method, stem, names, m, ks, kr)}
        with open(os.path.join(OUT, stem + ".xz"), "wb") as fh:
            fh.write(lzma.compress(m.tobytes(), preset=6))
        index[stem] = e
        print(stem, m.shape, "%.1f s" % (time.time() - t0),
              {k: v.get("expected", "error") for k, v in e["methods"].items()}, flush=True)
    with open(os.path.join(OUT, "INDEX.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden_corpus.py",
                   "reference": "pytrimal 0.8.5 / vendored trimAl 2.0 RC, AVX2 platform, oracle/_ref",
                   "inputs": index}, fh, indent=0)


if __name__ == "__main__":
    main()
