"""GPU debug helper (not product code): identity hit/dst counts vs the C oracle, separately."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
import pytrimal_b200 as pb
from pytrimal_b200.synthetic import synthetic_msa

port = oracle.Port()
X = ord("X")
bad = 0
for (n, L, seed) in [(6, 46, 1), (130, 90, 2), (300, 700, 3), (700, 1000, 4), (1500, 300, 5), (150, 66000, 6)]:
    m = synthetic_msa(n, L, seed)
    oi, oh, od = port.identity(m, X, counts=True)
    with pb.DeviceAlignment(m) as d:
        gi, gh, gd = d.identity(X, counts=True)
        t = d.timings
    eh, ed = int((gh != oh).sum()), int((gd != od).sum())
    ei = int((gi.view(np.uint32) != oi.view(np.uint32)).sum())
    print(f"n={n} L={L}: pairs={oh.size} hit_mismatch={eh} dst_mismatch={ed} ident_mismatch={ei} kernel_ms={t['kernel_ms']:.3f}")
    if ed:
        w = np.flatnonzero(gd != od)[:8]
        print("   dst got", gd[w], "want", od[w], "at", w)
    if eh:
        w = np.flatnonzero(gh != oh)[:8]
        print("   hit got", gh[w], "want", oh[w], "at", w)
    bad += eh + ed + ei
print("OK" if bad == 0 else "MISMATCH")
