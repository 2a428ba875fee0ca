import sys, numpy as np
sys.path.insert(0, "/root/repo")
import pytrimal_b200 as pb
from pytrimal_b200.synthetic import CONFIGS, synthetic_msa
n, L, seed = CONFIGS["C4"]
m = synthetic_msa(n, L, seed)
with pb.DeviceAlignment(m) as d:
    for _ in range(2):
        reps = d.representatives(0.8, indet=ord("X"))
    print(len(reps), d.timings)
