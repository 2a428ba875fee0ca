import sys; sys.path.insert(0, ".")
import pytrimal_b200 as pb
from pytrimal_b200.synthetic import CONFIGS, synthetic_msa
n, L, seed = CONFIGS["C4"]
m = synthetic_msa(n, L, seed)
with pb.DeviceAlignment(m) as d:
    for _ in range(3):
        d.identity_on_device(ord("X")); print(d.timings["kernel_ms"])
