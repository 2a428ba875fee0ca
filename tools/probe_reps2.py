import sys; sys.path.insert(0, "/root/repo")
import pytrimal_b200 as pb
from pytrimal_b200.synthetic import CONFIGS, synthetic_msa
n, L, seed = CONFIGS["C4"]
m = synthetic_msa(n, L, seed)
with pb.DeviceAlignment(m) as d:
    d.representatives(0.8, indet=ord("X"))
    d.identity_on_device(ord("X"))
    d.representatives(0.8, indet=ord("X"))   # launch index 2: threshold mode
    d.identity_on_device(ord("X"))           # launch index 3: float mode
