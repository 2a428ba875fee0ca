"""K1 / K0 / clustering timings on the seeded C4 alignment (device events), for A/B runs."""
import sys, json
sys.path.insert(0, "/root/repo")
import pytrimal_b200 as pb
from pytrimal_b200.synthetic import CONFIGS, synthetic_msa
n, L, seed = CONFIGS["C4"]
m = synthetic_msa(n, L, seed)
out = {}
with pb.DeviceAlignment(m) as d:
    ks, ps = [], []
    fs = []
    for _ in range(4):
        d.identity_on_device(ord("X"))
        fs.append(round(d.timings["kernel_ms"], 3))
    out["identity_float"] = {"kernel_ms(K1)": fs, "pack_ms(K0)": round(d.timings["pack_ms"], 3)}
    for _ in range(4):
        reps = d.representatives(0.8, indet=ord("X"))
        ks.append(round(d.timings["kernel_ms"], 3)); ps.append(round(d.timings["pack_ms"], 3))
    out["representatives"] = {"n": len(reps), "kernel_ms(K1+greedy)": ks, "pack_ms(K0+mirror)": ps}
print(json.dumps(out))
