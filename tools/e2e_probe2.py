#!/usr/bin/env python
"""Where the time of one e2e step (create + tcu_representatives + destroy) goes."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pytrimal_b200 as pb
from pytrimal_b200 import _lib
from pytrimal_b200.synthetic import CONFIGS, synthetic_msa

use_torch = "--torch" in sys.argv
n, L, seed = CONFIGS["C4"]
m = synthetic_msa(n, L, seed)
lib = pb.load()
if use_torch:
    import torch
    host_rows = torch.from_numpy(m).pin_memory()
    ptr = host_rows.data_ptr()
    big = torch.empty(1250000000, dtype=torch.float32, device="cuda")
else:
    ptr = m.ctypes.data
reps = np.zeros(n, np.int32)
k = C.c_int(0)
for it in range(4):
    t0 = time.perf_counter()
    h = C.c_void_p()
    _lib.check(lib.tcu_msa_create_strided(C.c_void_p(ptr), n, L, L, 0, C.byref(h)))
    t1 = time.perf_counter()
    _lib.check(lib.tcu_representatives(h, None, ord("X"), C.c_float(0.8),
                                       reps.ctypes.data_as(C.POINTER(C.c_int)), C.byref(k)))
    t2 = time.perf_counter()
    t = _lib.Timings(); lib.tcu_msa_timings(h, C.byref(t))
    lib.tcu_msa_destroy(h)
    t3 = time.perf_counter()
    print(f"iter {it} torch={use_torch}: create {1e3*(t1-t0):.1f} ms, representatives {1e3*(t2-t1):.1f} ms "
          f"(h2d {t.h2d_ms:.2f} pack {t.pack_ms:.2f} kernel {t.kernel_ms:.2f} d2h {t.d2h_ms:.2f}), "
          f"destroy {1e3*(t3-t2):.1f} ms, reps {k.value}", flush=True)
