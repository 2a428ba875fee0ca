"""GPU probe (not product code): where the end-to-end identity call spends its time, raw
PCIe D2H rates, and int8/fp8 GEMM peaks of the box (torch/cuBLASLt) for the roofline."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pytrimal_b200 as pb
from pytrimal_b200 import _lib
from pytrimal_b200.synthetic import CONFIGS, synthetic_msa

out = {}
lib = pb.load()
n, L, seed = CONFIGS["C4"]
if len(sys.argv) > 1:
    n = int(sys.argv[1])
m = synthetic_msa(n, L, seed)
pairs = n * (n - 1) // 2
host_rows = torch.from_numpy(m).pin_memory()
host_out = torch.empty(pairs, dtype=torch.float32).pin_memory()
out_ptr = C.cast(host_out.data_ptr(), C.POINTER(C.c_float))
X = ord("X")


def wall(f):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = f()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, r


steps = []
for it in range(3):
    h = C.c_void_p()
    t_create, _ = wall(lambda: _lib.check(lib.tcu_msa_create_strided(
        C.c_void_p(host_rows.data_ptr()), n, L, L, 0, C.byref(h))))
    t_band, _ = wall(lambda: _lib.check(lib.tcu_identity_band(h, None, None, X, 0, -1, out_ptr)))
    tm = _lib.Timings()
    lib.tcu_msa_timings(h, C.byref(tm))
    t_band2, _ = wall(lambda: _lib.check(lib.tcu_identity_band(h, None, None, X, 0, -1, out_ptr)))
    tm2 = _lib.Timings()
    lib.tcu_msa_timings(h, C.byref(tm2))
    t_destroy, _ = wall(lambda: lib.tcu_msa_destroy(h))
    steps.append({"create_s": t_create, "band_first_s": t_band, "band_again_s": t_band2,
                  "destroy_s": t_destroy,
                  "first": {"h2d": tm.h2d_ms, "pack": tm.pack_ms, "kernel": tm.kernel_ms, "d2h": tm.d2h_ms},
                  "again": {"h2d": tm2.h2d_ms, "pack": tm2.pack_ms, "kernel": tm2.kernel_ms, "d2h": tm2.d2h_ms}})
out["e2e_breakdown"] = steps

# raw D2H rates
dev = torch.empty(pairs, dtype=torch.float32, device="cuda")
dev.fill_(1.0)
for name, dst in (("pinned", host_out), ("pageable", torch.empty(pairs, dtype=torch.float32))):
    best = 1e9
    for _ in range(2):
        t, _ = wall(lambda: dst.copy_(dev, non_blocking=True))
        best = min(best, t)
    out[f"d2h_{name}_GBps"] = 4 * pairs / best / 1e9
# chunked pinned D2H on two streams
s = [torch.cuda.Stream(), torch.cuda.Stream()]
chunk = 64 << 20
def chunked():
    k = 0
    for o in range(0, pairs, chunk // 4):
        e = min(pairs, o + chunk // 4)
        with torch.cuda.stream(s[k & 1]):
            host_out[o:e].copy_(dev[o:e], non_blocking=True)
        k += 1
t, _ = wall(chunked)
out["d2h_pinned_chunked_2streams_GBps"] = 4 * pairs / t / 1e9
# H2D
t, _ = wall(lambda: dev.copy_(host_out, non_blocking=True))
out["h2d_pinned_GBps"] = 4 * pairs / t / 1e9
# host memcpy rate (single thread) for a staged pageable path
a = np.empty(1 << 28, np.uint8); b = np.ones(1 << 28, np.uint8)
t0 = time.perf_counter(); a[:] = b; out["host_memcpy_1thread_GBps"] = (1 << 28) / (time.perf_counter() - t0) / 1e9
# cudaHostRegister cost on a fresh pageable buffer
buf = np.empty(pairs, np.float32)
cudart = torch.cuda.cudart()
t0 = time.perf_counter(); rc = cudart.cudaHostRegister(buf.ctypes.data, buf.nbytes, 0); t_reg = time.perf_counter() - t0
out["cudaHostRegister_5GB_s"] = t_reg; out["cudaHostRegister_rc"] = int(rc)
if int(rc) == 0:
    tb = torch.from_numpy(buf)
    t, _ = wall(lambda: tb.copy_(dev, non_blocking=True))
    out["d2h_registered_GBps"] = 4 * pairs / t / 1e9
    t0 = time.perf_counter(); cudart.cudaHostUnregister(buf.ctypes.data); out["cudaHostUnregister_s"] = time.perf_counter() - t0
del dev

# tensor peaks for the roofline denominator
def best_of(f, reps=10):
    f(); torch.cuda.synchronize()
    b = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        b = min(b, e0.elapsed_time(e1))
    return b
N = 8192
try:
    A = torch.randint(-2, 2, (N, N), dtype=torch.int8, device="cuda")
    B = torch.randint(-2, 2, (N, N), dtype=torch.int8, device="cuda").t()
    ms = best_of(lambda: torch._int_mm(A, B))
    out["int8_tops_cublaslt_8192"] = 2 * N**3 / (ms * 1e-3) / 1e12
except Exception as e:
    out["int8_error"] = repr(e)
try:
    A = torch.randn(N, N, device="cuda").to(torch.float8_e4m3fn)
    B = torch.randn(N, N, device="cuda").to(torch.float8_e4m3fn).t()
    one = torch.tensor(1.0, device="cuda")
    ms = best_of(lambda: torch._scaled_mm(A, B, scale_a=one, scale_b=one, out_dtype=torch.bfloat16))
    out["fp8_tflops_cublaslt_8192"] = 2 * N**3 / (ms * 1e-3) / 1e12
except Exception as e:
    out["fp8_error"] = repr(e)
A = torch.randn(N, N, device="cuda", dtype=torch.bfloat16); B = torch.randn(N, N, device="cuda", dtype=torch.bfloat16)
ms = best_of(lambda: A @ B)
out["bf16_tflops_8192"] = 2 * N**3 / (ms * 1e-3) / 1e12
print(json.dumps(out, indent=1))
