"""Kernel timeline of one rand_svd step (CUPTI through torch.profiler): name, start offset, duration and the idle gap
before each kernel, for everything that is not a streaming GEMM pass.  Usage: python tools/trace_step.py [rows]"""
import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
lib = _lib.load(); rt.init(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
n = 20000
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 5, 9, m, n, 0, pA, lda)); rt.synchronize()
for _ in range(2):
    ld.rand_svd_dev(dA, 100, 10); rt.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ld.rand_svd_dev(dA, 100, 10); rt.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev_end = t0
tot_gap = 0.0; tot_small = 0.0
for e in ev:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = e.time_range.start - prev_end
    prev_end = max(prev_end, e.time_range.end)
    big = d > 5000
    if not big:
        tot_small += d
    tot_gap += max(gap, 0)
    print(f"{s/1e3:10.3f} ms  dur {d:9.1f} us  gap {gap:8.1f} us  {e.name[:90]}")
print("total span ms", (prev_end - t0) / 1e3, "sum gaps ms", tot_gap / 1e3, "sum small kernels ms", tot_small / 1e3)
