"""Time of the digit split of A alone (i8:rowmax + i8:split phases of a level-2 rand_svd) on the bench matrix."""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
import bench
lib = _lib.load(); rt.init(0)
m, n = 200000, 20000
sig = bench.planted_sigma()
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_generate_lowrank_dev(pA, lda, m, n, 0, m, bench.R0, sig.ctypes.data_as(C.c_void_p), 1e-7, 1234))
for level in (3, 1):
    opts = rt.make_options(range_passes_int8=level)
    best = None
    for _ in range(4):
        U, S, Vt = ld.rand_svd_dev(dA, 100, 10, opts); rt.synchronize()
        ph = dict(rt.timings())
        t = ph.get("i8:rowmax(A)", 0.0) + ph["i8:split(A)"]
        best = t if best is None else min(best, t)
    print(f"level {level}: split {best:.2f} ms; sigma[0..2] {S.cpu().numpy()[:3]}", flush=True)
