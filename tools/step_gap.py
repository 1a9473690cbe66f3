"""per-call wall time against the sum of the library's phases (what of a rand_svd call is not inside a phase?)  argv: rows"""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
import bench
lib = _lib.load(); rt.init(0)
m, n = int(sys.argv[1]), 20000
sig = bench.planted_sigma()
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_generate_lowrank_dev(pA, lda, m, n, 0, m, bench.R0, sig.ctypes.data_as(C.c_void_p), 1e-7, 1234))
opts = rt.make_options()
for i in range(7):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = ld.rand_svd_dev(dA, 100, 10, opts); rt.synchronize(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    ph = rt.timings()
    print(f"call {i}: wall {1e3 * (t1 - t0):7.2f} ms, phases {sum(v for _, v in ph):7.2f} ms", flush=True)
