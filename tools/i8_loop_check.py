import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
import bench
lib = _lib.load(); rt.init(0)
m, n = 200000, 20000
sig = bench.planted_sigma()
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_generate_lowrank_dev(pA, lda, m, n, 0, m, bench.R0, sig.ctypes.data_as(C.c_void_p), 1e-7, 1234))
if "roofs" in sys.argv:
    fp64 = C.c_double(0); hbm = C.c_double(0)
    _lib.check(lib.rnla_measure_roofs(C.byref(fp64), C.byref(hbm), 4 << 30))
for level in (2, 1, 2):
    opts = rt.make_options(range_passes_int8=level)
    ts = []
    for it in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        U, S, Vt = ld.rand_svd_dev(dA, 100, 10, opts); rt.synchronize(); torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    free, total = torch.cuda.mem_get_info()
    print(level, [f"{t:.1f}" for t in ts], f"free {free/2**30:.1f} GiB of {total/2**30:.1f}", flush=True)
