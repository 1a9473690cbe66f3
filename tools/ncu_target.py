"""Small target for `ncu --set full`: rand_svd at m x 20000 (default m = 40000), k=100, s=10, materialised then fused Omega."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
lib = _lib.load(); rt.init(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
n = 20000
sig = np.concatenate([np.logspace(0, -3, 100), np.full(100, 1e-5)])
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 5, 9, m, n, 0, pA, lda))    # plain Gaussian data: no GEMM launches before the driver
for fused in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,0,1").split(",")]:
    U, S, Vt = ld.rand_svd_dev(dA, 100, 10, rt.make_options(fused_sketch=fused)); rt.synchronize()
print("done", rt.timings())
