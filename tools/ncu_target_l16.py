"""ncu target: the streaming GEMMs in the HBM-bound regime (l = 16) at 200000 x 20000."""
import sys
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
lib = _lib.load(); rt.init(0)
m, n = 200000, 20000
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 5, 9, m, n, 0, pA, lda)); rt.synchronize()
U, S, Vt = ld.rand_svd_dev(dA, 10, 6, rt.make_options(fused_sketch=0)); rt.synchronize()
print(rt.timings())
