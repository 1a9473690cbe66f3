"""ncu target: rand_svd in the library's default mode (auto: every pass on the INT8 tensor cores, 55-bit split) at 200000 x 20000;
argv[1] = mode (auto | fp64 | level2)."""
import sys, ctypes as C
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
import bench
lib = _lib.load(); rt.init(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
m, n = 200000, 20000
sig = bench.planted_sigma()
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_generate_lowrank_dev(pA, lda, m, n, 0, m, bench.R0, sig.ctypes.data_as(C.c_void_p), 1e-7, 1234))
U, S, Vt = ld.rand_svd_dev(dA, 100, 10, rt.make_options(range_passes_int8=bench.MODES[mode])); rt.synchronize()
print(rt.timings())
