"""time of one A B / A^T B product on the integer tensor cores at the headline size, per sweep (kernel-level timing entries)"""
import sys, ctypes as C
import torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
import bench
lib = _lib.load(); rt.init(0)
m, n, N = 200000, 20000, int(sys.argv[1]) if len(sys.argv) > 1 else 110
planes = int(sys.argv[2]) if len(sys.argv) > 2 else 7
sig = bench.planted_sigma()
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_generate_lowrank_dev(pA, lda, m, n, 0, m, bench.R0, sig.ctypes.data_as(C.c_void_p), 1e-7, 1234))
_lib.check(lib.rnla_set_kernel_timing(1))
for trans in (0, 1):
    kb = m if trans else n
    B = rt.empty_colmajor(kb, N); B.copy_(torch.randn((kb, N), device="cuda", dtype=torch.float64))
    Cm = rt.empty_colmajor(n if trans else m, N)
    pb, ldb = rt.dev_ptr_ld(B); pc, ldc = rt.dev_ptr_ld(Cm)
    torch.cuda.synchronize()
    _lib.check(lib.rnla_i8_gemm_dev(trans, planes, 1, pA, lda, m, n, pb, ldb, N, pc, ldc, 3))
    print("trans", trans, [(a, round(b, 3)) for a, b in rt.timings() if a.startswith("k:") or a.startswith("i8:A")], flush=True)
