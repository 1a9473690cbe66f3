"""Target for `ncu --set full`: the dominant kernels at full size, one process.
  1. block sparse-sign sketch at BASELINE config 4 (1M x 2000): d = 8000 and d = 4000 (zeta = 8, w = 4)
  2. one rand_svd at BASELINE config 2 (200000 x 20000, k = 100, s = 10, materialised Omega)
Use -k regex:"gemm_.._kernel|saso_block_kernel" and a launch count of ~16."""
import sys
sys.path.insert(0, ".")
import torch
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
lib = _lib.load(); rt.init(0)
m, n = 1000000, 2000
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 77, 9, m, n, 0, pA, lda)); rt.synchronize()
for d in (8000, 4000):
    dS = rt.empty_colmajor(d, n); pS, lds = rt.dev_ptr_ld(dS)
    _lib.check(lib.rnla_sketch_apply_dev(2, 4, 5, d, 8, pA, lda, m, n, 0, pS, lds)); rt.synchronize()
    print("saso_block d =", d, rt.timings()[-1], flush=True)
del dA, dS; torch.cuda.empty_cache()
m, n = 200000, 20000
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 5, 9, m, n, 0, pA, lda)); rt.synchronize()
U, S, Vt = ld.rand_svd_dev(dA, 100, 10, rt.make_options(fused_sketch=0)); rt.synchronize()
print("rand_svd", rt.timings(), flush=True)
