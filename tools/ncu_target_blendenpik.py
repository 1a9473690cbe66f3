"""ncu target: one blendenpik call (block sparse-sign sketch, d = 4 n) on 1M x 2000 with 3 CGLS iterations: kernels of the preconditioner"""
import sys, ctypes as C
import torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
m, n = 1_000_000, 2000
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 77, 9, m, n, 0, pA, lda)); rt.synchronize()
xt = torch.rand(n, 1, dtype=torch.float64, device="cuda") * 200 - 100
db = rt.empty_colmajor(m, 1); db.copy_(dA @ xt + 1e-2 * torch.randn(m, 1, dtype=torch.float64, device="cuda"))
dx = rt.empty_colmajor(n, 1); it = C.c_int64(0); cv = C.c_int32(0)
torch.cuda.synchronize()
_lib.check(lib.rnla_blendenpik_overdetermined_dev(pA, lda, m, n, C.c_void_p(db.data_ptr()), 1e-8, 3, 4.0, 2, 0, 8, C.c_void_p(dx.data_ptr()), C.byref(it), C.byref(cv)))
rt.synchronize()
print("phases", rt.timings())
