"""ncu target: three iterations of rnla_lsqr_dev on a 1M x 2000 matrix (16 GB): gemv_n / gemv_t and the fused vector kernels."""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
m, n = 1_000_000, 2000
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 31, 4, m, n, 0, pA, lda)); rt.synchronize()
db = torch.randn(m, dtype=torch.float64, device="cuda"); dx = torch.zeros(n, dtype=torch.float64, device="cuda")
res = _lib.LsqrResult(); hist = np.zeros(8)
_lib.check(lib.rnla_lsqr_dev(pA, lda, m, n, C.c_void_p(db.data_ptr()), 0.0, 0.0, 0.0, 0.0, 3, 0, None, C.c_void_p(dx.data_ptr()),
                             C.byref(res), C.c_void_p(hist.ctypes.data), hist.size, None))
rt.synchronize()
print("itn", res.itn)
