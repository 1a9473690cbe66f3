"""ncu target: the one-pass kernel (csrc/normal_pass.cu) twice on a 1M x 2000 matrix (16 GB), then the two streaming mat-vec kernels."""
import sys, ctypes as C
import torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
P = lambda t: C.c_void_p(t.data_ptr())
m, n = 1_000_000, 2000
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 31, 4, m, n, 0, pA, lda)); rt.synchronize()
x = torch.randn(n, dtype=torch.float64, device="cuda"); y = torch.randn(m, dtype=torch.float64, device="cuda")
t = torch.empty(n + 1, dtype=torch.float64, device="cuda"); u = torch.empty(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    _lib.check(lib.rnla_normal_pass_dev(pA, lda, m, n, P(x), 1.0, None, 0.0, None, P(t)))
_lib.check(lib.rnla_normal_pass_dev(pA, lda, m, n, P(x), -1.0, P(y), 1.0, P(y), P(t)))
_lib.check(lib.rnla_gemv_dev(pA, lda, m, n, 0, P(x), P(y)))
_lib.check(lib.rnla_gemv_dev(pA, lda, m, n, 1, P(y), P(u)))
rt.synchronize()
print("done")
