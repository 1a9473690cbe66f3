"""The one-pass kernel against the two streaming mat-vec kernels at a given shape (default 1M x 2000): python tools/onepass_exp.py [m n]"""
import sys, os, json, ctypes as C
import torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
P = lambda t: C.c_void_p(t.data_ptr())
m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, int(sys.argv[2]) if len(sys.argv) > 2 else 2000
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 77, 9, m, n, 0, pA, lda)); rt.synchronize()
x = torch.randn(n, dtype=torch.float64, device="cuda"); t = torch.empty(n + 1, dtype=torch.float64, device="cuda")
def ev(fn, reps=10):
    fn(); rt.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(lib.rnla_stream())
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(reps): fn()
        e1.record()
    rt.synchronize(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
os.environ["RNLA_NP_VERBOSE"] = "1"
_lib.check(lib.rnla_normal_pass_dev(pA, lda, m, n, P(x), 1.0, None, 0.0, None, P(t)))
del os.environ["RNLA_NP_VERBOSE"]
ms = ev(lambda: _lib.check(lib.rnla_normal_pass_dev(pA, lda, m, n, P(x), 1.0, None, 0.0, None, P(t))))
print(f"one pass (A x, A^T (A x), |A x|^2): {ms:.3f} ms = {8.0 * m * n / ms * 1e-6:.0f} GB/s of A", flush=True)
y = torch.empty(m, dtype=torch.float64, device="cuda"); u2 = torch.empty(n, dtype=torch.float64, device="cuda")
def two():
    _lib.check(lib.rnla_gemv_dev(pA, lda, m, n, 0, P(x), P(y)))
    _lib.check(lib.rnla_gemv_dev(pA, lda, m, n, 1, P(y), P(u2)))
ms2 = ev(two)
print(f"two streaming kernels: {ms2:.3f} ms = {16.0 * m * n / ms2 * 1e-6:.0f} GB/s per kernel", flush=True)
# the same product on an operand with an ODD leading dimension (even / odd columns through a tensor map each)
if m % 2 == 0:
    dB = rt.empty_colmajor(m + 1, n); pB, ldb = rt.dev_ptr_ld(dB)
    _lib.check(lib.rnla_sketch_fill_dev(0, 0, 77, 9, m + 1, n, 0, pB, ldb)); rt.synchronize()
    ms3 = ev(lambda: _lib.check(lib.rnla_normal_pass_dev(pB, ldb, m + 1, n, P(x), 1.0, None, 0.0, None, P(t))))
    print(f"one pass, odd leading dimension: {ms3:.3f} ms = {8.0 * (m + 1) * n / ms3 * 1e-6:.0f} GB/s of A", flush=True)
