"""one wave of the integer MMA kernels on k SMs (m = 128 k rows, n = 20000, N = 110): per-tile time against the number of active SMs
tells whether the sweeps are bound per SM (L2 -> SM port) or globally (L2 / HBM)"""
import sys, ctypes as C
import torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
n, N = 20000, 110
_lib.check(lib.rnla_set_kernel_timing(1))
trans = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for k in (18, 37, 74, 111, 148, 296):
    m = 128 * k
    if trans:                                      # A^T B: tiles are 128-column blocks of A, one chunk of rows per CTA
        A = rt.empty_colmajor(n, m); A.copy_(torch.randn((n, m), device="cuda", dtype=torch.float64))
        B = rt.empty_colmajor(n, N); B.copy_(torch.randn((n, N), device="cuda", dtype=torch.float64))
        Cm = rt.empty_colmajor(m, N)
    else:
        A = rt.empty_colmajor(m, n); A.copy_(torch.randn((m, n), device="cuda", dtype=torch.float64))
        B = rt.empty_colmajor(n, N); B.copy_(torch.randn((n, N), device="cuda", dtype=torch.float64))
        Cm = rt.empty_colmajor(m, N)
    pa, lda = rt.dev_ptr_ld(A); pb, ldb = rt.dev_ptr_ld(B); pc, ldc = rt.dev_ptr_ld(Cm)
    torch.cuda.synchronize()
    _lib.check(lib.rnla_i8_gemm_dev(trans, 7, 1, pa, lda, A.shape[0], A.shape[1], pb, ldb, N, pc, ldc, 3))
    t = [(a, b) for a, b in rt.timings() if a.startswith("k:")]
    lo = min(b for a, b in t if "planes 3" in a); hi = min(b for a, b in t if "planes 7" in a)
    by_lo = 313 * (3 * 8192 + 3 * 7 * 1024) * k; by_hi = 313 * (7 * 8192 + 7 * 7 * 1024) * k
    print(f"SMs {k:4d}: sweep(3 planes) {lo:.3f} ms = {by_lo / lo * 1e-6 / k:.1f} GB/s per SM; sweep(7 planes) {hi:.3f} ms = {by_hi / hi * 1e-6 / k:.1f} GB/s per SM, "
          f"{by_hi / hi * 1e-9:.2f} TB/s total", flush=True)
