"""Target for `ncu` on the block sparse-sign sketch kernel: C4 shape (1M x 2000 f64), one launch per (d, zeta, w) given as
d:zeta:w arguments (default 8000:8:4)."""
import sys
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
m, n = 1000000, 2000
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 77, 9, m, n, 0, pA, lda)); rt.synchronize()
for spec in (sys.argv[1:] or ["8000:8:4"]):
    d, zeta, w = (int(x) for x in spec.split(":"))
    dS = rt.empty_colmajor(d, n); pS, lds = rt.dev_ptr_ld(dS)
    _lib.check(lib.rnla_sketch_apply_dev(2, w, 5, d, zeta, pA, lda, m, n, 0, pS, lds)); rt.synchronize()
    print(spec, rt.timings()[-1])
