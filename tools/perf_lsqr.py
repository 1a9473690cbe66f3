"""lsqr (reference src/solvers.rs:115-278) on device buffers at BASELINE config-4 scale (1M x 2000 f64 = 16 GB) and on a
thinner system: time per iteration against the HBM floor of the two streamed passes over A (2 x 8 m n bytes per iteration)."""
import sys, json, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
out = {}
for (m, n, iters) in [(1_000_000, 2000, 30), (1_000_000, 500, 30), (200_000, 20_000, 20)]:
    dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
    _lib.check(lib.rnla_sketch_fill_dev(0, 0, 31, 4, m, n, 0, pA, lda)); rt.synchronize()
    db = torch.randn(m, dtype=torch.float64, device="cuda")
    dx = torch.zeros(n, dtype=torch.float64, device="cuda")
    res = _lib.LsqrResult(); hist = np.zeros(iters)
    def run():
        _lib.check(lib.rnla_lsqr_dev(pA, lda, m, n, C.c_void_p(db.data_ptr()), 0.0, 0.0, 0.0, 0.0, iters, 0, None,
                                     C.c_void_p(dx.data_ptr()), C.byref(res), C.c_void_p(hist.ctypes.data), hist.size, None))
        rt.synchronize()
    run()
    best = 1e30
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); run(); best = min(best, time.perf_counter() - t0)
    per_it = best / (int(res.itn) + 0.5)     # iterations actually run (the machine-precision tests may stop early); the start-up applies A^T once
    key = f"lsqr_{m}x{n}"
    out[key] = {"iterations": int(res.itn), "ms_total": best * 1e3, "ms_per_iteration": per_it * 1e3,
                "a_stream_gbs": 2 * 8.0 * m * n / per_it / 1e9}
    print(key, out[key], flush=True)
    del dA, db; torch.cuda.empty_cache()
print(json.dumps(out))
