"""The one-pass kernel (csrc/normal_pass.cu) on a B200: values against torch, timing at the BASELINE config-4 shape against the two
streaming mat-vec kernels, and blendenpik with one pass per iteration against the two-pass iteration (RNLA_ONEPASS=0)."""
import sys, os, json, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
out = {}
P = lambda t: C.c_void_p(t.data_ptr())

def normal_pass(dA, x, cq, y, cy, store):
    m, n = dA.shape
    pA, lda = rt.dev_ptr_ld(dA)
    t = torch.empty(n + 1, dtype=torch.float64, device="cuda")
    u = torch.empty(m, dtype=torch.float64, device="cuda") if store else None
    torch.cuda.synchronize()
    _lib.check(lib.rnla_normal_pass_dev(pA, lda, m, n, P(x), cq, P(y) if y is not None else None, cy, P(u) if store else None, P(t)))
    rt.synchronize()
    return t, u

worst = 0.0
for (m, n) in [(1, 1), (31, 7), (32, 8), (33, 9), (1000, 100), (4099, 255), (4096, 256), (5000, 257), (7777, 500), (3001, 1000), (9000, 2000), (2500, 2048), (100000, 640), (60001, 640), (9001, 2000)]:
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    dA = rt.empty_colmajor(m, n); dA.copy_(torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g))
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    pA, lda = rt.dev_ptr_ld(dA)
    if not lib.rnla_normal_pass_supported(pA, lda, m, n):
        print("unsupported", m, n, lda); continue
    for (cq, yy, cy, store) in [(1.0, None, 0.0, False), (-1.0, y, 1.0, True), (1.0, y, -0.37, True)]:
        t, u = normal_pass(dA, x, cq, yy, cy, store)
        ur = cq * (dA @ x) + (cy * yy if yy is not None else 0.0)
        tr = dA.t() @ ur
        scale_t = (dA.abs().t() @ ur.abs()).max().item() + 1e-300
        e_t = ((t[:n] - tr).abs().max().item()) / scale_t
        e_uu = abs(t[n].item() - (ur @ ur).item()) / ((ur @ ur).item() + 1e-300)
        e_u = 0.0
        if store:
            e_u = (u - ur).abs().max().item() / ((dA.abs() @ x.abs()).max().item() + 1.0)
        worst = max(worst, e_t, e_uu, e_u)
        if max(e_t, e_uu, e_u) > 1e-13:
            print("MISMATCH", m, n, cq, cy, store, e_t, e_uu, e_u)
    # reproducible
    t1, _ = normal_pass(dA, x, 1.0, None, 0.0, False); t2, _ = normal_pass(dA, x, 1.0, None, 0.0, False)
    assert torch.equal(t1, t2), "not reproducible"
    del dA
out["values_worst_rel_err"] = worst
print("values: worst relative error", worst, flush=True)

if "small" not in sys.argv:
    m, n = 1000000, 2000
    dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
    _lib.check(lib.rnla_sketch_fill_dev(0, 0, 77, 9, m, n, 0, pA, lda)); rt.synchronize()
    x = torch.randn(n, dtype=torch.float64, device="cuda"); y = torch.empty(m, dtype=torch.float64, device="cuda")
    t = torch.empty(n + 1, dtype=torch.float64, device="cuda"); u2 = torch.empty(n, dtype=torch.float64, device="cuda")
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
    ms1 = ev(lambda: _lib.check(lib.rnla_normal_pass_dev(pA, lda, m, n, P(x), 1.0, None, 0.0, None, P(t))))
    def two():
        _lib.check(lib.rnla_gemv_dev(pA, lda, m, n, 0, P(x), P(y)))
        _lib.check(lib.rnla_gemv_dev(pA, lda, m, n, 1, P(y), P(u2)))
    ms2 = ev(two)
    out["pass_1Mx2000"] = {"one_pass_ms": ms1, "one_pass_GBps": 8.0 * m * n / ms1 * 1e-6, "two_kernels_ms": ms2, "two_kernels_GBps_per_kernel": 16.0 * m * n / ms2 * 1e-6}
    print(out["pass_1Mx2000"], flush=True)
    rel = ((t[:n] - u2).abs().max() / u2.abs().max()).item()
    print("one-pass vs two kernels:", rel, flush=True)
    # blendenpik end to end (tools/perf_c4_c5.py c4solve)
    dA.mul_(torch.logspace(0, -4, n, dtype=torch.float64, device="cuda"))
    xt = (torch.rand(n, 1, dtype=torch.float64, device="cuda") * 200 - 100)
    db = rt.empty_colmajor(m, 1); db.copy_(dA @ xt + 1e-4 * torch.randn(m, 1, dtype=torch.float64, device="cuda"))
    res = {}
    for mode in ["0", "1"]:
        os.environ["RNLA_ONEPASS"] = mode
        dx = rt.empty_colmajor(n, 1)
        it = C.c_int64(0); cv = C.c_int32(0)
        def run():
            _lib.check(lib.rnla_blendenpik_overdetermined_dev(pA, lda, m, n, P(db), 1e-8, 200, 4.0, 2, 0, 8, P(dx), C.byref(it), C.byref(cv)))
        run(); rt.synchronize()
        best = 1e30
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); run(); rt.synchronize(); best = min(best, time.perf_counter() - t0)
        r = dA.t() @ (db - dA @ dx)
        res[mode] = dx.clone()
        out["blendenpik_onepass_" + mode] = {"ms": best * 1e3, "iterations": int(it.value), "converged": bool(cv.value),
                                             "rel_err_x": float(torch.linalg.vector_norm(dx - xt) / torch.linalg.vector_norm(xt)),
                                             "normal_eq_residual": float(torch.linalg.vector_norm(r) / torch.linalg.vector_norm(dA.t() @ db)),
                                             "phases": rt.timings()}
        print(mode, out["blendenpik_onepass_" + mode], flush=True)
    out["blendenpik_x_rel_diff_one_vs_two_pass"] = float(torch.linalg.vector_norm(res["0"] - res["1"]) / torch.linalg.vector_norm(res["0"]))
    print("x one-pass vs two-pass:", out["blendenpik_x_rel_diff_one_vs_two_pass"])
print(json.dumps(out))
