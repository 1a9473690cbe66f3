"""Next rows at BASELINE config-4 scale on one B200, device-resident timing:
  cqrrpt:  sap_chol_qrcp of a 1M x 512 and a 1M x 2000 matrix, block sparse-sign sketch d = 2n
  sas:     sketched_least_squares_qr on 65536 x 2000 (d = rows / 4 = 16384, block sparse-sign sketch)
  id:      osid_randomised (Column) on 200k x 20k (the config-2 matrix), k = 100
  cur:     cur_randomised on 200k x 2000, k = 100
  saddle:  sketch_saddle_point_precondition on 1M x 1000, sf = 2
  qrcp:    the pivoted QR kernel alone on sketch-sized matrices (us per step)
"""
import sys, json, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
out = {}
P = lambda t: C.c_void_p(t.data_ptr())
def timed(fn, reps=2):
    fn(); rt.synchronize()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); rt.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
which = sys.argv[1:] or ["qrcp", "cqrrpt", "sas", "id", "cur", "saddle"]
def gauss(m, n, seed):
    dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
    _lib.check(lib.rnla_sketch_fill_dev(0, 0, seed, 9, m, n, 0, pA, lda)); rt.synchronize()
    return dA, pA, lda
if "qrcp" in which:
    for (m, n, steps) in [(4000, 2000, 2000), (1024, 512, 512), (100, 20000, 100), (8000, 2000, 2000)]:
        dA, pA, lda = gauss(m, n, 3)
        W = rt.empty_colmajor(m, n); pW, ldw = rt.dev_ptr_ld(W)
        perm = torch.zeros(n, dtype=torch.int64, device="cuda")
        def run():
            W.copy_(dA)
            _lib.check(lib.rnla_qrcp_dev(pW, ldw, m, n, steps, P(perm), None, m, 0))
        t = timed(run)
        out[f"qrcp_{m}x{n}_{steps}"] = {"ms": t * 1e3, "us_per_step": t * 1e6 / steps}
        print(f"qrcp_{m}x{n}_{steps}", out[f"qrcp_{m}x{n}_{steps}"], flush=True)
        del dA, W
if "cqrrpt" in which:
    for (m, n) in [(1000000, 512), (1000000, 2000)]:
        dA, pA, lda = gauss(m, n, 77)
        dA.mul_(torch.logspace(0, -3, n, dtype=torch.float64, device="cuda"))
        Q = rt.empty_colmajor(m, n); pQ, ldq = rt.dev_ptr_ld(Q)
        R = rt.empty_colmajor(n, n); pR, ldr = rt.dev_ptr_ld(R)
        J = torch.zeros(n, dtype=torch.int64, device="cuda"); k = C.c_int64(0)
        d = 2 * n
        def run():
            _lib.check(lib.rnla_sap_chol_qrcp_dev(pA, lda, m, n, d, 2, 0, 8, pQ, ldq, pR, ldr, P(J), C.byref(k)))
        t = timed(run, reps=1)
        ph = rt.timings()
        kk = int(k.value)
        # checks on a row sample (the full product would need another 16 GB)
        rows = torch.randint(0, m, (4096,), device="cuda")
        rec = Q[rows, :kk] @ R[:kk, :] - dA[rows][:, J]
        orth = Q[:, :64].t() @ Q[:, :kk]; orth[:, :64] -= torch.eye(64, dtype=torch.float64, device="cuda")
        out[f"cqrrpt_{m}x{n}"] = {"ms": t * 1e3, "k": kk, "recon_rel": float(rec.norm() / dA[rows].norm()), "orth_max": float(orth.abs().max()),
                                  "phases": ph}
        print(f"cqrrpt_{m}x{n}", out[f"cqrrpt_{m}x{n}"], flush=True)
        del dA, Q, R; torch.cuda.empty_cache()
if "sas" in which:
    m, n = 65536, 2000          # d = rows / 4 = 16384, the largest block sparse-sign sketch
    dA, pA, lda = gauss(m, n, 78)
    xt = torch.rand(n, 1, dtype=torch.float64, device="cuda") * 200 - 100
    db = rt.empty_colmajor(m, 1); db.copy_(dA @ xt + 1e-2 * torch.randn(m, 1, dtype=torch.float64, device="cuda"))
    dx = rt.empty_colmajor(n, 1)
    for which_s, name in ((0, "qr"),):
        def run():
            _lib.check(lib.rnla_sketched_least_squares_dev(which_s, pA, lda, m, n, P(db), 2, 0, 8, P(dx)))
        t = timed(run, reps=1)
        out[f"sas_{name}_65536x2000"] = {"ms": t * 1e3, "rel_err_x": float((dx - xt).norm() / xt.norm()), "phases": rt.timings()}
        print(f"sas_{name}", out[f"sas_{name}_65536x2000"], flush=True)
    del dA, db; torch.cuda.empty_cache()
if "id" in which:
    m, n, k = 200000, 20000, 100
    L, _, _ = gauss(m, k, 5); Rm, _, _ = gauss(k, n, 6)
    dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
    pL, ldl = rt.dev_ptr_ld(L); pRm, ldrm = rt.dev_ptr_ld(Rm)
    _lib.check(lib.rnla_gemm_nn_dev(pL, ldl, m, k, pRm, ldrm, n, pA, lda)); rt.synchronize()
    X = rt.empty_colmajor(k, n); pX, ldx = rt.dev_ptr_ld(X)
    J = torch.zeros(k, dtype=torch.int64, device="cuda")
    def run():
        _lib.check(lib.rnla_osid_randomised_dev(pA, lda, m, n, k, 1, None, pX, ldx, P(J)))
    t = timed(run)
    rows = torch.randint(0, m, (2048,), device="cuda")
    S = dA[rows]
    out["osid_randomised_col_200kx20k_k100"] = {"ms": t * 1e3, "A_stream_GBps": 8.0 * m * n / t * 1e-9,
                                                "recon_rel_rowsample": float((S[:, J] @ X - S).norm() / S.norm()), "phases": rt.timings()}
    print(out["osid_randomised_col_200kx20k_k100"], flush=True)
    del dA, L, Rm, X; torch.cuda.empty_cache()
if "cur" in which:
    m, n, k = 200000, 2000, 100
    L, _, _ = gauss(m, k, 5); Rm, _, _ = gauss(k, n, 6)
    dA = rt.empty_colmajor(m, n); dA.copy_(L @ Rm); pA, lda = rt.dev_ptr_ld(dA)
    U = rt.empty_colmajor(k, k); pU, ldu = rt.dev_ptr_ld(U)
    I = torch.zeros(k, dtype=torch.int64, device="cuda"); J = torch.zeros(k, dtype=torch.int64, device="cuda")
    def run():
        _lib.check(lib.rnla_cur_dev(pA, lda, m, n, k, 1, None, P(J), pU, ldu, P(I)))
    t = timed(run)
    err = float((dA[:, J] @ U @ dA[I, :] - dA).norm() / dA.norm())
    out["cur_randomised_200kx2000_k100"] = {"ms": t * 1e3, "recon_rel": err, "phases": rt.timings()}
    print(out["cur_randomised_200kx2000_k100"], flush=True)
    del dA, L, Rm; torch.cuda.empty_cache()
if "saddle" in which:
    m, n = 1000000, 1000
    dA, pA, lda = gauss(m, n, 79)
    xt = torch.rand(n, 1, dtype=torch.float64, device="cuda") * 200 - 100
    db = rt.empty_colmajor(m, 1); db.copy_(dA @ xt + 1e-2 * torch.randn(m, 1, dtype=torch.float64, device="cuda"))
    dc = rt.empty_colmajor(n, 1); dc.copy_(torch.rand(n, 1, dtype=torch.float64, device="cuda") * 20 - 10)
    dx = rt.empty_colmajor(n, 1); dy = rt.empty_colmajor(m, 1)
    it = C.c_int64(0); cv = C.c_int32(0)
    def run():
        _lib.check(lib.rnla_sketch_saddle_point_precondition_dev(pA, lda, m, n, P(db), P(dc), 0.0, 1e-6, 200, 2.0, P(dx), P(dy),
                                                                 C.byref(it), C.byref(cv)))
    t = timed(run, reps=1)
    g = dA.t() @ (dA @ dx - db) + dc       # gradient of the objective at x (mu = 0)
    out["saddle_1Mx1000_sf2"] = {"ms": t * 1e3, "iterations": int(it.value), "converged": bool(cv.value),
                                 "grad_rel": float(g.norm() / (dA.t() @ db).norm()), "phases": rt.timings()}
    print(out["saddle_1Mx1000_sf2"], flush=True)
print(json.dumps(out))
