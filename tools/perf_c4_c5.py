"""BASELINE configs 4 and 5 on one B200 (device-resident timing, CUDA events inside the library):
  C4: sketch step S*A for A = 1M x 2000 (16 GB): SASO (zeta=8, d=8000) and dense Gaussian (d=4000)
  C5: rand_evd2 (Nystrom) on a 50k x 50k SPD matrix, k=200, s=10
"""
import sys, json, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
lib = _lib.load(); rt.init(0)
out = {}
def timed(fn, reps=3):
    fn(); rt.synchronize()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); rt.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("c4", "both"):
    m, n = 1000000, 2000
    dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
    _lib.check(lib.rnla_sketch_fill_dev(0, 0, 77, 9, m, n, 0, pA, lda)); rt.synchronize()
    for d, zeta in [(8000, 8)]:
        dS = rt.empty_colmajor(d, n); pS, lds = rt.dev_ptr_ld(dS)
        t = timed(lambda: _lib.check(lib.rnla_sketch_apply_dev(1, 0, 5, d, zeta, pA, lda, m, n, 0, pS, lds)))
        out[f"c4_saso_d{d}_z{zeta}"] = {"ms": t * 1e3, "A_stream_GBps": 8.0 * m * n / t * 1e-9, "frob_ratio": float(torch.linalg.vector_norm(dS) / torch.linalg.vector_norm(dA))}
        print(out[f"c4_saso_d{d}_z{zeta}"], flush=True)
    for d, zeta, w in [(8000, 8, 4), (8000, 8, 2), (8000, 8, 1), (4000, 8, 4), (8000, 4, 4), (2000, 8, 4)]:
        dS = rt.empty_colmajor(d, n); pS, lds = rt.dev_ptr_ld(dS)
        t = timed(lambda: _lib.check(lib.rnla_sketch_apply_dev(2, w, 5, d, zeta, pA, lda, m, n, 0, pS, lds)), reps=5)
        key = f"c4_saso_block_d{d}_z{zeta}_w{w}"
        out[key] = {"ms": t * 1e3, "A_stream_GBps": 8.0 * m * n / t * 1e-9, "frob_ratio": float(torch.linalg.vector_norm(dS) / torch.linalg.vector_norm(dA)), "phases": rt.timings()[-3:]}
        print(key, out[key], flush=True)
    if "nodense" in sys.argv:
        print(json.dumps(out)); sys.exit(0)
    d = 4000
    dS = rt.empty_colmajor(d, n); pS, lds = rt.dev_ptr_ld(dS)
    t = timed(lambda: _lib.check(lib.rnla_sketch_apply_dev(0, 0, 5, d, 0, pA, lda, m, n, 0, pS, lds)), reps=2)
    out["c4_dense_d4000"] = {"ms": t * 1e3, "tflops": 2.0 * d * m * n / t * 1e-12, "phases": rt.timings()}
    print(out["c4_dense_d4000"], flush=True)
    del dA, dS; torch.cuda.empty_cache()
if which == "c4solve":
    # BASELINE config 4 end to end: blendenpik on a 1M x 2000 least-squares problem (test_assist.rs:71-93 shape: hypothesis
    # ~ U(-100, 100), Gaussian data with a 1e4 column scaling, small noise), block sparse-sign sketch, d = 4n
    m, n = 1000000, 2000
    dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
    _lib.check(lib.rnla_sketch_fill_dev(0, 0, 77, 9, m, n, 0, pA, lda)); rt.synchronize()
    dA.mul_(torch.logspace(0, -4, n, dtype=torch.float64, device="cuda"))
    xt = (torch.rand(n, 1, dtype=torch.float64, device="cuda") * 200 - 100)
    db = rt.empty_colmajor(m, 1); db.copy_(dA @ xt + 1e-4 * torch.randn(m, 1, dtype=torch.float64, device="cuda"))
    dx = rt.empty_colmajor(n, 1)
    it = C.c_int64(0); cv = C.c_int32(0)
    for kind, zeta, sf, name in [(2, 8, 4.0, "saso_block_sf4"), (2, 8, 2.0, "saso_block_sf2"), (0, 0, 2.0, "dense_gaussian_sf2")]:
        def run():
            _lib.check(lib.rnla_blendenpik_overdetermined_dev(pA, lda, m, n, C.c_void_p(db.data_ptr()), 1e-8, 200, sf, kind, 0, zeta,
                                                              C.c_void_p(dx.data_ptr()), C.byref(it), C.byref(cv)))
        t = timed(run, reps=2)
        res = dA.t() @ (db - dA @ dx)
        out[f"c4_blendenpik_{name}"] = {"ms": t * 1e3, "iterations": int(it.value), "converged": bool(cv.value),
                                        "rel_err_x": float(torch.linalg.vector_norm(dx - xt) / torch.linalg.vector_norm(xt)),
                                        "normal_eq_residual": float(torch.linalg.vector_norm(res) / torch.linalg.vector_norm(dA.t() @ db)),
                                        "phases": rt.timings()}
        print(name, out[f"c4_blendenpik_{name}"], flush=True)
    print(json.dumps(out)); sys.exit(0)
if which in ("c5", "both"):
    n, r0, k, s = 50000, 400, 200, 10
    V0 = rt.empty_colmajor(n, r0); pV, ldv = rt.dev_ptr_ld(V0)
    _lib.check(lib.rnla_sketch_fill_dev(0, 0, 31, 9, n, r0, 0, pV, ldv))
    _lib.check(lib.rnla_orth_dev(pV, ldv, n, r0, 0, None, None)); rt.synchronize()
    lam = np.concatenate([np.logspace(1, -2, 200), np.full(200, 1e-4)])
    Vs = rt.empty_colmajor(n, r0); Vs.copy_(V0 * torch.from_numpy(lam).cuda())
    V0t = rt.empty_colmajor(r0, n); V0t.copy_(V0.t())
    dA = rt.empty_colmajor(n, n); pA, lda = rt.dev_ptr_ld(dA)
    pVs, ldvs = rt.dev_ptr_ld(Vs); pVt, ldvt = rt.dev_ptr_ld(V0t)
    _lib.check(lib.rnla_gemm_nn_dev(pVs, ldvs, n, r0, pVt, ldvt, n, pA, lda)); rt.synchronize()
    # exact symmetry + a small ridge
    dA.copy_(0.5 * (dA + dA.t())); dA.diagonal().add_(1e-8)
    torch.cuda.synchronize()
    res = {}
    def run():
        res["V"], res["L"] = ld.rand_evd2_dev(dA, k, s)
    t = timed(run, reps=2)
    L = res["L"].cpu().numpy()
    out["c5_rand_evd2_50k"] = {"ms": t * 1e3, "A_stream_GBps": 4 * 8.0 * n * n / t * 1e-9, "tflops": 4 * 2.0 * n * n * (k + s) / t * 1e-12,
                               "r": int(len(L)), "max_rel_lambda_err_vs_planted": float(np.max(np.abs(L - (lam[:len(L)] + 1e-8)) / lam[:len(L)])), "phases": rt.timings()}
    print(out["c5_rand_evd2_50k"], flush=True)
print(json.dumps(out))
