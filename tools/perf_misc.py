"""Side measurements on one B200: rand_evd1 on a 20000 x 20000 symmetric matrix, rand_svd at intermediate sketch widths."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
lib = _lib.load(); rt.init(0)
out = {}
def timed(fn, reps=3):
    fn(); rt.synchronize()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); rt.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
n, r0 = 20000, 200
V0 = rt.empty_colmajor(n, r0); pV, ldv = rt.dev_ptr_ld(V0)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 31, 9, n, r0, 0, pV, ldv)); _lib.check(lib.rnla_orth_dev(pV, ldv, n, r0, 0, None, None)); rt.synchronize()
lam = np.concatenate([np.logspace(1, -2, 100) * np.where(np.arange(100) % 3 == 0, -1.0, 1.0), np.full(100, 1e-5)])
A = rt.empty_colmajor(n, n); A.copy_((V0 * torch.from_numpy(lam).cuda()) @ V0.t()); A.copy_(0.5 * (A + A.t())); torch.cuda.synchronize()
res = {}
def run():
    res["V"], res["L"] = ld.rand_evd1_dev(A, 100, 10)
t = timed(run)
L = res["L"].cpu().numpy(); ref = lam[:100][np.argsort(-np.abs(lam[:100]))]
out["rand_evd1_20k_k100"] = {"ms": t * 1e3, "max_rel_lambda_err": float(np.max(np.abs(L - ref) / np.abs(ref))), "phases": rt.timings()}
print(out["rand_evd1_20k_k100"], flush=True)
del A, V0; torch.cuda.empty_cache()
m, n = 200000, 20000
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 5, 9, m, n, 0, pA, lda)); rt.synchronize()
for k, s in [(2, 6), (14, 10), (22, 10), (30, 10)]:
    t = timed(lambda: ld.rand_svd_dev(dA, k, s), reps=2)
    ph = rt.timings(); pm = [v for kk, v in ph if kk.startswith("pass:")]
    l = k + s
    out[f"rand_svd_l{l}"] = {"ms": t * 1e3, "mean_pass_ms": float(np.mean(pm)), "tflops_per_pass": 2.0 * m * n * l / (np.mean(pm) * 1e-3) * 1e-12,
                             "A_stream_GBps_per_pass": 8.0 * m * n / (np.mean(pm) * 1e-3) * 1e-9, "other_ms": float(sum(v for kk, v in ph if not kk.startswith("pass:"))), "passes": [round(v, 2) for v in pm]}
    print(l, out[f"rand_svd_l{l}"], flush=True)
print(json.dumps(out))
