"""rand_svd on the bench workload (low-rank + noise, planted spectrum 1 .. 1e-3 then 1e-5) at every level of
rnla_options.range_passes_int8 against the all-FP64 path: singular values, orthogonality, time and phases."""
import sys, json, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
import bench
lib = _lib.load(); rt.init(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sig = bench.planted_sigma()
dA = rt.empty_colmajor(m, n); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_generate_lowrank_dev(pA, lda, m, n, 0, m, bench.R0, sig.ctypes.data_as(C.c_void_p), 1e-7, 1234))
out = {"m": m, "n": n}
res = {}
for name, flag in (("fp64", 0), ("level1", 1), ("level2", 2), ("level3", 3), ("auto", -1)):
    opts = rt.make_options(range_passes_int8=flag)
    U, S, Vt = ld.rand_svd_dev(dA, 100, 10, opts); rt.synchronize()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        U, S, Vt = ld.rand_svd_dev(dA, 100, 10, opts); rt.synchronize()
        best = min(best, time.perf_counter() - t0)
    res[name] = (U.clone(), S.clone().cpu().numpy(), Vt.clone())
    eye = torch.eye(100, dtype=torch.float64, device="cuda")
    out[name] = {"ms": best * 1e3, "phases": rt.timings(), "max_rel_sigma_diff_vs_fp64": float(np.max(np.abs(res[name][1] - res["fp64"][1]) / res["fp64"][1])),
                 "orth_err": float((U.t() @ U - eye).abs().max()), "sigma_vs_planted": float(np.max(np.abs(res[name][1] - sig[:100]) / sig[:100]))}
    print(name, f"{best*1e3:.2f} ms  dsigma {out[name]['max_rel_sigma_diff_vs_fp64']:.2e}", [(a, round(b, 2)) for a, b in rt.timings()], flush=True)
print(json.dumps(out))
