"""rand_svd on the bench workload (low-rank + noise, planted spectrum 1 .. 1e-3 then 1e-5) with the range-finder passes on the
INT8 tensor cores (rnla_options.range_passes_int8 = 1) against the all-FP64 path: singular values, subspace, time."""
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
for name, flag in (("fp64", 0), ("int8_range", 1), ("int8_all", 2)):
    opts = rt.make_options(range_passes_int8=flag)
    U, S, Vt = ld.rand_svd_dev(dA, 100, 10, opts); rt.synchronize()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        U, S, Vt = ld.rand_svd_dev(dA, 100, 10, opts); rt.synchronize()
        best = min(best, time.perf_counter() - t0)
    res[name] = (U.clone(), S.clone(), Vt.clone())
    out[name] = {"ms": best * 1e3, "phases": rt.timings()}
    print(name, f"{best*1e3:.2f} ms", rt.timings(), flush=True)
S0 = res["fp64"][1].cpu().numpy(); S1 = res["int8_range"][1].cpu().numpy(); S2 = res["int8_all"][1].cpu().numpy()
out["max_rel_sigma_diff_all"] = float(np.max(np.abs(S0 - S2) / S0))
out["rel_sigma_diff_all_last5"] = [float(x) for x in (np.abs(S0 - S2) / S0)[-5:]]
U2, V2 = res["int8_all"][0], res["int8_all"][2]
out["orth_err_all"] = float((U2.t() @ U2 - torch.eye(100, dtype=torch.float64, device="cuda")).abs().max())
out["VtV_err_all"] = float((V2 @ V2.t() - torch.eye(100, dtype=torch.float64, device="cuda")).abs().max())
out["sigma_vs_planted_all"] = float(np.max(np.abs(S2 - sig[:100]) / sig[:100]))
out["max_rel_sigma_diff"] = float(np.max(np.abs(S0 - S1) / S0))
out["rel_sigma_diff_last10"] = [float(x) for x in (np.abs(S0 - S1) / S0)[-10:]]
U0, U1 = res["fp64"][0], res["int8_range"][0]
G = U0.t() @ U1
out["subspace_sin_max"] = float(torch.linalg.svdvals(U1 - U0 @ G).max())
out["orth_err_int8"] = float((U1.t() @ U1 - torch.eye(100, dtype=torch.float64, device="cuda")).abs().max())
out["sigma_vs_planted_fp64"] = float(np.max(np.abs(S0 - sig[:100]) / sig[:100]))
out["sigma_vs_planted_int8"] = float(np.max(np.abs(S1 - sig[:100]) / sig[:100]))
print(json.dumps(out))
