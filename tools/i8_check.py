"""INT8 tensor-core range-finder products (csrc/i8gemm.cu) against torch FP64 matmul, and their rate at the config-2 size."""
import sys, json, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
big = "big" in sys.argv
out = {}
def run(trans, m, n, N, reps=1, scale_rows=False):
    g = torch.Generator(device="cuda").manual_seed(m + n + N)
    A = rt.empty_colmajor(m, n); A.copy_(torch.randn((m, n), generator=g, device="cuda", dtype=torch.float64))
    if scale_rows:
        A.mul_(torch.logspace(-6, 6, m, dtype=torch.float64, device="cuda").reshape(-1, 1))
    kb = m if trans else n
    B = rt.empty_colmajor(kb, N); B.copy_(torch.randn((kb, N), generator=g, device="cuda", dtype=torch.float64))
    Cm = rt.empty_colmajor(n if trans else m, N); Cm.fill_(float("nan"))
    pa, lda = rt.dev_ptr_ld(A); pb, ldb = rt.dev_ptr_ld(B); pc, ldc = rt.dev_ptr_ld(Cm)
    torch.cuda.synchronize()
    _lib.check(lib.rnla_i8_range_gemm_dev(1 if trans else 0, pa, lda, m, n, pb, ldb, N, pc, ldc, reps))
    torch.cuda.synchronize()
    ref = (A.t() @ B) if trans else (A @ B)
    den = (A.abs().t() @ B.abs()) if trans else (A.abs() @ B.abs())
    err = float(((Cm - ref).abs() / den).max())
    fro = float((Cm - ref).norm() / ref.norm())
    return err, fro, rt.timings()
cases = [(1, 300, 200, 110, False, -1), (1, 70000, 1000, 110, False, -1), (1, 1000, 500, 110, True, -1), (0, 300, 200, 110, False, -1), (0, 5000, 3000, 60, False, -1), (0, 1000, 500, 110, True, -1), (0, 128, 64, 128), (0, 256, 128, 16), (0, 300, 200, 110), (1, 128, 128, 128), (1, 300, 200, 110), (0, 5000, 3000, 60), (1, 70000, 1000, 110), (0, 1000, 500, 110, True), (1, 1000, 500, 110, True)]
if big:
    cases = []
for cs in cases:
    trans, m, n, N = cs[:4]
    err, fro, ph = run(trans, m, n, N, cs[5] if len(cs) > 5 else 1, len(cs) > 4 and cs[4])
    print(f"trans={trans} {m}x{n} N={N} scaled={len(cs) > 4 and cs[4]} reps={cs[5] if len(cs) > 5 else 1}: componentwise err {err:.3e}  fro {fro:.3e}", flush=True)
    out[f"{trans}_{m}x{n}_{N}_{len(cs)}"] = [err, fro]
if big:
    m, n, N = 200000, 20000, 110
    for trans, rp in ((0, 3), (1, 3), (0, -3), (1, -3)):
        err, fro, ph = run(trans, m, n, N, rp)
        print(f"trans={trans} {m}x{n} N={N}: err {err:.3e} fro {fro:.3e} phases {ph}", flush=True)
        out[f"big_{trans}_{rp}"] = {"err": err, "fro": fro, "phases": ph}
print(json.dumps(out))
