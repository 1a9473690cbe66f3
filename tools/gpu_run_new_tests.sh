mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_next_rows.py -m gpu -q -k "lupp or conjugate" ) > gpurun_out/t_lupp.log 2>&1
( time timeout 600 python -m pytest tests/test_cpp_mirror.py -m gpu -q ) > gpurun_out/t_cpp.log 2>&1
python - > gpurun_out/perf_lupp.log 2>&1 <<'PY'
import sys, time, numpy as np, torch, ctypes as C
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
for n in (500, 2000, 4000):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    W = rt.empty_colmajor(n, n); L = rt.empty_colmajor(n, n); U = rt.empty_colmajor(n, n)
    p = torch.zeros(n, dtype=torch.int64, device="cuda")
    pw, ldw = rt.dev_ptr_ld(W); pl, ldl = rt.dev_ptr_ld(L); pu, ldu = rt.dev_ptr_ld(U)
    best = 1e30
    for _ in range(3):
        W.copy_(A); torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.check(lib.rnla_lupp_dev(pw, ldw, n, pl, ldl, pu, ldu, C.c_void_p(p.data_ptr()))); rt.synchronize()
        best = min(best, time.perf_counter() - t0)
    err = float((L @ U - A[p]).abs().max())
    print(f"lupp n={n}: {best*1e3:.2f} ms, {best*1e6/(n-1):.2f} us per step, max |LU - PA| = {err:.2e}", flush=True)
PY
tail -n 6 gpurun_out/t_lupp.log gpurun_out/t_cpp.log gpurun_out/perf_lupp.log
