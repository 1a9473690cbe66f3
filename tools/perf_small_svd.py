"""Timing of the small dense core (rnla_small_svd_dev = CholeskyQR2 + blocked Jacobi on R^T + U = Q Ur) at the panel
widths of the BASELINE configs; matrices with a graded spectrum.  Usage: python tools/perf_small_svd.py [p ...]"""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
for p in [int(x) for x in (sys.argv[1:] or [60, 110, 210])]:
    rng = np.random.default_rng(p)
    Q1, _ = np.linalg.qr(rng.standard_normal((p, p))); Q2, _ = np.linalg.qr(rng.standard_normal((p, p)))
    M = np.asfortranarray((Q1 * np.logspace(0, -6, p)) @ Q2.T)
    dM = rt.to_device_colmajor(M); dU = rt.empty_colmajor(p, p); dV = rt.empty_colmajor(p, p)
    dS = torch.empty(p, dtype=torch.float64, device="cuda")
    pM, ldm = rt.dev_ptr_ld(dM)
    call = lambda: _lib.check(lib.rnla_small_svd_dev(pM, ldm, p, C.c_void_p(dU.data_ptr()), C.c_void_p(dS.data_ptr()), C.c_void_p(dV.data_ptr())))
    call(); rt.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): call()
    rt.synchronize()
    dt = (time.perf_counter() - t0) / 10
    err = np.abs(dS.cpu().numpy() - np.linalg.svd(M, compute_uv=False)).max()
    print(f"p={p}: {dt*1e3:.3f} ms per small_svd, {lib.rnla_last_jacobi_sweeps()} sweeps, max abs sigma err {err:.2e}", flush=True)
