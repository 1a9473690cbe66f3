"""First-contact GPU sanity checks against numpy (not a test: the pytest suite checks against oracle/)."""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, ".")
import randnla_b200 as rb
from randnla_b200 import runtime as rt, _lib
import ctypes as C
lib = _lib.load()
rt.init(0)
rng = np.random.default_rng(0)
ok = True
def report(name, err, tol):
    global ok
    good = bool(err <= tol)
    ok &= good
    print(f"{'PASS' if good else 'FAIL'} {name}: err={err:.3e} tol={tol:.1e}", flush=True)

# philox KAT
ctr = np.zeros((10, 4), np.uint32); ctr[:, 0] = np.arange(10)
out = rb.sketch.philox4x32_10(ctr, np.array([[0x11111111, 0x22222222]], np.uint32))
print("philox[0] =", [hex(x) for x in out[0]], "expect cc7d356a 5e7dedd7 76798bc3 6c05818c")
report("philox KAT row0", float(out[0, 0] != 0xcc7d356a or out[1, 3] != 0xc5c86681), 0)

# gemm_nn / gemm_tn / sketch_gemm with ragged sizes
def dev(a): return rt.to_device_colmajor(a)
for (m, K, N) in [(300, 70, 9), (2000, 1000, 60), (1025, 515, 110), (777, 333, 210), (4096, 2048, 128)]:
    A = rng.standard_normal((m, K)); B = rng.standard_normal((K, N))
    dA, dB = dev(A), dev(B); dC = rt.empty_colmajor(m, N)
    pA, lda = rt.dev_ptr_ld(dA); pB, ldb = rt.dev_ptr_ld(dB); pC, ldc = rt.dev_ptr_ld(dC)
    _lib.check(lib.rnla_gemm_nn_dev(pA, lda, m, K, pB, ldb, N, pC, ldc)); rt.synchronize()
    ref = A @ B
    report(f"gemm_nn {m}x{K}x{N}", np.abs(dC.cpu().numpy() - ref).max() / np.abs(ref).max(), 1e-13)
    Q = rng.standard_normal((m, N)); dQ = dev(Q); dZ = rt.empty_colmajor(K, N)
    pQ, ldq = rt.dev_ptr_ld(dQ); pZ, ldz = rt.dev_ptr_ld(dZ)
    _lib.check(lib.rnla_gemm_tn_dev(pA, lda, m, K, pQ, ldq, N, pZ, ldz, 0)); rt.synchronize()
    ref = A.T @ Q
    report(f"gemm_tn {m}x{K}x{N}", np.abs(dZ.cpu().numpy() - ref).max() / np.abs(ref).max(), 1e-13)
    Om = rb.sketch.sketch_fill(0, K, N, seed=7, stream=1)
    _lib.check(lib.rnla_sketch_gemm_dev(pA, lda, m, K, 0, 7, 1, N, pC, ldc)); rt.synchronize()
    ref = A @ Om
    report(f"sketch_gemm {m}x{K}x{N}", np.abs(dC.cpu().numpy() - ref).max() / np.abs(ref).max(), 1e-13)
print("omega sample stats: mean %.4f std %.4f" % (Om.mean(), Om.std()))
# unaligned (odd lda) path
A = rng.standard_normal((301, 71)); B = rng.standard_normal((71, 33))
dA, dB = dev(A), dev(B); dC = rt.empty_colmajor(301, 33)
pA, lda = rt.dev_ptr_ld(dA); pB, ldb = rt.dev_ptr_ld(dB); pC, ldc = rt.dev_ptr_ld(dC)
_lib.check(lib.rnla_gemm_nn_dev(pA, lda, 301, 71, pB, ldb, 33, pC, ldc)); rt.synchronize()
report("gemm_nn odd lda", np.abs(dC.cpu().numpy() - A @ B).max(), 1e-12)

# orth
for (rows, cols, rank) in [(500, 40, 40), (2000, 60, 50), (300, 20, 0), (64, 64, 64)]:
    X = rng.standard_normal((rows, cols))
    if rank == 0: X[:] = 0
    elif rank < cols: X = rng.standard_normal((rows, rank)) @ rng.standard_normal((rank, cols))
    Q, R = rb.lora_helpers.Orth(X, return_r=True)
    report(f"orth {rows}x{cols} r{rank} QtQ", np.abs(Q.T @ Q - np.eye(cols)).max(), 1e-13)
    report(f"orth {rows}x{cols} r{rank} QR=X", np.abs(Q @ R - X).max() / max(np.abs(X).max(), 1), 1e-12)
    report(f"orth {rows}x{cols} r{rank} R upper, diag>=0", max(np.abs(np.tril(R, -1)).max(), max(0, -np.diag(R).min())), 0)
    if rank == 0: report("orth(0)=I", np.abs(Q - np.eye(rows, cols)).max(), 0)
    if rank == cols:
        Qn, Rn = np.linalg.qr(X); sg = np.sign(np.diag(Rn)); Qn = Qn * sg
        report(f"orth {rows}x{cols} vs householder Q", np.abs(Q - Qn).max(), 1e-11)

# rand_svd C1
def rank_k(m, n, k, seed):
    r = np.random.default_rng(seed); return r.standard_normal((m, k)) @ r.standard_normal((k, n))
A = rank_k(2000, 1000, 50, 1)
t0 = time.time(); U, S, Vt = rb.lora_drivers.rand_svd(A, 50, 1e-6, 10); t1 = time.time()
sv = np.linalg.svd(A, compute_uv=False)[:50]
report("rand_svd C1 sigma rel", np.abs(np.diag(S) - sv).max() / sv[0], 1e-12)
report("rand_svd C1 sigma rel each", (np.abs(np.diag(S) - sv) / sv).max(), 1e-10)
report("rand_svd C1 recon", np.linalg.norm(U @ S @ Vt - A) / np.linalg.norm(A), 1e-12)
report("rand_svd C1 UtU", np.abs(U.T @ U - np.eye(50)).max(), 1e-12)
print("timings:", rt.timings(), "wall", t1 - t0)
# zero matrix
U, S, Vt = rb.lora_drivers.rand_svd(np.zeros((10, 8)), 3, 0.1, 2)
report("rand_svd(0) U=I", np.abs(U - np.eye(10, 3)).max(), 1e-6)
report("rand_svd(0) S=0", np.abs(S).max(), 1e-6)
report("rand_svd(0) Vt=I", np.abs(Vt - np.eye(3, 8)).max(), 1e-6)
# evd1
G = rng.standard_normal((300, 300)); H = 0.5 * (G + G.T)
V, lam = rb.lora_drivers.rand_evd1(H, 10, 0.1, 290)
w = np.linalg.eigvalsh(H); w = w[np.argsort(-np.abs(w))][:10]
report("rand_evd1 full-l eigenvalues", np.abs(np.array(lam) - w).max(), 1e-10)
report("rand_evd1 VtV", np.abs(V.T @ V - np.eye(10)).max(), 1e-12)
# evd2
Bm = rng.standard_normal((5, 5)); Psd = Bm @ Bm.T
V, lam = rb.lora_drivers.rand_evd2(Psd, 3, 2)
w = np.sort(np.linalg.eigvalsh(Psd))[::-1][:3]
report("rand_evd2 5x5", np.abs(np.array(lam) - w).max(), 1e-6)
# literal
with rt.options(mode=rt.MODE_LITERAL):
    U, S, Vt = rb.lora_drivers.rand_svd(A, 50, 1e-6, 10)
    report("literal rand_svd C1 sigma", (np.abs(np.diag(S) - sv) / sv).max(), 1e-9)
# saso
Am = rng.standard_normal((5000, 37))
Ask = rb.sketch_and_precondition.sketch_apply(Am, None, 200, kind=1, zeta=8, seed=3)
print("saso: ||SA||_F/||A||_F =", np.linalg.norm(Ask) / np.linalg.norm(Am))
Ask = rb.sketch_and_precondition.sketch_apply(Am, None, 200, kind=0, seed=3)
print("dense: ||SA||_F/||A||_F / sqrt(d) =", np.linalg.norm(Ask) / np.linalg.norm(Am) / np.sqrt(200))
print("ALL OK" if ok else "SOME FAILED")
