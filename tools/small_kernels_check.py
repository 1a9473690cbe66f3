"""Blocked Cholesky / triangular inverse (csrc/panel.cu) against the unblocked kernels (RNLA_SMALL_OLD=1): Orth of tall panels,
well conditioned, ill conditioned, rank deficient and with zero columns -- Q and R must agree to rounding."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
def orth(X):
    m, p = X.shape
    dX = rt.to_device_colmajor(X); pX, ld = rt.dev_ptr_ld(dX)
    dR = rt.empty_colmajor(p, p); pR, ldr = rt.dev_ptr_ld(dR)
    df = C.c_int64(0)
    torch.cuda.synchronize()
    _lib.check(lib.rnla_orth_dev(pX, ld, m, p, 0, pR, C.byref(df)))
    rt.synchronize()
    return dX.cpu().numpy(), dR.cpu().numpy(), int(df.value)
rng = np.random.default_rng(0)
worst = 0.0
for (m, p, kind) in [(3000, 5, "rand"), (4000, 16, "rand"), (5000, 17, "ill"), (8000, 110, "rand"), (8000, 128, "ill"), (6000, 97, "rankdef"), (7000, 144, "rand"),
                     (9000, 145, "rand"), (5000, 210, "ill"), (6000, 200, "rand"), (5000, 177, "rankdef"), (4000, 64, "zero"), (5000, 33, "rankdef")]:
    X = rng.standard_normal((m, p))
    if kind == "ill": X = X * np.logspace(0, -6, p)
    if kind == "rankdef": X[:, p // 2] = X[:, 0] + X[:, 1]; X[:, p - 1] = 2 * X[:, 3]
    if kind == "zero": X[:, 7] = 0.0; X[:, p - 2] = 0.0
    X = np.asfortranarray(X)
    os.environ["RNLA_SMALL_OLD"] = "1"; Q0, R0, d0 = orth(X)
    os.environ["RNLA_SMALL_OLD"] = "0"; Q1, R1, d1 = orth(X)
    eq = np.abs(Q0 - Q1).max(); er = np.abs(R0 - R1).max() / np.abs(R0).max()
    oo = np.abs(Q1.T @ Q1 - np.eye(p)).max() if kind not in ("rankdef",) else 0.0
    rec = np.abs(Q1 @ R1 - X).max() / np.abs(X).max()
    print(f"{m} x {p} {kind:8s}: |Q_old - Q_new| {eq:.2e}  |R_old - R_new|/|R| {er:.2e}  deficient {d0}/{d1}  |Q^T Q - I| {oo:.2e}  |Q R - X| {rec:.2e}", flush=True)
    assert d0 == d1
    if kind in ("rand",): assert eq < 1e-12 and er < 1e-12
    assert rec < 1e-10
    worst = max(worst, rec)
print("ok", worst)
