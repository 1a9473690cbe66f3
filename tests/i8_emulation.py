"""CPU emulation of the fixed-point passes of randnla_b200/csrc/i8gemm.cu: the same digits (widths [7, 8, 8, ...], balanced,
carry from the trailing part), the same digit-pair groups, exact integer accumulation (float64 matmuls of small integers are exact).
Test helper, numpy only: pins the digit arithmetic on CPU and carries the accuracy contract of rnla_options.range_passes_int8
(tests/test_host_logic.py); tools/i8_emulate.py prints the sweep tables quoted in DESIGN.md section 5c."""
import numpy as np

FLUSH_STAGES = {3: 682, 4: 682, 6: 408, 7: 340}      # stages of 64 contraction indices per int32 accumulation (i8gemm.cu FLUSH_P*)


def scales(mx):
    """scales_from_max_bits, vectorised: up = 2^(e+1) with max < 2^e, down = 2^31 / up; zero / unscalable maxima -> 0"""
    mx = np.ascontiguousarray(mx, dtype=np.float64)
    E = ((mx.view(np.uint64) >> np.uint64(52)) & np.uint64(0x7ff)).astype(np.int64)
    ok = (E >= 64) & (E < 2045)
    up = np.where(ok, np.ldexp(1.0, np.where(ok, E - 1021, 0)), 0.0)
    down = np.where(ok, np.ldexp(1.0, np.where(ok, 1052 - E, 0)), 0.0)
    return up, down


def sext8(v):
    return ((v + 128) & 255) - 128


def planes(X, down, P):
    """digits<P> of i8gemm.cu on an array: P integer planes (as float64), most significant first"""
    y = X * down
    yr = np.rint(y)
    vh = yr.astype(np.int64)
    d = [None] * P
    if P > 4:
        vl = np.rint((y - yr) * (65536.0 if P == 6 else 16777216.0)).astype(np.int64)
        for t in range(P - 1, 4, -1):
            d[t] = sext8(vl); vl = (vl - d[t]) >> 8
        c = (vl + 128) >> 8
        d[4] = vl - (c << 8)
        vh = vh + c
    d[3] = sext8(vh); vh = (vh - d[3]) >> 8
    d[2] = sext8(vh); vh = (vh - d[2]) >> 8
    d[1] = sext8(vh); vh = (vh - d[1]) >> 8
    d[0] = vh
    return [p.astype(np.float64) for p in d]


def pairs(P, g0, ng):
    return [(ta, tb) for ta in range(P) for tb in range(P) if g0 <= ta + tb < g0 + ng]


def sweeps(P, all_pairs):
    """(planes used, first group, groups) of the sweeps of one product at precision (P, all_pairs): run_sweeps of i8gemm.cu"""
    if P == 7:
        return [(3, 0, 3), (7, 3, 4)]
    out = [(4, 0, 4)]
    if P == 6: out.append((6, 4, 2))
    elif all_pairs: out.append((4, 4, 3))
    return out


def _product(PA, PB, P, all_pairs, trans):
    C = None
    for pu, g0, ng in sweeps(P, all_pairs):
        acc = {}
        for ta, tb in pairs(pu, g0, ng):
            t = (PA[ta].T @ PB[tb]) if trans else (PA[ta] @ PB[tb])
            acc[ta + tb] = acc.get(ta + tb, 0.0) + t
        v = acc[g0 + ng - 1]
        for g in range(g0 + ng - 2, g0 - 1, -1):
            v = v * 0.00390625 + acc[g]
        v = v * 2.0 ** -(14 + 8 * g0)
        C = v if C is None else C + v
    return C


def i8_nn(A, B, P=4, all_pairs=False):
    """C = A B: A split per row, B per column"""
    ua, da = scales(np.abs(A).max(axis=1)); ub, db = scales(np.abs(B).max(axis=0))
    C = _product(planes(A, da[:, None], P), planes(B, db[None, :], P), P, all_pairs, False)
    return C * (ua[:, None] * ub[None, :])


def i8_tn(A, Q, P=4, all_pairs=False):
    """Z = A^T Q: A split per row (the same split as i8_nn), the row scale folded into Q, Q split per column"""
    ua, da = scales(np.abs(A).max(axis=1))
    Qs = Q * ua[:, None]
    ub, db = scales(np.abs(Qs).max(axis=0))
    Z = _product(planes(A, da[:, None], P), planes(Qs, db[None, :], P), P, all_pairs, True)
    return Z * ub[None, :]


def qr_pos(X):
    Q, R = np.linalg.qr(X)
    s = np.sign(np.diag(R)); s[s == 0] = 1
    return Q * s


def plan(level):
    """I8Plan of csrc/drivers.cu::i8_plan: (early planes, early all-pairs, last planes, last all-pairs, carry planes or 0 = FP64)"""
    return {0: None, 1: (4, False, 4, True, 0), 2: (4, False, 4, True, 7), 3: (7, True, 7, True, 7)}[level]


def rand_svd_emulated(A, Om, k, level, q=2):
    """intended-mode rand_svd (q even) with every product at the precision rnla_options.range_passes_int8 = level gives it"""
    pl = plan(level)
    if pl is None:
        nn = lambda X, Y, last: X @ Y
        tn = lambda X, Y, carry: X.T @ Y
    else:
        e, ea, l, la, c = pl
        nn = lambda X, Y, last: i8_nn(X, Y, l, la) if last else i8_nn(X, Y, e, ea)
        tn = lambda X, Y, carry: (X.T @ Y if c == 0 else i8_tn(X, Y, c, True)) if carry else i8_tn(X, Y, e, ea)
    S = Om
    for _ in range(q // 2):
        Y = qr_pos(nn(A, S, False))
        S = qr_pos(tn(A, Y, False))
    Q = qr_pos(nn(A, S, True))
    Bt = tn(A, Q, True)
    return np.linalg.svd(Bt, compute_uv=False)[:k]


def spectrum_matrix(m, n, k, kappa, gap, noise=1e-12, seed=0, r0=None):
    """sigma_i geometric from 1 to 1/kappa over i < k; then either a gap (tail = gap / kappa) or, gap = None, the same decay
    continued; plus white noise"""
    rng = np.random.default_rng(seed)
    r0 = r0 or 3 * k
    U, _ = np.linalg.qr(rng.standard_normal((m, r0))); V, _ = np.linalg.qr(rng.standard_normal((n, r0)))
    sig = kappa ** (-np.arange(r0) / (k - 1.0))
    if gap is not None:
        sig[k:] = gap / kappa
    return np.asfortranarray((U * sig) @ V.T + noise * rng.standard_normal((m, n)) / np.sqrt(m)), sig
