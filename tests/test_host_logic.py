"""Host-side logic of the multi-GPU path, on CPU (-m "not gpu"):
  * the row-shard layout used by bench.py / runtime.shard_rows
  * the row-sharded dataflow of SURVEY.md §8e (local A*S and A^T Q, all-reduce of the n x l partials and of the
    l x l Gram matrices) reproduces the single-process oracle -- run over world_size-2 gloo, with the ORACLE
    doing the arithmetic of each rank (this is a test of the sharding scheme, not of the product kernels).
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partition():
    from randnla_b200.runtime import shard_rows
    for m in [1, 7, 8, 100, 2000, 200000, 2000001]:
        for w in [1, 2, 4, 8]:
            spans = [shard_rows(m, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(s[1] >= s[0] for s in spans)
            if m >= 4 * w:
                assert all(s[0] % 4 == 0 for s in spans)       # Philox row-quads never straddle ranks


def test_launch_planning_is_wave_aware():
    """host logic of the launches (include/rnla.h rnla_plan_*), no GPU: split-K of gemm_nn and the chunk count of gemm_tn
    fill whole waves of 148 one-CTA SMs at the BASELINE shapes; small problems are left alone"""
    import ctypes as C
    from randnla_b200 import _lib
    lib = _lib.load()
    out = (C.c_int32 * 4)()
    lib.rnla_plan_gemm(200000, 20000, 110, 148, out)          # config 2
    ksplit, chunks, chunk_rows, tiles = list(out)
    assert ksplit == 3 and tiles == 157 and chunks == 16
    waves = tiles * chunks / 148
    assert 0.985 < waves / np.ceil(waves) <= 1.0
    nn_waves = 1563 * ksplit / 148
    assert nn_waves / np.ceil(nn_waves) > 0.985
    lib.rnla_plan_gemm(50000, 50000, 210, 148, out)           # config 5
    assert out[0] == 3 and (782 * out[1] / 148) / np.ceil(782 * out[1] / 148) > 0.98
    lib.rnla_plan_gemm(2000, 1000, 60, 148, out)              # config 1: too short to split
    assert out[0] == 1 and out[1] >= 1
    lib.rnla_plan_gemm(200000, 110, 110, 148, out)            # Gram of a tall panel: K = 110 is never split
    assert out[0] == 1 and out[1] * out[2] >= 200000


@pytest.mark.parametrize("d,zeta,width,n,nchunks", [(8000, 8, 4, 2000, 489), (4000, 8, 4, 2000, 489), (8000, 8, 2, 2000, 489),
                                                    (200, 8, 0, 37, 3), (16384, 1, 1, 5, 40), (64, 8, 4, 600, 1),
                                                    (1200, 8, 0, 3000, 24), (8000, 4, 4, 301, 977), (2000, 8, 4, 1, 100)])
def test_saso_block_work_list_covers_every_unit_once(d, zeta, width, n, nchunks):
    """host logic of the block sparse-sign launch, no GPU: every (column group, chunk) unit is covered by exactly one CTA,
    a column group is either written directly by one CTA or split into slot-carrying fragments in chunk order, slots are
    unique, and the fragments are dispatched longest first"""
    import ctypes as C
    from randnla_b200 import _lib
    lib = _lib.load()
    shape = (C.c_int32 * 4)(); cap = 20000
    desc = (C.c_int32 * (4 * cap))(); nslots = C.c_int32(0)
    nctas = lib.rnla_plan_saso_block(d, zeta, width, n, nchunks, 148, shape, desc, cap, C.byref(nslots))
    assert 0 < nctas <= cap
    bpt, cb, parts, ncg = list(shape)
    assert ncg == -(-n // cb) and bpt * (width or min(zeta, 4)) * cb <= 48
    cover = np.zeros((ncg, nchunks), dtype=np.int32)
    seen_slots = set(); direct = set(); last_len = None; in_frags = False
    for i in range(nctas):
        cg, lo, hi, slot = desc[4 * i:4 * i + 4]
        assert 0 <= cg < ncg and 0 <= lo < hi <= nchunks
        cover[cg, lo:hi] += 1
        if slot < 0:
            assert lo == 0 and hi == nchunks
            direct.add(cg)
        else:
            assert slot not in seen_slots and not (lo == 0 and hi == nchunks)
            seen_slots.add(slot)
            if in_frags:
                assert hi - lo <= last_len
            in_frags = True; last_len = hi - lo
    assert (cover == 1).all()
    assert len(seen_slots) == nslots.value == max(seen_slots, default=-1) + 1
    # work is balanced: no CTA-sized tail beyond the ideal share (whole groups fill complete waves)
    whole = len(direct)
    if nchunks >= 16 and ncg >= 148:
        assert whole >= (ncg // 148) * 148
    # unsupported shapes are refused
    assert lib.rnla_plan_saso_block(20000, 8, 4, 10, 10, 148, shape, desc, cap, C.byref(nslots)) == -1
    assert lib.rnla_plan_saso_block(100, 3, 0, 10, 10, 148, shape, desc, cap, C.byref(nslots)) == -1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _cholqr2_sharded(X_local, dist, sharded):
    """CholeskyQR2 with an all-reduced Gram matrix: the distributed orthonormalisation of SURVEY.md §8e."""
    import torch
    for _ in range(2):
        G = X_local.T @ X_local
        if sharded:
            t = torch.from_numpy(G); dist.all_reduce(t); G = t.numpy()
        R = np.linalg.cholesky(G).T
        X_local = np.linalg.solve(R.T, X_local.T).T
    return X_local


def _worker(rank, world, port, m, n, k, s, q, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    from randnla_b200.runtime import shard_rows
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((m, 12)) @ rng.standard_normal((12, n)) + 1e-3 * rng.standard_normal((m, n)))
    r0, r1 = shard_rows(m, world, rank)
    Al = np.asfortranarray(A[r0:r1])
    l = k + s

    def allreduce(Z):
        t = torch.from_numpy(np.ascontiguousarray(Z)); dist.all_reduce(t); return t.numpy()

    done = 0
    if q % 2 == 0:
        S = orc.omega_fill(0, n, l, seed=0, stream=1)                         # replicated: regenerated from the seed
    else:
        Om = orc.omega_fill(0, r1 - r0, l, seed=0, stream=2, row_off=r0)      # row shard of Omega (m x l)
        S = allreduce(orc.gemm_tn(Al, Om)); done = 1
        S = _cholqr2_sharded(S, dist, False)
    while q - done >= 2:
        Y = _cholqr2_sharded(orc.gemm_nn(Al, S), dist, True); done += 1
        S = _cholqr2_sharded(allreduce(orc.gemm_tn(Al, Y)), dist, False); done += 1
    Q = _cholqr2_sharded(orc.gemm_nn(Al, S), dist, True)
    Bt = allreduce(orc.gemm_tn(Al, Q))
    sig = np.linalg.svd(Bt, compute_uv=False)[:k]
    if rank == 0:
        np.save(out, sig)
    dist.destroy_process_group()


@pytest.mark.parametrize("q", [2, 3])
def test_row_sharded_dataflow_matches_single_process_oracle(tmp_path, orc, q):
    import torch.multiprocessing as mp
    m, n, k, s = 203, 60, 8, 4
    out = str(tmp_path / "sig.npy")
    mp.spawn(_worker, args=(2, _free_port(), m, n, k, s, q, out), nprocs=2, join=True)
    sig = np.load(out)
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((m, 12)) @ rng.standard_normal((12, n)) + 1e-3 * rng.standard_normal((m, n)))
    _, S, _ = orc.rand_svd(A, k, 0.1, s, orc.make_opts(mode=0, num_passes=q))
    assert (np.abs(sig - np.diag(S)) / np.diag(S)).max() < 1e-10


# ---- the fixed-point split behind the INT8 tensor-core passes (randnla_b200/csrc/i8gemm.cu), emulated in tests/i8_emulation.py ----
def test_int8_digit_split_is_exact_bounded_and_accumulates_in_int32():
    """digits<P> of i8gemm.cu: plane 0 holds 7 bits (|d_0| <= 64), the others balanced 8-bit digits in [-128, 127] (int8 without
    slack, hence the carry out of the trailing part); the digits reproduce the rounded fixed-point value exactly; the
    representation error is 2^-(8P) of the row scale `up`; and every digit-pair group of every sweep stays below 2^31 over the
    stages between two drains of the accumulators (FLUSH_P*)."""
    from fractions import Fraction
    import i8_emulation as em
    rng = np.random.default_rng(0)
    for scale in (1.0, 3e-7, 9.1e11):
        row = rng.standard_normal(4000) * scale
        mx = np.abs(row).max()
        row[:4] = [mx * (1 - 2 ** -52), -mx, 0.0, scale * 2 ** -60]
        up, down = em.scales(np.array([mx]))
        up, down = float(up[0]), float(down[0])
        assert mx < up / 2 <= 2 * mx and up * down == 2.0 ** 31
        # entries whose scaled value sits exactly half-way between two integers: the trailing part rounds to +-2^(8(P-4)-1), whose
        # top digit would be +128 without the carry
        row[4:8] = [(12345 + 0.5) / down, -(777 + 0.5) / down, 0.5 / down, (2 ** 29 - 0.5) / down]
        for P in (4, 6, 7):
            pl = em.planes(row[:600], down, P)
            assert np.abs(pl[0]).max() <= 64
            for t in range(1, P):
                assert pl[t].min() >= -128 and pl[t].max() <= 127
            for i in range(600):
                v = sum(int(pl[t][i]) << (8 * (P - 1 - t)) for t in range(P))              # exact integer value of the digits
                exact = Fraction(float(row[i])) * Fraction(2) ** (8 * P - 1) / Fraction(up)
                assert abs(Fraction(v) - exact) <= Fraction(1, 2)
    bound = lambda t: 64 if t == 0 else 128
    for P, all_pairs in ((4, False), (4, True), (6, True), (7, True)):
        for pu, g0, ng in em.sweeps(P, all_pairs):
            for g in range(g0, g0 + ng):
                per_index = sum(bound(ta) * bound(tb) for ta, tb in em.pairs(pu, g0, ng) if ta + tb == g)
                assert per_index * 64 * em.FLUSH_STAGES[pu] < 2 ** 31
    assert [len(em.pairs(pu, g0, ng)) for pu, g0, ng in em.sweeps(7, True)] == [6, 22]
    assert [len(em.pairs(pu, g0, ng)) for pu, g0, ng in em.sweeps(6, True)] == [10, 11]
    assert [len(em.pairs(pu, g0, ng)) for pu, g0, ng in em.sweeps(4, True)] == [10, 6]
    u0, d0 = em.scales(np.array([0.0, 1e-300, np.inf]))
    assert not u0.any() and not d0.any()


def test_int8_products_reach_their_stated_accuracy():
    """emulated products (same digits, same pair groups, exact integer accumulation) against numpy: 4 planes / 10 pairs 2^-27,
    all 16 pairs 2^-30, 6 planes 2^-43, 7 planes 2^-50.5 (numpy's own FP64 product is no better) -- relative to (row max) x
    (column max) x sqrt(K); the scales `up` are 2 to 4 times the maxima, which is where the bits below 31 / 47 / 55 go"""
    import i8_emulation as em
    rng = np.random.default_rng(1)
    A = rng.standard_normal((300, 500)) * np.logspace(-6, 6, 300)[:, None]
    B = rng.standard_normal((500, 24)); Q = rng.standard_normal((300, 24))
    for P, all_pairs, tol in ((4, False, 2.0 ** -25), (4, True, 2.0 ** -28), (6, True, 2.0 ** -41), (7, True, 2.0 ** -49)):
        C = em.i8_nn(A, B, P, all_pairs)
        bound = np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=0)[None, :] * np.sqrt(A.shape[1])
        assert (np.abs(C - A @ B) / bound).max() < tol
        Z = em.i8_tn(A, Q, P, all_pairs)
        bound = (np.abs(A).max(axis=1)[:, None] * np.abs(Q)).max(axis=0)[None, :] * np.sqrt(A.shape[0])
        assert (np.abs(Z - A.T @ Q) / bound).max() < tol


@pytest.mark.parametrize("kappa", [1e2, 1e4, 1e6])
def test_int8_accuracy_contract_on_the_emulated_sweep(kappa):
    """The accuracy contract of rnla_options.range_passes_int8 (include/rnla.h), on the CPU emulation with the same
    Omega as the all-FP64 run: level 3 (every pass on the 55-bit split; what `auto` selects) reproduces the FP64 singular values to
    the north_star tolerance whatever the tail of the spectrum does; levels 1 and 2 (31-bit range passes) only when the part of A
    outside the captured range is small -- with a flat tail at sigma_k they are off by far more than 1e-10, which is why `auto`
    never selects them."""
    import i8_emulation as em
    m, n, k, s = 1500, 600, 20, 8
    Om = np.random.default_rng(1).standard_normal((n, k + s))
    for gap in (1e-2, 1.0, None):
        A, sig = em.spectrum_matrix(m, n, k, kappa, gap)
        s0 = em.rand_svd_emulated(A, Om, k, 0)
        d3 = np.max(np.abs(em.rand_svd_emulated(A, Om, k, 3) - s0) / s0)
        assert d3 < 1e-10, (kappa, gap, d3)
        d2 = np.max(np.abs(em.rand_svd_emulated(A, Om, k, 2) - s0) / s0)
        if gap == 1e-2 and kappa <= 1e4:
            assert d2 < 1e-10, (kappa, gap, d2)
        if gap == 1.0:
            assert d2 > 1e-10, (kappa, gap, d2)


def test_bench_reference_arm_prints_its_json_line():
    """`bench.py --impl reference` (the oracle's literal-mode rand_svd on host cores, no GPU) keeps the contract's keys"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1")               # what torch.distributed.run exports: the arm must not inherit it
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-rows", "1200", "--cols", "400"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "rand_svd_A_stream_GBps" and line["unit"] == "GB/s"
    assert line["value"] > 0 and line["dtype"] == "f64" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["rows_timed"] == 1200 and line["cpu_baseline"]["full_workload"] is False
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


# ---- the row-sharded dataflow of the iterative solvers (csrc/solve.cu: dev_lsqr, cgls_operator): u / r sharded with A, the n-vectors
# ---- replicated, A^T u all-reduced, every norm of an m-vector an all-reduced scalar -- over world_size-2 gloo
def _solver_worker(rank, world, port, m, n, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from randnla_b200.runtime import shard_rows
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    A = rng.standard_normal((m, n)) * np.logspace(0, -1, n); b = rng.standard_normal(m)
    r0, r1 = shard_rows(m, world, rank)
    Al, bl = A[r0:r1], b[r0:r1]

    def allsum(x):
        t = torch.from_numpy(np.atleast_1d(np.asarray(x, dtype=np.float64)).copy()); dist.all_reduce(t); return t.numpy()

    # LSQR, damp = 0 (src/solvers.rs:140-232), six iterations, local arithmetic on the shard + the collectives of dev_lsqr
    x = np.zeros(n); u = bl.copy()
    beta = float(np.sqrt(allsum(u @ u)[0])); u /= beta
    v = allsum(Al.T @ u); alfa = float(np.linalg.norm(v)); v /= alfa
    w = v.copy(); rhobar, phibar = alfa, beta
    for _ in range(6):
        u = Al @ v - alfa * u
        beta = float(np.sqrt(allsum(u @ u)[0])); u /= beta
        v = allsum(Al.T @ u) - beta * v
        alfa = float(np.linalg.norm(v)); v /= alfa
        rho = np.hypot(rhobar, beta); cs, sn = rhobar / rho, beta / rho
        theta = sn * alfa; rhobar = -cs * alfa; phi = cs * phibar; phibar = sn * phibar
        x = x + (phi / rho) * w
        w = v - (theta / rho) * w
    # CGLS (src/cg.rs:29-52), five iterations
    xc = np.zeros(n); r = bl - Al @ xc
    s = allsum(Al.T @ r); p = s.copy(); ns = float(s @ s)
    for _ in range(5):
        ap = Al @ p
        alpha = ns / float(allsum(ap @ ap)[0])
        xc = xc + alpha * p; r = r - alpha * ap
        s = allsum(Al.T @ r); nn = float(s @ s)
        p = s + (nn / ns) * p; ns = nn
    # the same two solvers on the ONE-PASS dataflow (csrc/normal_pass.cu): per iteration one local pass gives u = cq A_l x + cy y_l,
    # A_l^T u and u . u, and ONE all-reduce of n + 1 doubles combines the shards
    def normal_pass(xv, cq, y, cy):
        ul = cq * (Al @ xv) + (cy * y if y is not None else 0.0)
        t = allsum(np.concatenate([Al.T @ ul, [ul @ ul]]))
        return ul, t[:n], float(t[n])
    # LSQR: u stays unnormalised in memory, its factor goes into the next coefficient (dev_lsqr, `uscale`)
    x1 = np.zeros(n); u = bl.copy()
    beta = float(np.sqrt(allsum(u @ u)[0])); u /= beta
    v = allsum(Al.T @ u); alfa = float(np.linalg.norm(v)); v /= alfa
    w = v.copy(); rhobar, phibar = alfa, beta; uscale = 1.0
    for _ in range(6):
        u, t, uu = normal_pass(v, 1.0, u, -alfa * uscale)
        beta = float(np.sqrt(uu)); uscale = 1.0 / beta
        v = t / beta - beta * v
        alfa = float(np.linalg.norm(v)); v /= alfa
        rho = np.hypot(rhobar, beta); cs, sn = rhobar / rho, beta / rho
        theta = sn * alfa; rhobar = -cs * alfa; phi = cs * phibar; phibar = sn * phibar
        x1 = x1 + (phi / rho) * w
        w = v - (theta / rho) * w
    # CGLS: s_new = s - alpha A^T (A p) (cgls_operator_onepass); r is never formed
    xc1 = np.zeros(n)
    _, s, _ = normal_pass(xc1, -1.0, bl, 1.0); p = s.copy(); ns = float(s @ s)
    for _ in range(5):
        _, t, qq = normal_pass(p, 1.0, None, 0.0)
        alpha = ns / qq
        xc1 = xc1 + alpha * p; s = s - alpha * t
        nn = float(s @ s)
        p = s + (nn / ns) * p; ns = nn
    if rank == 0:
        np.savez(out, x=x, xc=xc, x1=x1, xc1=xc1)
    dist.destroy_process_group()


def test_row_sharded_solver_dataflow_matches_single_process_oracle(tmp_path, orc):
    import torch.multiprocessing as mp
    m, n = 157, 23
    out = str(tmp_path / "solvers.npz")
    mp.spawn(_solver_worker, args=(2, _free_port(), m, n, out), nprocs=2, join=True)
    got = np.load(out)
    rng = np.random.default_rng(5)
    A = np.asfortranarray(rng.standard_normal((m, n)) * np.logspace(0, -1, n)); b = rng.standard_normal(m)
    xl = orc.lsqr(A, b, 0.0, 0.0, 0.0, 0.0, 6, False, None)[0][:, 0]
    xc, it, conv = orc.cgls(A, b, 1e-300, 5)
    assert np.abs(got["x"] - xl).max() <= 1e-12 * np.abs(xl).max()
    assert np.abs(got["xc"] - xc[:, 0]).max() <= 1e-12 * np.abs(xc).max()
    # the one-pass dataflow is the same iteration up to rounding
    assert np.abs(got["x1"] - xl).max() <= 1e-12 * np.abs(xl).max()
    assert np.abs(got["xc1"] - xc[:, 0]).max() <= 1e-12 * np.abs(xc).max()


# ---- the CGLS recurrence of the one-pass iteration (csrc/solve.cu cgls_operator_onepass) against the reference's (src/cg.rs:29-52) ----
def _cgls_r_recurrence(A, b, tol, maxit):
    """src/cg.rs:29-52 statement for statement: r updated, s = a^T r recomputed every iteration"""
    x = np.zeros(A.shape[1]); r = b - A @ x; s = A.T @ r; p = s.copy(); ns = s @ s
    hist = []
    for i in range(maxit):
        ap = A @ p
        alpha = ns / (ap @ ap)
        x = x + alpha * p; r = r - alpha * ap
        s = A.T @ r; nn = s @ s
        hist.append(np.sqrt(nn))
        if np.sqrt(nn) < tol:
            return x, i + 1, hist
        p = s + (nn / ns) * p; ns = nn
    return x, maxit, hist


def _cgls_s_recurrence(A, b, tol, maxit):
    """what the device runs when a p and a^T (a p) come from one pass over a: s_new = s - alpha a^T (a p); r is never formed"""
    x = np.zeros(A.shape[1]); s = A.T @ (b - A @ x); p = s.copy(); ns = s @ s
    hist = []
    for i in range(maxit):
        q = A @ p; t = A.T @ q
        alpha = ns / (q @ q)
        x = x + alpha * p; s = s - alpha * t
        nn = s @ s
        hist.append(np.sqrt(nn))
        if np.sqrt(nn) < tol:
            return x, i + 1, hist
        p = s + (nn / ns) * p; ns = nn
    return x, maxit, hist


@pytest.mark.parametrize("cond", [1.5, 3.0, 10.0])
def test_one_pass_cgls_recurrence_is_the_reference_iteration_on_preconditioned_operators(cond):
    """On a preconditioned operator (cond(A M) = O(1): what blendenpik / LSRN / the saddle-point driver hand to cgls) the s recurrence
    of the one-pass iteration and the reference's r recurrence are the same iteration to rounding: same stopping iteration, same solution to
    1e-11, residual histories equal to rounding over the first ten iterations.  (Why plain cgls(a, ...) keeps the reference's recurrence: at cond 1e4 the two histories
    part after a few dozen iterations -- asserted below as a fact about the recurrences, not about the device.)"""
    rng = np.random.default_rng(int(cond * 10))
    m, n = 400, 60
    U, _ = np.linalg.qr(rng.standard_normal((m, n))); V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (U * np.linspace(1.0, 1.0 / cond, n)) @ V.T
    b = A @ rng.uniform(-100, 100, n) + 1e-2 * rng.standard_normal(m)
    xr, itr, hr = _cgls_r_recurrence(A, b, 1e-9, 200)
    xs, its, hs = _cgls_s_recurrence(A, b, 1e-9, 200)
    assert itr == its and itr < 120
    assert np.abs(xr - xs).max() <= 1e-11 * np.abs(xr).max()
    # residual histories: equal to rounding while the Krylov basis is still orthogonal; later any two CG implementations drift apart
    # at the rate rounding errors are amplified (not a property of the recurrence: two summation orders of the SAME recurrence do so too)
    assert np.abs(np.array(hr[:10]) - np.array(hs[:10])).max() <= 1e-12 * hr[0]
    assert np.abs(np.array(hr) - np.array(hs)).max() <= 1e-5 * hr[0]
    assert np.linalg.norm(A.T @ (b - A @ xs)) < 1e-8


def test_recurrences_differ_only_at_the_rounding_floor():
    """Unpreconditioned, cond(a) = 1e3, run far past convergence: both recurrences reach the SAME true residual a^T (b - a x) (the
    rounding floor), but what they carry differs there -- the s recurrence's carried residual keeps shrinking (it is never
    re-anchored on r) while the reference's stays within a few hundred times the true one.  So a tolerance below the floor stops the
    two at different iterations; above the floor they are the same iteration.  This is why the one-pass iteration is the default only
    where the caller's tolerance sits above the floor by construction (the preconditioned drivers) and plain cgls(a, ...) keeps the
    reference's recurrence (RNLA_ONEPASS=2 opts it in; RNLA_ONEPASS_CONFIRM=1 re-anchors a stop on the true residual)."""
    rng = np.random.default_rng(3)
    m, n = 400, 60
    U, _ = np.linalg.qr(rng.standard_normal((m, n))); V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (U * np.logspace(0, -3, n)) @ V.T
    b = A @ rng.uniform(-100, 100, n) + 1e-2 * rng.standard_normal(m)
    s0 = np.linalg.norm(A.T @ b)
    # above the floor: same carried residuals, equal to the true ones
    xr, _, hr = _cgls_r_recurrence(A, b, 1e-30, 400)
    xs, _, hs = _cgls_s_recurrence(A, b, 1e-30, 400)
    for x, h in ((xr, hr), (xs, hs)):
        assert abs(np.linalg.norm(A.T @ (b - A @ x)) - h[-1]) <= 1e-6 * h[-1] + 1e-13 * s0
    # at the floor
    xr, _, hr = _cgls_r_recurrence(A, b, 1e-30, 2000)
    xs, _, hs = _cgls_s_recurrence(A, b, 1e-30, 2000)
    true_r = np.linalg.norm(A.T @ (b - A @ xr)); true_s = np.linalg.norm(A.T @ (b - A @ xs))
    assert true_r < 1e-11 * s0 and true_s < 1e-11 * s0 and 0.01 < true_s / true_r < 100.0      # the same floor
    assert hr[-1] > 1e-4 * true_r                                                                  # the reference's carried residual stays near it
    assert hs[-1] < 1e-8 * true_s                                                                  # the s recurrence's does not


# ---- host logic of the one-pass kernel's launch (csrc/normal_pass.cu, rnla_plan_normal_pass): which tensor map serves which column ----
@pytest.mark.parametrize("addr", [0x7f0000000000, 0x7f0000000008])
@pytest.mark.parametrize("lda", [1000, 1001, 4096, 4097])
@pytest.mark.parametrize("n", [1, 2, 7, 255, 256, 257, 500, 999, 1000, 2000, 2047, 2048])
def test_one_pass_launch_geometry(addr, lda, n):
    """every column of every CTA is served by exactly one stage position; the map that serves it knows whether the column starts on a
    16-byte boundary (TMA's requirement for every row of a box) or 8 bytes off it (then the box starts one element early and the
    stage keeps 34 rows per column); the stage fits; an odd leading dimension makes every CTA start on an even column"""
    import ctypes as C
    from randnla_b200 import _lib
    lib = _lib.load()
    out = (C.c_int32 * 7)()
    assert lib.rnla_plan_normal_pass(addr, lda, n, out) == 1
    cl, ncb, ne, she, sho, pitch, stage_bytes = list(out)
    assert cl in (1, 2, 4, 8) and cl * ncb >= n and ncb <= 256 and (cl == 1 or (cl // 2) * 256 < n)
    odd = lda % 2 == 1
    if odd and cl > 1:
        assert ncb % 2 == 0
    any_off = False
    for rank in range(cl):
        col0 = rank * ncb
        seen = set()
        for pos in range(ncb):
            lc = pos if ne == ncb else (2 * pos if pos < ne else 2 * (pos - ne) + 1)
            assert 0 <= lc < ncb and lc not in seen
            seen.add(lc)
            j = col0 + lc
            off = ((addr >> 3) + j * lda) & 1                  # 1: column j starts 8 bytes off a 16-byte boundary
            assert off == (she if pos < ne else sho), (rank, pos, lc)
            any_off |= bool(off) and j < n
        assert seen == set(range(ncb))
    assert pitch == (34 if (odd or (addr >> 3) & 1) else 32)
    if any_off:
        assert pitch == 34
    second_box = (ne * pitch + 15) // 16 * 16                     # doubles; the second box starts on a 128-byte boundary
    assert (second_box + (256 - ne) * pitch) * 8 <= stage_bytes
    assert 3 * stage_bytes + 8 * 8 * 32 * 8 + 2 * 8 * 32 * 8 + 256 + 128 <= 227 * 1024


def test_one_pass_launch_geometry_declines_wide_operands():
    import ctypes as C
    from randnla_b200 import _lib
    out = (C.c_int32 * 7)()
    assert _lib.load().rnla_plan_normal_pass(0x7f0000000000, 1000, 2049, out) == 0
    assert _lib.load().rnla_plan_normal_pass(0x7f0000000004, 1000, 100, out) == 0
