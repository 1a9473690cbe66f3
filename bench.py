#!/usr/bin/env python
"""bench.py -- headline benchmark of the sketch-and-factor path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one `rand_svd` (k=100, s=10, q=2, Gaussian sketch) of a synthetic low-rank-plus-noise f64 matrix with
200 000 rows per GPU and 20 000 columns (N=1: BASELINE config 2, 200k x 20k = 32 GB; N>1: the rows are sharded,
weak scaling).  The metric is the algorithmic A-stream rate: 4 passes x 8*m*n bytes per step / time
(SURVEY.md §8d), whole job.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_RANK, S_OVER, Q_PASSES = 100, 10, 2
N_COLS = 20000
ROWS_PER_GPU = 200000
R0 = 200
METRIC = "rand_svd_A_stream_GBps"


def planted_sigma():
    # geometric 1 -> 1e-3 over the first 100, then 1e-5 (gap >= 100 at k = 100)
    return np.concatenate([np.logspace(0, -3, 100), np.full(R0 - 100, 1e-5)])


def algorithmic_bytes(m, n):
    return 4.0 * 8.0 * m * n          # q + 2 = 4 passes, A read exactly once per pass


def algorithmic_flops(m, n, l):
    return 4.0 * 2.0 * m * n * l


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe), read through NVML in a thread of this
    process (nvidia_ml_py).  An `nvidia-smi -lms 100` child did the same in round 1, but its queries were seen to stall the timed steps
    now and then (one 20 - 150 ms gap in some runs, with normal phase sums): the NVML calls below take microseconds.  The sampler is
    started before the warm-up; only samples taken between begin() and end() are reported.  Falls back to nvidia-smi if NVML is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_s=0.05):
        self.rows = []                      # (timestamp, sm_mhz, sm_max_mhz, set of reasons)
        self.proc = None
        self.idx = gpu_index
        self.period = period_s
        self.t0 = self.t1 = None
        self.stop_flag = threading.Event()
        self.th = None
        self.source = None

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {pynvml.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        self.rows.append((time.perf_counter(), sm, smax, {nm for bit, nm in names.items() if mask & bit}))
                    except Exception:
                        pass
                    self.stop_flag.wait(self.period)
            self.th = threading.Thread(target=loop, daemon=True)
            self.th.start()
            self.source = "NVML (nvidia_ml_py), in-process, every %d ms" % int(self.period * 1e3)
            return
        except Exception:
            self.th = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            self.source = "nvidia-smi -lms 100"
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                reasons = {name for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9])
                           if v.lower().startswith("active")}
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]), reasons))
            except Exception:
                pass

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif self.th:
            self.th.join(timeout=1)
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        sm, smax, reasons = [], [], set()
        for ts, a, b, rs in self.rows:
            if (self.t0 is not None and ts < self.t0) or (self.t1 is not None and ts > self.t1):
                continue
            sm.append(a); smax.append(b); reasons |= rs
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def host_threads():
    """all host cores this process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)"""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _blas_threads(n):
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def workload_matrix(rows, n, seed=1234):
    """the bench workload on the host: planted low-rank part of the device generator's matrix (same spectrum)"""
    rng = np.random.default_rng(seed)
    sig = planted_sigma()
    with _blas_threads(host_threads()):
        U0, _ = np.linalg.qr(rng.standard_normal((rows, R0)))
        V0, _ = np.linalg.qr(rng.standard_normal((n, R0)))
        A = np.empty((rows, n), order="F")
        np.matmul(U0 * sig, V0.T, out=A)
    return A


def cpu_rand_svd_seconds(A, k, s, threads, steps=1, warmup=0):
    """The reference's CPU dataflow (oracle `literal` mode: src/lora_helpers.rs:17-146 + src/lora_drivers.rs:30-69 statement for
    statement, every product a real GEMM, the transpose materialised) on `threads` host threads; mean seconds per call."""
    from oracle import oracle as orc
    orc.load()
    orc.set_threads(int(threads))
    o = orc.make_opts(mode=orc.MODE_LITERAL)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        orc.rand_svd(A, k, 1e-6, s, o)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.mean(times)), times


def cpu_baseline(n, sample_rows):
    """cpu_baseline leg of the GPU arm (rank 0, N = 1): the oracle port on a bounded row sample of the workload with ALL host cores
    (`value`), the same on ONE thread (faithful to the reference: nalgebra / matrixmultiply are single-threaded, Cargo.lock:618-619,
    644-645) on a smaller sample, and BASELINE config 1 (2000 x 1000 rank-50, k = 50, s = 10) in full, both ways."""
    cores = host_threads()
    A = workload_matrix(sample_rows, n)
    t_all, _ = cpu_rand_svd_seconds(A, K_RANK, S_OVER, cores)
    rows1 = max(1024, sample_rows // 8)
    t_one, _ = cpu_rand_svd_seconds(A[:rows1].copy(order="F"), K_RANK, S_OVER, 1)
    del A
    rng = np.random.default_rng(0)
    C1 = np.zeros((2000, 1000), order="F")
    for _ in range(50):                                        # src/test_assist.rs:7-31 rank_k_matrix
        C1 += np.outer(rng.standard_normal(2000), rng.standard_normal(1000))
    c1_all = float(np.median(cpu_rand_svd_seconds(C1, 50, 10, cores, steps=5, warmup=1)[1]))
    c1_one = float(np.median(cpu_rand_svd_seconds(C1, 50, 10, 1, steps=5, warmup=1)[1]))
    return {"value": algorithmic_bytes(sample_rows, n) / t_all * 1e-9, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"oracle literal-mode rand_svd (k={K_RANK}, s={S_OVER}) on a {sample_rows} x {n} row sample of the workload, "
                      f"{t_all:.2f} s per call on {cores} threads, all products are real GEMMs as in the reference",
            "seconds_per_call": t_all, "tflops": algorithmic_flops(sample_rows, n, K_RANK + S_OVER) / t_all * 1e-12,
            "one_thread": {"value": algorithmic_bytes(rows1, n) / t_one * 1e-9, "unit": "GB/s", "cores": 1, "sample_rows": rows1,
                           "seconds_per_call": t_one, "tflops": algorithmic_flops(rows1, n, K_RANK + S_OVER) / t_one * 1e-12,
                           "note": "the reference itself is single-threaded (nalgebra 0.33 / matrixmultiply 0.3.9 without the threading feature)"},
            "config1_2000x1000_k50": {"ms_all_cores": c1_all * 1e3, "ms_one_thread": c1_one * 1e3, "cores": cores,
                                      "note": "BASELINE config 1 timed in full (median of 5), oracle literal mode"}}


def secondary_configs(torch, rt, _lib, lib, dA_headline, hbm_peak):
    """BASELINE configs 4 and 5 on the same GPU, reusing the headline matrix's memory: the block sparse-sign sketch step and
    the blendenpik solve on a 1M x 2000 problem, rand_evd2 (Nystrom) on a 50k x 50k SPD matrix."""
    import ctypes as C
    from randnla_b200 import lora_drivers as ld
    out = {}

    def timed(fn, reps):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # the headline matrix once more with a sketch narrow enough (l = k + p = 16 < the FP64/HBM ridge of ~21 columns) that the
    # four A-streaming passes are HBM-bound: this is the regime north_star's "70 % of HBM roofline" can physically refer to
    mh, nh = dA_headline.shape
    o_fp64 = rt.make_options(range_passes_int8=0)
    o_auto = rt.make_options(range_passes_int8=-1)
    ms = timed(lambda: ld.rand_svd_dev(dA_headline, 10, 6, o_fp64), 3)
    ph = rt.timings()
    pass_ms = [v for k_, v in ph if k_.startswith("pass:")]
    gbs = 8.0 * mh * nh / (float(np.mean(pass_ms)) * 1e-3) * 1e-9
    assert not any(k_.startswith("i8:") for k_, _ in ph), "the HBM-bound FP64 leg ran integer passes"
    out["c2_matrix_k10_p6_hbm_bound_regime"] = {"mode": "fp64 (thin FP64 kernels, l = 16)", "ms": ms, "mean_pass_ms": float(np.mean(pass_ms)), "A_stream_GBps_per_pass": gbs,
                                                "hbm_frac": gbs / hbm_peak, "whole_call_A_stream_GBps": 4 * 8.0 * mh * nh / (ms * 1e-3) * 1e-9,
                                                "phases_ms": [[k_, v] for k_, v in ph]}
    m, n = 1000000, 2000
    flat = dA_headline.t().reshape(-1)                              # the 32 GB buffer of the headline matrix, column-major
    A4 = flat[: m * n].view(n, m).t()                               # 1M x 2000 view, lda = m (its content: low-rank + noise columns)
    pA, lda = rt.dev_ptr_ld(A4)
    _lib.check(lib.rnla_sketch_fill_dev(0, 0, 77, 9, m, n, 0, pA, lda)); rt.synchronize()
    for d in (8000, 4000):
        dS = rt.empty_colmajor(d, n); pS, lds = rt.dev_ptr_ld(dS)
        ms = timed(lambda: _lib.check(lib.rnla_sketch_apply_dev(2, 0, 5, d, 8, pA, lda, m, n, 0, pS, lds)), 5)
        gbs = 8.0 * m * n / (ms * 1e-3) * 1e-9
        out[f"c4_sketch_step_block_sparse_sign_d{d}"] = {"ms": ms, "A_stream_GBps": gbs, "hbm_frac": gbs / hbm_peak, "zeta": 8, "shape": [m, n]}
    xt = torch.rand(n, 1, dtype=torch.float64, device="cuda") * 200 - 100
    db = rt.empty_colmajor(m, 1); db.copy_(A4 @ xt + 1e-2 * torch.randn(m, 1, dtype=torch.float64, device="cuda"))   # inconsistent system
    dx = rt.empty_colmajor(n, 1); it = C.c_int64(0); cv = C.c_int32(0)
    ms = timed(lambda: _lib.check(lib.rnla_blendenpik_overdetermined_dev(pA, lda, m, n, C.c_void_p(db.data_ptr()), 1e-8, 100, 4.0, 2, 0, 8,
                                                                         C.c_void_p(dx.data_ptr()), C.byref(it), C.byref(cv))), 1)
    def c4_entry(ms):
        ph = rt.timings()
        cg_ms = sum(v for k_, v in ph if k_ == "cgls")
        passes = int(it.value) + 1                                  # one pass per iteration + the initial residual (two-pass mode: 2 x that)
        return {"ms": ms, "cgls_iterations": int(it.value), "converged": bool(cv.value),
                "rel_err_vs_planted": float(torch.linalg.vector_norm(dx - xt) / torch.linalg.vector_norm(xt)),
                "normal_eq_residual": float(torch.linalg.vector_norm(A4.t() @ (db - A4 @ dx)) / torch.linalg.vector_norm(A4.t() @ db)),
                "cgls_ms_per_iteration": cg_ms / passes,
                "iteration_GBps_of_A": 8.0 * m * n / (cg_ms / passes * 1e-3) * 1e-9,
                "phases_ms": [[k_, v] for k_, v in ph]}
    out["c4_blendenpik_block_sparse_sign_sf4"] = c4_entry(ms)
    out["c4_blendenpik_block_sparse_sign_sf4"]["iteration"] = ("one pass over A per CGLS iteration (csrc/normal_pass.cu: clusters hold 32-row slabs in distributed "
                                                               "shared memory; a p and a^T (a p) together): iteration_GBps_of_A counts 8 m n bytes per iteration")
    # the same call on the two streaming mat-vec kernels (A read twice per iteration, the reference's recurrence): RNLA_ONEPASS=0
    x_one = dx.clone()
    os.environ["RNLA_ONEPASS"] = "0"
    try:
        ms2 = timed(lambda: _lib.check(lib.rnla_blendenpik_overdetermined_dev(pA, lda, m, n, C.c_void_p(db.data_ptr()), 1e-8, 100, 4.0, 2, 0, 8,
                                                                              C.c_void_p(dx.data_ptr()), C.byref(it), C.byref(cv))), 1)
        e2 = c4_entry(ms2)
        e2["x_rel_diff_vs_one_pass"] = float(torch.linalg.vector_norm(dx - x_one) / torch.linalg.vector_norm(x_one))
        out["c4_blendenpik_two_pass_iteration"] = e2
    finally:
        del os.environ["RNLA_ONEPASS"]
    # lsqr (src/solvers.rs:115-278) on the same system: ten iterations with the stopping tests off
    res_l = _lib.LsqrResult(); hist = np.zeros(16)
    def run_lsqr():
        _lib.check(lib.rnla_lsqr_dev(pA, lda, m, n, C.c_void_p(db.data_ptr()), 0.0, 0.0, 0.0, 0.0, 10, 0, None, C.c_void_p(dx.data_ptr()),
                                     C.byref(res_l), C.c_void_p(hist.ctypes.data), hist.size, None))
    ms_l1 = timed(run_lsqr, 2)
    os.environ["RNLA_ONEPASS"] = "0"
    try:
        ms_l2 = timed(run_lsqr, 2)
    finally:
        del os.environ["RNLA_ONEPASS"]
    out["c4_lsqr_1Mx2000_10_iterations"] = {"ms_per_iteration_one_pass": ms_l1 / 10.0, "ms_per_iteration_two_pass": ms_l2 / 10.0,
                                            "iteration_GBps_of_A_one_pass": 8.0 * m * n / (ms_l1 / 10.0 * 1e-3) * 1e-9, "itn": int(res_l.itn)}
    nn_, r0, k, s = 50000, 400, 200, 10
    A5 = flat[: nn_ * nn_].view(nn_, nn_).t()
    V0 = rt.empty_colmajor(nn_, r0); pV, ldv = rt.dev_ptr_ld(V0)
    _lib.check(lib.rnla_sketch_fill_dev(0, 0, 31, 9, nn_, r0, 0, pV, ldv))
    _lib.check(lib.rnla_orth_dev(pV, ldv, nn_, r0, 0, None, None)); rt.synchronize()
    lam = np.concatenate([np.logspace(1, -2, 200), np.full(200, 1e-4)])
    Vs = rt.empty_colmajor(nn_, r0); Vs.copy_(V0 * torch.from_numpy(lam).cuda())
    V0t = rt.empty_colmajor(r0, nn_); V0t.copy_(V0.t())
    p5, ld5 = rt.dev_ptr_ld(A5); pVs, ldvs = rt.dev_ptr_ld(Vs); pVt, ldvt = rt.dev_ptr_ld(V0t)
    _lib.check(lib.rnla_gemm_nn_dev(pVs, ldvs, nn_, r0, pVt, ldvt, nn_, p5, ld5)); rt.synchronize()
    A5.diagonal().add_(1e-8)                                        # V0 diag(lam) V0^T is symmetric to rounding; symmetrise exactly in place
    res = {}
    # exact symmetry is a precondition of rand_evd2 (reference :106 style check): mirror the upper triangle blockwise
    B = 5000
    for i0 in range(0, nn_, B):
        for j0 in range(i0, nn_, B):
            blk = A5[i0:i0 + B, j0:j0 + B]
            if i0 == j0:
                blk.copy_(0.5 * (blk + blk.t()))
            else:
                A5[j0:j0 + B, i0:i0 + B].copy_(blk.t())
    torch.cuda.synchronize()
    for tag, oo in (("fp64", o_fp64), ("auto", o_auto)):
        def run5():
            res["V"], res["L"] = ld.rand_evd2_dev(A5, k, s, oo)
        ms = timed(run5, 2)
        L = res["L"].cpu().numpy()
        ph5 = rt.timings()
        assert any(k_.startswith("i8:") for k_, _ in ph5) == (tag == "auto"), f"config 5 leg {tag}: phases {ph5}"
        res["L_" + tag] = L
        out["c5_rand_evd2_50k_k200_" + tag] = {"ms": ms, "tflops_fp64_equivalent": 4 * 2.0 * nn_ * nn_ * (k + s) / (ms * 1e-3) * 1e-12, "r": int(len(L)),
                                               "max_rel_lambda_err_vs_planted": float(np.max(np.abs(L - (lam[:len(L)] + 1e-8)) / lam[:len(L)])),
                                               "phases_ms": [[k_, v] for k_, v in ph5]}
    La, Lf = res["L_auto"], res["L_fp64"]
    out["c5_rand_evd2_50k_k200_auto"]["max_rel_lambda_diff_vs_fp64_path"] = float(np.max(np.abs(La - Lf[:len(La)]) / Lf[:len(La)]))
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port: the Rust reference cannot be built in
    this image) on all host cores, on the GPU arm's config.  Every step is one full `rand_svd` of the N = 1 workload (200 000 x 20 000)
    when the host has the memory for it (A + the transposed copy the reference makes + panels: ~70 GB) and the whole run fits a few
    minutes; otherwise a bounded row sample, and the line says which."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import psutil
    n = args.cols
    cores = host_threads()
    rows_full = args.rows
    # calibrate on a small sample, then take the largest row count whose (warmup + steps) calls stay within the time budget
    cal_rows = 6250
    t_cal, _ = cpu_rand_svd_seconds(workload_matrix(cal_rows, n), K_RANK, S_OVER, cores)
    per_row = t_cal / cal_rows
    calls = max(args.steps, 1) + max(args.warmup, 0)
    fit_rows = int(args.ref_budget_s / (calls * per_row))
    mem_rows = int(0.55 * psutil.virtual_memory().available / (2.2 * 8 * n))
    rows = rows_full if args.ref_rows <= 0 else args.ref_rows
    rows = max(1000, min(rows, fit_rows, mem_rows))
    rows -= rows % 8
    A = workload_matrix(rows, n)
    t, times = cpu_rand_svd_seconds(A, K_RANK, S_OVER, cores, steps=max(args.steps, 1), warmup=max(args.warmup, 0))
    value = algorithmic_bytes(rows, n) / t * 1e-9
    sample = (f"oracle literal-mode rand_svd (k={K_RANK}, s={S_OVER}) on {rows} x {n}"
              + (" = the full N = 1 workload" if rows == rows_full else f" row sample of the {rows_full} x {n} workload (time budget {args.ref_budget_s:.0f} s, "
                 f"host RAM {psutil.virtual_memory().total >> 30} GiB)") + f", {t:.2f} s per call on {cores} threads")
    cb = {"value": value, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample, "seconds_per_call": t,
          "tflops": algorithmic_flops(rows, n, K_RANK + S_OVER) / t * 1e-12, "rows_timed": rows, "full_workload": rows == rows_full}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"rand_svd f64 {args.rows * args.gpus}x{n} low-rank+noise, k={K_RANK}, p={S_OVER}, q={Q_PASSES}, Gaussian sketch",
                   "timed_on": f"{rows}x{n} (CPU, {cores} threads), throughput is per byte of A streamed"},
        "cpu_baseline": cb, "gpu_launches": 0,
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


MODES = {"auto": -1, "fp64": 0, "level1": 1, "level2": 2, "level3": 3}
MODE_DTYPE = {
    "auto": "f64 operands and results; every pass over A FP64-grade on the int8 tensor cores: 55-bit balanced-digit fixed-point split, "
            "exact int32 accumulation in TMEM (rnla_options.range_passes_int8 = -1 'auto', the library default)",
    "level3": "f64 operands and results; every pass over A FP64-grade on the int8 tensor cores: 55-bit balanced-digit fixed-point split, "
              "exact int32 accumulation in TMEM (rnla_options.range_passes_int8 = 3)",
    "fp64": "f64 (every pass on the FP64 DMMA kernels, rnla_options.range_passes_int8 = 0)",
    "level1": "f64 results; A Omega, A^T Y, A S on 31-bit int8 digit planes, Q^T A in FP64 (opt-in, spectrum-conditional: range_passes_int8 = 1)",
    "level2": "f64 results; A Omega, A^T Y, A S on 31-bit int8 digit planes, Q^T A on the 55-bit split (opt-in, spectrum-conditional: range_passes_int8 = 2)",
}


def gpu_numa_cpus(local_rank, policy):
    """CPUs of the NUMA node a rank's pinned staging buffer should live on.  'gpu': the node the GPU hangs off (sysfs); 'spread':
    node = rank mod nodes (used when every GPU reports the same node, so that one socket's memory does not feed all uploads)"""
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        if len(nodes) < 2 or policy == "none":
            return None, None
        bdf = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        node = -1
        if bdf:
            bdf = bdf[-12:] if len(bdf) > 12 else bdf              # sysfs uses a 4-digit domain
            f = f"/sys/bus/pci/devices/{bdf}/numa_node"
            if os.path.exists(f):
                node = int(open(f).read().strip())
        if policy == "spread" or node < 0:
            node = nodes[local_rank % len(nodes)]
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return node, cpus
    except Exception:
        return None, None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ctypes as C
    import randnla_b200 as rb  # noqa: F401
    from randnla_b200 import runtime as rt, _lib, lora_drivers as ld

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: randnla_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    rt.init(local_rank)
    if world > 1:
        rt.init_comm_from_torch()
    # a non-default torch stream: the library launches on it, so torch.cuda.Event brackets see its kernels
    torch.cuda.set_stream(torch.cuda.Stream())
    rt.use_torch_stream()

    n = args.cols
    strong = args.scaling == "strong"
    if strong:
        m_global = args.rows_total
        r_start, r_stop = rt.shard_rows(m_global, world, rank)
        m_local, row_off = r_stop - r_start, r_start
    else:
        m_local = args.rows
        m_global = m_local * world
        row_off = rank * m_local
    l = K_RANK + S_OVER
    sig = planted_sigma()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(x):
        """the value of every rank, in rank order"""
        if world == 1:
            return [float(x)]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = float(x)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    # ---- synthetic input, generated on the device shard by shard ----
    dA = rt.empty_colmajor(m_local, n)
    pA, lda = rt.dev_ptr_ld(dA)
    _lib.check(lib.rnla_generate_lowrank_dev(pA, lda, m_local, n, row_off, m_global, R0, sig.ctypes.data_as(C.c_void_p), 1e-7, 1234))

    def make_opts(mode):
        return rt.make_options(fused_sketch=args.fused, range_passes_int8=MODES[mode])

    # ---- roofs of this box, this run ----
    fp64 = C.c_double(0); hbm = C.c_double(0); i8_burst = C.c_double(0); i8_sust = C.c_double(0)
    _lib.check(lib.rnla_measure_roofs(C.byref(fp64), C.byref(hbm), 4 << 30))
    _lib.check(lib.rnla_measure_int8_roof(C.byref(i8_burst), C.byref(i8_sust)))
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src, bf16_sust = 6650.0, "fallback (B200_PROFILING.md)", None
    if os.path.exists(peaks_file):
        try:
            pk = json.load(open(peaks_file))
            hbm_peak = float(pk["hbm_gbs"]); hbm_src = "MEASURED_PEAKS.json (driver copy benchmark)"
            bf16_sust = float(pk.get("bf16_tflops_sustained", 0.0)) or None
        except Exception:
            pass

    def timed_steps(opts, warmup, steps, collect=False):
        """`steps` calls bracketed by CUDA events on the launching stream (barrier + synchronize on both sides), max over ranks"""
        for _ in range(warmup):
            out = ld.rand_svd_dev(dA, K_RANK, S_OVER, opts)
        barrier()
        acc = {}
        launches0 = rt.kernel_launches()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        walls = []
        for _ in range(steps):
            tw = time.perf_counter()
            out = ld.rand_svd_dev(dA, K_RANK, S_OVER, opts)
            walls.append(1e3 * (time.perf_counter() - tw))      # the call returns with the library's stream drained
            if collect:
                for name, ms in rt.timings():     # library-side CUDA events of this step (stream already drained by the call)
                    acc.setdefault(name, []).append(ms)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1) / steps)
        timed_steps.last_walls = walls
        return ms, out, {k_: float(np.mean(v)) for k_, v in acc.items()}, rt.kernel_launches() - launches0

    # ---- device-resident timing of the headline mode ----
    opts = make_opts(args.mode)
    _lib.check(lib.rnla_set_kernel_timing(1))
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("RNLA_BENCH_NO_CLOCKS"):      # (debug switch: is a stall the sampler's doing?)
        sampler.start()
    for _ in range(args.warmup):                                   # warm-up outside the clock sampling window
        ld.rand_svd_dev(dA, K_RANK, S_OVER, opts)
    barrier()
    sampler.begin()
    ms_step, (U, S, Vt), phases_all, launches = timed_steps(opts, 0, args.steps, collect=True)
    sampler.end()
    headline_walls = list(timed_steps.last_walls)             # host wall time of every timed step (diagnostic: a stalled step shows here)
    clocks = sampler.stop() if rank == 0 else None
    _lib.check(lib.rnla_set_kernel_timing(0))
    kernels = {k_: v for k_, v in phases_all.items() if k_.startswith("k:")}
    phases = {k_: v for k_, v in phases_all.items() if not k_.startswith("k:")}
    value = algorithmic_bytes(m_global, n) / (ms_step * 1e-3) * 1e-9
    int8_ran = any(k_.startswith("i8:") for k_ in phases)
    per_rank = {"int8_ran": all_ranks(1.0 if int8_ran else 0.0), "sum_of_phases_ms": all_ranks(sum(phases.values())),
                "free_hbm_gib": all_ranks(torch.cuda.mem_get_info()[0] / 2**30)}
    int8_ran = all(v > 0.5 for v in per_rank["int8_ran"])

    # ---- accuracy of the timed result (size-independent properties) ----
    Sg = S.cpu().numpy()
    gram = rt.empty_colmajor(K_RANK, K_RANK)
    pU, ldu = rt.dev_ptr_ld(U); pG, ldg = rt.dev_ptr_ld(gram)
    _lib.check(lib.rnla_gemm_tn_dev(pU, ldu, m_local, K_RANK, pU, ldu, K_RANK, pG, ldg, 1))     # U^T U, all-reduced over the shards
    rt.synchronize()
    orth_err = float((gram - torch.eye(K_RANK, dtype=torch.float64, device="cuda")).abs().max().item())
    sigma_vs_planted = float(np.max(np.abs(Sg - sig[:K_RANK]) / sig[:K_RANK]))

    # ---- the other modes beside the headline: all-FP64 is a first-class second value; levels 1 / 2 are the opt-in fast modes ----
    def side(mode, steps):
        oo = make_opts(mode)
        msx, (Ux, Sx, Vx), ph, _ = timed_steps(oo, 2, steps, collect=True)
        Sxh = Sx.cpu().numpy()
        ph = {k_: v for k_, v in ph.items() if not k_.startswith("k:")}
        has_i8 = any(k_.startswith("i8:") for k_ in ph)
        assert has_i8 == (mode != "fp64") or not int8_ran, f"mode {mode}: phases {list(ph)} do not match the requested arithmetic"
        return {"ms_per_step": msx, "value": algorithmic_bytes(m_global, n) / (msx * 1e-3) * 1e-9, "unit": "GB/s", "steps": steps,
                "max_rel_sigma_diff_vs_headline": float(np.max(np.abs(Sg - Sxh) / Sxh)), "phases_ms": ph,
                "dtype": MODE_DTYPE[mode]}
    sides = {}
    for mode in ("fp64", "level1", "level2") if args.mode in ("auto", "level3") else ("fp64",):
        if mode != args.mode:
            sides[mode] = side(mode, min(args.steps, 5) if mode == "fp64" else 3)

    # ---- end to end through the host-buffer C ABI (pinned host memory -> device -> host), headline mode and all-FP64 ----
    e2e = None
    e2e_fp64 = None
    if args.e2e_steps > 0:
        host_bytes = 8 * m_local * n
        import psutil
        avail = psutil.virtual_memory().available
        full = host_bytes * world < 0.7 * avail
        numa_node = None
        if full:
            numa_node, cpus = gpu_numa_cpus(local_rank, args.pin_numa) if world > 1 else (None, None)
            old_aff = None
            if cpus:
                try:
                    old_aff = os.sched_getaffinity(0); os.sched_setaffinity(0, cpus)       # first-touch: the pinned pages land on this node
                except Exception:
                    old_aff = None
            hA = torch.empty((n, m_local), dtype=torch.float64, pin_memory=True)      # column-major m_local x n
            hA.copy_(dA.t())
            if old_aff:
                os.sched_setaffinity(0, old_aff)
            kk = K_RANK
            hU = np.empty((m_local, kk), order="F"); hS = np.empty((kk, kk), order="F"); hVt = np.empty((kk, n), order="F")
            r = C.c_int64(0)

            def e2e_step():
                _lib.check(lib.rnla_rand_svd(C.c_void_p(hA.data_ptr()), m_local, n, kk, 1e-6, S_OVER, rt.ptr(hU), rt.ptr(hS), rt.ptr(hVt), C.byref(r)))
            staging = "full shard in pinned host memory" + (f" on NUMA node {numa_node} ({args.pin_numa})" if numa_node is not None else "")
        else:
            win_rows = 25000
            hW = torch.empty((n, win_rows), dtype=torch.float64, pin_memory=True)
            hW.copy_(dA[:win_rows].t())
            dB = rt.empty_colmajor(m_local, n)
            cur = {"opts": opts}

            def e2e_step():
                for r0 in range(0, m_local, win_rows):
                    rr = min(win_rows, m_local - r0)
                    dB[r0:r0 + rr].t().copy_(hW[:, :rr], non_blocking=True)
                Ue, Se, Vte = ld.rand_svd_dev(dB, K_RANK, S_OVER, cur["opts"])
                Ue.cpu(); Se.cpu(); Vte.cpu()
            staging = f"pinned {win_rows}-row window reused cyclically (host RAM cannot hold {world} x {host_bytes >> 30} GiB)"
        d2h = (m_local * K_RANK + K_RANK + K_RANK * n) * 8

        def run_e2e(mode):
            # the host-buffer entry point reads the process-wide options: set them for this leg only, restored on exit
            with rt.options(fused_sketch=args.fused, range_passes_int8=MODES[mode]):
                if not full:
                    cur["opts"] = make_opts(mode)
                e2e_step()
                names = [nm for nm, _ in rt.timings()]
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    e2e_step()
                barrier()
                dt = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
            assert any("i8:" in nm for nm in names) == (mode != "fp64") or not int8_ran, f"e2e leg {mode}: phases {names}"
            return {"value": algorithmic_bytes(m_global, n) / dt * 1e-9, "unit": "GB/s", "h2d_bytes_per_step": int(host_bytes),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "staging": staging, "mode": mode,
                    "api": "rnla_rand_svd (host buffers, include/rnla.h)" if full else "runtime copy + lora_drivers.rand_svd_dev"}
        e2e = run_e2e(args.mode)
        if args.mode != "fp64":
            e2e_fp64 = run_e2e("fp64")
        if full:
            # what the host can feed: the same pinned shard uploaded by every rank at once, no compute (the floor of e2e)
            dUp = rt.empty_colmajor(m_local, n)
            dUp.t().copy_(hA, non_blocking=True); barrier()
            t0 = time.perf_counter()
            dUp.t().copy_(hA, non_blocking=True)
            barrier()
            up = max_over_ranks(time.perf_counter() - t0)
            for blk in (e2e, e2e_fp64):
                if blk:
                    blk["upload_only_ms"] = up * 1e3
                    blk["upload_only_GBps_per_gpu"] = host_bytes / up * 1e-9
                    blk["upload_only_GBps_all_gpus"] = host_bytes * world / up * 1e-9
            del dUp
        if full:
            del hA
        else:
            del dB
    assert rt.get_options().range_passes_int8 == -1 or os.environ.get("RNLA_RANGE_INT8"), "process-wide options were not restored"

    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    # ---- roofline of the dominant kernel (per launch, CUDA events on the launching stream, timed region) ----
    step_ms = sum(phases.values())
    fl_pass = 2.0 * m_local * n * l
    by_pass = 8.0 * m_local * n
    pass_ms = {k_: v for k_, v in phases.items() if k_.startswith("pass:")}
    ncu = {}
    ncu_file = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if os.path.exists(ncu_file) and m_local == ROWS_PER_GPU and n == N_COLS:
        try:
            ncu = json.load(open(ncu_file))
        except Exception:
            ncu = {}
    step_block = {"algorithmic_bytes": algorithmic_bytes(m_local, n), "ms": ms_step, "achieved_GBps": algorithmic_bytes(m_local, n) / (ms_step * 1e-3) * 1e-9,
                  "frac_hbm": algorithmic_bytes(m_local, n) / (ms_step * 1e-3) * 1e-9 / hbm_peak, "dram_bytes": ncu.get("step_dram_bytes_" + args.mode),
                  "note": "4 passes x 8 m n bytes (SURVEY 8d) over the whole step; dram_bytes = ncu dram read + write summed over the step's kernels"}
    hbm_block = lambda t_ms: {"achieved": by_pass / (t_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s", "frac": by_pass / (t_ms * 1e-3) * 1e-9 / hbm_peak,
                              "peak_source": hbm_src, "read_only_stream_measured_gbs": hbm.value, "bytes": by_pass,
                              "note": "algorithmic bytes of one pass (8 m n: A read once, SURVEY 8d) over the duration of the pass"}
    if int8_ran and kernels:
        # the dominant launch: the sweep over all seven digit planes (groups 3..6, 22 digit-pair MMAs per 32 contraction indices)
        dom = max(kernels, key=lambda k_: kernels[k_])
        pairs = 22 if "planes 7" in dom else (6 if "planes 3" in dom else (11 if "planes 6" in dom else (10 if "groups 0..3" in dom else 6)))
        nmma = 16 * ((l + 15) // 16)
        ops = 2.0 * m_local * n * nmma * pairs
        t_ms = kernels[dom]
        mean_pass = float(np.mean(list(pass_ms.values())))
        roofline = {
            "bound": "tensor", "kernel": dom[2:] + " (CTA pairs: tcgen05.mma.cta_group::2 kind::i8 with M = 256, TMEM int32 accumulators, tiled TMA loads of the pre-tiled digit planes)",
            "achieved": ops / (t_ms * 1e-3) * 1e-12, "peak": i8_sust.value, "unit": "TFLOP/s", "frac": ops / (t_ms * 1e-3) * 1e-12 / i8_sust.value,
            "peak_source": "int8 tensor-core rate measured live by rnla_measure_int8_roof on this GPU, 0.25 s back-to-back under the power cap "
                           "(MEASURED_PEAKS.json holds bf16 only; int8 dense = 2 x bf16 on this part)",
            "peak_burst": i8_burst.value, "measured_peaks_bf16_sustained_x2": 2.0 * bf16_sust if bf16_sust else None,
            "ops_note": "int8 multiply-accumulates executed by the launch, 2 ops each: 2 m n x 112 columns x digit pairs",
            "launch_ms": t_ms, "digit_pairs": pairs, "share_of_step": float(4 * t_ms / step_ms) if "planes 7" in dom else None,
            "traffic": ncu.get("dominant_kernel_dram_bytes_per_launch"),
            "hbm": hbm_block(mean_pass), "fp64_equivalent_tflops_per_pass": fl_pass / (mean_pass * 1e-3) * 1e-12,
            "fp64_dmma_peak_tflops": fp64.value, "per_kernel_ms": kernels, "per_pass_ms": pass_ms,
            "split_of_A": {"rowmax_ms": phases.get("i8:rowmax(A)"), "split_ms": phases.get("i8:split(A)"),
                           "bytes": (8.0 + 8.0 + 7.0) * m_local * n,
                           "note": "row maxima (A read once) + digit split (A read once, seven tiled digit planes written once)"},
            "step": step_block,
        }
    else:
        gemm_ms = list(pass_ms.values())
        per_launch_ms = float(np.mean(gemm_ms)) if gemm_ms else float("nan")
        roofline = {
            "bound": "tensor", "kernel": "gemm_nn_kernel / gemm_tn_kernel (FP64 DMMA.8x8x4, 2 launches each per step)",
            "achieved": fl_pass / (per_launch_ms * 1e-3) * 1e-12, "peak": fp64.value, "unit": "TFLOP/s",
            "frac": fl_pass / (per_launch_ms * 1e-3) * 1e-12 / fp64.value,
            "peak_source": "FP64 DMMA peak measured live by rnla_measure_roofs on this GPU (MEASURED_PEAKS.json holds no FP64 number)",
            "traffic": ncu.get("fp64_gemm_dram_bytes_per_launch"), "hbm": hbm_block(per_launch_ms),
            "note": "l = k+p = 110 makes an FP64 pass FP64-pipe bound (27.5 flop/B vs a 5.7 flop/B ridge), SURVEY.md section 0 fact 3",
            "per_pass_ms": pass_ms, "share_of_step": float(sum(gemm_ms) / step_ms), "step": step_block,
        }

    cb = None
    if world == 1 and args.cpu_rows > 0:
        cb = cpu_baseline(n, args.cpu_rows)

    # ---- other BASELINE configs, one short measurement each (not the headline; device-resident, CUDA events) ----
    secondary = None
    if world == 1 and args.secondary and not strong:
        try:
            secondary = secondary_configs(torch, rt, _lib, lib, dA, hbm_peak)
        except Exception as exc:                                   # never let a side measurement break the headline line
            secondary = {"error": repr(exc)[:300]}

    fp64_first_class = None
    if "fp64" in sides:
        f = sides["fp64"]
        fp64_first_class = {"value": f["value"], "unit": "GB/s", "ms_per_step": f["ms_per_step"], "dtype": f["dtype"], "steps": f["steps"],
                            "e2e": e2e_fp64, "max_rel_sigma_diff_vs_headline": f["max_rel_sigma_diff_vs_headline"], "phases_ms": f["phases_ms"],
                            "fp64_dmma_frac_of_peak": fl_pass / (float(np.mean([v for k_, v in f["phases_ms"].items() if k_.startswith("pass:")])) * 1e-3) * 1e-12 / fp64.value}
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": MODE_DTYPE[args.mode] if int8_ran or args.mode == "fp64" else MODE_DTYPE["fp64"] + " [the int8 workspace did not fit: FP64 fallback]",
        "data": "synthetic",
        "config": {"workload": f"rand_svd f64 {m_global}x{n} low-rank+noise, k={K_RANK}, p={S_OVER}, q={Q_PASSES}, Gaussian sketch",
                   "rows_per_gpu": m_local, "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
                   "sketch": ("in-kernel Philox: the int8 digit planes of Omega are formed from the Philox blocks inside the operand kernels of A*Omega; "
                              "Omega is never materialised in FP64" if any("Philox inside" in k_ for k_ in phases) else
                              "auto (materialised while Omega is L2-resident)" if args.fused == 2 else ("fused in-kernel Philox" if args.fused == 1 else "materialised")),
                   "mode": args.mode, "range_passes_int8": MODES[args.mode],
                   "l2": f"inputs larger than L2 (A shard = {8 * m_local * n / 2**30:.1f} GiB per GPU, streamed 4x per step)"},
        "rand_svd_ms": ms_step, "step_wall_ms": [round(w, 2) for w in headline_walls], "tflops_fp64_equivalent": algorithmic_flops(m_global, n, l) / (ms_step * 1e-3) * 1e-12,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cb, "phases_ms": phases, "secondary": secondary,
        "accuracy": {"max_abs_UtU_minus_I": orth_err, "max_rel_sigma_vs_planted(noise-limited)": sigma_vs_planted},
        "fp64": fp64_first_class, "other_modes": {k_: v for k_, v in sides.items() if k_ != "fp64"},
        "per_rank": per_rank,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="rows per GPU (debug)")
    ap.add_argument("--cols", type=int, default=N_COLS)
    ap.add_argument("--fused", type=int, default=2, help="0 materialise Omega, 1 in-kernel Philox, 2 auto")
    ap.add_argument("--mode", default="auto", choices=sorted(MODES),
                    help="rnla_options.range_passes_int8 of the headline: auto = the library default (every pass FP64-grade on the int8 tensor cores, "
                         "55-bit split); fp64 = the FP64 DMMA kernels; level1 / level2 = the opt-in, spectrum-conditional 31-bit range passes")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --rows per GPU; strong: --rows-total sharded over the ranks")
    ap.add_argument("--rows-total", type=int, default=2000000, help="global rows of --scaling strong (BASELINE config 3: 2M x 20k)")
    ap.add_argument("--pin-numa", default="gpu", choices=["gpu", "spread", "none"],
                    help="N > 1: NUMA node of each rank's pinned staging buffer (gpu: the node its GPU hangs off; spread: rank mod nodes)")
    ap.add_argument("--ref-budget-s", type=float, default=300.0, help="time budget of --impl reference; the rows per step shrink to fit it")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-rows", type=int, default=25000, help="row sample of the cpu_baseline leg (0 = skip)")
    ap.add_argument("--secondary", type=int, default=1, help="1: also time BASELINE configs 4 and 5 once (N=1 only, reported under 'secondary')")
    ap.add_argument("--ref-rows", type=int, default=0, help="rows per step of --impl reference (0: the full N = 1 workload, memory and budget permitting)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
