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
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(n, sample_rows, threads=None, steps=1, warmup=0):
    """The reference's CPU dataflow (oracle `literal` mode: src/lora_helpers.rs:17-146 + src/lora_drivers.rs:30-69
    statement for statement) timed on the host cores, on a bounded row sample of the same workload."""
    from oracle import oracle as orc
    orc.load()
    if threads:
        orc.set_threads(threads)
    rng = np.random.default_rng(1234)
    sig = planted_sigma()
    U0, _ = np.linalg.qr(rng.standard_normal((sample_rows, R0)))
    V0, _ = np.linalg.qr(rng.standard_normal((n, R0)))
    A = np.asfortranarray((U0 * sig) @ V0.T)
    del U0
    o = orc.make_opts(mode=orc.MODE_LITERAL)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        U, S, Vt = orc.rand_svd(A, K_RANK, 1e-6, S_OVER, o)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    return {"value": algorithmic_bytes(sample_rows, n) / t * 1e-9, "unit": "GB/s", "cores": orc.get_threads(), "kind": "port",
            "sample": f"oracle literal-mode rand_svd (k={K_RANK}, s={S_OVER}) on {sample_rows} x {n} rows-sample of the workload, "
                      f"{t:.2f} s per call, all products are real GEMMs as in the reference",
            "seconds_per_call": t, "tflops": algorithmic_flops(sample_rows, n, K_RANK + S_OVER) / t * 1e-12}


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
    ms = timed(lambda: ld.rand_svd_dev(dA_headline, 10, 6), 3)
    ph = rt.timings()
    pass_ms = [v for k_, v in ph if k_.startswith("pass:")]
    gbs = 8.0 * mh * nh / (float(np.mean(pass_ms)) * 1e-3) * 1e-9
    out["c2_matrix_k10_p6_hbm_bound_regime"] = {"ms": ms, "mean_pass_ms": float(np.mean(pass_ms)), "A_stream_GBps_per_pass": gbs,
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
    out["c4_blendenpik_block_sparse_sign_sf4"] = {"ms": ms, "cgls_iterations": int(it.value), "converged": bool(cv.value),
                                                  "rel_err_vs_planted": float(torch.linalg.vector_norm(dx - xt) / torch.linalg.vector_norm(xt)),
                                                  "normal_eq_residual": float(torch.linalg.vector_norm(A4.t() @ (db - A4 @ dx)) / torch.linalg.vector_norm(A4.t() @ db)),
                                                  "phases_ms": [[k, v] for k, v in rt.timings()]}
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
    def run5():
        res["V"], res["L"] = ld.rand_evd2_dev(A5, k, s)
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
    ms = timed(run5, 2)
    L = res["L"].cpu().numpy()
    out["c5_rand_evd2_50k_k200"] = {"ms": ms, "tflops_fp64": 4 * 2.0 * nn_ * nn_ * (k + s) / (ms * 1e-3) * 1e-12, "r": int(len(L)),
                                    "max_rel_lambda_err_vs_planted": float(np.max(np.abs(L - (lam[:len(L)] + 1e-8)) / lam[:len(L)])),
                                    "phases_ms": [[k_, v] for k_, v in rt.timings()]}
    # the same call with every pass over A on the integer tensor cores (range_passes_int8 = 2: power iteration on the 28-bit
    # split, Y = A S as A^T S on the 49-bit split; l = 210 columns = two 128-column MMA tiles)
    try:
        o8 = rt.make_options(range_passes_int8=2)
        def run5i():
            res["V8"], res["L8"] = ld.rand_evd2_dev(A5, k, s, o8)
        ms8 = timed(run5i, 2)
        L8 = res["L8"].cpu().numpy()
        out["c5_rand_evd2_50k_k200_int8"] = {"ms": ms8, "r": int(len(L8)),
                                             "max_rel_lambda_diff_vs_fp64_path": float(np.max(np.abs(L8 - L[:len(L8)]) / L[:len(L8)])),
                                             "max_rel_lambda_err_vs_planted": float(np.max(np.abs(L8 - (lam[:len(L8)] + 1e-8)) / lam[:len(L8)])),
                                             "phases_ms": [[k_, v] for k_, v in rt.timings()]}
    except Exception as exc:
        out["c5_rand_evd2_50k_k200_int8"] = {"error": repr(exc)[:200]}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cols
    sample_rows = args.ref_rows
    cb = cpu_baseline(n, sample_rows, steps=max(args.steps, 1), warmup=max(args.warmup, 0))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_call"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"rand_svd f64 {args.rows * args.gpus}x{n} low-rank+noise, k={K_RANK}, p={S_OVER}, q={Q_PASSES}, Gaussian sketch",
                   "timed_on": f"{sample_rows}x{n} row sample (CPU), throughput is per byte of A streamed"},
        "cpu_baseline": cb, "gpu_launches": 0,
        "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ctypes as C
    import randnla_b200 as rb
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

    # ---- synthetic input, generated on the device shard by shard ----
    dA = rt.empty_colmajor(m_local, n)
    pA, lda = rt.dev_ptr_ld(dA)
    _lib.check(lib.rnla_generate_lowrank_dev(pA, lda, m_local, n, row_off, m_global, R0, sig.ctypes.data_as(C.c_void_p), 1e-7, 1234))
    use_i8 = args.range != "fp64"
    i8_level = {"fp64": 0, "int8": 1, "int8-all": 2}[args.range]
    opts = rt.make_options(fused_sketch=args.fused, range_passes_int8=i8_level)

    # ---- roofs of this box, this run ----
    fp64 = C.c_double(0); hbm = C.c_double(0)
    _lib.check(lib.rnla_measure_roofs(C.byref(fp64), C.byref(hbm), 4 << 30))
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_file):
        try:
            hbm_peak = float(json.load(open(peaks_file))["hbm_gbs"]); hbm_src = "MEASURED_PEAKS.json (driver copy benchmark)"
        except Exception:
            pass

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        U, S, Vt = ld.rand_svd_dev(dA, K_RANK, S_OVER, opts)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = rt.kernel_launches()
    phase_acc = {}
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        U, S, Vt = ld.rand_svd_dev(dA, K_RANK, S_OVER, opts)
        for name, ms in rt.timings():             # library-side CUDA events of this step (stream already drained by the call)
            phase_acc.setdefault(name, []).append(ms)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = rt.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(ms_total / args.steps)
    value = algorithmic_bytes(m_global, n) / (ms_step * 1e-3) * 1e-9

    # ---- accuracy of the timed result (size-independent properties) ----
    Sg = S.cpu().numpy()
    gram = rt.empty_colmajor(K_RANK, K_RANK)
    pU, ldu = rt.dev_ptr_ld(U); pG, ldg = rt.dev_ptr_ld(gram)
    _lib.check(lib.rnla_gemm_tn_dev(pU, ldu, m_local, K_RANK, pU, ldu, K_RANK, pG, ldg, 1))     # U^T U, all-reduced over the shards
    rt.synchronize()
    orth_err = float((gram - torch.eye(K_RANK, dtype=torch.float64, device="cuda")).abs().max().item())
    sigma_vs_planted = float(np.max(np.abs(Sg - sig[:K_RANK]) / sig[:K_RANK]))

    # ---- the same call with all four passes in FP64, beside the headline (sigma agreement between the two, time) ----
    fp64_side = None
    if use_i8:
        def side(level):
            oo = rt.make_options(fused_sketch=args.fused, range_passes_int8=level)
            for _ in range(2):
                Ux, Sx, Vx = ld.rand_svd_dev(dA, K_RANK, S_OVER, oo)
            per = []
            for _ in range(3):
                barrier()
                f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
                f0.record()
                Ux, Sx, Vx = ld.rand_svd_dev(dA, K_RANK, S_OVER, oo)
                f1.record()
                barrier()
                per.append(max_over_ranks(f0.elapsed_time(f1)))
            msx = float(np.median(per))
            Sxh = Sx.cpu().numpy()
            return {"ms_per_step": msx, "A_stream_GBps": algorithmic_bytes(m_global, n) / (msx * 1e-3) * 1e-9,
                    "max_rel_sigma_diff_vs_headline": float(np.max(np.abs(Sg - Sxh) / Sxh)), "phases_ms": dict(rt.timings())}
        fp64_side = {"all_fp64": side(0)}
        if i8_level == 2:
            fp64_side["int8_range_passes_fp64_QtA"] = side(1)

    # ---- end to end through the host-buffer C ABI (pinned host memory -> device -> host) ----
    e2e = None
    if args.e2e_steps > 0:
        rt.set_options(range_passes_int8=i8_level)          # the host-buffer entry point reads the process-wide options
        host_bytes = 8 * m_local * n
        import psutil
        avail = psutil.virtual_memory().available
        full = host_bytes * world < 0.7 * avail if world > 1 else host_bytes < 0.7 * avail
        if full:
            hA = torch.empty((n, m_local), dtype=torch.float64, pin_memory=True)      # column-major m_local x n
            hA.copy_(dA.t())
            hAn = hA.numpy().T
            kk = K_RANK
            hU = np.empty((m_local, kk), order="F"); hS = np.empty((kk, kk), order="F"); hVt = np.empty((kk, n), order="F")
            r = C.c_int64(0)

            def e2e_step():
                _lib.check(lib.rnla_rand_svd(C.c_void_p(hA.data_ptr()), m_local, n, kk, 1e-6, S_OVER, rt.ptr(hU), rt.ptr(hS), rt.ptr(hVt), C.byref(r)))
            staging = "full shard in pinned host memory"
            d2h = (m_local * kk + kk + kk * n) * 8
        else:
            win_rows = 25000
            hW = torch.empty((n, win_rows), dtype=torch.float64, pin_memory=True)
            hW.copy_(dA[:win_rows].t())
            dB = rt.empty_colmajor(m_local, n)

            def e2e_step():
                for r0 in range(0, m_local, win_rows):
                    rr = min(win_rows, m_local - r0)
                    dB[r0:r0 + rr].t().copy_(hW[:, :rr], non_blocking=True)
                Ue, Se, Vte = ld.rand_svd_dev(dB, K_RANK, S_OVER, opts)
                Ue.cpu(); Se.cpu(); Vte.cpu()
            staging = f"pinned {win_rows}-row window reused cyclically (host RAM cannot hold {world} x {host_bytes >> 30} GiB)"
            d2h = (m_local * K_RANK + K_RANK + K_RANK * n) * 8
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e = {"value": algorithmic_bytes(m_global, n) / dt * 1e-9, "unit": "GB/s", "h2d_bytes_per_step": int(host_bytes),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "staging": staging,
               "api": "rnla_rand_svd (host buffers, include/rnla.h)" if full else "runtime copy + lora_drivers.rand_svd_dev"}
        if full:
            del hA
        else:
            del dB

    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    # ---- roofline of the dominant kernels (per launch, CUDA events on the launching stream, timed region) ----
    phases = {k: float(np.mean(v)) for k, v in phase_acc.items()}
    fl = 2.0 * m_local * n * l
    by = 8.0 * m_local * n
    # FP64 DMMA passes of the step: all four with --range fp64, only "pass:At*Q" (the one that carries sigma) with --range int8
    i8_names = (("pass:A*Omega", "pass:At*Y", "pass:A*S", "pass:At*Q") if i8_level == 2 else ("pass:A*Omega", "pass:At*Y", "pass:A*S")) if use_i8 else ()
    fp64_ms = {k: v for k, v in phases.items() if k.startswith("pass:") and not k.startswith(i8_names)} if use_i8 else \
              {k: v for k, v in phases.items() if k.startswith("pass:")}
    i8_ms = {k: v for k, v in phases.items() if use_i8 and k.startswith(i8_names)}
    nn_ms = [v for k, v in fp64_ms.items() if k.startswith("pass:A*")]
    tn_ms = [v for k, v in fp64_ms.items() if k.startswith("pass:At*")]
    gemm_ms = nn_ms + tn_ms
    step_ms = sum(phases.values())
    fp64_block = None
    if gemm_ms:
        per_launch_ms = float(np.mean(gemm_ms))
        fp64_block = {
            "bound": "tensor", "kernel": "gemm_tn_kernel (FP64 DMMA.8x8x4): B = Q^T A, the pass that carries the singular values" if use_i8
                     else "gemm_nn_kernel / gemm_tn_kernel (FP64 DMMA.8x8x4, 2 launches each per step)",
            "achieved": fl / (per_launch_ms * 1e-3) * 1e-12, "peak": fp64.value, "unit": "TFLOP/s",
            "frac": fl / (per_launch_ms * 1e-3) * 1e-12 / fp64.value,
            "peak_source": "FP64 DMMA peak measured live by rnla_measure_roofs on this GPU (MEASURED_PEAKS.json holds no FP64 number)",
            "traffic": None,
            "hbm": {"achieved": by / (per_launch_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": by / (per_launch_ms * 1e-3) * 1e-9 / hbm_peak, "peak_source": hbm_src, "read_only_stream_measured_gbs": hbm.value},
            "note": "l = k+p = 110 makes an FP64 pass FP64-pipe bound (27.5 flop/B vs a 5.7 flop/B ridge), SURVEY.md §0 fact 3",
            "per_kernel_ms": {"gemm_nn": float(np.mean(nn_ms)) if nn_ms else None, "gemm_tn": float(np.mean(tn_ms)) if tn_ms else None},
            "share_of_step": float(sum(gemm_ms) / step_ms),
        }
    if i8_level == 2:
        # no FP64 pass left: the longest single launch of the step is the digit split of A (HBM-bound: A read once, the two tiled
        # 4-plane images and the 3-plane image of the trailing digits written: 8 + 4 + 4 + 3 = 19 bytes per element)
        sp_ms = phases.get("i8:split(A)", float("nan"))
        sp_by = 19.0 * m_local * n
        roofline = {"bound": "hbm", "kernel": "slice_a_kernel<7 digits> (FP64 -> seven balanced 7-bit digit planes, pre-tiled MMA images)",
                    "achieved": sp_by / (sp_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s", "frac": sp_by / (sp_ms * 1e-3) * 1e-9 / hbm_peak,
                    "peak_source": hbm_src, "traffic": None, "bytes_per_launch": sp_by, "share_of_step": float(sp_ms / step_ms),
                    "note": "with --range int8-all no FP64 GEMM is left in the step; per-kernel blocks for the integer passes below, "
                            "the FP64 DMMA kernels are measured in all_fp64_passes / --range fp64 (95 % of the DMMA peak)"}
    else:
        roofline = fp64_block
    if use_i8 and i8_ms:
        # the integer passes stream the 4 digit planes of A: 4 bytes per element per sweep (A S makes two sweeps), HBM-bound
        # bytes of digit planes streamed: 4 per element per sweep; A S makes two sweeps; the 49-bit Q^T A reads 4 + 7
        sweeps = {k: (2.0 if k.startswith("pass:A*S") else (2.75 if k.startswith("pass:At*Q") else 1.0)) for k in i8_ms}
        pairs = {k: (16.0 if k.startswith("pass:A*S") else (28.0 if k.startswith("pass:At*Q") else 10.0)) for k in i8_ms}
        tot_ms = sum(i8_ms.values()); tot_by = sum(4.0 * m_local * n * sweeps[k] for k in i8_ms)
        split_ms = phases.get("i8:rowmax(A)", 0.0) + phases.get("i8:split(A)", 0.0)
        roofline["int8_passes"] = {
            "bound": "hbm", "kernel": "i8_mma_kernel (tcgen05.mma kind::i8, TMEM accumulators, cp.async.bulk of pre-tiled digit planes)",
            "achieved": tot_by / (tot_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s", "frac": tot_by / (tot_ms * 1e-3) * 1e-9 / hbm_peak,
            "per_pass_ms": i8_ms, "bytes_per_sweep": 4.0 * m_local * n,
            "tensor_TOPS": sum(pairs.values()) * 2.0 * m_local * n * 128 / (tot_ms * 1e-3) * 1e-12, "digit_pair_mmas": pairs,
            "f64_equivalent_A_stream_GBps": by * len(i8_ms) / (tot_ms * 1e-3) * 1e-9,
            "split_of_A": {"ms": split_ms, "bytes": (16.0 + (11.0 if i8_level == 2 else 8.0)) * m_local * n,
                           "GBps": (16.0 + (11.0 if i8_level == 2 else 8.0)) * m_local * n / (split_ms * 1e-3) * 1e-9 if split_ms else None,
                           "note": "row maxima (A read once) + digit split (A read once, tiled digit-plane images written)"},
        }
    if i8_level == 2 and fp64_block is None:
        roofline["fp64_dmma_kernels"] = "see all_fp64_passes"
        ncu8 = os.path.join(ROOT, "profiles", "r01_ncu_traffic_int8.json")
        if os.path.exists(ncu8) and m_local == ROWS_PER_GPU and n == N_COLS:
            try:
                roofline["traffic"] = json.load(open(ncu8)).get("split_bytes_per_launch")     # dram read + write of the split kernel (ncu)
            except Exception:
                pass
    ncu_traffic = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(ncu_traffic) and i8_level != 2:
        try:
            roofline["traffic"] = json.load(open(ncu_traffic)).get("gemm_bytes_per_launch")
        except Exception:
            pass

    cb = None
    if world == 1 and args.cpu_rows > 0:
        cb = cpu_baseline(n, args.cpu_rows)

    # ---- other BASELINE configs, one short measurement each (not the headline; device-resident, CUDA events) ----
    secondary = None
    if world == 1 and args.secondary:
        try:
            secondary = secondary_configs(torch, rt, _lib, lib, dA, hbm_peak)
        except Exception as exc:                                   # never let a side measurement break the headline line
            secondary = {"error": repr(exc)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": ["f64", "f64 (range-finder passes: int8 tensor cores on a 28-bit fixed-point split of A; Q^T A, factorisations and outputs f64)",
                                        "f64 results; every pass over A on the int8 tensor cores with exact int32 accumulation: 28-bit balanced-digit split for the "
                                        "range-finder passes, 49-bit split (28 digit pairs) for Q^T A; factorisations, small products and outputs f64"][i8_level],
        "data": "synthetic",
        "config": {"workload": f"rand_svd f64 {m_global}x{n} low-rank+noise, k={K_RANK}, p={S_OVER}, q={Q_PASSES}, Gaussian sketch",
                   "rows_per_gpu": m_local, "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
                   "sketch": "auto (materialised while Omega is L2-resident)" if args.fused == 2 else ("fused in-kernel Philox" if args.fused == 1 else "materialised"),
                   "range_passes": ["fp64", "int8 tensor cores, Q^T A in FP64 (rnla_options.range_passes_int8 = 1)",
                                    "int8 tensor cores, Q^T A on a 49-bit split (rnla_options.range_passes_int8 = 2)"][i8_level],
                   "l2": f"inputs larger than L2 (A shard = {8 * m_local * n / 2**30:.1f} GiB per GPU, streamed 4x per step)"},
        "rand_svd_ms": ms_step, "tflops_fp64": algorithmic_flops(m_global, n, l) / (ms_step * 1e-3) * 1e-12,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cb, "phases_ms": phases, "secondary": secondary,
        "accuracy": {"max_abs_UtU_minus_I": orth_err, "max_rel_sigma_vs_planted(noise-limited)": sigma_vs_planted},
        "all_fp64_passes": fp64_side,
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
    ap.add_argument("--range", default="int8-all", choices=["int8-all", "int8", "fp64"],
                    help="int8: the three range-finder passes (A Omega, A^T Y, A S) run on the INT8 tensor cores from a 4 x 7-bit split of A "
                         "(rnla_options.range_passes_int8 = 1), the pass that carries the singular values (Q^T A) in FP64; int8-all: Q^T A too, "
                         "on a 49-bit split with exact int32 accumulation (= 2); fp64: all four passes on the FP64 DMMA kernels")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-rows", type=int, default=25000, help="row sample of the cpu_baseline leg (0 = skip)")
    ap.add_argument("--secondary", type=int, default=1, help="1: also time BASELINE configs 4 and 5 once (N=1 only, reported under 'secondary')")
    ap.add_argument("--ref-rows", type=int, default=25000, help="row sample per step of --impl reference")
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
