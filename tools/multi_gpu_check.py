"""Row-sharded drivers on real GPUs against the single-GPU result for the SAME global matrix.
    python tools/multi_gpu_check.py                      # single process: writes gpurun_out/mgpu_ref.npz
    torchrun --nproc-per-node N tools/multi_gpu_check.py # N ranks: compares with the file, prints one JSON line on rank 0
Checks rand_svd (sigma to 1e-10 relative, U^T U = I, residual), the block sparse-sign sketch (sum of the shards' sketches
= sketch of the whole), blendenpik (same x), and lsqr / plain cgls on the row-sharded system (same x, same iteration counts, same
norm estimates)."""
import ctypes as C, json, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import torch.distributed as dist
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = _lib.load(); rt.init(lr)
if world > 1:
    rt.init_comm_from_torch()
M, N, K, S = 48000, 3000, 40, 10
ml = M // world; off = rank * ml
sig = np.concatenate([np.logspace(0, -3, 40), np.full(40, 1e-6)])
dA = rt.empty_colmajor(ml, N); pA, lda = rt.dev_ptr_ld(dA)
_lib.check(lib.rnla_generate_lowrank_dev(pA, lda, ml, N, off, M, 80, sig.ctypes.data_as(C.c_void_p), 1e-9, 77))
U, Sg, Vt = ld.rand_svd_dev(dA, K, S); rt.synchronize()
s = Sg.cpu().numpy()
gram = U.t() @ U
res2 = torch.linalg.matrix_norm(dA - (U * Sg) @ Vt) ** 2; a2 = torch.linalg.matrix_norm(dA) ** 2
if world > 1:
    dist.all_reduce(gram); dist.all_reduce(res2); dist.all_reduce(a2)
orth = float((gram - torch.eye(K, dtype=torch.float64, device="cuda")).abs().max())
relres = float(torch.sqrt(res2 / a2))
# block sparse-sign sketch of the shard, summed over ranks by the library's all-reduce
d = 1200
dS = rt.empty_colmajor(d, N); pS, lds = rt.dev_ptr_ld(dS)
_lib.check(lib.rnla_sketch_apply_dev(2, 0, 5, d, 8, pA, lda, ml, N, off, pS, lds)); rt.synchronize()
sk = dS.cpu().numpy()
# blendenpik on a least-squares problem built from the same shard.  The solvers take 1600 columns: within the width (n <= 2048) at which
# their iterations read A once per iteration (csrc/normal_pass.cu), so the sharded one-pass dataflow (one all-reduce of n + 1 doubles per
# iteration) is what runs here
N_RANDSVD = N
N = 1600
torch.manual_seed(1)
xt = torch.rand(N, 1, dtype=torch.float64, device="cuda") * 2 - 1
dB = rt.empty_colmajor(ml, N); pB, ldb = rt.dev_ptr_ld(dB)
_lib.check(lib.rnla_sketch_fill_dev(0, 0, 9, 9, ml, N, off, pB, ldb)); rt.synchronize()
one_pass = bool(lib.rnla_normal_pass_supported(pB, ldb, ml, N))
db = rt.empty_colmajor(ml, 1); db.copy_(dB @ xt)
dx = rt.empty_colmajor(N, 1); it = C.c_int64(0); cv = C.c_int32(0)
_lib.check(lib.rnla_blendenpik_overdetermined_dev(pB, ldb, ml, N, C.c_void_p(db.data_ptr()), 1e-9, 100, 2.0, 2, 0, 8,
                                                  C.c_void_p(dx.data_ptr()), C.byref(it), C.byref(cv)))
x = dx.cpu().numpy(); xerr = float(torch.linalg.vector_norm(dx - xt) / torch.linalg.vector_norm(xt))
# lsqr and plain cgls (reference src/solvers.rs:115-278, src/cg.rs:18-61) on the same row-sharded system, inconsistent right-hand side
torch.manual_seed(100 + rank)
db2 = rt.empty_colmajor(ml, 1); db2.copy_(db)
g = torch.Generator(device="cuda").manual_seed(7)
noise = torch.randn(M, 1, dtype=torch.float64, device="cuda", generator=g)[off:off + ml] * 1e-3      # same global vector on every layout
db2.add_(noise)
dxl = rt.empty_colmajor(N, 1); res = _lib.LsqrResult(); hist = np.zeros(2 * N)
_lib.check(lib.rnla_lsqr_dev(pB, ldb, ml, N, C.c_void_p(db2.data_ptr()), 0.0, 1e-12, 1e-12, 1e8, -1, 0, None, C.c_void_p(dxl.data_ptr()),
                             C.byref(res), C.c_void_p(hist.ctypes.data), hist.size, None))
xl = dxl.cpu().numpy()
dxc = rt.empty_colmajor(N, 1); dxc.zero_(); itc = C.c_int64(0); cvc = C.c_int32(0)
_lib.check(lib.rnla_cgls_dev(pB, ldb, ml, N, C.c_void_p(db2.data_ptr()), 1e-9, 200, C.c_void_p(dxc.data_ptr()), C.byref(itc), C.byref(cvc)))
xc = dxc.cpu().numpy()
lsq = {"solver_iterations_read_A_once": one_pass, "lsqr_istop": int(res.istop), "lsqr_itn": int(res.itn), "lsqr_r1norm": res.r1norm, "lsqr_anorm": res.anorm, "lsqr_xnorm": res.xnorm,
       "cgls_plain_iterations": int(itc.value), "cgls_plain_converged": bool(cvc.value),
       "lsqr_vs_cgls_x_rel_diff": float(np.linalg.norm(xl - xc) / np.linalg.norm(xc))}
os.makedirs("gpurun_out", exist_ok=True)
if world == 1:
    np.savez("gpurun_out/mgpu_ref.npz", s=s, sk=sk, x=x, xl=xl, xc=xc, lsq=np.array([res.itn, res.r1norm, res.anorm, res.xnorm, itc.value]))
    print(json.dumps({"n_gpus": 1, **lsq, "sigma_head": s[:3].tolist(), "orth": orth, "relres": relres, "cgls_iterations": int(it.value), "x_rel_err": xerr}))
elif rank == 0:
    ref = np.load("gpurun_out/mgpu_ref.npz")
    out = {"n_gpus": world, "max_rel_sigma_diff_vs_1gpu": float(np.max(np.abs(s - ref["s"]) / ref["s"])), "orth": orth, "relres": relres,
           "saso_block_max_abs_diff_vs_1gpu": float(np.abs(sk - ref["sk"]).max() / np.abs(ref["sk"]).max()),
           "blendenpik_x_rel_diff_vs_1gpu": float(np.linalg.norm(x - ref["x"]) / np.linalg.norm(ref["x"])), "cgls_iterations": int(it.value),
           "converged": bool(cv.value), "x_rel_err": xerr, **lsq,
           "lsqr_x_rel_diff_vs_1gpu": float(np.linalg.norm(xl - ref["xl"]) / np.linalg.norm(ref["xl"])),
           "cgls_plain_x_rel_diff_vs_1gpu": float(np.linalg.norm(xc - ref["xc"]) / np.linalg.norm(ref["xc"])),
           "lsqr_itn_1gpu": int(ref["lsq"][0]), "lsqr_r1norm_rel_diff_vs_1gpu": float(abs(res.r1norm - ref["lsq"][1]) / ref["lsq"][1]),
           "cgls_plain_iterations_1gpu": int(ref["lsq"][4])}
    out["pass"] = bool(out["lsqr_x_rel_diff_vs_1gpu"] < 1e-9 and out["cgls_plain_x_rel_diff_vs_1gpu"] < 1e-9 and out["lsqr_r1norm_rel_diff_vs_1gpu"] < 1e-9 and
                       abs(out["lsqr_itn"] - out["lsqr_itn_1gpu"]) <= 1 and out["max_rel_sigma_diff_vs_1gpu"] < 1e-10 and orth < 1e-12 and out["saso_block_max_abs_diff_vs_1gpu"] < 1e-12
                       and out["blendenpik_x_rel_diff_vs_1gpu"] < 1e-8 and out["converged"])
    print(json.dumps(out))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
