"""ctypes binding of librnla.so (include/rnla.h).  No compute happens in Python and there is no CPU
fallback: if the shared library is missing, loading fails loudly; if no B200 is visible, every compute
entry point returns RNLA_ERR_COMPUTATION, surfaced as `ComputationError`."""
import ctypes as C
import os

from .errors import STATUS_TO_ERROR, ComputationError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librnla.so")

c_i32, c_i64, c_u32, c_u64, c_f64 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
P = C.c_void_p


class Options(C.Structure):
    """`rnla_options` (include/rnla.h)."""
    _fields_ = [("mode", c_i32), ("dist", c_i32), ("seed", c_u64), ("num_passes", c_i32),
                ("passes_per_stab", c_i32), ("fused_sketch", c_i32), ("range_passes_int8", c_i32), ("generator", c_i32),
                ("reserved_", c_i32)]


class LsqrResult(C.Structure):
    """`rnla_lsqr_result` (include/rnla.h)."""
    _fields_ = [("istop", c_i64), ("itn", c_i64), ("r1norm", c_f64), ("r2norm", c_f64), ("anorm", c_f64), ("acond", c_f64),
                ("xnorm", c_f64), ("n_arnorms", c_i64)]


# name -> (restype, argtypes); every symbol include/rnla.h declares
SIGNATURES = {
    "rnla_version": (c_i32, []),
    "rnla_last_error_message": (C.c_char_p, []),
    "rnla_init": (c_i32, [c_i32]),
    "rnla_shutdown": (None, []),
    "rnla_stream": (P, []),
    "rnla_set_stream": (c_i32, [P]),
    "rnla_synchronize": (c_i32, []),
    "rnla_release_workspace": (c_i32, []),
    "rnla_default_options": (None, [C.POINTER(Options)]),
    "rnla_set_options": (c_i32, [C.POINTER(Options)]),
    "rnla_get_options": (None, [C.POINTER(Options)]),
    "rnla_kernel_launches": (c_u64, []),
    "rnla_set_kernel_timing": (c_i32, [c_i32]),
    "rnla_get_timings": (c_i32, [C.POINTER(C.c_char_p), C.POINTER(c_f64), c_i32]),
    "rnla_comm_unique_id": (c_i32, [P]),
    "rnla_comm_init": (c_i32, [c_i32, c_i32, P]),
    "rnla_comm_destroy": (c_i32, []),
    "rnla_comm_size": (c_i32, []),
    "rnla_comm_rank": (c_i32, []),
    "rnla_philox4x32_10": (c_i32, [c_i64, P, P, P]),
    "rnla_threefry2x64_20": (c_i32, [c_i64, P, P, P]),
    "rnla_sketching_operator": (c_i32, [c_i32, c_i64, c_i64, P]),
    "rnla_sketch_fill": (c_i32, [c_i32, c_i32, c_u64, c_u32, c_i64, c_i64, c_i64, P, c_i64]),
    "rnla_sketch_fill_dev": (c_i32, [c_i32, c_i32, c_u64, c_u32, c_i64, c_i64, c_i64, P, c_i64]),
    "rnla_ziggurat_tables": (c_i32, [P, P]),
    "rnla_haar_sample": (c_i32, [c_i64, c_i64, c_i32, P]),
    "rnla_orth": (c_i32, [P, c_i64, c_i64, P, P, C.POINTER(c_i64)]),
    "rnla_stabilizer": (c_i32, [P, c_i64, c_i64, P, C.POINTER(c_i64)]),
    "rnla_tsog1": (c_i32, [P, c_i64, c_i64, c_i64, c_i32, c_i32, P]),
    "rnla_rf1": (c_i32, [P, c_i64, c_i64, c_i64, P, C.POINTER(c_i64)]),
    "rnla_qb1": (c_i32, [P, c_i64, c_i64, c_i64, c_f64, P, P, C.POINTER(c_i64)]),
    "rnla_rand_svd": (c_i32, [P, c_i64, c_i64, c_i64, c_f64, c_i64, P, P, P, C.POINTER(c_i64)]),
    "rnla_rand_evd1": (c_i32, [P, c_i64, c_i64, c_f64, c_i64, P, P, C.POINTER(c_i64)]),
    "rnla_rand_evd2": (c_i32, [P, c_i64, c_i64, c_i64, P, P, C.POINTER(c_i64)]),
    "rnla_rand_svd_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i64, c_i64, C.POINTER(Options), P, c_i64, P, P, c_i64, C.POINTER(c_i64)]),
    "rnla_rand_evd1_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i64, C.POINTER(Options), P, c_i64, P, C.POINTER(c_i64)]),
    "rnla_rand_evd2_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i64, C.POINTER(Options), P, c_i64, P, C.POINTER(c_i64)]),
    "rnla_sketch_dim": (c_i64, [c_i64, c_i64, c_f64, c_i32]),
    "rnla_sketch_apply": (c_i32, [c_i32, c_i32, c_u64, c_i64, c_i32, P, c_i64, c_i64, P, c_i64, P, P]),
    "rnla_sketch_apply_dev": (c_i32, [c_i32, c_i32, c_u64, c_i64, c_i32, P, c_i64, c_i64, c_i64, c_i64, P, c_i64]),
    "rnla_gemm_nn_dev": (c_i32, [P, c_i64, c_i64, c_i64, P, c_i64, c_i64, P, c_i64]),
    "rnla_sketch_gemm_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i32, c_u64, c_u32, c_i64, P, c_i64]),
    "rnla_gemm_tn_dev": (c_i32, [P, c_i64, c_i64, c_i64, P, c_i64, c_i64, P, c_i64, c_i32]),
    "rnla_orth_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i32, P, C.POINTER(c_i64)]),
    "rnla_small_svd_dev": (c_i32, [P, c_i64, c_i64, P, P, P]),
    "rnla_last_jacobi_sweeps": (c_i32, []),
    "rnla_lsrn_overdetermined": (c_i32, [P, c_i64, c_i64, P, C.c_double, c_i64, C.c_double, c_i32, c_i32, c_i32, P, P, P]),
    "rnla_lsrn_overdetermined_dev": (c_i32, [P, c_i64, c_i64, c_i64, P, C.c_double, c_i64, C.c_double, c_i32, c_i32, c_i32, P, P, P]),
    "rnla_cgls": (c_i32, [P, c_i64, c_i64, P, c_f64, c_i64, P, P, C.POINTER(c_i64), C.POINTER(c_i32)]),
    "rnla_cgls_dev": (c_i32, [P, c_i64, c_i64, c_i64, P, c_f64, c_i64, P, C.POINTER(c_i64), C.POINTER(c_i32)]),
    "rnla_conjugate_grad": (c_i32, [P, c_i64, P, P, P, C.POINTER(c_i64), C.POINTER(c_i32)]),
    "rnla_conjugate_grad_dev": (c_i32, [P, c_i64, c_i64, P, P, C.POINTER(c_i64), C.POINTER(c_i32)]),
    "rnla_verify_solution": (c_i32, [P, c_i64, c_i64, P, P, C.POINTER(c_f64)]),
    "rnla_lsqr": (c_i32, [P, c_i64, c_i64, P, c_f64, c_f64, c_f64, c_f64, c_i64, c_i32, P, P, C.POINTER(LsqrResult), P, c_i64, P]),
    "rnla_lsqr_dev": (c_i32, [P, c_i64, c_i64, c_i64, P, c_f64, c_f64, c_f64, c_f64, c_i64, c_i32, P, P, C.POINTER(LsqrResult), P,
                      c_i64, P]),
    "rnla_plan_gemm": (None, [c_i64, c_i64, c_i64, c_i32, P]),
    "rnla_plan_saso_block": (c_i32, [c_i64, c_i32, c_i32, c_i64, c_i64, c_i32, P, P, c_i32, P]),
    "rnla_gemv_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i32, P, P]),
    "rnla_normal_pass_supported": (c_i32, [P, c_i64, c_i64, c_i64]),
    "rnla_plan_normal_pass": (c_i32, [c_u64, c_i64, c_i64, P]),
    "rnla_normal_pass_dev": (c_i32, [P, c_i64, c_i64, c_i64, P, c_f64, P, c_f64, P, P]),
    "rnla_blendenpik_overdetermined": (c_i32, [P, c_i64, c_i64, P, C.c_double, c_i64, C.c_double, c_i32, c_i32, c_i32, P, P, P]),
    "rnla_blendenpik_overdetermined_dev": (c_i32, [P, c_i64, c_i64, c_i64, P, C.c_double, c_i64, C.c_double, c_i32, c_i32, c_i32, P, P, P]),
    "rnla_qrcp": (c_i32, [P, c_i64, c_i64, c_i64, c_i64, P, P, P]),
    "rnla_qrcp_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i64, P, P, c_i64, c_i64]),
    "rnla_lupp": (c_i32, [P, c_i64, c_i64, P, P, P]),
    "rnla_lupp_dev": (c_i32, [P, c_i64, c_i64, P, c_i64, P, c_i64, P]),
    "rnla_sap_chol_qrcp": (c_i32, [P, c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, P, P, P, C.POINTER(c_i64)]),
    "rnla_sap_chol_qrcp_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, P, c_i64, P, c_i64, P, C.POINTER(c_i64)]),
    "rnla_sketched_least_squares_qr": (c_i32, [P, c_i64, c_i64, P, c_i32, c_i32, c_i32, P]),
    "rnla_sketched_least_squares_svd": (c_i32, [P, c_i64, c_i64, P, c_i32, c_i32, c_i32, P]),
    "rnla_sketched_least_squares_dev": (c_i32, [c_i32, P, c_i64, c_i64, c_i64, P, c_i32, c_i32, c_i32, P]),
    "rnla_osid_qrcp": (c_i32, [P, c_i64, c_i64, c_i64, c_i32, P, P]),
    "rnla_osid_randomised": (c_i32, [P, c_i64, c_i64, c_i64, c_i32, P, P]),
    "rnla_osid_randomised_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i64, c_i32, C.POINTER(Options), P, c_i64, P]),
    "rnla_two_sided_id": (c_i32, [P, c_i64, c_i64, c_i64, c_i32, P, P, P, P]),
    "rnla_cur": (c_i32, [P, c_i64, c_i64, c_i64, c_i32, P, P, P]),
    "rnla_cur_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i64, c_i32, C.POINTER(Options), P, P, c_i64, P]),
    "rnla_sketch_saddle_point_precondition": (c_i32, [P, c_i64, c_i64, P, P, c_f64, c_f64, c_i64, c_f64, P, P, P, P]),
    "rnla_sketch_saddle_point_precondition_dev": (c_i32, [P, c_i64, c_i64, c_i64, P, P, c_f64, c_f64, c_i64, c_f64, P, P, P, P]),
    "rnla_i8_gemm_dev": (c_i32, [c_i32, c_i32, c_i32, P, c_i64, c_i64, c_i64, P, c_i64, c_i64, P, c_i64, c_i32]),
    "rnla_debug_i8_flush": (c_i32, [c_i32]),
    "rnla_i8_range_gemm_dev": (c_i32, [c_i32, P, c_i64, c_i64, c_i64, P, c_i64, c_i64, P, c_i64, c_i32]),
    "rnla_small_eigh_dev": (c_i32, [P, c_i64, c_i64, P, P]),
    "rnla_generate_lowrank_dev": (c_i32, [P, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, P, c_f64, c_u64]),
    "rnla_measure_int8_roof": (c_i32, [C.POINTER(c_f64), C.POINTER(c_f64)]),
    "rnla_measure_roofs": (c_i32, [C.POINTER(c_f64), C.POINTER(c_f64), C.c_size_t]),
    "rnla_malloc": (c_i32, [C.POINTER(P), C.c_size_t]),
    "rnla_free": (c_i32, [P]),
    "rnla_memcpy_h2d": (c_i32, [P, P, C.c_size_t]),
    "rnla_memcpy_d2h": (c_i32, [P, P, C.c_size_t]),
}

_lib = None


def load():
    """Load librnla.so (once) and attach the prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(randnla_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    msg = load().rnla_last_error_message()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status):
    """Translate an rnla_status into the matching RandNLAError subclass."""
    if status == 0:
        return
    raise STATUS_TO_ERROR.get(int(status), ComputationError)(last_error())
