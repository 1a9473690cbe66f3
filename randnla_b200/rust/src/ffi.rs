//! Raw bindings of include/rnla.h (only what the reference-signature functions need).
use std::os::raw::{c_char, c_double, c_int};

extern "C" {
    pub fn rnla_last_error_message() -> *const c_char;
    pub fn rnla_sketching_operator(dist: c_int, rows: i64, cols: i64, out: *mut c_double) -> c_int;
    pub fn rnla_haar_sample(rows: i64, cols: i64, attr: c_int, out: *mut c_double) -> c_int;
    pub fn rnla_orth(x: *const c_double, rows: i64, cols: i64, q: *mut c_double, r: *mut c_double, qcols: *mut i64) -> c_int;
    pub fn rnla_stabilizer(x: *const c_double, rows: i64, cols: i64, l: *mut c_double, lcols: *mut i64) -> c_int;
    pub fn rnla_tsog1(a: *const c_double, m: i64, n: i64, k: i64, num_passes: c_int, passes_per_stab: c_int, s: *mut c_double) -> c_int;
    pub fn rnla_rf1(a: *const c_double, m: i64, n: i64, k: i64, q: *mut c_double, qcols: *mut i64) -> c_int;
    pub fn rnla_qb1(a: *const c_double, m: i64, n: i64, k: i64, epsilon: c_double, q: *mut c_double, b: *mut c_double, qcols: *mut i64) -> c_int;
    pub fn rnla_rand_svd(a: *const c_double, m: i64, n: i64, k: i64, epsilon: c_double, s: i64,
                         u: *mut c_double, sig: *mut c_double, vt: *mut c_double, r: *mut i64) -> c_int;
    pub fn rnla_rand_evd1(a: *const c_double, n: i64, k: i64, epsilon: c_double, s: i64, v: *mut c_double, lambda: *mut c_double, r: *mut i64) -> c_int;
    pub fn rnla_rand_evd2(a: *const c_double, n: i64, k: i64, s: i64, v: *mut c_double, lambda: *mut c_double, r: *mut i64) -> c_int;
    pub fn rnla_sketch_dim(m: i64, n: i64, sampling_factor: c_double, rule: c_int) -> i64;
    pub fn rnla_sketch_apply(kind: c_int, dist: c_int, seed: u64, d: i64, zeta: c_int, a: *const c_double, m: i64, n: i64,
                             b: *const c_double, nrhs: i64, a_sk: *mut c_double, b_sk: *mut c_double) -> c_int;
    pub fn rnla_blendenpik_overdetermined(a: *const c_double, m: i64, n: i64, b: *const c_double, epsilon: c_double, l: i64,
                                          sampling_factor: c_double, kind: c_int, dist: c_int, zeta: c_int, x: *mut c_double,
                                          iterations: *mut i64, converged: *mut c_int) -> c_int;
    pub fn rnla_lsrn_overdetermined(a: *const c_double, m: i64, n: i64, b: *const c_double, epsilon: c_double, l: i64,
                                    sampling_factor: c_double, kind: c_int, dist: c_int, zeta: c_int, x: *mut c_double,
                                    iterations: *mut i64, converged: *mut c_int) -> c_int;
    pub fn rnla_sketch_saddle_point_precondition(a: *const c_double, m: i64, n: i64, b: *const c_double, c: *const c_double, mu: c_double,
                                                 epsilon: c_double, l: i64, sampling_factor: c_double, x: *mut c_double, y: *mut c_double,
                                                 iterations: *mut i64, converged: *mut c_int) -> c_int;
    // rows after the hot path: pivoted QR, CQRRPT, sketch-and-solve, ID / CUR
    pub fn rnla_qrcp(a: *const c_double, m: i64, n: i64, steps: i64, qcols: i64, q: *mut c_double, r: *mut c_double, perm: *mut i64) -> c_int;
    pub fn rnla_sap_chol_qrcp(a: *const c_double, m: i64, n: i64, d: i64, kind: c_int, dist: c_int, zeta: c_int,
                              q: *mut c_double, r: *mut c_double, j: *mut i64, k: *mut i64) -> c_int;
    pub fn rnla_sketched_least_squares_qr(a: *const c_double, m: i64, n: i64, b: *const c_double, kind: c_int, dist: c_int, zeta: c_int, x: *mut c_double) -> c_int;
    pub fn rnla_sketched_least_squares_svd(a: *const c_double, m: i64, n: i64, b: *const c_double, kind: c_int, dist: c_int, zeta: c_int, x: *mut c_double) -> c_int;
    pub fn rnla_osid_qrcp(y: *const c_double, l: i64, w: i64, k: i64, attr: c_int, x: *mut c_double, j: *mut i64) -> c_int;
    pub fn rnla_osid_randomised(a: *const c_double, m: i64, n: i64, k: i64, attr: c_int, x: *mut c_double, j: *mut i64) -> c_int;
    pub fn rnla_two_sided_id(a: *const c_double, m: i64, n: i64, k: i64, randomised: c_int, z: *mut c_double, i: *mut i64, j: *mut i64, x: *mut c_double) -> c_int;
    pub fn rnla_cur(a: *const c_double, m: i64, n: i64, k: i64, randomised: c_int, j: *mut i64, u: *mut c_double, i: *mut i64) -> c_int;
    pub fn rnla_lupp(a: *const c_double, rows: i64, cols: i64, l: *mut c_double, u: *mut c_double, perm: *mut i64) -> c_int;
    // reference src/cg.rs
    pub fn rnla_cgls(a: *const c_double, m: i64, n: i64, b: *const c_double, tolerance: c_double, num_iterations: i64,
                     x0: *const c_double, x: *mut c_double, iterations: *mut i64, converged: *mut c_int) -> c_int;
    pub fn rnla_conjugate_grad(a: *const c_double, n: i64, b: *const c_double, x0: *const c_double, x: *mut c_double,
                               iterations: *mut i64, converged: *mut c_int) -> c_int;
    pub fn rnla_verify_solution(a: *const c_double, m: i64, n: i64, b: *const c_double, x: *const c_double,
                                residual_norm: *mut c_double) -> c_int;
    // lsqr (reference src/solvers.rs:115-278)
    pub fn rnla_lsqr(a: *const c_double, m: i64, n: i64, b: *const c_double, damp: c_double, atol: c_double, btol: c_double,
                     conlim: c_double, iter_lim: i64, calc_var: c_int, x0: *const c_double, x: *mut c_double,
                     result: *mut RnlaLsqrResult, arnorms: *mut c_double, arnorms_cap: i64, var: *mut c_double) -> c_int;
}

pub fn last_message() -> String {
    unsafe {
        let p = rnla_last_error_message();
        if p.is_null() { String::new() } else { std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned() }
    }
}

/// `rnla_lsqr_result` (include/rnla.h)
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct RnlaLsqrResult {
    pub istop: i64,
    pub itn: i64,
    pub r1norm: c_double,
    pub r2norm: c_double,
    pub anorm: c_double,
    pub acond: c_double,
    pub xnorm: c_double,
    pub n_arnorms: i64,
}
