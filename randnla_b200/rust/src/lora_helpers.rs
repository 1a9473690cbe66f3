//! reference src/lora_helpers.rs: infallible signatures (the reference unwraps); a device failure panics.
use crate::errors::from_status;
use crate::ffi;
use nalgebra::DMatrix;

fn ok(code: i32) { if let Err(e) = from_status(code) { panic!("{}", e); } }

pub fn QB1(A: &DMatrix<f64>, k: usize, epsilon: f64) -> (DMatrix<f64>, DMatrix<f64>) {
    let (m, n) = A.shape();
    let l = k.min(m).min(n);
    let mut Q = DMatrix::<f64>::zeros(m, l);
    let mut B = DMatrix::<f64>::zeros(l, n);
    let mut qc: i64 = 0;
    ok(unsafe { ffi::rnla_qb1(A.as_ptr(), m as i64, n as i64, k as i64, epsilon, Q.as_mut_ptr(), B.as_mut_ptr(), &mut qc) });
    (Q, B)
}

pub fn RF1(A: &DMatrix<f64>, k: usize) -> DMatrix<f64> {
    let (m, n) = A.shape();
    let mut Q = DMatrix::<f64>::zeros(m, k.min(m).min(n));
    let mut qc: i64 = 0;
    ok(unsafe { ffi::rnla_rf1(A.as_ptr(), m as i64, n as i64, k as i64, Q.as_mut_ptr(), &mut qc) });
    Q
}

pub fn tsog1(A: &DMatrix<f64>, k: usize, num_passes: i32, passes_per_stab: i32) -> DMatrix<f64> {
    let (m, n) = A.shape();
    let mut S = DMatrix::<f64>::zeros(n, k);
    ok(unsafe { ffi::rnla_tsog1(A.as_ptr(), m as i64, n as i64, k as i64, num_passes, passes_per_stab, S.as_mut_ptr()) });
    S
}

pub fn Orth(X: &DMatrix<f64>) -> DMatrix<f64> {
    let (r, c) = X.shape();
    let mut Q = DMatrix::<f64>::zeros(r, r.min(c));
    let mut qc: i64 = 0;
    ok(unsafe { ffi::rnla_orth(X.as_ptr(), r as i64, c as i64, Q.as_mut_ptr(), std::ptr::null_mut(), &mut qc) });
    Q
}

pub fn Stabilizer(X: &DMatrix<f64>) -> DMatrix<f64> {
    let (r, c) = X.shape();
    let mut L = DMatrix::<f64>::zeros(r, r.min(c));
    let mut lc: i64 = 0;
    ok(unsafe { ffi::rnla_stabilizer(X.as_ptr(), r as i64, c as i64, L.as_mut_ptr(), &mut lc) });
    L
}
