//! Drop-in for `sap_chol_qrcp` of reference src/cqrrpt.rs (:27-58) on the GPU.
use crate::errors::from_status;
use crate::ffi;
use nalgebra::DMatrix;

pub fn sap_chol_qrcp(a: &DMatrix<f64>, d: usize) -> (DMatrix<f64>, DMatrix<f64>, Vec<usize>) {
    let (m, n) = a.shape();
    assert!(n <= d && d <= m, "d must satisfy n ≤ d ≪ m");
    let mut q = DMatrix::<f64>::zeros(m, n);
    let mut r = vec![0.0f64; n * n];
    let mut j = vec![0i64; n.max(1)];
    let mut k = 0i64;
    from_status(unsafe { ffi::rnla_sap_chol_qrcp(a.as_ptr(), m as i64, n as i64, d as i64, 0, 0, 0, q.as_mut_ptr(), r.as_mut_ptr(), j.as_mut_ptr(), &mut k) })
        .unwrap_or_else(|e| panic!("{}", e));                    // the reference `expect`s / unwraps (:47, :51)
    let k = k as usize;
    (q.columns(0, k).into_owned(), DMatrix::from_column_slice(k, n, &r[..k * n]), j[..n].iter().map(|&v| v as usize).collect())
}
