//! Drop-ins for reference src/sketch_and_solve.rs (:24-33, :54-66) on the GPU.
use crate::errors::from_status;
use crate::ffi;
use nalgebra::DMatrix;

pub fn sketched_least_squares_qr(a: &DMatrix<f64>, b: &DMatrix<f64>) -> DMatrix<f64> {
    let (m, n) = a.shape();
    let mut x = DMatrix::<f64>::zeros(n, 1);
    from_status(unsafe { ffi::rnla_sketched_least_squares_qr(a.as_ptr(), m as i64, n as i64, b.as_ptr(), 0, 0, 0, x.as_mut_ptr()) })
        .unwrap_or_else(|e| panic!("{}", e));
    x
}

pub fn sketched_least_squares_svd(a: &DMatrix<f64>, b: &DMatrix<f64>) -> DMatrix<f64> {
    let (m, n) = a.shape();
    let mut x = DMatrix::<f64>::zeros(n, 1);
    from_status(unsafe { ffi::rnla_sketched_least_squares_svd(a.as_ptr(), m as i64, n as i64, b.as_ptr(), 0, 0, 0, x.as_mut_ptr()) })
        .unwrap_or_else(|e| panic!("{}", e));                    // "Panics when SVD fails" (reference doc comment)
    x
}
