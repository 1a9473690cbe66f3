//! Drop-ins for reference src/cg.rs: `cgls` (:18-61), `conjugate_grad` (:77-112), `verify_solution` (:115-117) -- same
//! signatures, same `println!`s; the iterations run on the GPU (a streamed once or twice per iteration).
use crate::errors::{from_status, RandNLAError};
use crate::ffi;
use nalgebra::{DMatrix, DVector};

pub fn cgls(a: &DMatrix<f64>, b: &DMatrix<f64>, tolerance: f64, num_iterations: usize, x: Option<DMatrix<f64>>) -> DMatrix<f64> {
    let (m, n) = a.shape();
    assert_eq!(b.nrows(), m, "cgls: b must have as many rows as a");                    // nalgebra panics at :30
    if let Some(v) = &x { assert_eq!(v.nrows(), n, "cgls: x must have as many rows as a has columns"); }
    let mut out = DMatrix::<f64>::zeros(n, 1);
    let (mut iters, mut converged) = (0i64, 0i32);
    from_status(unsafe {
        ffi::rnla_cgls(a.as_ptr(), m as i64, n as i64, b.as_ptr(), tolerance, num_iterations as i64,
                       x.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()), out.as_mut_ptr(), &mut iters, &mut converged)
    })
    .unwrap_or_else(|e| panic!("{}", e));
    if converged != 0 { println!("CGLS converged after {} iterations", iters); }
    else { println!("CGLS failed to converged after {} iterations", num_iterations); }
    out
}

pub fn conjugate_grad(a: &DMatrix<f64>, b: &DVector<f64>, x: Option<DVector<f64>>) -> Result<DVector<f64>, RandNLAError> {
    let n = b.len();
    assert!(a.nrows() == n && a.ncols() == n, "conjugate_grad: a must be n x n");
    let mut out = DVector::<f64>::zeros(n);
    let (mut iters, mut converged) = (0i64, 0i32);
    from_status(unsafe {
        ffi::rnla_conjugate_grad(a.as_ptr(), n as i64, b.as_ptr(), x.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()),
                                 out.as_mut_ptr(), &mut iters, &mut converged)
    })?;
    if converged != 0 { println!("Converged after {} iterations", iters); }
    Ok(out)
}

pub fn verify_solution(a: &DMatrix<f64>, b: &DVector<f64>, x: &DVector<f64>) -> f64 {
    let (m, n) = a.shape();
    let mut r = 0.0f64;
    from_status(unsafe { ffi::rnla_verify_solution(a.as_ptr(), m as i64, n as i64, b.as_ptr(), x.as_ptr(), &mut r) })
        .unwrap_or_else(|e| panic!("{}", e));
    r
}
