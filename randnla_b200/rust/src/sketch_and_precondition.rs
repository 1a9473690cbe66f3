//! Sketch step of reference src/sketch_and_precondition.rs (:26-52, :82-107, :150-176) on the GPU.
//! A maintainer replaces the three lines `let s = sketching_operator(..); let a_sk = &s*a; let b_sk = &s*b;`
//! of each solver with a call to `sketch_step`; the QR/SVD preconditioner and CGLS that follow stay as they are.
use crate::errors::{from_status, RandNLAError};
use crate::ffi;
use nalgebra::DMatrix;
use std::error::Error;

pub enum SketchKind { Dense, SparseSign { zeta: i32 } }

/// d of :49 / :105 (`saddle = false`) or :172 (`saddle = true`)
pub fn sketch_dim(m: usize, n: usize, sampling_factor: f64, saddle: bool) -> usize {
    unsafe { ffi::rnla_sketch_dim(m as i64, n as i64, sampling_factor, if saddle { 1 } else { 0 }) as usize }
}

pub fn validate(a: &DMatrix<f64>, epsilon: f64, l: usize, sampling_factor: f64) -> Result<(), Box<dyn Error>> {
    if a.nrows() < a.ncols() {
        return Err(Box::new(RandNLAError::NotOverdetermined(format!("Need more columns than rows, found {} rows and {} columns", a.nrows(), a.ncols()))));
    }
    if sampling_factor < 1.0 { return Err(Box::new(RandNLAError::InvalidParameters(format!("Sampling factor must be greater than 1, current input is {}", sampling_factor)))); }
    if epsilon <= 0.0 { return Err(Box::new(RandNLAError::InvalidParameters(format!("Epsilon must be positive, current input is {}", epsilon)))); }
    if l == 0 { return Err(Box::new(RandNLAError::InvalidParameters(format!("Number of iterations must be positive, current input is {}", l)))); }
    Ok(())
}

/// (S a, S b) for a d x m sketching operator S
pub fn sketch_step(a: &DMatrix<f64>, b: Option<&DMatrix<f64>>, d: usize, kind: SketchKind) -> Result<(DMatrix<f64>, Option<DMatrix<f64>>), Box<dyn Error>> {
    let (m, n) = a.shape();
    let (k, zeta) = match kind { SketchKind::Dense => (0, 0), SketchKind::SparseSign { zeta } => (1, zeta) };
    let mut a_sk = DMatrix::<f64>::zeros(d, n);
    let mut b_sk = b.map(|bb| DMatrix::<f64>::zeros(d, bb.ncols()));
    let (bp, nrhs, bskp) = match (b, b_sk.as_mut()) {
        (Some(bb), Some(o)) => (bb.as_ptr(), bb.ncols() as i64, o.as_mut_ptr()),
        _ => (std::ptr::null(), 0, std::ptr::null_mut()),
    };
    from_status(unsafe { ffi::rnla_sketch_apply(k, 0, 0, d as i64, zeta, a.as_ptr(), m as i64, n as i64, bp, nrhs, a_sk.as_mut_ptr(), bskp) })?;
    Ok((a_sk, b_sk))
}
