//! Sketch step of reference src/sketch_and_precondition.rs (:26-52, :82-107, :150-176) on the GPU.
//! A maintainer replaces the three lines `let s = sketching_operator(..); let a_sk = &s*a; let b_sk = &s*b;`
//! of each solver with a call to `sketch_step`; the QR/SVD preconditioner and CGLS that follow stay as they are.
use crate::errors::{from_status, RandNLAError};
use crate::ffi;
use nalgebra::DMatrix;
use std::error::Error;

/// Dense: i.i.d. Gaussian (what the reference draws).  SparseSign: textbook SASO.  BlockSparseSign: the block
/// sparse-sign operator (include/rnla.h RNLA_SKETCH_SASO_BLOCK), `width` = 0 for the default block width.
pub enum SketchKind { Dense, SparseSign { zeta: i32 }, BlockSparseSign { zeta: i32, width: i32 } }

impl SketchKind {
    fn abi(&self) -> (i32, i32, i32) {      // (kind, dist slot, zeta)
        match *self {
            SketchKind::Dense => (0, 0, 0),
            SketchKind::SparseSign { zeta } => (1, 0, zeta),
            SketchKind::BlockSparseSign { zeta, width } => (2, width, zeta),
        }
    }
}

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
    let (k, dist, zeta) = kind.abi();
    let mut a_sk = DMatrix::<f64>::zeros(d, n);
    let mut b_sk = b.map(|bb| DMatrix::<f64>::zeros(d, bb.ncols()));
    let (bp, nrhs, bskp) = match (b, b_sk.as_mut()) {
        (Some(bb), Some(o)) => (bb.as_ptr(), bb.ncols() as i64, o.as_mut_ptr()),
        _ => (std::ptr::null(), 0, std::ptr::null_mut()),
    };
    from_status(unsafe { ffi::rnla_sketch_apply(k, dist, 0, d as i64, zeta, a.as_ptr(), m as i64, n as i64, bp, nrhs, a_sk.as_mut_ptr(), bskp) })?;
    Ok((a_sk, b_sk))
}

/// Drop-in for `blendenpik_overdetermined` (reference src/sketch_and_precondition.rs:26-59), end to end on the GPU:
/// sketch, QR of the sketch, z0, R^-1, CGLS on A R^-1 in operator form, x = R^-1 z.  Same validation and errors.
pub fn blendenpik_overdetermined(a: &DMatrix<f64>, b: &DMatrix<f64>, epsilon: f64, l: usize, sampling_factor: f64) -> Result<DMatrix<f64>, Box<dyn Error>> {
    blendenpik_overdetermined_with(a, b, epsilon, l, sampling_factor, SketchKind::Dense)
}

pub fn blendenpik_overdetermined_with(a: &DMatrix<f64>, b: &DMatrix<f64>, epsilon: f64, l: usize, sampling_factor: f64, kind: SketchKind) -> Result<DMatrix<f64>, Box<dyn Error>> {
    validate(a, epsilon, l, sampling_factor)?;
    let (m, n) = a.shape();
    let (k, dist, zeta) = kind.abi();
    let mut x = DMatrix::<f64>::zeros(n, 1);
    let (mut iters, mut converged) = (0i64, 0i32);
    from_status(unsafe {
        ffi::rnla_blendenpik_overdetermined(a.as_ptr(), m as i64, n as i64, b.as_ptr(), epsilon, l as i64, sampling_factor,
                                            k, dist, zeta, x.as_mut_ptr(), &mut iters, &mut converged)
    })?;
    // src/cg.rs:46-59 prints the same two lines
    if converged != 0 { println!("CGLS converged after {} iterations", iters); } else { println!("CGLS failed to converged after {} iterations", l); }
    Ok(x)
}

/// Drop-in for `lsrn_overdetermined` (reference src/sketch_and_precondition.rs:82-119), end to end on the GPU
/// (n <= 1024: size of the on-device SVD core).
pub fn lsrn_overdetermined(a: &DMatrix<f64>, b: &DMatrix<f64>, epsilon: f64, l: usize, sampling_factor: f64) -> Result<DMatrix<f64>, Box<dyn Error>> {
    validate(a, epsilon, l, sampling_factor)?;
    let (m, n) = a.shape();
    let mut x = DMatrix::<f64>::zeros(n, 1);
    let (mut iters, mut converged) = (0i64, 0i32);
    from_status(unsafe {
        ffi::rnla_lsrn_overdetermined(a.as_ptr(), m as i64, n as i64, b.as_ptr(), epsilon, l as i64, sampling_factor,
                                      0, 0, 0, x.as_mut_ptr(), &mut iters, &mut converged)
    })?;
    if converged != 0 { println!("CGLS converged after {} iterations", iters); } else { println!("CGLS failed to converged after {} iterations", l); }
    Ok(x)
}

/// Drop-in for `sketch_saddle_point_precondition` (reference src/sketch_and_precondition.rs:150-216), end to end on the GPU
/// (dense Gaussian sketch as in the reference; n <= 1024).  `c` may be empty (`c.is_empty()`, :195).
pub fn sketch_saddle_point_precondition(a: &DMatrix<f64>, b: &DMatrix<f64>, c: &DMatrix<f64>, mu: f64, epsilon: f64, l: usize,
                                        sampling_factor: f64) -> Result<(DMatrix<f64>, DMatrix<f64>), Box<dyn Error>> {
    validate(a, epsilon, l, sampling_factor)?;
    let (m, n) = a.shape();
    let mut x = DMatrix::<f64>::zeros(n, 1);
    let mut y = DMatrix::<f64>::zeros(m, 1);
    let (mut iters, mut converged) = (0i64, 0i32);
    let cp = if c.is_empty() { std::ptr::null() } else { c.as_ptr() };
    from_status(unsafe {
        ffi::rnla_sketch_saddle_point_precondition(a.as_ptr(), m as i64, n as i64, b.as_ptr(), cp, mu, epsilon, l as i64, sampling_factor,
                                                   x.as_mut_ptr(), y.as_mut_ptr(), &mut iters, &mut converged)
    })?;
    if converged != 0 { println!("CGLS converged after {} iterations", iters); } else { println!("CGLS failed to converged after {} iterations", l); }
    Ok((x, y))
}
