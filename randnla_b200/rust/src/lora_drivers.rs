//! reference src/lora_drivers.rs: same signatures, same validation order and messages, same stdout side effects.
use crate::errors::{from_status, RandNLAError};
use crate::ffi;
use nalgebra::DMatrix;

pub fn rand_svd(A: &DMatrix<f64>, k: usize, epsilon: f64, s: usize) -> Result<(DMatrix<f64>, DMatrix<f64>, DMatrix<f64>), RandNLAError> {
    if k == 0 { return Err(RandNLAError::InvalidParameters(format!("Rank k must be positive, current input is {}", k))); }
    if epsilon <= 0.0 { return Err(RandNLAError::InvalidParameters(format!("Epsilon must be positive, current input is {}", epsilon))); }
    if s == 0 { return Err(RandNLAError::InvalidParameters(format!("Oversampling parameter s must be positive, current input is {}", s))); }
    println!("Running RSVD");
    let (m, n) = A.shape();
    let r = k.min((k + s).min(m).min(n));
    let mut U = DMatrix::<f64>::zeros(m, r);
    let mut S = DMatrix::<f64>::zeros(r, r);
    let mut Vt = DMatrix::<f64>::zeros(r, n);
    let mut rr: i64 = 0;
    from_status(unsafe { ffi::rnla_rand_svd(A.as_ptr(), m as i64, n as i64, k as i64, epsilon, s as i64, U.as_mut_ptr(), S.as_mut_ptr(), Vt.as_mut_ptr(), &mut rr) })?;
    Ok((U, S, Vt))
}

pub fn rand_evd1(A: &DMatrix<f64>, k: usize, epsilon: f64, s: usize) -> Result<(DMatrix<f64>, Vec<f64>), RandNLAError> {
    if k == 0 { return Err(RandNLAError::InvalidParameters(format!("Rank k must be positive, current input is {}", k))); }
    if epsilon <= 0.0 { return Err(RandNLAError::InvalidParameters(format!("Epsilon must be positive, current input is {}", epsilon))); }
    if s == 0 { return Err(RandNLAError::InvalidParameters(format!("Oversampling parameter s must be positive, current input is {}", s))); }
    if A.nrows() != A.ncols() { return Err(RandNLAError::NotHermitian("Input matrix is not Hermitian".to_string())); }
    println!("Running REVD1");
    let n = A.nrows();
    let cap = k.min(n);
    let mut V = DMatrix::<f64>::zeros(n, cap);
    let mut lambda = vec![0.0f64; cap];
    let mut r: i64 = 0;
    from_status(unsafe { ffi::rnla_rand_evd1(A.as_ptr(), n as i64, k as i64, epsilon, s as i64, V.as_mut_ptr(), lambda.as_mut_ptr(), &mut r) })?;
    lambda.truncate(r as usize);
    Ok((V.columns(0, r as usize).into_owned(), lambda))
}

pub fn rand_evd2(A: &DMatrix<f64>, k: usize, s: usize) -> Result<(DMatrix<f64>, Vec<f64>), RandNLAError> {
    if k == 0 { return Err(RandNLAError::InvalidParameters(format!("Rank k must be positive, current input is {}", k))); }
    println!("Running REVD2");
    let n = A.nrows();
    let cap = k.min(n);
    let mut V = DMatrix::<f64>::zeros(n, cap);
    let mut lambda = vec![0.0f64; cap];
    let mut r: i64 = 0;
    from_status(unsafe { ffi::rnla_rand_evd2(A.as_ptr(), n as i64, k as i64, s as i64, V.as_mut_ptr(), lambda.as_mut_ptr(), &mut r) })?;
    lambda.truncate(r as usize);
    Ok((V.columns(0, r as usize).into_owned(), lambda))
}
