//! Drop-in for `lsqr` of reference src/solvers.rs:115-278 (its translation of scipy 1.14.1 sparse.linalg.lsqr): same signature,
//! same 10-tuple.  The Golub-Kahan bidiagonalisation runs on the GPU (A streamed twice per iteration), the scalar
//! recurrences in librnla on the host.
use crate::errors::from_status;
use crate::ffi;
use nalgebra::{DMatrix, DVector};

#[allow(clippy::too_many_arguments, clippy::type_complexity)]
pub fn lsqr(
    a: &DMatrix<f64>,
    b: &DVector<f64>,
    damp: f64,
    atol: f64,
    btol: f64,
    conlim: f64,
    iter_lim: Option<usize>,
    calc_var: bool,
    x0: Option<&DVector<f64>>,
) -> (DVector<f64>, usize, usize, f64, f64, f64, f64, Vec<f64>, f64, DVector<f64>) {
    let (m, n) = a.shape();
    assert_eq!(b.len(), m, "lsqr: b must have as many entries as a has rows");          // nalgebra panics at :195
    if let Some(v) = x0 { assert_eq!(v.len(), n, "lsqr: x0 must have as many entries as a has columns"); }
    let cap = iter_lim.unwrap_or(2 * n).max(1);
    let mut x = DVector::<f64>::zeros(n);
    let mut var = DVector::<f64>::zeros(n);
    let mut arnorms = vec![0.0f64; cap];
    let mut res = ffi::RnlaLsqrResult::default();
    from_status(unsafe {
        ffi::rnla_lsqr(a.as_ptr(), m as i64, n as i64, b.as_ptr(), damp, atol, btol, conlim, iter_lim.map_or(-1, |v| v as i64),
                       calc_var as i32, x0.map_or(std::ptr::null(), |v| v.as_ptr()), x.as_mut_ptr(), &mut res, arnorms.as_mut_ptr(),
                       cap as i64, var.as_mut_ptr())
    })
    .unwrap_or_else(|e| panic!("{}", e));
    arnorms.truncate((res.n_arnorms as usize).min(cap));
    (x, res.istop as usize, res.itn as usize, res.r1norm, res.r2norm, res.anorm, res.acond, arnorms, res.xnorm, var)
}
