//! Drop-ins for `lupp`, `qrcp` and `economic_qrcp` of reference src/pivot_decompositions.rs (:21-86, :105-180, :196-269) on
//! the GPU.
use crate::errors::from_status;
use std::error::Error;
use crate::ffi;
use nalgebra::DMatrix;

fn run(a: &DMatrix<f64>, steps: usize, qcols: usize) -> (DMatrix<f64>, DMatrix<f64>, Vec<usize>) {
    let (m, n) = a.shape();
    let mut q = DMatrix::<f64>::zeros(m, qcols);
    let mut r = DMatrix::<f64>::zeros(m, n);
    let mut p = vec![0i64; n.max(1)];
    // the reference asserts / panics on bad input; so does the shim
    from_status(unsafe { ffi::rnla_qrcp(a.as_ptr(), m as i64, n as i64, steps as i64, qcols as i64, q.as_mut_ptr(), r.as_mut_ptr(), p.as_mut_ptr()) })
        .unwrap_or_else(|e| panic!("{}", e));
    (q, r, p[..n].iter().map(|&v| v as usize).collect())
}

pub fn qrcp(a: &DMatrix<f64>) -> (DMatrix<f64>, DMatrix<f64>, Vec<usize>) {
    let (m, n) = a.shape();
    run(a, m.min(n), m)
}

pub fn economic_qrcp(a: &DMatrix<f64>, k: usize) -> (DMatrix<f64>, DMatrix<f64>, Vec<usize>) {
    let (m, n) = a.shape();
    assert!(k <= m.min(n), "k must be <= min(m,n)");
    assert!(k > 0, "k must be positive");
    let (q, r, p) = run(a, k, k);
    (q, r.rows(0, k).into_owned(), p)
}

/// `lupp` (reference :21-86): same errors, and the same bits in l, u, p (first-maximum pivot, the reference's operation order).
pub fn lupp(matrix: &DMatrix<f64>) -> Result<(DMatrix<f64>, DMatrix<f64>, Vec<usize>), Box<dyn Error>> {
    let (rows, cols) = matrix.shape();
    let n = rows.max(1);
    let mut l = DMatrix::<f64>::zeros(n, n);
    let mut u = DMatrix::<f64>::zeros(n, n);
    let mut p = vec![0i64; n];
    from_status(unsafe { ffi::rnla_lupp(matrix.as_ptr(), rows as i64, cols as i64, l.as_mut_ptr(), u.as_mut_ptr(), p.as_mut_ptr()) })?;
    Ok((l, u, p.iter().map(|&v| v as usize).collect()))
}
