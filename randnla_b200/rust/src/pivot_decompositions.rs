//! Drop-ins for `qrcp` and `economic_qrcp` of reference src/pivot_decompositions.rs (:105-180, :196-269) on the GPU.
//! (`lupp` :21-86 is not on any sketch path and stays as it is.)
use crate::errors::from_status;
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
