//! Drop-ins for reference src/id.rs on the GPU: cur (:34-71), two_sided_id_randomised (:94-101), two_sided_id (:118-129),
//! cur_randomised (:154-193), osid_randomised (:217-249), osid_qrcp (:272-318).  Invalid k panics, as in the reference.
use crate::errors::from_status;
use crate::ffi;
use crate::sketch::MatrixAttribute;
use nalgebra::DMatrix;

fn idx(v: Vec<i64>, k: usize) -> Vec<usize> { v[..k].iter().map(|&t| t as usize).collect() }
fn attr_code(attr: &MatrixAttribute) -> i32 { match attr { MatrixAttribute::Row => 0, MatrixAttribute::Column => 1 } }

pub fn osid_qrcp(y: &DMatrix<f64>, k: usize, attr: MatrixAttribute) -> (DMatrix<f64>, Vec<usize>) {
    let (l, w) = y.shape();
    assert!(k > 0, "k must be positive)");
    assert!(k <= l.min(w), "k must be <= min(l,w)");
    let mut x = match attr { MatrixAttribute::Column => DMatrix::<f64>::zeros(k, w), MatrixAttribute::Row => DMatrix::<f64>::zeros(l, k) };
    let mut j = vec![0i64; k];
    from_status(unsafe { ffi::rnla_osid_qrcp(y.as_ptr(), l as i64, w as i64, k as i64, attr_code(&attr), x.as_mut_ptr(), j.as_mut_ptr()) })
        .unwrap_or_else(|e| panic!("{}", e));
    (x, idx(j, k))
}

pub fn osid_randomised(a: &DMatrix<f64>, k: usize, attr: MatrixAttribute) -> (DMatrix<f64>, Vec<usize>) {
    let (m, n) = a.shape();
    assert!(k > 0, "k must be positive)");
    assert!(k <= m.min(n), "k must be <= min(l,w)");
    let mut x = match attr { MatrixAttribute::Column => DMatrix::<f64>::zeros(k, n), MatrixAttribute::Row => DMatrix::<f64>::zeros(m, k) };
    let mut j = vec![0i64; k];
    from_status(unsafe { ffi::rnla_osid_randomised(a.as_ptr(), m as i64, n as i64, k as i64, attr_code(&attr), x.as_mut_ptr(), j.as_mut_ptr()) })
        .unwrap_or_else(|e| panic!("{}", e));
    (x, idx(j, k))
}

fn two_sided(a: &DMatrix<f64>, k: usize, randomised: i32) -> (DMatrix<f64>, Vec<usize>, Vec<usize>, DMatrix<f64>) {
    let (m, n) = a.shape();
    assert!(k > 0, "k must be positive)");
    assert!(k <= m.min(n), "k must be <= min(l,w)");
    let (mut z, mut x) = (DMatrix::<f64>::zeros(m, k), DMatrix::<f64>::zeros(k, n));
    let (mut i, mut j) = (vec![0i64; k], vec![0i64; k]);
    from_status(unsafe { ffi::rnla_two_sided_id(a.as_ptr(), m as i64, n as i64, k as i64, randomised, z.as_mut_ptr(), i.as_mut_ptr(), j.as_mut_ptr(), x.as_mut_ptr()) })
        .unwrap_or_else(|e| panic!("{}", e));
    (z, idx(i, k), idx(j, k), x)
}
pub fn two_sided_id(a: &DMatrix<f64>, k: usize) -> (DMatrix<f64>, Vec<usize>, Vec<usize>, DMatrix<f64>) { two_sided(a, k, 0) }
pub fn two_sided_id_randomised(a: &DMatrix<f64>, k: usize) -> (DMatrix<f64>, Vec<usize>, Vec<usize>, DMatrix<f64>) { two_sided(a, k, 1) }

fn cur_impl(a: &DMatrix<f64>, k: usize, randomised: i32) -> (Vec<usize>, DMatrix<f64>, Vec<usize>) {
    let (m, n) = a.shape();
    assert!(k > 0, "k must be positive)");
    assert!(k <= m.min(n), "k must be <= min(l,w)");
    let mut u = DMatrix::<f64>::zeros(k, k);
    let (mut i, mut j) = (vec![0i64; k], vec![0i64; k]);
    from_status(unsafe { ffi::rnla_cur(a.as_ptr(), m as i64, n as i64, k as i64, randomised, j.as_mut_ptr(), u.as_mut_ptr(), i.as_mut_ptr()) })
        .unwrap_or_else(|e| panic!("{}", e));
    (idx(j, k), u, idx(i, k))
}
pub fn cur(a: &DMatrix<f64>, k: usize) -> (Vec<usize>, DMatrix<f64>, Vec<usize>) { cur_impl(a, k, 0) }
pub fn cur_randomised(a: &DMatrix<f64>, k: usize) -> (Vec<usize>, DMatrix<f64>, Vec<usize>) { cur_impl(a, k, 1) }
