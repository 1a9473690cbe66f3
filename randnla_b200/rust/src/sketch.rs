//! reference src/sketch.rs: same enums and signatures.
use crate::errors::from_status;
use crate::ffi;
use nalgebra::DMatrix;
use std::error::Error;

pub enum DistributionType { Gaussian, Uniform, Rademacher }
pub enum MatrixAttribute { Row, Column }

pub fn sketching_operator(dist_type: DistributionType, rows: usize, cols: usize) -> Result<DMatrix<f64>, Box<dyn Error>> {
    let d = match dist_type { DistributionType::Gaussian => 0, DistributionType::Uniform => 1, DistributionType::Rademacher => 2 };
    let mut out = DMatrix::<f64>::zeros(rows, cols);
    from_status(unsafe { ffi::rnla_sketching_operator(d, rows as i64, cols as i64, out.as_mut_ptr()) })?;
    Ok(out)
}

pub fn haar_sample(rows: usize, columns: usize, attr: MatrixAttribute) -> Result<DMatrix<f64>, Box<dyn Error>> {
    let a = match attr { MatrixAttribute::Row => 0, MatrixAttribute::Column => 1 };
    let mut out = DMatrix::<f64>::zeros(rows, columns);
    from_status(unsafe { ffi::rnla_haar_sample(rows as i64, columns as i64, a, out.as_mut_ptr()) })?;
    Ok(out)
}
