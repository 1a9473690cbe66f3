//! reference src/errors.rs, plus the status-code mapping of include/rnla.h.
use std::error::Error;
#[derive(Debug)]
pub enum RandNLAError {
    InvalidParameters(String),
    InvalidDimensions(String),
    NegativeDimensions(String),
    NotOverdetermined(String),
    NotSquare(String),
    SingularMatrix(String),
    MatrixDecompositionError(String),
    NotHermitian(String),
    NotPositiveSemiDefinite(String),
    ComputationError(String),
}

impl std::fmt::Display for RandNLAError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        use RandNLAError::*;
        match self {
            InvalidParameters(m) | InvalidDimensions(m) | NegativeDimensions(m) | NotOverdetermined(m) | SingularMatrix(m) | NotSquare(m) => write!(f, "{}", m),
            MatrixDecompositionError(m) => write!(f, "Matrix decomposition error: {}", m),
            NotHermitian(m) => write!(f, "Not a Hermitian matrix: {}", m),
            NotPositiveSemiDefinite(m) => write!(f, "Not a positive semi-definite matrix: {}", m),
            ComputationError(m) => write!(f, "Computation error: {}", m),
        }
    }
}
impl Error for RandNLAError {}

/// rnla_status -> RandNLAError (0 = Ok)
pub fn from_status(code: i32) -> Result<(), RandNLAError> {
    use RandNLAError::*;
    let m = crate::ffi::last_message();
    match code {
        0 => Ok(()),
        1 => Err(InvalidParameters(m)),
        2 => Err(InvalidDimensions(m)),
        3 => Err(NegativeDimensions(m)),
        4 => Err(NotOverdetermined(m)),
        5 => Err(NotSquare(m)),
        6 => Err(SingularMatrix(m)),
        7 => Err(MatrixDecompositionError(m)),
        8 => Err(NotHermitian(m)),
        9 => Err(NotPositiveSemiDefinite(m)),
        _ => Err(ComputationError(m)),
    }
}
