"""`RandNLAError` hierarchy -- mirrors the reference enum (reference src/errors.rs:3-14, Display :16-31).

Status codes of the C ABI (include/rnla.h `rnla_status`) map 1:1 onto the variants.
"""


class RandNLAError(Exception):
    """Base class; `variant` is the reference's enum variant name."""
    variant = "RandNLAError"
    prefix = ""

    def __init__(self, msg=""):
        self.msg = msg
        super().__init__(msg)

    def __str__(self):  # reference Display impl, src/errors.rs:16-31
        return f"{self.prefix}{self.msg}"


class InvalidParameters(RandNLAError):
    variant = "InvalidParameters"


class InvalidDimensions(RandNLAError):
    variant = "InvalidDimensions"


class NegativeDimensions(RandNLAError):
    variant = "NegativeDimensions"


class NotOverdetermined(RandNLAError):
    variant = "NotOverdetermined"


class NotSquare(RandNLAError):
    variant = "NotSquare"


class SingularMatrix(RandNLAError):
    variant = "SingularMatrix"


class MatrixDecompositionError(RandNLAError):
    variant = "MatrixDecompositionError"
    prefix = "Matrix decomposition error: "


class NotHermitian(RandNLAError):
    variant = "NotHermitian"
    prefix = "Not a Hermitian matrix: "


class NotPositiveSemiDefinite(RandNLAError):
    variant = "NotPositiveSemiDefinite"
    prefix = "Not a positive semi-definite matrix: "


class ComputationError(RandNLAError):
    variant = "ComputationError"
    prefix = "Computation error: "


STATUS_TO_ERROR = {
    1: InvalidParameters,
    2: InvalidDimensions,
    3: NegativeDimensions,
    4: NotOverdetermined,
    5: NotSquare,
    6: SingularMatrix,
    7: MatrixDecompositionError,
    8: NotHermitian,
    9: NotPositiveSemiDefinite,
    10: ComputationError,
}
