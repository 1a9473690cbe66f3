"""randnla_b200 -- B200-native (sm_100a) sketch-and-factor hot path of divijkhaitan/randnla.

Python mirror of the reference's Rust modules over the C ABI in include/rnla.h:

    randnla_b200.sketch                   <- reference src/sketch.rs
    randnla_b200.lora_helpers             <- reference src/lora_helpers.rs
    randnla_b200.lora_drivers             <- reference src/lora_drivers.rs
    randnla_b200.sketch_and_precondition  <- reference src/sketch_and_precondition.rs (sketch step and the three drivers)
    randnla_b200.pivot_decompositions     <- reference src/pivot_decompositions.rs (qrcp, economic_qrcp)
    randnla_b200.cqrrpt                   <- reference src/cqrrpt.rs
    randnla_b200.sketch_and_solve         <- reference src/sketch_and_solve.rs
    randnla_b200.id                       <- reference src/id.rs
    randnla_b200.solvers                  <- reference src/solvers.rs (lsqr)
    randnla_b200.cg                       <- reference src/cg.rs (cgls, conjugate_grad, verify_solution)
    randnla_b200.errors                   <- reference src/errors.rs

All arithmetic runs in hand-written CUDA inside librnla.so; this package only marshals numpy / torch
buffers.  There is no CPU implementation in here.
"""
from . import errors  # noqa: F401
from .errors import RandNLAError  # noqa: F401
from . import _lib  # noqa: F401
from . import runtime  # noqa: F401
from . import sketch, lora_helpers, lora_drivers, sketch_and_precondition  # noqa: F401
from . import pivot_decompositions, cqrrpt, sketch_and_solve, id, solvers, cg  # noqa: F401

__all__ = ["errors", "RandNLAError", "runtime", "sketch", "lora_helpers", "lora_drivers", "sketch_and_precondition",
           "pivot_decompositions", "cqrrpt", "sketch_and_solve", "id", "solvers", "cg"]
