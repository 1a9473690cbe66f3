//! `randblas` with the sketch-and-factor hot path on a B200 (reference src/lib.rs:8-19 keeps these module names).
//! Every function below has the reference's exact signature and forwards to the C ABI in include/rnla.h.
#![allow(non_snake_case)]
pub mod errors;
pub mod ffi;
pub mod lora_drivers;
pub mod lora_helpers;
pub mod sketch;
pub mod sketch_and_precondition;
pub mod pivot_decompositions;
pub mod cqrrpt;
pub mod sketch_and_solve;
pub mod id;
pub mod solvers;
pub mod cg;
