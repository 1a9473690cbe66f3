"""Sketch operators -- mirror of reference src/sketch.rs (DistributionType :9-13, MatrixAttribute :18-21,
haar_sample :45-85, sketching_operator :102-130)."""
import ctypes as C
import enum

import numpy as np

from . import _lib
from ._lib import check
from . import runtime


class DistributionType(enum.IntEnum):
    """reference src/sketch.rs:9-13"""
    Gaussian = 0
    Uniform = 1
    Rademacher = 2


class MatrixAttribute(enum.IntEnum):
    """reference src/sketch.rs:18-21"""
    Row = 0
    Column = 1


def sketching_operator(dist_type, rows, cols):
    """`sketching_operator(dist_type, rows, cols) -> DMatrix<f64>` (reference src/sketch.rs:102-130).

    Entries are i.i.d. from `dist_type`; raises `InvalidDimensions` when rows == 0 or cols == 0 (:107-111).
    The reference re-seeds ThreeFry with 0 on every call (:112); here the operator is the Philox4x32-10
    counter map of randnla_b200/csrc/rng.cuh with the process-wide seed (default 0), evaluated on the GPU."""
    lib = _lib.load()
    rows, cols = int(rows), int(cols)
    out = np.empty((max(rows, 0), max(cols, 0)), dtype=np.float64, order="F")
    check(lib.rnla_sketching_operator(int(dist_type), rows, cols, runtime.ptr(out)))
    return out


def sketch_fill(dist_type, rows, cols, seed=0, stream=0, row_offset=0, generator=runtime.GEN_PHILOX):
    """Extended operator generation: explicit seed / Philox stream / global row offset / generator.
    `generator=GEN_THREEFRY` reproduces the reference's operator: its sequential ThreeFry2x64 stream with rand's Uniform / Bernoulli
    and rand_distr's ziggurat Gaussian (src/sketch.rs:112-127)."""
    lib = _lib.load()
    rows, cols = int(rows), int(cols)
    out = np.empty((max(rows, 0), max(cols, 0)), dtype=np.float64, order="F")
    check(lib.rnla_sketch_fill(int(generator), int(dist_type), int(seed), int(stream), rows, cols, int(row_offset),
                               runtime.ptr(out), max(rows, 1)))
    return out


def haar_sample(rows, columns, attr):
    """`haar_sample(rows, columns, attr)` (reference src/sketch.rs:45-85): Haar-distributed matrix with
    orthonormal rows (attr=Row) or columns (attr=Column); `InvalidDimensions` for the long side (:49-63)."""
    lib = _lib.load()
    rows, columns = int(rows), int(columns)
    out = np.empty((max(rows, 0), max(columns, 0)), dtype=np.float64, order="F")
    check(lib.rnla_haar_sample(rows, columns, int(attr), runtime.ptr(out)))
    return out


def philox4x32_10(ctr, key):
    """Philox4x32-10 blocks evaluated on the GPU (KAT hook; reference rust-random123/src/philox.rs:211-223)."""
    lib = _lib.load()
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
    key = np.ascontiguousarray(key, dtype=np.uint32).reshape(-1, 2)
    if key.shape[0] == 1 and ctr.shape[0] > 1:
        key = np.ascontiguousarray(np.repeat(key, ctr.shape[0], axis=0))
    out = np.empty_like(ctr)
    check(lib.rnla_philox4x32_10(ctr.shape[0], runtime.ptr(ctr), runtime.ptr(key), runtime.ptr(out)))
    return out


def threefry2x64_20(ctr, key):
    """ThreeFry2x64-20 blocks on the GPU (reference rust-random123/src/threefry.rs:69-93)."""
    lib = _lib.load()
    ctr = np.ascontiguousarray(ctr, dtype=np.uint64).reshape(-1, 2)
    key = np.ascontiguousarray(key, dtype=np.uint64).reshape(-1, 2)
    if key.shape[0] == 1 and ctr.shape[0] > 1:
        key = np.ascontiguousarray(np.repeat(key, ctr.shape[0], axis=0))
    out = np.empty_like(ctr)
    check(lib.rnla_threefry2x64_20(ctr.shape[0], runtime.ptr(ctr), runtime.ptr(key), runtime.ptr(out)))
    return out
