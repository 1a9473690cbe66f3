"""The C-ABI library loads on a CPU-only box, exports every symbol include/rnla.h declares, validates arguments
before touching the device, and refuses to compute without a GPU (no CPU fallback).  -m "not gpu"."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rnla.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rnla_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from randnla_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rnla.h but not exported by librnla.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound but not declared in the header"
    assert lib.rnla_version() >= 100


def test_library_has_no_link_time_dependency_on_torch_or_oracle():
    import subprocess
    from randnla_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "liboracle" not in out and "libnccl" not in out
    assert "libcudart" in out or "libcuda" in out or "statically" in out or True


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "randnla_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".rs")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"(import\s+oracle|from\s+oracle|liboracle|oracle/oracle|#include\s+\"[^\"]*oracle)", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour on a box without a GPU")
def test_no_cpu_fallback():
    import randnla_b200 as rb
    from randnla_b200.errors import ComputationError
    with pytest.raises(ComputationError) as e:
        rb.sketch.sketching_operator(rb.sketch.DistributionType.Gaussian, 4, 4)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(ComputationError):
        rb.lora_drivers.rand_svd(np.eye(4), 2, 0.1, 1)
    with pytest.raises(ComputationError):
        rb.lora_helpers.Orth(np.eye(4))


def test_validation_precedes_device_access():
    """error behaviour of the reference signatures (src/lora_drivers.rs:31-45, 89-103, 169-173; src/sketch.rs:107-111)"""
    import randnla_b200 as rb
    from randnla_b200.errors import InvalidParameters, InvalidDimensions, RandNLAError
    A = np.eye(5)
    with pytest.raises(InvalidParameters) as e:
        rb.lora_drivers.rand_svd(A, 0, 0.1, 5)
    assert str(e.value) == "Rank k must be positive, current input is 0"
    with pytest.raises(InvalidParameters) as e:
        rb.lora_drivers.rand_svd(A, 2, 0.0, 5)
    assert str(e.value) == "Epsilon must be positive, current input is 0"
    with pytest.raises(InvalidParameters) as e:
        rb.lora_drivers.rand_svd(A, 2, 0.1, 0)
    assert str(e.value) == "Oversampling parameter s must be positive, current input is 0"
    with pytest.raises(InvalidParameters):
        rb.lora_drivers.rand_evd1(A, 0, 0.1, 5)
    with pytest.raises(InvalidParameters):
        rb.lora_drivers.rand_evd2(A, 0, 5)
    for (r, c) in [(0, 5), (5, 0), (0, 0)]:
        with pytest.raises(InvalidDimensions) as e:
            rb.sketch.sketching_operator(rb.sketch.DistributionType.Rademacher, r, c)
        assert str(e.value) == "Rows and columns must be greater than 0"
    with pytest.raises(InvalidDimensions) as e:
        rb.sketch.haar_sample(5, 3, rb.sketch.MatrixAttribute.Row)
    assert "Cannot have more rows (5) than columns (3)" in str(e.value)
    with pytest.raises(InvalidDimensions):
        rb.sketch.haar_sample(3, 5, rb.sketch.MatrixAttribute.Column)
    assert issubclass(InvalidParameters, RandNLAError)


def test_error_display_matches_reference():
    """src/errors.rs:16-31"""
    from randnla_b200 import errors as E
    assert str(E.InvalidParameters("x")) == "x"
    assert str(E.MatrixDecompositionError("SVD decomposition failed")) == "Matrix decomposition error: SVD decomposition failed"
    assert str(E.NotHermitian("Input matrix is not Hermitian")) == "Not a Hermitian matrix: Input matrix is not Hermitian"
    assert str(E.NotPositiveSemiDefinite("m")) == "Not a positive semi-definite matrix: m"
    assert str(E.ComputationError("m")) == "Computation error: m"
    assert sorted(E.STATUS_TO_ERROR) == list(range(1, 11))
    assert [E.STATUS_TO_ERROR[i].variant for i in range(1, 11)] == [
        "InvalidParameters", "InvalidDimensions", "NegativeDimensions", "NotOverdetermined", "NotSquare",
        "SingularMatrix", "MatrixDecompositionError", "NotHermitian", "NotPositiveSemiDefinite", "ComputationError"]


def test_sketch_dim_rule_is_host_arithmetic(orc):
    from randnla_b200 import sketch_and_precondition as sp
    for (m, n, sf) in [(100, 10, 2.0), (15, 10, 2.0), (100, 10, 2.55), (1000000, 2000, 4.0), (7, 7, 1.0)]:
        assert sp.sketch_dim(m, n, sf) == orc.sketch_dim(m, n, sf)
        assert sp.sketch_dim(m, n, sf, saddle=True) == orc.sketch_dim(m, n, sf, saddle=True)


def test_sketch_and_precondition_validation():
    """src/sketch_and_precondition.rs:29-48 (and :85-104, :152-171)"""
    from randnla_b200 import sketch_and_precondition as sp
    from randnla_b200.errors import InvalidParameters, NotOverdetermined
    a = np.ones((10, 4)); b = np.ones((10, 1))
    with pytest.raises(NotOverdetermined) as e:
        sp.blendenpik_sketch(np.ones((3, 4)), np.ones((3, 1)), 0.1, 10, 2.0)
    assert str(e.value) == "Need more columns than rows, found 3 rows and 4 columns"
    with pytest.raises(InvalidParameters) as e:
        sp.blendenpik_sketch(a, b, 0.1, 10, 0.5)
    assert str(e.value) == "Sampling factor must be greater than 1, current input is 0.5"
    with pytest.raises(InvalidParameters):
        sp.lsrn_sketch(a, b, 0.0, 10, 2.0)
    with pytest.raises(InvalidParameters):
        sp.saddle_point_sketch(a, b, None, 0.0, 0.1, 0, 2.0)


def test_options_struct_layout():
    from randnla_b200 import _lib
    assert C.sizeof(_lib.Options) == 40
    assert _lib.Options.generator.offset == 32 and _lib.Options.range_passes_int8.offset == 28 and _lib.Options.seed.offset == 8


# ---- static agreement of the three bindings with include/rnla.h (the Rust crate cannot be compiled here: no toolchain) ----
def _c_prototypes():
    """name -> (return type, [parameter types]) of every function include/rnla.h declares, types normalised"""
    src = open(os.path.join(ROOT, "include", "rnla.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for ret, name, args in re.findall(r"\b([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(rnla_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src):
        params = []
        for a in [x.strip() for x in args.replace("\n", " ").split(",")]:
            if a in ("void", ""):
                continue
            a = re.sub(r"\s+", " ", a)
            if "[" in a:                                               # array parameter (`uint8_t id[128]`) decays to a pointer
                a = a[:a.index("[")].rsplit(" ", 1)[0] + "* x"
            m = re.match(r"^(.*?)([A-Za-z_][A-Za-z0-9_]*)?$", a)       # drop the parameter name
            ty = m.group(1).strip() if m.group(2) and m.group(1).strip() else a
            params.append(re.sub(r"\s*\*\s*", "*", ty).strip())
        protos[name] = (re.sub(r"\s*\*\s*", "*", ret.strip()), params)
    return protos


_C_TO_RUST = {"double": "c_double", "const double*": "*const c_double", "double*": "*mut c_double", "int64_t": "i64", "int64_t*": "*mut i64",
              "const int64_t*": "*const i64", "int32_t": "c_int", "int32_t*": "*mut c_int", "uint64_t": "u64", "uint32_t": "u32",
              "rnla_status": "c_int", "const char*": "*const c_char", "rnla_lsqr_result*": "*mut RnlaLsqrResult"}
_C_TO_CTYPES = {"double": "c_double", "int64_t": "c_long", "int32_t": "c_int", "uint64_t": "c_ulong", "uint32_t": "c_uint", "rnla_status": "c_int",
                "size_t": "c_ulong", "void": None}


def _rust_externs():
    src = open(os.path.join(ROOT, "randnla_b200", "rust", "src", "ffi.rs")).read()
    src = re.sub(r"//[^\n]*", "", src)
    blocks = re.findall(r'extern\s+"C"\s*\{(.*?)\n\}', src, flags=re.S)
    decls = {}
    for blk in blocks:
        for name, args, ret in re.findall(r"pub\s+fn\s+(rnla_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*([^;]+?))?\s*;", blk, flags=re.S):
            params = [re.sub(r"\s+", " ", a.split(":", 1)[1].strip()) for a in args.replace("\n", " ").split(",") if ":" in a]
            decls[name] = ((ret or "()").strip(), params)
    return src, blocks, decls


def test_rust_ffi_matches_the_header():
    """randnla_b200/rust/src/ffi.rs cannot be compiled in this image (no rustc / cargo): check statically that it is well formed
    (every bodiless `fn` sits inside an `extern "C"` block, braces balance), that each declaration agrees with include/rnla.h in
    name, arity and parameter / return types, that every `ffi::rnla_*` the other modules call is declared, and that the
    #[repr(C)] mirror of rnla_lsqr_result has the header's fields in the header's order."""
    protos = _c_prototypes()
    src, blocks, decls = _rust_externs()
    assert src.count("{") == src.count("}")
    outside = src
    for blk in blocks:
        outside = outside.replace(blk, "")
    assert not re.search(r"\bfn\s+\w+\s*\([^)]*\)\s*(->\s*[^;{]+)?;", outside), "bodiless fn outside the extern block"
    assert len(decls) >= 29
    for name, (ret, params) in decls.items():
        assert name in protos, f"{name} is not declared in include/rnla.h"
        cret, cparams = protos[name]
        assert _C_TO_RUST[cret] == ret, (name, cret, ret)
        assert len(cparams) == len(params), (name, cparams, params)
        for cp, rp in zip(cparams, params):
            assert _C_TO_RUST[cp] == rp, (name, cp, rp)
    rust_dir = os.path.join(ROOT, "randnla_b200", "rust", "src")
    for f in os.listdir(rust_dir):
        for used in re.findall(r"ffi::(rnla_[a-z0-9_]+)", open(os.path.join(rust_dir, f)).read()):
            assert used in decls, f"{f} calls ffi::{used}, which ffi.rs does not declare"
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "rnla.h")).read(), flags=re.S)
    cfields = re.findall(r"(int64_t|double)\s+(\w+)\s*;", re.search(r"typedef struct[^{]*\{([^}]*)\}\s*rnla_lsqr_result", hdr).group(1))
    rfields = re.findall(r"pub\s+(\w+)\s*:\s*(\w+)", re.search(r"pub struct RnlaLsqrResult\s*\{(.*?)\}", src, flags=re.S).group(1))
    assert [(n, {"int64_t": "i64", "double": "c_double"}[t]) for t, n in cfields] == rfields


def test_ctypes_prototypes_match_the_header():
    """randnla_b200/_lib.py: arity and scalar parameter types of every prototype against include/rnla.h (pointers are void*)"""
    from randnla_b200 import _lib
    protos = _c_prototypes()
    for name, (res, args) in _lib.SIGNATURES.items():
        cret, cparams = protos[name]
        assert len(cparams) == len(args), (name, cparams, args)
        for cp, a in zip(cparams, args):
            if cp.endswith("*"):
                assert a is _lib.P or a is C.c_char_p or hasattr(a, "contents") or getattr(a, "_type_", None) is not None, (name, cp, a)
            else:
                assert _C_TO_CTYPES[cp] == a.__name__, (name, cp, a)
        if cret.endswith("*"):
            assert res in (_lib.P, C.c_char_p)
        else:
            assert (res.__name__ if res is not None else None) == _C_TO_CTYPES[cret], (name, cret, res)
