"""Process-level plumbing around librnla.so: options, stream binding, phase timings, launch counter,
device-buffer helpers (torch is used for device memory / streams / torch.distributed only)."""
import contextlib
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Options, check

MODE_INTENDED, MODE_LITERAL = 0, 1
GAUSSIAN, UNIFORM, RADEMACHER = 0, 1, 2
GEN_PHILOX, GEN_THREEFRY = 0, 1


def init(device=None):
    lib = _lib.load()
    check(lib.rnla_init(-1 if device is None else int(device)))


def get_options():
    o = Options()
    _lib.load().rnla_get_options(C.byref(o))
    return o


def set_options(**kw):
    """Update the process-wide defaults used by the reference-signature functions
    (mode, dist, seed, num_passes, passes_per_stab, fused_sketch)."""
    lib = _lib.load()
    o = get_options()
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k!r}")
        setattr(o, k, int(v))
    check(lib.rnla_set_options(C.byref(o)))
    return o


@contextlib.contextmanager
def options(**kw):
    """Temporarily override options (e.g. `with options(mode=MODE_LITERAL): ...`)."""
    lib = _lib.load()
    old = get_options()
    set_options(**kw)
    try:
        yield
    finally:
        check(lib.rnla_set_options(C.byref(old)))


def make_options(**kw):
    o = Options()
    _lib.load().rnla_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k!r}")
        setattr(o, k, int(v))
    return o


def kernel_launches():
    return int(_lib.load().rnla_kernel_launches())


def timings():
    """Per-phase CUDA-event timings (ms) of the last driver call, in order."""
    lib = _lib.load()
    cap = 256
    names = (C.c_char_p * cap)()
    ms = (C.c_double * cap)()
    n = min(lib.rnla_get_timings(names, ms, cap), cap)
    return [(names[i].decode(), float(ms[i])) for i in range(n)]


def synchronize():
    check(_lib.load().rnla_synchronize())


def use_torch_stream():
    """Run the library on torch's current CUDA stream (so torch.cuda.Event brackets see its kernels)."""
    import torch
    check(_lib.load().rnla_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))


# ---- host arrays ------------------------------------------------------------------------------
def as_f(a):
    """float64, column-major (nalgebra DMatrix layout), 2-D."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError("expected a matrix")
    return np.asfortranarray(a)


def ptr(a):
    return C.c_void_p(a.ctypes.data)


# ---- device (torch) matrices ---------------------------------------------------------------------
def empty_colmajor(rows, cols, device=None):
    """Column-major f64 device matrix as a torch tensor of shape (rows, cols), strides (1, rows)."""
    import torch
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.empty((cols, max(rows, 1)), dtype=torch.float64, device=dev).t()[:rows, :]


def to_device_colmajor(a):
    import torch
    a = as_f(a)
    t = empty_colmajor(a.shape[0], a.shape[1])
    t.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    return t


def dev_ptr_ld(t):
    """(void*, ld) of a column-major f64 CUDA tensor."""
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.dim() == 2):
        raise TypeError("expected a 2-D float64 CUDA tensor")
    rows, cols = t.shape
    if rows > 1 and t.stride(0) != 1:
        raise ValueError("matrix must be column-major (stride(0) == 1); use runtime.empty_colmajor")
    ld = t.stride(1) if cols > 1 else max(rows, 1)
    if ld < max(rows, 1):
        raise ValueError("bad leading dimension")
    # The library launches on its own (non-blocking) stream unless use_torch_stream() was called: work torch has queued on ITS
    # stream for this tensor (a fill, an upload's transpose kernel) must have finished before the library reads or writes it.
    cur = torch.cuda.current_stream(t.device)
    if (_lib.load().rnla_stream() or 0) != cur.cuda_stream:
        cur.synchronize()
    return C.c_void_p(t.data_ptr()), int(ld)


# ---- multi-GPU -----------------------------------------------------------------------------------
def init_comm_from_torch():
    """Create the library's NCCL communicator for the ranks of torch.distributed's default group.
    Rank 0 draws the NCCL unique id; it travels through torch.distributed (plumbing only)."""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return
    uid = (C.c_uint8 * 128)()
    if rank == 0:
        check(lib.rnla_comm_unique_id(uid))
    obj = [bytes(uid) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    buf = (C.c_uint8 * 128).from_buffer_copy(obj[0])
    check(lib.rnla_comm_init(world, rank, buf))


def shard_rows(m_global, world, rank):
    """Contiguous row shard [start, stop) owned by `rank`; shard sizes are multiples of 4 where possible
    so that Philox row-quads never straddle ranks."""
    base = (m_global // world) // 4 * 4
    start = rank * base
    stop = m_global if rank == world - 1 else start + base
    if base == 0:
        start = min(rank, m_global)
        stop = min(rank + 1, m_global) if rank < world - 1 else m_global
    return start, stop
