"""probe one shape of rnla_normal_pass_dev in a fresh process: python tools/np_probe.py m n ld off"""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from randnla_b200 import runtime as rt, _lib
lib = _lib.load(); rt.init(0)
m, n, ld, off = (int(v) for v in sys.argv[1:5])
rng = np.random.default_rng(0)
parent = np.asfortranarray(rng.standard_normal((ld, n)))
dP = rt.to_device_colmajor(parent); rt.dev_ptr_ld(dP)
A = parent[off:off + m]
x = rng.standard_normal(n)
dx = torch.from_numpy(x).cuda(); dt = torch.empty(n + 1, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
_lib.check(lib.rnla_normal_pass_dev(C.c_void_p(dP.data_ptr() + 8 * off), ld, m, n, C.c_void_p(dx.data_ptr()), 1.0, None, 0.0, None, C.c_void_p(dt.data_ptr())))
rt.synchronize()
t = dt.cpu().numpy(); ur = A @ x
print("ok", np.abs(t[:n] - A.T @ ur).max() / (np.abs(A).T @ np.abs(ur)).max())
