"""Quick device-resident timing of rand_svd at the headline size (C2: 200k x 20k, k=100, s=10, q=2)."""
import sys, json, time
import numpy as np, torch
sys.path.insert(0, ".")
import randnla_b200 as rb
from randnla_b200 import runtime as rt, _lib, lora_drivers as ld
import ctypes as C
lib = _lib.load(); rt.init(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
k, s = 100, 10
r0 = 200
sigma = np.concatenate([np.logspace(0, -3, 100), np.full(r0 - 100, 1e-5)])
dA = rt.empty_colmajor(m, n)
pA, lda = rt.dev_ptr_ld(dA)
t0 = time.time()
_lib.check(lib.rnla_generate_lowrank_dev(pA, lda, m, n, 0, m, r0, sigma.ctypes.data_as(C.c_void_p), 1e-7, 1234)); rt.synchronize()
print("generate: %.2f s" % (time.time() - t0), flush=True)
for fused in (1, 0):
    o = rt.make_options(fused_sketch=fused)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        U, S, Vt = ld.rand_svd_dev(dA, k, s, o); rt.synchronize()
        dt = time.time() - t0
        tm = rt.timings()
        tot = sum(x[1] for x in tm)
        print(f"fused={fused} it={it} wall={dt*1e3:.2f} ms  phases_sum={tot:.2f} ms  A-stream {4*8*m*n/dt*1e-9:.0f} GB/s  {4*2*m*n*110/dt*1e-12:.2f} TF/s")
    print(json.dumps(tm))
    Sg = S.cpu().numpy()
    print("sigma rel err vs planted (noise-limited):", float(np.abs(Sg - sigma[:k]).max() / 1.0), "min ratio", float((Sg / sigma[:k]).min()), float((Sg/sigma[:k]).max()))
print("launches", rt.kernel_launches())
