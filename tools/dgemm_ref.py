# cuBLAS dgemm sanity ceiling: 8192^3 and the tall-skinny 200k x 20k x 112 shape (scaled to fit quickly).
import torch, json, time
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
a = torch.randn(8192, 8192, dtype=torch.float64, device='cuda'); b = torch.randn(8192, 8192, dtype=torch.float64, device='cuda')
ms = t(lambda: a @ b); out['cublas_dgemm_8192_tflops'] = 2 * 8192**3 / ms * 1e-9
del a, b
m, n, l = 100000, 20000, 112
A = torch.randn(n, m, dtype=torch.float64, device='cuda').t()   # column-major m x n
S = torch.randn(l, n, dtype=torch.float64, device='cuda').t()   # column-major n x l
ms = t(lambda: A @ S); out['cublas_AS_100kx20kx112_ms'] = ms; out['cublas_AS_tflops'] = 2.0 * m * n * l / ms * 1e-9; out['cublas_AS_gbs'] = 8.0 * m * n / ms * 1e-6
Q = torch.randn(l, m, dtype=torch.float64, device='cuda').t()
ms = t(lambda: A.t() @ Q); out['cublas_AtQ_100kx20kx112_ms'] = ms; out['cublas_AtQ_tflops'] = 2.0 * m * n * l / ms * 1e-9
print(json.dumps(out))
