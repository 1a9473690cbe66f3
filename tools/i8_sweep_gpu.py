"""The accuracy sweep of rnla_options.range_passes_int8 on the GPU against the oracle (same matrices as
tests/test_gpu_parity.py::test_int8_accuracy_contract_sweep_against_the_oracle): prints the table quoted in DESIGN.md section 5c."""
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import i8_emulation as em
from oracle import oracle as orc
from randnla_b200 import runtime as rt, lora_drivers as ld
orc.load(); rt.init(0)
m, n, k, s = 6000, 1500, 32, 10
print("max relative deviation of sigma_1..k from the oracle (all FP64, same Omega); 6000 x 1500, k = 32, s = 10")
print("sigma1/sigmak  tail         | fp64 kernels  level 1    level 2    level 3 = auto")
for kappa in (1e2, 1e3, 1e4, 1e6, 1e8):
    for gap in (1e-2, 1.0, None):
        A, sig = em.spectrum_matrix(m, n, k, kappa, gap, seed=int(np.log10(kappa)))
        so = np.diag(orc.rand_svd(A, k, 1e-6, s, orc.make_opts(mode=0))[1])
        row = []
        for level in (0, 1, 2, 3):
            with rt.options(range_passes_int8=level):
                S = ld.rand_svd(A, k, 1e-6, s)[1]
            row.append(float(np.max(np.abs(np.diag(S) - so) / so)))
        tail = "continues" if gap is None else f"{gap:g} sigma_k"
        print(f"{kappa:12.0e}  {tail:<12} | " + "  ".join(f"{x:9.2e}" for x in row), flush=True)
