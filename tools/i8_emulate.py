"""Prints the CPU-emulated accuracy sweep of rnla_options.range_passes_int8 (levels 1-3 against all-FP64 with the same Omega) that
DESIGN.md section 5c quotes.  Not product code; the emulation itself is tests/i8_emulation.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from i8_emulation import rand_svd_emulated, spectrum_matrix  # noqa: E402

if __name__ == "__main__":
    m, n, k, s = 3000, 1000, 40, 10
    Om = np.random.default_rng(1).standard_normal((n, k + s))
    print("max relative difference of sigma_1..k from the all-FP64 run with the same Omega")
    print("sigma1/sigmak  tail        | level 1   level 2   level 3")
    for kappa in (1e2, 1e3, 1e4, 1e6, 1e8):
        for gap in (1e-2, 1.0, None):
            A, sig = spectrum_matrix(m, n, k, kappa, gap)
            s0 = rand_svd_emulated(A, Om, k, 0)
            row = [np.max(np.abs(rand_svd_emulated(A, Om, k, lvl) - s0) / s0) for lvl in (1, 2, 3)]
            tail = "continues" if gap is None else f"{gap:g} sigma_k"
            print(f"{kappa:12.0e}  {tail:<11} | " + " ".join(f"{x:9.2e}" for x in row))
