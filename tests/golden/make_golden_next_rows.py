"""Writes tests/golden/next_rows_golden.npz from the CPU oracle (run from the repo root:
python tests/golden/make_golden_next_rows.py): small inputs and outputs of the rows after the hot path (SURVEY.md section 8f:
qrcp, sap_chol_qrcp, sketch-and-solve, ID / CUR, saddle point).  The Rust reference cannot be executed here (no rustc/cargo),
so these are oracle outputs: a committed, travel-safe target that both the oracle and the CUDA path are held to."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc  # noqa: E402
from conftest import rank_k_matrix, random_matrix  # noqa: E402

out = {}
A = random_matrix(30, 12, seed=31)
Q, R, p = orc.qrcp(A)
out["qrcp_A"] = A; out["qrcp_R"] = R; out["qrcp_p"] = p; out["qrcp_Q"] = Q
T = np.asfortranarray(np.random.default_rng(32).uniform(-1, 1, (200, 8)))
Qc, Rc, Jc = orc.sap_chol_qrcp(T, 24)
out["cq_A"] = T; out["cq_d"] = 24; out["cq_R"] = Rc; out["cq_J"] = Jc; out["cq_Q"] = Qc
La = random_matrix(320, 9, seed=33); Lb = random_matrix(320, 1, seed=34)
out["sas_A"] = La; out["sas_b"] = Lb
out["sas_x_qr"] = orc.sketched_least_squares(0, La, Lb)
out["sas_x_svd"] = orc.sketched_least_squares(1, La, Lb)
M = rank_k_matrix(40, 33, 10, seed=35)
out["id_A"] = M; out["id_k"] = 10
X, J = orc.osid_qrcp(M, 10, orc.COLUMN); out["id_col_X"] = X; out["id_col_J"] = J
X, I = orc.osid_qrcp(M, 10, orc.ROW); out["id_row_X"] = X; out["id_row_I"] = I
X, J = orc.osid_randomised(M, 10, orc.COLUMN); out["id_rand_X"] = X; out["id_rand_J"] = J
for name, rnd in (("det", False), ("rand", True)):
    J, U, I = orc.cur(M, 10, rnd); out[f"cur_{name}_J"] = J; out[f"cur_{name}_U"] = U; out[f"cur_{name}_I"] = I
    Z, I, J, X = orc.two_sided_id(M, 10, rnd, orc.make_opts(mode=1))
    out[f"tsid_{name}_Z"] = Z; out[f"tsid_{name}_I"] = I; out[f"tsid_{name}_J"] = J; out[f"tsid_{name}_X"] = X
Sa = random_matrix(150, 7, seed=36); Sb = random_matrix(150, 1, seed=37); Sc = random_matrix(7, 1, seed=38)
out["sp_A"] = Sa; out["sp_b"] = Sb; out["sp_c"] = Sc
out["sp_x_mu0"] = orc.saddle_point(Sa, Sb, Sc, 0.0, 1e-12, 200, 2.0)[0]
out["sp_x_mu2"] = orc.saddle_point(Sa, Sb, Sc, 2.0, 1e-12, 200, 2.0)[0]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "next_rows_golden.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
