"""Writes tests/golden/path_golden.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).
The Rust reference cannot be executed in this environment (no rustc/cargo), so these vectors are oracle outputs:
they guard against regressions of BOTH sides and give the GPU tests a committed, travel-safe target."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc  # noqa: E402
from conftest import lowrank_plus_noise, random_matrix  # noqa: E402

out = {}
A, sig = lowrank_plus_noise(240, 90, seed=21, k=12)
out["svd_A"] = A; out["svd_k"] = 12; out["svd_s"] = 6
for mode in (0, 1):
    U, S, Vt = orc.rand_svd(A, 12, 1e-6, 6, orc.make_opts(mode=mode))
    out[f"svd_sigma_mode{mode}"] = np.diag(S).copy()
out["omega_seed"] = 424242
out["omega_gauss"] = orc.omega_fill(0, 16, 5, seed=424242, stream=1)
rng = np.random.default_rng(22)
Q, _ = np.linalg.qr(rng.standard_normal((60, 60)))
ev = np.concatenate([[4.0, -3.0, 2.5, 2.0, -1.5, 1.0], 1e-7 * rng.standard_normal(54)])
H = (Q * ev) @ Q.T; H = 0.5 * (H + H.T)
out["evd1_A"] = H
out["evd1_lambda"] = orc.rand_evd1(H, 6, 0.1, 6, orc.make_opts(mode=0))[1]
ev2 = np.concatenate([np.logspace(0.5, -0.5, 6), 1e-8 * np.abs(rng.standard_normal(54))])
P = (Q * ev2) @ Q.T; P = 0.5 * (P + P.T)
out["evd2_A"] = P
out["evd2_lambda"] = orc.rand_evd2(P, 6, 4, orc.make_opts(mode=0))[1]
# block sparse-sign operator: S itself (from S I) for two shapes of (zeta, width), and a sketch of a matrix
out["sbs_S_z8w4"] = orc.sketch_apply_saso_block(np.eye(2100), 48, zeta=8, seed=11, width=4)[:, ::7].copy()
out["sbs_S_z4w1"] = orc.sketch_apply_saso_block(np.eye(2100), 48, zeta=4, seed=11, width=1)[:, ::7].copy()
T = random_matrix(2100, 5, seed=24)
out["sbs_T"] = T; out["sbs_SA"] = orc.sketch_apply_saso_block(T, 48, zeta=8, seed=11)
# blendenpik end to end (block sparse-sign sketch and dense Gaussian sketch)
La = random_matrix(900, 12, seed=25) * np.logspace(0, -3, 12); Lb = random_matrix(900, 1, seed=26)
out["lsq_A"] = La; out["lsq_b"] = Lb
out["lsq_x_block"] = orc.blendenpik(La, Lb, 1e-12, 100, 4.0, kind=2, zeta=8)[0]
out["lsq_x_dense"] = orc.blendenpik(La, Lb, 1e-12, 100, 4.0, kind=0)[0]
X = random_matrix(40, 9, seed=23)
out["stab_X"] = X; out["stab_L"] = orc.Stabilizer(X)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "path_golden.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
