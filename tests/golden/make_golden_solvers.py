"""Writes tests/golden/solvers_golden.npz (run from the repo root: python tests/golden/make_golden_solvers.py): small inputs and
known answers for the iterative solvers and lupp (reference src/solvers.rs:115-278, src/cg.rs, src/pivot_decompositions.rs:21-86).
Independent generators where they exist: scipy's `lsqr` (the algorithm the reference's `lsqr` translates, its doc comment
:108-110) for six iterations with the stopping tests off; LAPACK (`numpy.linalg.solve / lstsq`, `scipy.linalg.lu`) for the
solutions and the pivot order; the CPU oracle for the bit-exact LU factors and the cgls / conjugate_grad iteration counts."""
import os
import sys

import numpy as np
import scipy.linalg as sl
from scipy.sparse.linalg import lsqr as sp_lsqr

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc  # noqa: E402
from conftest import random_matrix  # noqa: E402

out = {}
rng = np.random.default_rng(41)
A = np.asfortranarray(rng.standard_normal((90, 24)) * np.logspace(0, -2, 24))
b = rng.standard_normal(90)
x0 = rng.standard_normal(24)
out["lsqr_A"] = A; out["lsqr_b"] = b; out["lsqr_x0"] = x0
for tag, kw in (("plain", dict(damp=0.0, x0=None)), ("damped_x0", dict(damp=0.3, x0=x0))):
    r = sp_lsqr(A, b, damp=kw["damp"], atol=0.0, btol=0.0, conlim=0.0, iter_lim=6, calc_var=True, x0=kw["x0"])
    out[f"lsqr_{tag}_x"] = r[0]
    out[f"lsqr_{tag}_scalars"] = np.array([r[1], r[2], r[3], r[4], r[5], r[6], r[8]], dtype=np.float64)   # istop itn r1 r2 anorm acond xnorm
    out[f"lsqr_{tag}_var"] = r[9]
out["lsqr_lstsq_x"] = np.linalg.lstsq(A, b, rcond=None)[0]
# cgls / conjugate_grad
xs, it, conv = orc.cgls(A, b, 1e-11, 500)
out["cgls_x"] = xs[:, 0]; out["cgls_iterations"] = np.array([it, int(conv)])
G = rng.standard_normal((30, 30)); Sm = np.asfortranarray(G @ G.T + 30 * np.eye(30)); sb = rng.standard_normal(30)
out["cg_A"] = Sm; out["cg_b"] = sb; out["cg_x_lapack"] = np.linalg.solve(Sm, sb)
xo, ito, convo = orc.conjugate_grad(Sm, sb)
out["cg_iterations"] = np.array([ito, int(convo)])
# lupp
L_in = random_matrix(40, 40, seed=42)
Lo, Uo, po = orc.lupp(L_in)
P, Ls, Us = sl.lu(L_in)
assert np.array_equal(P.T @ L_in, L_in[po, :]), "LAPACK picked different pivots"
out["lupp_A"] = L_in; out["lupp_L"] = Lo; out["lupp_U"] = Uo; out["lupp_p"] = po; out["lupp_L_lapack"] = Ls; out["lupp_U_lapack"] = Us
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "solvers_golden.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
