/* oracle_next.c -- CPU restatement of the rows SURVEY.md section 8(f) lists after the hot path: the callers and
 * neighbours of the sketch.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Reference files followed (read-only tree /root/reference):
 *   src/pivot_decompositions.rs:105-180 (qrcp), :196-269 (economic_qrcp)
 *   src/cqrrpt.rs:27-58                    sap_chol_qrcp
 *   src/sketch_and_solve.rs:24-65          sketched_least_squares_qr / _svd
 *   src/solvers.rs:22-69                   solve_upper_triangular_system / solve_diagonal_system
 *   src/id.rs:34-331                       cur, two_sided_id(_randomised), cur_randomised, osid_randomised, osid_qrcp
 *   src/sketch_and_precondition.rs:150-216 sketch_saddle_point_precondition
 * Everything here draws its sketching operators from this build's Philox map (orc_omega_fill), exactly where the reference
 * calls `sketching_operator` -- the reference's Gaussian values are parity-unpinned (oracle_core.c header) -- and follows the
 * reference literally after that.  nalgebra 0.33 pieces restated from the published algorithms: `solve_upper_triangular`,
 * `pseudo_inverse(eps)` (SVD, reciprocal of the singular values > eps, 0 otherwise), `select_columns/rows`.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define AT(M, ld, i, j) (M)[(int64_t)(i) + (int64_t)(j) * (int64_t)(ld)]
static double* dalloc(int64_t n) { return (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }
static int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }

/* pivot_decompositions.rs:105-180 (steps = min(m, n)) and :196-269 (steps = k): Householder QR with column pivoting on
 * exactly recomputed trailing column norms, first maximum wins (strict >, :122-127).  R: m x n work matrix (all rows, as the
 * reference keeps them); Q: m x m accumulated product of the reflectors, or NULL; perm: n. */
void orc_qrcp_steps(const double* A, int64_t m, int64_t n, int64_t steps, double* R, double* Q, int64_t* perm) {
    memcpy(R, A, (size_t)(m * n) * sizeof(double));
    if (Q) { memset(Q, 0, (size_t)(m * m) * sizeof(double)); for (int64_t i = 0; i < m; ++i) AT(Q, m, i, i) = 1.0; }
    double* norms = dalloc(n); double* v = dalloc(m); double* w = dalloc(m);
    for (int64_t j = 0; j < n; ++j) {
        perm[j] = j;
        double s = 0.0; for (int64_t i = 0; i < m; ++i) s += AT(R, m, i, j) * AT(R, m, i, j);
        norms[j] = sqrt(s);                                                                /* :113-115 */
    }
    for (int64_t k = 0; k < steps; ++k) {
        double mx = norms[k]; int64_t mi = k;
        for (int64_t j = k + 1; j < n; ++j) if (norms[j] > mx) { mx = norms[j]; mi = j; }  /* :118-127 */
        if (mi != k) {                                                                     /* :130-134 */
            for (int64_t i = 0; i < m; ++i) { const double t = AT(R, m, i, k); AT(R, m, i, k) = AT(R, m, i, mi); AT(R, m, i, mi) = t; }
            const int64_t tp = perm[k]; perm[k] = perm[mi]; perm[mi] = tp;
            const double tn = norms[k]; norms[k] = norms[mi]; norms[mi] = tn;
        }
        const int64_t len = m - k;
        double s = 0.0;
        for (int64_t i = 0; i < len; ++i) { v[i] = AT(R, m, k + i, k); s += v[i] * v[i]; } /* :137-140 */
        const double norm_x = sqrt(s);
        if (norm_x == 0.0) continue;                                                       /* :143 */
        v[0] += (v[0] >= 0.0) ? norm_x : -norm_x;                                          /* :145 */
        double s2 = 0.0; for (int64_t i = 0; i < len; ++i) s2 += v[i] * v[i];
        const double nv = sqrt(s2);
        for (int64_t i = 0; i < len; ++i) v[i] /= nv;                                      /* :146 */
#pragma omp parallel for schedule(static) if ((n - k) * len > 100000)
        for (int64_t j = k; j < n; ++j) {                                                  /* :149-154 */
            double dot = 0.0;
            for (int64_t i = 0; i < len; ++i) dot += v[i] * AT(R, m, k + i, j);
            for (int64_t i = 0; i < len; ++i) AT(R, m, k + i, j) -= 2.0 * v[i] * dot;
        }
        if (Q) {                                                                           /* :157-167: q = q (I - 2 v v^T) */
            for (int64_t r = 0; r < m; ++r) {
                double dot = 0.0;
                for (int64_t i = 0; i < len; ++i) dot += AT(Q, m, r, k + i) * v[i];
                w[r] = dot;
            }
            for (int64_t i = 0; i < len; ++i)
                for (int64_t r = 0; r < m; ++r) AT(Q, m, r, k + i) -= 2.0 * w[r] * v[i];
        }
        for (int64_t j = k + 1; j < n; ++j) {                                              /* :169-171 */
            double t = 0.0; for (int64_t i = k + 1; i < m; ++i) t += AT(R, m, i, j) * AT(R, m, i, j);
            norms[j] = sqrt(t);
        }
    }
    free(norms); free(v); free(w);
}

/* nalgebra solve_upper_triangular(&B): X = U^-1 B for the leading k x k block of U (ld ldu); returns -1 on a zero diagonal */
static int solve_upper(const double* U, int64_t ldu, int64_t k, const double* B, int64_t ldb, int64_t nrhs, double* X, int64_t ldx) {
    for (int64_t i = 0; i < k; ++i) if (AT(U, ldu, i, i) == 0.0) return -1;
    for (int64_t c = 0; c < nrhs; ++c) {
        for (int64_t i = 0; i < k; ++i) AT(X, ldx, i, c) = AT(B, ldb, i, c);
        for (int64_t i = k - 1; i >= 0; --i) {
            const double xi = AT(X, ldx, i, c) / AT(U, ldu, i, i);
            AT(X, ldx, i, c) = xi;
            for (int64_t r = 0; r < i; ++r) AT(X, ldx, r, c) -= xi * AT(U, ldu, r, i);
        }
    }
    return 0;
}

static void sketch_any(int kind, int dist_or_width, int zeta, uint64_t seed, int64_t d, const double* A, int64_t m, int64_t n, double* Ask) {
    if (kind == 0) orc_sketch_apply_dense(dist_or_width, seed, d, A, m, n, Ask);
    else if (kind == 1) orc_sketch_apply_saso(seed, d, zeta, A, m, n, Ask);
    else orc_sketch_apply_saso_block(seed, d, zeta, dist_or_width ? dist_or_width : (zeta < 4 ? zeta : 4), A, m, n, 0, Ask);
}

/* src/cqrrpt.rs:27-58.  Q: m x n buffer (first k columns valid, ld m), R: n x n buffer (k x n valid, ld k), J: n.
 * Returns 0, 1 if !(n <= d <= m) (the reference asserts :29), 6 if the rank-k block cannot be inverted, 7 if Cholesky fails. */
int orc_sap_chol_qrcp(const double* A, int64_t m, int64_t n, int64_t d, int kind, int dist_or_width, int zeta, uint64_t seed,
                      double* Q, double* R, int64_t* J, int64_t* k_out) {
    if (!(n <= d && d <= m)) return 1;
    double* Ask = dalloc(d * n); double* Rsk = dalloc(d * n);
    sketch_any(kind, dist_or_width, zeta, seed, d, A, m, n, Ask);                          /* :31-33 */
    orc_qrcp_steps(Ask, d, n, imin(d, n), Rsk, NULL, J);                                   /* :35 */
    int64_t k = 0;
    for (int64_t i = 0; i < imin(d, n); ++i) if (fabs(AT(Rsk, d, i, i)) > 1e-10) ++k;       /* :37-43 */
    *k_out = k;
    int rc = 0;
    double* I = dalloc(k * k); double* Rinv = dalloc(k * k); double* Apre = dalloc(m * k); double* G = dalloc(k * k); double* L = dalloc(k * k);
    for (int64_t i = 0; i < k; ++i) AT(I, k, i, i) = 1.0;
    if (solve_upper(Rsk, d, k, I, k, k, Rinv, k) != 0) rc = 6;                             /* :47 */
    if (!rc) {
        double* Aperm = dalloc(m * k);
        for (int64_t c = 0; c < k; ++c) memcpy(Aperm + c * m, A + J[c] * m, (size_t)m * sizeof(double));   /* :46 */
        orc_gemm_nn(Aperm, m, m, k, Rinv, k, k, Apre, m);                                  /* :48 */
        free(Aperm);
        orc_gemm_tn(Apre, m, m, k, Apre, m, k, G, k);                                      /* :50 */
        if (orc_cholesky_lower(G, k, L) != 0) rc = 7;                                      /* :51 */
    }
    if (!rc) {
        double* Rpre = dalloc(k * k); double* Rpinv = dalloc(k * k);
        for (int64_t i = 0; i < k; ++i) for (int64_t j = i; j < k; ++j) AT(Rpre, k, i, j) = AT(L, k, j, i);   /* :52 */
        if (solve_upper(Rpre, k, k, I, k, k, Rpinv, k) != 0) rc = 6;
        else {
            orc_gemm_nn(Apre, m, m, k, Rpinv, k, k, Q, m);                                 /* :53 */
            orc_gemm_nn(Rpre, k, k, k, Rsk, d, n, R, k);                                   /* :55 */
        }
        free(Rpre); free(Rpinv);
    }
    free(Ask); free(Rsk); free(I); free(Rinv); free(Apre); free(G); free(L);
    return rc;
}

/* src/solvers.rs:22-41: rows with a zero diagonal keep x_i = 0 */
static void solve_upper_triangular_system(const double* U, int64_t ldu, int64_t n, const double* y, double* x) {
    for (int64_t i = 0; i < n; ++i) x[i] = 0.0;
    for (int64_t i = n - 1; i >= 0; --i) {
        if (AT(U, ldu, i, i) != 0.0) {
            double sum = 0.0;
            for (int64_t j = i + 1; j < n; ++j) sum += AT(U, ldu, i, j) * x[j];
            x[i] = (y[i] - sum) / AT(U, ldu, i, i);
        }
    }
}

/* src/sketch_and_solve.rs:24-33 (which = 0) and :54-66 (which = 1); d = rows / 4 (:26, :56).  Returns 0, or 2 when the
 * sketch has fewer rows than columns (the reference then indexes out of bounds / returns a vector of the wrong length). */
int orc_sketched_least_squares(int which, const double* A, int64_t m, int64_t n, const double* b, int kind, int dist_or_width,
                               int zeta, uint64_t seed, double* x) {
    const int64_t d = m / 4;
    if (d < n || n <= 0) return 2;
    double* Ask = dalloc(d * n); double* bsk = dalloc(d);
    sketch_any(kind, dist_or_width, zeta, seed, d, A, m, n, Ask);
    sketch_any(kind, dist_or_width, zeta, seed, d, b, m, 1, bsk);
    int rc = 0;
    if (which == 0) {
        double* Q = dalloc(d * n); double* R = dalloc(n * n); double* z = dalloc(n);
        orc_qr(Ask, d, n, Q, R);                                                           /* :29 */
        orc_gemm_tn(Q, d, d, n, bsk, d, 1, z, n);                                          /* :30 */
        solve_upper_triangular_system(R, n, n, z, x);                                      /* :31 */
        free(Q); free(R); free(z);
    } else {
        double* U = dalloc(d * n); double* sg = dalloc(n); double* Vt = dalloc(n * n); double* z = dalloc(n);
        if (orc_svd(Ask, d, n, U, sg, Vt) != 0) rc = 7;                                    /* :59 */
        else {
            orc_gemm_tn(U, d, d, n, bsk, d, 1, z, n);                                      /* :63 */
            for (int64_t i = 0; i < n; ++i) z[i] = (sg[i] != 0.0) ? z[i] / sg[i] : 0.0;    /* :64, solvers.rs:57-69 */
            orc_gemm_tn(Vt, n, n, n, z, n, 1, x, n);                                       /* :65  v * x */
        }
        free(U); free(sg); free(Vt); free(z);
    }
    free(Ask); free(bsk);
    return rc;
}

/* src/id.rs:272-318.  attr 1 = Column: Y (l x w) ~ Y[:, J] X, X k x w; attr 0 = Row: Y ~ X Y[J, :], X l x k.
 * Returns 0, 1 for an invalid k (the reference asserts :278-279), 6 when R1 is singular (unwrap :290). */
int orc_osid_qrcp(const double* Y, int64_t l, int64_t w, int64_t k, int attr, double* X, int64_t* J) {
    if (k <= 0 || k > imin(l, w)) return 1;
    if (attr == 0) {                                                                       /* :309-314 */
        double* Yt = dalloc(l * w); double* Xt = dalloc(k * l);
        for (int64_t i = 0; i < l; ++i) for (int64_t j = 0; j < w; ++j) AT(Yt, w, j, i) = AT(Y, l, i, j);
        const int rc = orc_osid_qrcp(Yt, w, l, k, 1, Xt, J);
        if (!rc) for (int64_t i = 0; i < k; ++i) for (int64_t j = 0; j < l; ++j) AT(X, l, j, i) = AT(Xt, k, i, j);
        free(Yt); free(Xt);
        return rc;
    }
    double* R = dalloc(l * w); int64_t* p = (int64_t*)calloc((size_t)w, sizeof(int64_t));
    orc_qrcp_steps(Y, l, w, k, R, NULL, p);                                                /* :283 */
    double* T = dalloc(k * (w - k));
    int rc = 0;
    if (solve_upper(R, l, k, R + k * l, l, w - k, T, k) != 0) rc = 6;                      /* :285-290 */
    else {
        memset(X, 0, (size_t)(k * w) * sizeof(double));
        for (int64_t idx = 0; idx < k; ++idx) AT(X, k, idx, p[idx]) = 1.0;                 /* :297-301 */
        for (int64_t c = 0; c < w - k; ++c) for (int64_t i = 0; i < k; ++i) AT(X, k, i, p[k + c]) = AT(T, k, i, c);   /* :304-308 */
        for (int64_t i = 0; i < k; ++i) J[i] = p[i];
    }
    free(R); free(p); free(T);
    return rc;
}

/* src/id.rs:217-249.  Column: S = Gaussian k x m (this build's dense operator, seed), Y = S A, column ID of Y.
 * Row: S = tsog1(A, k, 2, 1) (n x k), Y = A S^T -- conformal only when n == k (:230-233; reached through
 * two_sided_id_randomised :99) -- then row ID of Y.  Returns 2 for the non-conformal case. */
int orc_osid_randomised(const double* A, int64_t m, int64_t n, int64_t k, int attr, const orc_opts* o, double* X, int64_t* J) {
    if (k <= 0 || k > imin(m, n)) return 1;
    if (attr == 1) {
        double* Y = dalloc(k * n);
        orc_sketch_apply_dense(0, o ? o->seed : 0, k, A, m, n, Y);                         /* :238-243 */
        const int rc = orc_osid_qrcp(Y, k, n, k, 1, X, J);                                 /* :246 */
        free(Y);
        return rc;
    }
    if (n != k) return 2;
    double* S = dalloc(n * k); double* St = dalloc(k * n); double* Y = dalloc(m * k);
    orc_tsog1(A, m, n, k, 2, 1, o, S);                                                     /* :230 */
    for (int64_t i = 0; i < n; ++i) for (int64_t j = 0; j < k; ++j) AT(St, k, j, i) = AT(S, n, i, j);
    orc_gemm_nn(A, m, m, n, St, k, n, Y, m);                                               /* :233 (k == n) */
    const int rc = orc_osid_qrcp(Y, m, k, k, 0, X, J);                                     /* :236 */
    free(S); free(St); free(Y);
    return rc;
}

static void select_columns(const double* A, int64_t m, const int64_t* J, int64_t k, double* out) {
    for (int64_t c = 0; c < k; ++c) memcpy(out + c * m, A + J[c] * m, (size_t)m * sizeof(double));
}
static void select_rows(const double* A, int64_t m, int64_t n, const int64_t* I, int64_t k, double* out) {
    for (int64_t j = 0; j < n; ++j) for (int64_t i = 0; i < k; ++i) AT(out, k, i, j) = AT(A, m, I[i], j);
}

/* src/id.rs:118-129 (randomised = 0) and :94-101 (randomised = 1): Z m x k, I k, J k, X k x n */
int orc_two_sided_id(int randomised, const double* A, int64_t m, int64_t n, int64_t k, const orc_opts* o,
                     double* Z, int64_t* I, int64_t* J, double* X) {
    int rc = randomised ? orc_osid_randomised(A, m, n, k, 1, o, X, J) : orc_osid_qrcp(A, m, n, k, 1, X, J);
    if (rc) return rc;
    double* Ac = dalloc(m * k);
    select_columns(A, m, J, k, Ac);
    rc = randomised ? orc_osid_randomised(Ac, m, k, k, 0, o, Z, I) : orc_osid_qrcp(Ac, m, k, k, 0, Z, I);
    free(Ac);
    return rc;
}

/* nalgebra pseudo_inverse(0.0) of M (rows x cols) -> cols x rows */
static int pinv(const double* M, int64_t rows, int64_t cols, double* out) {
    const int64_t p = imin(rows, cols);
    double* U = dalloc(rows * p); double* sg = dalloc(p); double* Vt = dalloc(p * cols);
    const int rc = orc_svd(M, rows, cols, U, sg, Vt);
    if (!rc) {
        for (int64_t i = 0; i < cols; ++i)
            for (int64_t j = 0; j < rows; ++j) {
                double s = 0.0;
                for (int64_t t = 0; t < p; ++t) if (sg[t] > 0.0) s += AT(Vt, p, t, i) * (1.0 / sg[t]) * AT(U, rows, j, t);
                AT(out, cols, i, j) = s;
            }
    }
    free(U); free(sg); free(Vt);
    return rc;
}

/* src/id.rs:34-71 (randomised = 0) and :154-193 (randomised = 1): J k, U k x k, I k */
int orc_cur(int randomised, const double* A, int64_t m, int64_t n, int64_t k, const orc_opts* o, int64_t* J, double* U, int64_t* I) {
    if (k <= 0 || k > imin(m, n)) return 1;
    int rc;
    if (m >= n) {
        double* X = dalloc(k * n); double* Ac = dalloc(m * k); double* Act = dalloc(k * m); double* R = dalloc(k * m);
        int64_t* p = (int64_t*)calloc((size_t)m, sizeof(int64_t));
        rc = randomised ? orc_osid_randomised(A, m, n, k, 1, o, X, J) : orc_osid_qrcp(A, m, n, k, 1, X, J);   /* :42 / :162 */
        if (!rc) {
            select_columns(A, m, J, k, Ac);                                                /* :44 */
            for (int64_t i = 0; i < m; ++i) for (int64_t j = 0; j < k; ++j) AT(Act, k, j, i) = AT(Ac, m, i, j);
            orc_qrcp_steps(Act, k, m, k, R, NULL, p);                                      /* :46 */
            for (int64_t i = 0; i < k; ++i) I[i] = p[i];                                   /* :48 */
            double* Ar = dalloc(k * n); double* Pi = dalloc(n * k);
            select_rows(A, m, n, I, k, Ar);                                                /* :51 */
            rc = pinv(Ar, k, n, Pi);
            if (!rc) orc_gemm_nn(X, k, k, n, Pi, n, k, U, k);                              /* :52 */
            free(Ar); free(Pi);
        }
        free(X); free(Ac); free(Act); free(R); free(p);
    } else {
        double* At = dalloc(m * n); double* Z = dalloc(k * m);
        for (int64_t i = 0; i < m; ++i) for (int64_t j = 0; j < n; ++j) AT(At, n, j, i) = AT(A, m, i, j);   /* :56 */
        rc = randomised ? orc_osid_randomised(At, n, m, k, 1, o, Z, I) : orc_osid_qrcp(At, n, m, k, 1, Z, I);   /* :57 / :177 */
        if (!rc) {
            double* Ar = dalloc(k * n); double* R = dalloc(k * n); int64_t* p = (int64_t*)calloc((size_t)n, sizeof(int64_t));
            select_rows(A, m, n, I, k, Ar);                                                /* :59 */
            orc_qrcp_steps(Ar, k, n, k, R, NULL, p);                                       /* :62 */
            for (int64_t i = 0; i < k; ++i) J[i] = p[i];                                   /* :65 */
            double* Ac = dalloc(m * k); double* Pi = dalloc(k * m); double* Zt = dalloc(m * k);
            select_columns(A, m, J, k, Ac);                                                /* :68 */
            rc = pinv(Ac, m, k, Pi);                                                       /* :69 */
            for (int64_t i = 0; i < k; ++i) for (int64_t j = 0; j < m; ++j) AT(Zt, m, j, i) = AT(Z, k, i, j);
            if (!rc) orc_gemm_nn(Pi, k, k, m, Zt, m, k, U, k);                             /* :70 */
            free(Ar); free(R); free(p); free(Ac); free(Pi); free(Zt);
        }
        free(At); free(Z);
    }
    return rc;
}

/* src/sketch_and_precondition.rs:150-216 with the dense operator S^T(j, i) = omega(row j, col i), stream 3 (what
 * orc_sketch_apply_dense applies).  c may be NULL (`c.is_empty()` :189).  x: n, y: m.  Returns 0, 4 / 1 for the validation
 * errors (:152-171), 7 if the SVD fails, 2 when mu == 0 and the sketch is rank deficient (the reference's shapes then do not
 * conform at :212). */
int orc_saddle_point(const double* A, int64_t m, int64_t n, const double* b, const double* c, double mu, double epsilon, int64_t l,
                     double sampling_factor, int dist, uint64_t seed, double* x, double* y, int64_t* iters_out, int* converged_out) {
    if (m < n) return 4;
    if (sampling_factor < 1.0 || epsilon <= 0.0 || l <= 0) return 1;
    const int64_t d = orc_sketch_dim(m, n, sampling_factor, 1);                            /* :172 */
    const int64_t r = imin(d, n);
    double* St = dalloc(m * d); double* Ask = dalloc(d * n);
    orc_omega_fill(dist, seed, 3u, m, d, 0, St, m);                                        /* :173 */
    orc_gemm_tn(St, m, m, d, A, m, n, Ask, d);                                             /* :176 */
    double* U = dalloc(d * r); double* sg = dalloc(r); double* Vt = dalloc(r * n);
    int rc = 0;
    if (orc_svd(Ask, d, n, U, sg, Vt) != 0) rc = 7;                                        /* :179-182 */
    int64_t kk = r;
    if (!rc && !(mu > 0.0)) {
        kk = 0; while (kk < r && sg[kk] > 1e-10) ++kk;                                     /* :188 */
        if (kk != r) rc = 2;
    }
    if (!rc) {
        double* M = dalloc(n * kk); double* w = dalloc(r);
        for (int64_t j = 0; j < kk; ++j) {
            w[j] = (mu > 0.0) ? 1.0 / sqrt(sg[j] * sg[j] + mu) : 1.0 / sg[j];              /* :186, :189 */
            for (int64_t i = 0; i < n; ++i) AT(M, n, i, j) = AT(Vt, r, j, i) * w[j];
        }
        double* Ap = dalloc(m * kk);
        orc_gemm_nn(A, m, m, n, M, n, kk, Ap, m);                                          /* :192 */
        double* bmod = dalloc(m);
        memcpy(bmod, b, (size_t)m * sizeof(double));                                       /* :194 */
        if (c) {                                                                           /* :195-205 */
            double* bs = dalloc(r); double* t = dalloc(d); double* tm = dalloc(m);
            orc_gemm_nn(Vt, r, r, n, c, n, 1, bs, r);                                      /* :196 */
            for (int64_t j = 0; j < r; ++j) bs[j] *= (mu > 0.0) ? 1.0 / sqrt(sg[j] * sg[j] + mu) : 1.0 / sg[j];
            orc_gemm_nn(U, d, d, r, bs, r, 1, t, d);
            orc_gemm_nn(St, m, m, d, t, d, 1, tm, m);                                      /* s^T (u (..)) */
            for (int64_t i = 0; i < m; ++i) bmod[i] -= tm[i];                              /* :205 */
            free(bs); free(t); free(tm);
        }
        double* sb = dalloc(d); double* z = dalloc(r);
        orc_gemm_tn(St, m, m, d, bmod, m, 1, sb, d);
        orc_gemm_tn(U, d, d, r, sb, d, 1, z, r);                                           /* :208 */
        int conv = 0;
        const int64_t it = orc_cgls(Ap, m, kk, bmod, epsilon, l, z, &conv);                /* :211 */
        orc_gemm_nn(M, n, n, kk, z, kk, 1, x, n);                                          /* :213 */
        orc_gemm_nn(A, m, m, n, x, n, 1, y, m);
        for (int64_t i = 0; i < m; ++i) y[i] = b[i] - y[i];                                /* :214 */
        if (iters_out) *iters_out = it;
        if (converged_out) *converged_out = conv;
        free(M); free(w); free(Ap); free(bmod); free(sb); free(z);
    }
    free(St); free(Ask); free(U); free(sg); free(Vt);
    return rc;
}

/* ======================================================================== src/solvers.rs:65-98 sym_ortho, :115-278 lsqr
 * (the reference's translation of scipy 1.14.1 sparse.linalg.lsqr), statement for statement on a dense a (m x n).
 * x0 may be NULL.  arnorms: iter_lim doubles (history of ||A^T r|| estimates); returns its length.  var: n doubles. */
static double sgn(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0); }
static void sym_ortho(double a, double b, double* c, double* s, double* r) {
    if (b == 0.0) { *c = sgn(a); *s = 0.0; *r = fabs(a); }
    else if (a == 0.0) { *c = 0.0; *s = sgn(b); *r = fabs(b); }
    else if (fabs(b) > fabs(a)) { const double tau = a / b; *s = sgn(b) / sqrt(1.0 + tau * tau); *c = *s * tau; *r = b / *s; }
    else { const double tau = b / a; *c = sgn(a) / sqrt(1.0 + tau * tau); *s = *c * tau; *r = a / *c; }
}
static double nrm2(const double* v, int64_t n) { double s = 0.0; for (int64_t i = 0; i < n; ++i) s += v[i] * v[i]; return sqrt(s); }

int64_t orc_lsqr(const double* a, int64_t m, int64_t n, const double* b, double damp, double atol, double btol, double conlim,
                 int64_t iter_lim, int calc_var, const double* x0, double* x, int64_t* istop_out, int64_t* itn_out,
                 double* r1norm_out, double* r2norm_out, double* anorm_out, double* acond_out, double* arnorms, double* xnorm_out,
                 double* var) {
    const double eps = 2.220446049250313e-16;
    if (iter_lim < 0) iter_lim = 2 * n;                                                  /* :140 */
    for (int64_t i = 0; i < n; ++i) { var[i] = 0.0; x[i] = x0 ? x0[i] : 0.0; }
    double* u = dalloc(m); double* v = dalloc(n); double* w = dalloc(n); double* t = dalloc(m > n ? m : n);
    memcpy(u, b, (size_t)m * sizeof(double));
    const double bnorm = nrm2(b, m);
    double beta;
    if (x0) { orc_gemm_nn(a, m, m, n, x, n, 1, t, m); for (int64_t i = 0; i < m; ++i) u[i] -= t[i]; beta = nrm2(u, m); }   /* :148-153 */
    else beta = bnorm;
    { const double sc = beta > 0.0 ? 1.0 / beta : 0.0; for (int64_t i = 0; i < m; ++i) u[i] *= sc; }                       /* :156 */
    orc_gemm_tn(a, m, m, n, u, m, 1, v, n);                                              /* :157 */
    double alfa = nrm2(v, n);
    { const double sc = alfa > 0.0 ? 1.0 / alfa : 0.0; for (int64_t i = 0; i < n; ++i) v[i] *= sc; }
    memcpy(w, v, (size_t)n * sizeof(double));
    double rhobar = alfa, phibar = beta, rnorm = beta, r1norm = rnorm, r2norm = rnorm;
    double anorm = 0.0, acond = 0.0, ddnorm = 0.0, res2 = 0.0, xnorm = 0.0, xxnorm = 0.0, z = 0.0, cs2 = -1.0, sn2 = 0.0;
    const double dampsq = damp * damp;
    double arnorm = alfa * beta;
    int64_t itn = 0, istop = 0, nhist = 0;
    if (arnorm == 0.0) {                                                                 /* :181-183 */
        arnorms[0] = 0.0; nhist = 1;
        *istop_out = 0; *itn_out = 0; *r1norm_out = beta; *r2norm_out = beta; *anorm_out = 0.0; *acond_out = 0.0; *xnorm_out = 0.0;
        free(u); free(v); free(w); free(t);
        return nhist;
    }
    const double ctol = conlim > 0.0 ? 1.0 / conlim : 0.0;
    while (itn < iter_lim) {
        arnorms[itn] = arnorm; nhist = itn + 1;
        ++itn;
        orc_gemm_nn(a, m, m, n, v, n, 1, t, m);
        for (int64_t i = 0; i < m; ++i) u[i] = t[i] - alfa * u[i];                       /* :195 */
        beta = nrm2(u, m);
        if (beta > 0.0) {
            for (int64_t i = 0; i < m; ++i) u[i] *= 1.0 / beta;
            anorm = sqrt(anorm * anorm + alfa * alfa + beta * beta + dampsq);
            orc_gemm_tn(a, m, m, n, u, m, 1, t, n);
            for (int64_t i = 0; i < n; ++i) v[i] = t[i] - beta * v[i];                   /* :202 */
            alfa = nrm2(v, n);
            { const double sc = alfa > 0.0 ? 1.0 / alfa : 0.0; for (int64_t i = 0; i < n; ++i) v[i] *= sc; }
        }
        const double rhobar1 = sqrt(rhobar * rhobar + dampsq);
        const double cs1 = rhobar / rhobar1, sn1 = damp / rhobar1;
        const double psi = sn1 * phibar;
        phibar *= cs1;
        double cs, sn, rho;
        sym_ortho(rhobar1, beta, &cs, &sn, &rho);
        const double theta = sn * alfa;
        rhobar = -cs * alfa;
        const double phi = cs * phibar;
        phibar *= sn;
        const double tau = sn * phi;
        const double t1 = phi / rho, t2 = -theta / rho;
        double dkn = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            const double dk = w[i] * (1.0 / rho);                                        /* :227 */
            x[i] += t1 * w[i];
            w[i] = v[i] + t2 * w[i];
            dkn += dk * dk;
            if (calc_var) var[i] += dk * dk;
        }
        ddnorm += dkn;
        const double delta = sn2 * rho, gambar = -cs2 * rho, rhs = phi - delta * z, zbar = rhs / gambar;
        xnorm = sqrt(xxnorm + zbar * zbar);
        const double gamma = sqrt(gambar * gambar + theta * theta);
        cs2 = gambar / gamma; sn2 = theta / gamma; z = rhs / gamma;
        xxnorm += z * z;
        acond = anorm * sqrt(ddnorm);
        const double res1 = phibar * phibar;
        res2 += psi * psi;
        rnorm = sqrt(res1 + res2);
        arnorm = alfa * fabs(tau);
        const double r1sq = rnorm * rnorm - dampsq * xxnorm;
        r1norm = sqrt(fabs(r1sq));
        r2norm = rnorm;
        const double test1 = rnorm / bnorm, test2 = arnorm / (anorm * rnorm + eps), test3 = 1.0 / (acond + eps);
        const double tt1 = test1 / (1.0 + anorm * xnorm / bnorm), rtol = atol + btol * (anorm * xnorm / bnorm);
        if (itn >= iter_lim) istop = 7;
        if (1.0 + test3 <= 1.0) istop = 6;
        if (1.0 + test2 <= 1.0) istop = 5;
        if (1.0 + tt1 <= 1.0) istop = 4;
        if (test3 <= ctol) istop = 3;
        if (test2 <= atol) istop = 2;
        if (test1 <= rtol) istop = 1;
        if (istop != 0) break;
    }
    *istop_out = istop; *itn_out = itn; *r1norm_out = r1norm; *r2norm_out = r2norm; *anorm_out = anorm; *acond_out = acond; *xnorm_out = xnorm;
    free(u); free(v); free(w); free(t);
    return nhist;
}

/* ======================================================================== src/cg.rs:77-112 conjugate_grad, statement for
 * statement.  x: in the initial guess (the reference's default is ones), out the solution.  Returns 0, or 9
 * (NotPositiveSemiDefinite) when an eigenvalue of a is negative (:80-86; beyond rounding noise of the Jacobi solver here).
 * iters_out: the loop index printed at convergence (2 n when the loop runs out). */
int orc_conjugate_grad(const double* a, int64_t n, const double* b, double* x, int64_t* iters_out, int* converged_out) {
    {
        double* W = dalloc(n * n); double* lam = dalloc(n);
        orc_symmetric_eigen(a, n, W, lam);
        double amax = 0.0; int bad = 0;
        for (int64_t i = 0; i < n; ++i) if (fabs(lam[i]) > amax) amax = fabs(lam[i]);
        for (int64_t i = 0; i < n; ++i) if (lam[i] < -1e-12 * (amax > 1e-300 ? amax : 1e-300) * (double)n) bad = 1;
        free(W); free(lam);
        if (bad) return 9;
    }
    double* r = dalloc(n); double* p = dalloc(n); double* ap = dalloc(n);
    orc_gemm_nn(a, n, n, n, x, n, 1, r, n);
    for (int64_t i = 0; i < n; ++i) { r[i] -= b[i]; p[i] = -r[i]; }                       /* :89-90 */
    double rk = 0.0; for (int64_t i = 0; i < n; ++i) rk += r[i] * r[i];                   /* :91 */
    int64_t it; int conv = 0;
    for (it = 0; it < 2 * n; ++it) {                                                      /* :93 */
        orc_gemm_nn(a, n, n, n, p, n, 1, ap, n);                                          /* :94 */
        double pap = 0.0; for (int64_t i = 0; i < n; ++i) pap += p[i] * ap[i];
        const double alpha = rk / pap;                                                    /* :95 */
        for (int64_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] += alpha * ap[i]; }  /* :96-97 */
        double rk1 = 0.0; for (int64_t i = 0; i < n; ++i) rk1 += r[i] * r[i];             /* :98 */
        if (rk1 < 1e-10) { conv = 1; break; }                                             /* :100-103 */
        const double beta = rk1 / rk;                                                     /* :105 */
        rk = rk1;
        for (int64_t i = 0; i < n; ++i) p[i] = beta * p[i] - r[i];                        /* :107 */
    }
    if (iters_out) *iters_out = it;
    if (converged_out) *converged_out = conv;
    free(r); free(p); free(ap);
    return 0;
}

/* ======================================================================== src/pivot_decompositions.rs:21-86 lupp, statement
 * for statement (this file is compiled with -ffp-contract=off: multiply and subtract round separately, as in the reference).
 * a: n x n column-major.  L, U: n x n.  perm: n.  Returns 0, or 6 (SingularMatrix) on an exactly zero pivot column (:44-48). */
int orc_lupp(const double* a, int64_t n, double* L, double* U, int64_t* perm) {
    double* lu = dalloc(n * n);
    memcpy(lu, a, (size_t)(n * n) * sizeof(double));
    for (int64_t i = 0; i < n; ++i) perm[i] = i;                                          /* :30 */
    for (int64_t k = 0; k + 1 < n; ++k) {                                                 /* :32 */
        int64_t pivot_row = k;
        double pivot_val = fabs(lu[k + k * n]);
        for (int64_t i = k + 1; i < n; ++i) {                                             /* :36-42 */
            const double val = fabs(lu[i + k * n]);
            if (val > pivot_val) { pivot_val = val; pivot_row = i; }
        }
        if (pivot_val == 0.0) { free(lu); return 6; }                                     /* :44-48 */
        if (pivot_row != k) {                                                             /* :50-60 */
            for (int64_t j = 0; j < n; ++j) { const double t = lu[k + j * n]; lu[k + j * n] = lu[pivot_row + j * n]; lu[pivot_row + j * n] = t; }
            const int64_t t = perm[k]; perm[k] = perm[pivot_row]; perm[pivot_row] = t;
        }
        for (int64_t i = k + 1; i < n; ++i) {                                             /* :62-70 */
            const double multiplier = lu[i + k * n] / lu[k + k * n];
            lu[i + k * n] = multiplier;
            for (int64_t j = k + 1; j < n; ++j) lu[i + j * n] -= multiplier * lu[k + j * n];
        }
    }
    for (int64_t j = 0; j < n; ++j)                                                       /* :73-84 */
        for (int64_t i = 0; i < n; ++i) {
            L[i + j * n] = i > j ? lu[i + j * n] : (i == j ? 1.0 : 0.0);
            U[i + j * n] = i > j ? 0.0 : lu[i + j * n];
        }
    free(lu);
    return 0;
}
