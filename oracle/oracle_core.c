/* oracle_core.c -- CPU restatement of the reference path.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Compile with -ffp-contract=off: the Gaussian transform and the full-pivot LU are specified operation by
 * operation (separately rounded mul/add, explicit fma) so that the GPU path can be compared bit for bit.
 *
 * Reference files followed (read-only tree /root/reference):
 *   rust-random123/src/philox.rs:3-7,23-27,149-154,173-176,211-223    Philox4x32-10
 *   rust-random123/src/threefry.rs:30-93                               ThreeFry2x64-20
 *   rust-random123/src/rng.rs:80-96                                    BlockRng64 wrapper (stream order)
 *   src/sketch.rs:102-130                                              sketching_operator
 *   src/lora_helpers.rs:17-146                                         QB1, RF1, tsog1, Orth, Stabilizer
 *   src/lora_drivers.rs:30-224                                         rand_svd, rand_evd1, rand_evd2
 *   src/sketch_and_precondition.rs:49-52,105-107,172-176               sketch step
 * Third-party arithmetic that is NOT in the tree and is restated from the published algorithms:
 *   nalgebra 0.33.0 (Cargo.lock:644-645): Householder `qr()` with the sign convention of
 *     `householder::reflection_axis_mut` / `QR::q()` / `QR::r()`; `FullPivLU::new` + `lu::gauss_step(_swap)` + `.l()`;
 *     `cholesky()`; `svd()` and `symmetric_eigen()` (values only are unique: computed here with Jacobi rotations)
 *   rand 0.8.5 (Cargo.lock:861-862): `Uniform<f64>` and `Bernoulli` sampling
 *   rand_core 0.6.4 (Cargo.lock:890-891): `SeedableRng::seed_from_u64` (PCG32 expansion), `BlockRng64`
 *   rand_distr 0.4.3 ziggurat tables are absent -> the reference's Gaussian values are "parity unpinned".
 */
#include "oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define AT(M, ld, i, j) (M)[(int64_t)(i) + (int64_t)(j) * (int64_t)(ld)]
static double* dalloc(int64_t n) { return (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }
static int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }

/* ======================================================================== L0: counter-based RNGs */
static void mul32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {          /* philox.rs:3-7 */
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    *hi = (uint32_t)(p >> 32); *lo = (uint32_t)p;
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int round = 0; round < 10; ++round) {                                 /* philox.rs:211-223 */
        if (round > 0) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }           /* philox.rs:173-176, 26-27 */
        uint32_t hi0, lo0, hi1, lo1;
        mul32(0xD2511F53u, c[0], &hi0, &lo0);                                  /* philox.rs:149-154, 24-25 */
        mul32(0xCD9E8D57u, c[2], &hi1, &lo1);
        const uint32_t n[4] = {hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0};
        memcpy(c, n, sizeof n);
    }
    memcpy(out, c, 4 * sizeof(uint32_t));
}

static uint64_t rotl(uint64_t x, unsigned r) { return (x << r) | (x >> (64 - r)); }
void orc_threefry2x64_20(const uint64_t ctr[2], const uint64_t key[2], uint64_t out[2]) {
    /* threefry.rs:69-93 */
    uint64_t ks[3] = {key[0], key[1], 0xA9FC1A22ull + ((uint64_t)0x1BD11BDAull << 32)};
    ks[2] ^= key[0]; ks[2] ^= key[1];
    uint64_t x[2] = {ctr[0], ctr[1]};
    static const unsigned R1[4] = {16, 42, 12, 31}, R2[4] = {16, 32, 24, 21};   /* threefry.rs:34-41 */
    x[0] += ks[0]; x[1] += ks[1];                                              /* sbox 0 */
    for (int s = 1; s <= 5; ++s) {
        const unsigned* R = (s & 1) ? R1 : R2;
        for (int i = 0; i < 4; ++i) { x[0] += x[1]; x[1] = rotl(x[1], R[i]); x[1] ^= x[0]; }
        x[0] += ks[s % 3]; x[1] += ks[(s + 1) % 3]; x[1] += (uint64_t)s;       /* sbox s, threefry.rs:61-67 */
    }
    out[0] = x[0]; out[1] = x[1];
}

void orc_seed_from_u64(uint64_t state, uint64_t key[2]) {
    /* rand_core 0.6.4 SeedableRng::seed_from_u64 -> 16 seed bytes -> le::read_u64_into (threefry.rs:23-27) */
    uint32_t w[4];
    for (int i = 0; i < 4; ++i) {
        state = state * 6364136223846793005ull + 11634580027462260723ull;
        const uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        const uint32_t rot = (uint32_t)(state >> 59);
        w[i] = (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
    }
    key[0] = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    key[1] = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
}

uint64_t orc_threefry_rng_u64(uint64_t seed, uint64_t t) {
    /* BlockRng64<ThreeFry2x64>: results of block b = threefry(ctr=(b,0)), handed out x[0], x[1]  (rng.rs:89,96; threefry.rs:12-21) */
    uint64_t key[2], ctr[2] = {t >> 1, 0}, x[2];
    orc_seed_from_u64(seed, key);
    orc_threefry2x64_20(ctr, key, x);
    return x[t & 1];
}

/* ======================================================================== L1: sketch entries */
/* this build's Gaussian transform, restated from its specification (DESIGN.md "Omega"):
 * every operation is an IEEE-754 binary32 round-to-nearest op; fmaf is a single-rounding fma. */
static float log_pos_f32(float t) {
    uint32_t bits; memcpy(&bits, &t, 4);
    int e = (int)(bits >> 23) - 127;
    uint32_t mb = (bits & 0x007fffffu) | 0x3f800000u;
    float m; memcpy(&m, &mb, 4);
    if (m > 1.41421354f) { m = m * 0.5f; e += 1; }
    const float f = m - 1.0f;
    static const float c[9] = {-0x1.4237fep-4f, 0x1.0696e4p-3f, -0x1.0c524cp-3f, 0x1.22973ap-3f, -0x1.548882p-3f,
                               0x1.99a3ecp-3f, -0x1.000206p-2f, 0x1.55554ep-2f, -0x1.fffffep-2f};
    float p = c[0];
    for (int i = 1; i < 9; ++i) p = fmaf(p, f, c[i]);
    const float ff = f * f;
    const float lm = fmaf(ff, p, f);
    return fmaf((float)e, 0.693147182f, lm);
}
float orc_gauss_from_u32(uint32_t k) {
    const uint32_t j = k & 0x7fffffffu;
    const float v = fmaf((float)j, 0x1p-31f, 0x1p-32f);
    const float two_minus_v = 2.0f - v;
    const float t = v * two_minus_v;
    const float x = 1.0f - v;
    float w = -log_pos_f32(t);
    float p;
    if (w < 5.0f) {
        static const float c[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f, 0.00021858087f,
                                   -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
        w = w - 2.5f;
        p = c[0];
        for (int i = 1; i < 9; ++i) p = fmaf(p, w, c[i]);
    } else {
        static const float c[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f, 0.00573950773f,
                                   -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
        w = sqrtf(w) - 3.0f;
        p = c[0];
        for (int i = 1; i < 9; ++i) p = fmaf(p, w, c[i]);
    }
    const float px = p * x;
    const float z = px * 1.41421354f;
    return (k >> 31) ? -z : z;
}
static double sample_from_u32(int dist, uint32_t k) {
    if (dist == 0) return (double)orc_gauss_from_u32(k);
    if (dist == 1) return ((double)k * 2.0 + 1.0) * 0x1p-32 - 1.0;
    return (k >> 31) ? -1.0 : 1.0;
}
static double omega_entry(int dist, uint64_t seed, uint32_t stream, uint64_t R, uint32_t c) {
    const uint64_t q = R >> 2;
    const uint32_t ctr[4] = {(uint32_t)q, (uint32_t)(q >> 32), c, stream};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    orc_philox4x32_10(ctr, key, o);
    return sample_from_u32(dist, o[R & 3]);
}
void orc_omega_fill(int dist, uint64_t seed, uint32_t stream, int64_t rows, int64_t cols, int64_t row_off, double* out, int64_t ld) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < cols; ++c)
        for (int64_t r = 0; r < rows; ++r) AT(out, ld, r, c) = omega_entry(dist, seed, stream, (uint64_t)(row_off + r), (uint32_t)c);
}

int orc_sketching_operator_ref(int dist, uint64_t seed, int64_t rows, int64_t cols, double* out) {
    /* src/sketch.rs:112-127: fresh ThreeFry2x64Rng::seed_from_u64(seed); DMatrix::from_fn fills column-major, one sample per entry */
    if (dist == 0) return -1;
    uint64_t key[2];
    orc_seed_from_u64(seed, key);
    const int64_t total = rows * cols;
    for (int64_t t = 0; t < total; ++t) {
        uint64_t ctr[2] = {(uint64_t)t >> 1, 0}, x[2];
        orc_threefry2x64_20(ctr, key, x);
        const uint64_t u = x[t & 1];
        double v;
        if (dist == 1) {
            /* rand 0.8.5 UniformFloat<f64>: value1_2 = (u >> 12) with exponent 0; (value1_2 - 1.0) * scale + low, scale = 2, low = -1 */
            const uint64_t b = (u >> 12) | 0x3FF0000000000000ull;
            double v12; memcpy(&v12, &b, 8);
            const double v01 = v12 - 1.0;
            const double sc = v01 * 2.0;
            v = sc + (-1.0);
        } else {
            v = (u < 0x8000000000000000ull) ? 1.0 : -1.0;    /* Bernoulli::new(0.5): p_int = 2^63, sample = u64 < p_int */
        }
        out[t] = v;   /* entry (t % rows, t / rows) */
    }
    return 0;
}

/* ======================================================================== L2: nalgebra-style dense kernels */
/* X.qr() -- Householder, nalgebra conventions (householder.rs reflection_axis_mut / clear_column_unchecked, qr.rs q() / r()) */
void orc_qr(const double* X, int64_t rows, int64_t cols, double* Q, double* R) {
    const int64_t p = imin(rows, cols);
    double* W = dalloc(rows * cols);
    double* diag = dalloc(p);
    memcpy(W, X, (size_t)(rows * cols) * sizeof(double));
    for (int64_t i = 0; i < p; ++i) {
        /* reflection_axis_mut on W[i.., i] */
        double sq = 0.0;
        for (int64_t r = i; r < rows; ++r) sq += AT(W, rows, r, i) * AT(W, rows, r, i);
        const double norm = sqrt(sq);
        const double x0 = AT(W, rows, i, i);
        const double modulus = fabs(x0), sign = (x0 >= 0.0) ? 1.0 : -1.0;
        const double signed_norm = sign * norm;
        const double factor = (sq + modulus * norm) * 2.0;
        AT(W, rows, i, i) = x0 + signed_norm;
        if (factor != 0.0) {
            const double fs = sqrt(factor);
            double nn = 0.0;
            for (int64_t r = i; r < rows; ++r) { AT(W, rows, r, i) /= fs; nn += AT(W, rows, r, i) * AT(W, rows, r, i); }
            nn = sqrt(nn);
            if (nn > 0.0) for (int64_t r = i; r < rows; ++r) AT(W, rows, r, i) /= nn;
            diag[i] = -signed_norm;
            /* reflect_with_sign on the trailing columns, sign = signum(diag) */
            const double sg = (diag[i] >= 0.0) ? 1.0 : -1.0;
#pragma omp parallel for schedule(static) if ((cols - i) * (rows - i) > 200000)
            for (int64_t c = i + 1; c < cols; ++c) {
                double dot = 0.0;
                for (int64_t r = i; r < rows; ++r) dot += AT(W, rows, r, i) * AT(W, rows, r, c);
                const double m_two = (-2.0 * sg) * dot;
                for (int64_t r = i; r < rows; ++r) AT(W, rows, r, c) = m_two * AT(W, rows, r, i) + sg * AT(W, rows, r, c);
            }
        } else {
            diag[i] = signed_norm;   /* zero column: no reflection */
            for (int64_t r = i; r < rows; ++r) AT(W, rows, r, i) = 0.0;   /* marks "identity reflector" for q() below */
        }
    }
    if (R) {
        for (int64_t c = 0; c < cols; ++c)
            for (int64_t r = 0; r < p; ++r) AT(R, p, r, c) = (r < c) ? AT(W, rows, r, c) : (r == c ? fabs(diag[r]) : 0.0);
    }
    if (Q) {
        for (int64_t c = 0; c < p; ++c)
            for (int64_t r = 0; r < rows; ++r) AT(Q, rows, r, c) = (r == c) ? 1.0 : 0.0;
        for (int64_t i = p - 1; i >= 0; --i) {
            double an = 0.0;
            for (int64_t r = i; r < rows; ++r) an += AT(W, rows, r, i) * AT(W, rows, r, i);
            if (an == 0.0) continue;
            const double sg = (diag[i] >= 0.0) ? 1.0 : -1.0;
#pragma omp parallel for schedule(static) if ((p - i) * (rows - i) > 200000)
            for (int64_t c = i; c < p; ++c) {
                double dot = 0.0;
                for (int64_t r = i; r < rows; ++r) dot += AT(W, rows, r, i) * AT(Q, rows, r, c);
                const double m_two = (-2.0 * sg) * dot;
                for (int64_t r = i; r < rows; ++r) AT(Q, rows, r, c) = m_two * AT(W, rows, r, i) + sg * AT(Q, rows, r, c);
            }
        }
    }
    free(W); free(diag);
}

/* X.full_piv_lu().l() -- nalgebra FullPivLU::new + lu::gauss_step(_swap), permutations dropped (lora_helpers.rs:144-146) */
void orc_stabilizer(const double* X, int64_t rows, int64_t cols, double* L) {
    const int64_t mn = imin(rows, cols);
    double* W = dalloc(rows * cols);
    memcpy(W, X, (size_t)(rows * cols) * sizeof(double));
    for (int64_t i = 0; i < mn; ++i) {
        /* icamax_full over W[i.., i..]: first maximum of |.| in column-major order */
        double best = -1.0; int64_t rp = i, cp = i;
        for (int64_t c = i; c < cols; ++c)
            for (int64_t r = i; r < rows; ++r) {
                const double v = fabs(AT(W, rows, r, c));
                if (v > best) { best = v; rp = r; cp = c; }
            }
        const double diag = AT(W, rows, rp, cp);
        if (diag == 0.0) break;
        if (cp != i) for (int64_t r = 0; r < rows; ++r) { const double t = AT(W, rows, r, i); AT(W, rows, r, i) = AT(W, rows, r, cp); AT(W, rows, r, cp) = t; }
        if (rp != i) for (int64_t c = 0; c < cols; ++c) { const double t = AT(W, rows, i, c); AT(W, rows, i, c) = AT(W, rows, rp, c); AT(W, rows, rp, c) = t; }
        const double inv = 1.0 / diag;
        for (int64_t r = i + 1; r < rows; ++r) AT(W, rows, r, i) = AT(W, rows, r, i) * inv;        /* coeffs *= inv_diag */
#pragma omp parallel for schedule(static) if ((cols - i) * (rows - i) > 200000)
        for (int64_t c = i + 1; c < cols; ++c) {
            const double a = -AT(W, rows, i, c);
            for (int64_t r = i + 1; r < rows; ++r) {
                const double prod = a * AT(W, rows, r, i);            /* axpy: y = a*x + y, separately rounded */
                AT(W, rows, r, c) = prod + AT(W, rows, r, c);
            }
        }
    }
    for (int64_t c = 0; c < mn; ++c)
        for (int64_t r = 0; r < rows; ++r) AT(L, rows, r, c) = (r > c) ? AT(W, rows, r, c) : (r == c ? 1.0 : 0.0);
    free(W);
}

int orc_cholesky_lower(const double* A, int64_t n, double* L) {
    memset(L, 0, (size_t)(n * n) * sizeof(double));
    for (int64_t j = 0; j < n; ++j) {
        double d = AT(A, n, j, j);
        for (int64_t k = 0; k < j; ++k) d -= AT(L, n, j, k) * AT(L, n, j, k);
        if (!(d > 0.0)) return -1;                     /* nalgebra cholesky(): None unless the pivot is positive */
        const double ljj = sqrt(d);
        AT(L, n, j, j) = ljj;
        for (int64_t i = j + 1; i < n; ++i) {
            double s = AT(A, n, i, j);
            for (int64_t k = 0; k < j; ++k) s -= AT(L, n, i, k) * AT(L, n, j, k);
            AT(L, n, i, j) = s / ljj;
        }
    }
    return 0;
}

/* one-sided Jacobi on the columns of G (rows x p): G <- G J, V <- V J until mutually orthogonal */
static int jacobi_cols(double* G, int64_t rows, int64_t p, double* V) {
    const double tol = sqrt((double)(rows > 1 ? rows : 1)) * DBL_EPSILON;
    for (int sweep = 0; sweep < 60; ++sweep) {
        int rot = 0;
        for (int64_t a = 0; a < p - 1; ++a)
            for (int64_t b = a + 1; b < p; ++b) {
                double al = 0, be = 0, ga = 0;
                for (int64_t r = 0; r < rows; ++r) { const double x = AT(G, rows, r, a), y = AT(G, rows, r, b); al += x * x; be += y * y; ga += x * y; }
                if (fabs(ga) > tol * sqrt(al * be) && fabs(ga) > DBL_MIN) {
                    rot = 1;
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                    for (int64_t r = 0; r < rows; ++r) { const double x = AT(G, rows, r, a), y = AT(G, rows, r, b); AT(G, rows, r, a) = cs * x - sn * y; AT(G, rows, r, b) = sn * x + cs * y; }
                    for (int64_t r = 0; r < p; ++r) { const double x = AT(V, p, r, a), y = AT(V, p, r, b); AT(V, p, r, a) = cs * x - sn * y; AT(V, p, r, b) = sn * x + cs * y; }
                }
            }
        if (!rot) return 0;
    }
    return -1;
}

static int cmp_desc(const void* a, const void* b) {
    const double x = ((const double*)a)[0], y = ((const double*)b)[0];
    if (x > y) return -1; if (x < y) return 1;
    const double i = ((const double*)a)[1], j = ((const double*)b)[1];
    return (i > j) - (i < j);
}

/* SVD of a tall T (rows x p, rows >= p): T = U diag(s) V^T.  Zero singular values get a unit-vector completion of U
 * (so that SVD(0) has identity factors, as the reference's tests on the zero matrix expect: lora_drivers.rs:341-357). */
static int svd_tall(const double* T, int64_t rows, int64_t p, double* U, double* s, double* V) {
    /* QR first (as one would for a tall matrix), then Jacobi on R^T ... keep it simple: Jacobi directly on T's columns */
    double* G = dalloc(rows * p); double* Vw = dalloc(p * p); double* key = dalloc(2 * p);
    memcpy(G, T, (size_t)(rows * p) * sizeof(double));
    for (int64_t i = 0; i < p; ++i) AT(Vw, p, i, i) = 1.0;
    const int rc = jacobi_cols(G, rows, p, Vw);
    for (int64_t j = 0; j < p; ++j) {
        double n2 = 0; for (int64_t r = 0; r < rows; ++r) n2 += AT(G, rows, r, j) * AT(G, rows, r, j);
        key[2 * j] = sqrt(n2); key[2 * j + 1] = (double)j;
    }
    qsort(key, (size_t)p, 2 * sizeof(double), cmp_desc);
    for (int64_t k = 0; k < p; ++k) {
        const int64_t j = (int64_t)key[2 * k + 1];
        s[k] = key[2 * k];
        for (int64_t r = 0; r < rows; ++r) AT(U, rows, r, k) = s[k] > 0.0 ? AT(G, rows, r, j) / s[k] : 0.0;
        for (int64_t r = 0; r < p; ++r) AT(V, p, r, k) = AT(Vw, p, r, j);
    }
    int64_t cand = 0;
    for (int64_t k = 0; k < p; ++k) {
        if (s[k] > 0.0) continue;
        for (; cand < rows; ++cand) {
            for (int64_t r = 0; r < rows; ++r) AT(U, rows, r, k) = (r == cand) ? 1.0 : 0.0;
            for (int pass = 0; pass < 2; ++pass)
                for (int64_t i = 0; i < p; ++i) {
                    if (i == k || (i > k && !(s[i] > 0.0))) continue;
                    double d = 0; for (int64_t r = 0; r < rows; ++r) d += AT(U, rows, r, i) * AT(U, rows, r, k);
                    for (int64_t r = 0; r < rows; ++r) AT(U, rows, r, k) -= d * AT(U, rows, r, i);
                }
            double nn = 0; for (int64_t r = 0; r < rows; ++r) nn += AT(U, rows, r, k) * AT(U, rows, r, k);
            if (nn > 0.25) { nn = sqrt(nn); for (int64_t r = 0; r < rows; ++r) AT(U, rows, r, k) /= nn; ++cand; break; }
        }
    }
    free(G); free(Vw); free(key);
    return rc;
}

/* tall matrices are first reduced by a Householder QR (T = Q R), then SVD(R): O(rows p^2) instead of Jacobi sweeps over all rows */
static int svd_tall_qr(const double* T, int64_t rows, int64_t p, double* U, double* s, double* V) {
    if (rows < 2 * p) return svd_tall(T, rows, p, U, s, V);
    double* Q = dalloc(rows * p); double* R = dalloc(p * p); double* Ur = dalloc(p * p);
    orc_qr(T, rows, p, Q, R);
    const int rc = svd_tall(R, p, p, Ur, s, V);
    orc_gemm_nn(Q, rows, rows, p, Ur, p, p, U, rows);
    free(Q); free(R); free(Ur);
    return rc;
}

int orc_svd(const double* M, int64_t rows, int64_t cols, double* U, double* sigma, double* Vt) {
    const int64_t p = imin(rows, cols);
    int rc;
    if (rows >= cols) {
        double* V = dalloc(p * p);
        rc = svd_tall_qr(M, rows, p, U, sigma, V);
        for (int64_t i = 0; i < p; ++i) for (int64_t j = 0; j < cols; ++j) AT(Vt, p, i, j) = AT(V, p, j, i);
        free(V);
    } else {
        /* M^T = U' s V'^T  ->  M = V' s U'^T */
        double* Mt = dalloc(rows * cols); double* Up = dalloc(cols * p); double* Vp = dalloc(p * p);
        for (int64_t i = 0; i < rows; ++i) for (int64_t j = 0; j < cols; ++j) AT(Mt, cols, j, i) = AT(M, rows, i, j);
        rc = svd_tall_qr(Mt, cols, p, Up, sigma, Vp);
        for (int64_t i = 0; i < rows; ++i) for (int64_t k = 0; k < p; ++k) AT(U, rows, i, k) = AT(Vp, p, i, k);
        for (int64_t k = 0; k < p; ++k) for (int64_t j = 0; j < cols; ++j) AT(Vt, p, k, j) = AT(Up, cols, j, k);
        free(Mt); free(Up); free(Vp);
    }
    return rc;
}

int orc_symmetric_eigen(const double* A, int64_t n, double* W, double* lambda) {
    /* cyclic two-sided Jacobi */
    double* S = dalloc(n * n); double* V = dalloc(n * n); double* key = dalloc(2 * n);
    double fro = 0;
    for (int64_t j = 0; j < n; ++j) for (int64_t i = 0; i < n; ++i) { AT(S, n, i, j) = 0.5 * (AT(A, n, i, j) + AT(A, n, j, i)); fro += AT(S, n, i, j) * AT(S, n, i, j); }
    for (int64_t i = 0; i < n; ++i) AT(V, n, i, i) = 1.0;
    const double afloor = 1e-20 * sqrt(fro);
    int rc = -1;
    for (int sweep = 0; sweep < 60; ++sweep) {
        int rot = 0;
        for (int64_t a = 0; a < n - 1; ++a)
            for (int64_t b = a + 1; b < n; ++b) {
                const double aa = AT(S, n, a, a), bb = AT(S, n, b, b), ab = AT(S, n, a, b);
                if (fabs(ab) > DBL_EPSILON * sqrt(fabs(aa * bb)) && fabs(ab) > afloor) {
                    rot = 1;
                    const double tau = (bb - aa) / (2.0 * ab);
                    const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    const double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
                    for (int64_t k = 0; k < n; ++k) { const double x = AT(S, n, k, a), y = AT(S, n, k, b); AT(S, n, k, a) = cs * x - sn * y; AT(S, n, k, b) = sn * x + cs * y; }
                    for (int64_t k = 0; k < n; ++k) { const double x = AT(S, n, a, k), y = AT(S, n, b, k); AT(S, n, a, k) = cs * x - sn * y; AT(S, n, b, k) = sn * x + cs * y; }
                    for (int64_t k = 0; k < n; ++k) { const double x = AT(V, n, k, a), y = AT(V, n, k, b); AT(V, n, k, a) = cs * x - sn * y; AT(V, n, k, b) = sn * x + cs * y; }
                }
            }
        if (!rot) { rc = 0; break; }
    }
    for (int64_t j = 0; j < n; ++j) { key[2 * j] = -AT(S, n, j, j); key[2 * j + 1] = (double)j; }
    qsort(key, (size_t)n, 2 * sizeof(double), cmp_desc);      /* descending in -lambda = ascending in lambda */
    for (int64_t k = 0; k < n; ++k) {
        const int64_t j = (int64_t)key[2 * k + 1];
        lambda[k] = AT(S, n, j, j);
        for (int64_t r = 0; r < n; ++r) AT(W, n, r, k) = AT(V, n, r, j);
    }
    free(S); free(V); free(key);
    return rc;
}

/* ======================================================================== L3: lora_helpers */
static void draw(const orc_opts* o, int which /* 1: n x k, 2: m x k */, int64_t rows, int64_t k, double* out) {
    const double* given = which == 1 ? o->omega_n : o->omega_m;
    if (given) memcpy(out, given, (size_t)(rows * k) * sizeof(double));
    else orc_omega_fill(o->dist, o->seed, (uint32_t)which, rows, k, 0, out, rows);
}
static void stab(int mode, double* X, int64_t rows, int64_t cols) {
    /* literal: Stabilizer = L factor (lora_helpers.rs:144-146); intended: any range-preserving conditioner -> thin Q */
    double* T = dalloc(rows * imin(rows, cols));
    if (mode == 1) orc_stabilizer(X, rows, cols, T); else orc_qr(X, rows, cols, T, NULL);
    memcpy(X, T, (size_t)(rows * imin(rows, cols)) * sizeof(double));
    free(T);
}

void orc_tsog1(const double* A, int64_t m, int64_t n, int64_t k, int num_passes, int passes_per_stab, const orc_opts* o, double* S) {
    int done = 0;
    double* tall = dalloc(m * k);
    if (o->mode == 1) {
        /* ---- literal: src/lora_helpers.rs:58-105, statement for statement ---- */
        double* S1 = dalloc(n * k);                                   /* :63 zeros */
        memset(S, 0, (size_t)(n * k) * sizeof(double));               /* :67 zeros */
        if (num_passes % 2 == 0) {
            draw(o, 1, n, k, S);                                      /* :71 (never read again if the loop runs) */
        } else {
            draw(o, 2, m, k, tall);                                   /* :74 */
            orc_gemm_tn(A, m, m, n, tall, m, k, S1, n);               /* :76 S1 = A^T S1 */
            done += 1;                                                /* :77; :78-81 result `_S2` discarded */
        }
        int diff = num_passes - done;                                 /* :85 */
        while (diff >= 2) {                                           /* :88 */
            orc_gemm_nn(A, m, m, n, S1, n, k, tall, m);               /* :89 S = A * S1  (S1, not S) */
            done += 1;
            if (done % passes_per_stab == 0) stab(1, tall, m, k);     /* :91-94 */
            orc_gemm_tn(A, m, m, n, tall, m, k, S, n);                /* :95 */
            done += 1;
            if (done % passes_per_stab == 0) stab(1, S, n, k);        /* :97-100 */
            diff -= 2;                                                /* :101 */
        }
        free(S1);
    } else {
        /* ---- intended: Murray et al. 2023, TSOG1 ---- */
        if (num_passes % 2 == 0) {
            draw(o, 1, n, k, S);
        } else {
            draw(o, 2, m, k, tall);
            orc_gemm_tn(A, m, m, n, tall, m, k, S, n);
            done = 1;
            if (done % passes_per_stab == 0) stab(0, S, n, k);
        }
        while (num_passes - done >= 2) {
            orc_gemm_nn(A, m, m, n, S, n, k, tall, m); done += 1;
            if (done % passes_per_stab == 0) stab(0, tall, m, k);
            orc_gemm_tn(A, m, m, n, tall, m, k, S, n); done += 1;
            if (done % passes_per_stab == 0) stab(0, S, n, k);
        }
    }
    free(tall);
}

static int eff_q(const orc_opts* o, int dflt) { return o->num_passes > 0 ? o->num_passes : dflt; }
static int eff_pps(const orc_opts* o) { return o->passes_per_stab > 0 ? o->passes_per_stab : 1; }

void orc_rf1(const double* A, int64_t m, int64_t n, int64_t k, const orc_opts* o, double* Q) {
    /* src/lora_helpers.rs:37-44: S = tsog1(A, k, 2, 1); Y = A*S; Q = Orth(Y) */
    double* S = dalloc(n * k); double* Y = dalloc(m * k);
    orc_tsog1(A, m, n, k, eff_q(o, 2), eff_pps(o), o, S);
    orc_gemm_nn(A, m, m, n, S, n, k, Y, m);
    orc_qr(Y, m, k, Q, NULL);
    free(S); free(Y);
}
void orc_qb1(const double* A, int64_t m, int64_t n, int64_t k, const orc_opts* o, double* Q, double* B) {
    /* src/lora_helpers.rs:17-23: Q = RF1(A, k); B = Q^T A */
    orc_rf1(A, m, n, k, o, Q);
    double* Bt = dalloc(n * k);
    orc_gemm_tn(A, m, m, n, Q, m, k, Bt, n);
    for (int64_t i = 0; i < k; ++i) for (int64_t j = 0; j < n; ++j) AT(B, k, i, j) = AT(Bt, n, j, i);
    free(Bt);
}

/* ======================================================================== L4: lora_drivers */
int orc_rand_svd(const double* A, int64_t m, int64_t n, int64_t k, double epsilon, int64_t s, const orc_opts* o,
                 double* U, double* S, double* Vt, int64_t* r_out) {
    if (k <= 0) return 1;                 /* :31-35 InvalidParameters */
    if (!(epsilon > 0.0)) return 1;       /* :36-40 */
    if (s <= 0) return 1;                 /* :41-45 */
    const int64_t l = imin(k + s, imin(m, n));       /* thin factors of nalgebra: Q.ncols() = min(k+s, m, n) */
    const int64_t r = imin(k, l);                    /* :51 */
    double* Q = dalloc(m * l); double* B = dalloc(l * n);
    orc_qb1(A, m, n, l, o, Q, B);                    /* :49 */
    double* Ub = dalloc(l * l); double* sg = dalloc(l); double* Vtb = dalloc(l * n);
    if (orc_svd(B, l, n, Ub, sg, Vtb) != 0) { free(Q); free(B); free(Ub); free(sg); free(Vtb); return 7; }   /* :53-58 */
    orc_gemm_nn(Q, m, m, l, Ub, l, r, U, m);         /* :60,66 U_final = Q * U[:, :r] */
    memset(S, 0, (size_t)(r * r) * sizeof(double));
    for (int64_t i = 0; i < r; ++i) AT(S, r, i, i) = sg[i];                                    /* :64 */
    for (int64_t i = 0; i < r; ++i) for (int64_t j = 0; j < n; ++j) AT(Vt, r, i, j) = AT(Vtb, l, i, j);   /* :62,68 */
    if (r_out) *r_out = r;
    free(Q); free(B); free(Ub); free(sg); free(Vtb);
    return 0;
}

static int cmp_abs_desc(const void* a, const void* b) {
    const double x = fabs(((const double*)a)[0]), y = fabs(((const double*)b)[0]);
    if (x > y) return -1; if (x < y) return 1;
    const double i = ((const double*)a)[1], j = ((const double*)b)[1];
    return (i > j) - (i < j);
}

int orc_rand_evd1(const double* A, int64_t n, int64_t k, double epsilon, int64_t s, const orc_opts* o, double* V, double* lambda, int64_t* r_out) {
    if (k <= 0 || !(epsilon > 0.0) || s <= 0) return 1;                    /* :89-103 */
    for (int64_t j = 0; j < n; ++j) for (int64_t i = 0; i < j; ++i) if (AT(A, n, i, j) != AT(A, n, j, i)) return 8;   /* :106-110 */
    const int64_t l = imin(k + s, n);
    double* Q = dalloc(n * l); double* B = dalloc(l * n); double* C = dalloc(l * l);
    orc_qb1(A, n, n, l, o, Q, B);                                          /* :114 */
    orc_gemm_nn(B, l, l, n, Q, n, l, C, l);                                /* :121 C = B * Q */
    double* W = dalloc(l * l); double* lam = dalloc(l); double* key = dalloc(2 * l);
    const int rc = orc_symmetric_eigen(C, l, W, lam);                      /* :129 */
    for (int64_t i = 0; i < l; ++i) { key[2 * i] = lam[i]; key[2 * i + 1] = (double)i; }
    qsort(key, (size_t)l, 2 * sizeof(double), cmp_abs_desc);               /* :134-138 */
    const int64_t r = imin(k, l);                                          /* :140 */
    double* Usel = dalloc(l * r);
    for (int64_t t = 0; t < r; ++t) {
        const int64_t i = (int64_t)key[2 * t + 1];
        lambda[t] = lam[i];                                                /* :144 */
        for (int64_t q = 0; q < l; ++q) AT(Usel, l, q, t) = AT(W, l, q, i);  /* :146 */
    }
    orc_gemm_nn(Q, n, n, l, Usel, l, r, V, n);                             /* :148 */
    if (r_out) *r_out = r;
    free(Q); free(B); free(C); free(W); free(lam); free(key); free(Usel);
    return rc ? 10 : 0;
}

int orc_rand_evd2(const double* A, int64_t n, int64_t k, int64_t s, const orc_opts* o, double* V, double* lambda, int64_t* r_out) {
    if (k <= 0) return 1;                                                  /* :169-173 */
    if (r_out) *r_out = 0;
    if (!o || !o->skip_psd_check) {   /* :178-184 PSD check through a full symmetric eigen-decomposition */
        double* W = dalloc(n * n); double* lam = dalloc(n);
        orc_symmetric_eigen(A, n, W, lam);
        int neg = 0; double amax = 0;
        for (int64_t i = 0; i < n; ++i) amax = fmax(amax, fabs(lam[i]));
        /* the reference tests `x < 0.0` on nalgebra's computed eigenvalues; rounding noise of a different eigen-solver
         * cannot be reproduced, so values within noise of zero are treated as zero */
        for (int64_t i = 0; i < n; ++i) if (lam[i] < -1e-12 * fmax(amax, 1e-300) * (double)n) neg = 1;
        free(W); free(lam);
        if (neg) return 9;
    }
    const int64_t l = imin(k + s, n);
    double* S = dalloc(n * l); double* Y = dalloc(n * l); double* SY = dalloc(l * l); double* L = dalloc(l * l);
    orc_tsog1(A, n, n, l, eff_q(o, 3), eff_pps(o), o, S);                  /* :186 */
    orc_gemm_nn(A, n, n, n, S, n, l, Y, n);                                /* :187 */
    double ss = 0; for (int64_t i = 0; i < n * l; ++i) ss += Y[i] * Y[i];
    const double nu = sqrt((double)n) * DBL_EPSILON * sqrt(ss);            /* :188-189 */
    for (int64_t i = 0; i < n * l; ++i) Y[i] = Y[i] + nu * S[i];           /* :190 */
    orc_gemm_tn(S, n, n, l, Y, n, l, SY, l);                               /* :191 */
    if (orc_cholesky_lower(SY, l, L) != 0) { free(S); free(Y); free(SY); free(L); return 7; }   /* :193-198 */
    /* R = L^T; B = Y_new R^-1  (:200-201): solve B R = Y, R upper */
    double* B = dalloc(n * l);
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < l; ++j) {
            double v = AT(Y, n, i, j);
            for (int64_t t = 0; t < j; ++t) v -= AT(B, n, i, t) * AT(L, l, j, t);    /* R[t][j] = L[j][t] */
            AT(B, n, i, j) = v / AT(L, l, j, j);
        }
    double* Ub = dalloc(n * l); double* sg = dalloc(l); double* Vtb = dalloc(l * l);
    if (orc_svd(B, n, l, Ub, sg, Vtb) != 0) { free(S); free(Y); free(SY); free(L); free(B); free(Ub); free(sg); free(Vtb); return 7; }   /* :208-213 */
    int64_t nl = 0, cnt = 0;
    double* lam = dalloc(l);
    for (int64_t i = 0; i < l; ++i) if (sg[i] > 0.0) lam[nl++] = sg[i] * sg[i];      /* :217 */
    for (int64_t i = 0; i < nl; ++i) if (lam[i] > nu) ++cnt;
    const int64_t r = imin(k, cnt);                                                   /* :219 */
    for (int64_t i = 0; i < r; ++i) lambda[i] = lam[i] - nu;                          /* :220 */
    for (int64_t j = 0; j < r; ++j) for (int64_t i = 0; i < n; ++i) AT(V, n, i, j) = AT(Ub, n, i, j);   /* :221 */
    if (r_out) *r_out = r;
    free(S); free(Y); free(SY); free(L); free(B); free(Ub); free(sg); free(Vtb); free(lam);
    return 0;
}

/* ======================================================================== haar_sample  (src/sketch.rs:45-85) */
/* rows x cols matrix with orthonormal rows (attr 0 = Row, rows <= cols) or columns (attr 1 = Column, cols <= rows): Q factor of the
 * Householder QR of an m x n Gaussian matrix (m >= n; :68-73), each column multiplied by signum(R_ii) (:74-80), transposed for Row
 * (:81-84).  The Gaussian entries are Omega(seed, stream 0) of this build (the reference's ziggurat values are parity-unpinned,
 * oracle.h); column-major fill as DMatrix::from_vec (:72).  Returns 2 (InvalidDimensions) as :49-63 do. */
int orc_haar_sample(int64_t rows, int64_t cols, int attr, uint64_t seed, double* out) {
    int64_t m, n;
    if (attr == 0) { if (rows > cols) return 2; m = cols; n = rows; }          /* :48-56 */
    else { if (cols > rows) return 2; m = rows; n = cols; }                     /* :57-65 */
    if (m <= 0 || n <= 0) return 2;
    double* G = dalloc(m * n); double* Q = dalloc(m * n); double* R = dalloc(n * n);
    orc_omega_fill(0, seed, 0u, m, n, 0, G, m);                                 /* :68-72 */
    orc_qr(G, m, n, Q, R);                                                      /* :73 */
    for (int64_t j = 0; j < n; ++j) {                                           /* :74-80 */
        const double d = AT(R, n, j, j);
        const double sg = (d > 0.0) ? 1.0 : (d < 0.0 ? -1.0 : (d == 0.0 ? (signbit(d) ? -1.0 : 1.0) : d));   /* f64::signum */
        for (int64_t i = 0; i < m; ++i) AT(Q, m, i, j) *= sg;
    }
    if (attr == 0) { for (int64_t j = 0; j < n; ++j) for (int64_t i = 0; i < m; ++i) AT(out, n, j, i) = AT(Q, m, i, j); }   /* :82 */
    else memcpy(out, Q, (size_t)(m * n) * sizeof(double));                      /* :83 */
    free(G); free(Q); free(R);
    return 0;
}

/* ======================================================================== sketch step */
int64_t orc_sketch_dim(int64_t m, int64_t n, double sf, int rule) {
    if (rule == 0) return (sf * (double)n > (double)m) ? m : (int64_t)floor(sf * (double)n);   /* sketch_and_precondition.rs:49,105 */
    int64_t d = (int64_t)floor(sf * (double)n); if (d < 1) d = 1; if (d > m) d = m; return d;  /* :172 */
}
void orc_sketch_apply_dense(int dist, uint64_t seed, int64_t d, const double* A, int64_t m, int64_t n, double* A_sk) {
    double* St = dalloc(m * d);
    orc_omega_fill(dist, seed, 3u, m, d, 0, St, m);     /* S^T(j, i) = omega(row j, col i) */
    orc_gemm_tn(St, m, m, d, A, m, n, A_sk, d);         /* S A = (S^T)^T A   (:51-52, :107, :176) */
    free(St);
}
void orc_sketch_apply_saso(uint64_t seed, int64_t d, int zeta, const double* A, int64_t m, int64_t n, double* A_sk) {
    memset(A_sk, 0, (size_t)(d * n) * sizeof(double));
    const double scale = 1.0 / sqrt((double)zeta);
    for (int64_t j = 0; j < m; ++j) {
        int64_t idx[8]; double sg[8];
        const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        if (d <= 32768) {
            const uint32_t ctr[4] = {(uint32_t)j, (uint32_t)((uint64_t)j >> 32), 0u, 4u};
            uint32_t w[4]; orc_philox4x32_10(ctr, key, w);
            for (int t = 0; t < zeta; ++t) {
                const uint32_t f = (w[t >> 1] >> ((t & 1) * 16)) & 0xffffu;
                idx[t] = (int64_t)(((uint64_t)(f >> 1) * (uint64_t)d) >> 15);
                sg[t] = (f & 1u) ? -1.0 : 1.0;
            }
        } else {
            for (int blk = 0; blk * 4 < zeta; ++blk) {
                const uint32_t ctr[4] = {(uint32_t)j, (uint32_t)((uint64_t)j >> 32), (uint32_t)blk, 4u};
                uint32_t w[4]; orc_philox4x32_10(ctr, key, w);
                for (int e = 0; e < 4 && blk * 4 + e < zeta; ++e) {
                    idx[blk * 4 + e] = (int64_t)(((uint64_t)(w[e] >> 1) * (uint64_t)d) >> 31);
                    sg[blk * 4 + e] = (w[e] & 1u) ? -1.0 : 1.0;
                }
            }
        }
        for (int64_t c = 0; c < n; ++c) {
            const double v = AT(A, m, j, c);
            for (int t = 0; t < zeta; ++t) AT(A_sk, d, idx[t], c) += sg[t] * v;
        }
    }
    for (int64_t i = 0; i < d * n; ++i) A_sk[i] *= scale;
}

/* Block sparse-sign operator (kind RNLA_SKETCH_SASO_BLOCK; not in the reference, SURVEY.md Appendix A.9 -- the
 * operator is defined by this build, DESIGN.md section 5, and restated here independently of the CUDA source).
 * zeta = g * w non-zeros per column: g groups ("stripes") of w consecutive rows each.  Stripe t owns output rows
 * [t*nbs*w, (t+1)*nbs*w), nbs = d / (g*w) blocks of width w.  Global row gr: chunk q = gr / 2048, x = gr % 2048.
 * Two Philox blocks per (chunk, stripe), ctr = (q_lo, q_hi, 2t, 5) -> k and (q_lo, q_hi, 2t+1, 5) -> k2, key a
 * bijection sigma on [0, 2048): with y = 16 yh + yl,  sigma(y) = 16 tau(yh) + (mo * yl + rho(yh)) mod 16, where tau is
 * three multiply-xorshift rounds on 7 bits keyed by k0..k2, mo = (k2[0] & 15) | 1, rho(yh) = (((yh+1) * (k2[1]|1)) >> 11) & 15
 * (sixteen consecutive slots hit sixteen different residues mod 16: shared-memory banks on the device).  The block
 * offset is off = 16 * ((k3 % nbs) / 16).  Slot y holds row x = sigma(y) and feeds block (y + off) % nbs of stripe t;
 * the w entries are rows (t*nbs + block)*w + r with sign bit (t*w + r) of word 0 of philox(ctr = (gr_lo, gr_hi, 1, 5));
 * value 1/sqrt(zeta).
 * Rows of A are global rows row_off .. row_off + m - 1 (so a row shard reproduces its slice of the operator). */
void orc_sketch_apply_saso_block(uint64_t seed, int64_t d, int zeta, int w, const double* A, int64_t m, int64_t n,
                                 int64_t row_off, double* A_sk) {
    memset(A_sk, 0, (size_t)(d * n) * sizeof(double));
    if (m <= 0) return;
    const int g = zeta / w;
    const int64_t R = 2048, nbs = d / zeta;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const double scale = 1.0 / sqrt((double)zeta);
    for (int64_t q = row_off / R; q <= (row_off + m - 1) / R; ++q)
        for (int t = 0; t < g; ++t) {
            const uint32_t ctr[4] = {(uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)(2 * t), 5u};
            const uint32_t ctr2[4] = {(uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)(2 * t + 1), 5u};
            uint32_t k[4], k2[4]; orc_philox4x32_10(ctr, key, k); orc_philox4x32_10(ctr2, key, k2);
            const int64_t off = (int64_t)(((k[3] % (uint32_t)nbs) >> 4) << 4);
            const uint32_t mo = (k2[0] & 15u) | 1u, rk = k2[1] | 1u;
            for (int64_t y = 0; y < R; ++y) {
                const uint32_t yh = (uint32_t)y >> 4, yl = (uint32_t)y & 15u;
                uint32_t h = yh;
                h = (h * (k[0] | 1u) + (k[0] >> 16)) & 127u; h ^= h >> 3;
                h = (h * (k[1] | 1u) + (k[1] >> 16)) & 127u; h ^= h >> 4;
                h = (h * (k[2] | 1u) + (k[2] >> 16)) & 127u; h ^= h >> 2;
                const uint32_t rho = (((yh + 1u) * rk) >> 11) & 15u;
                const uint32_t x = (h << 4) | ((yl * mo + rho) & 15u);
                const int64_t gr = q * R + (int64_t)x;
                if (gr < row_off || gr >= row_off + m) continue;
                const int64_t blk = (int64_t)t * nbs + (y + off) % nbs;
                const uint32_t cs[4] = {(uint32_t)gr, (uint32_t)((uint64_t)gr >> 32), 1u, 5u};
                uint32_t sw[4]; orc_philox4x32_10(cs, key, sw);
                for (int64_t c = 0; c < n; ++c) {
                    const double v = AT(A, m, gr - row_off, c);
                    for (int r = 0; r < w; ++r) AT(A_sk, d, blk * w + r, c) += ((sw[0] >> (t * w + r)) & 1u) ? -v : v;
                }
            }
        }
    for (int64_t i = 0; i < d * n; ++i) A_sk[i] *= scale;
}

/* ======================================================================== blendenpik + cgls
 * src/cg.rs:18-61, statement for statement on a dense a (m x n) */
int64_t orc_cgls(const double* a, int64_t m, int64_t n, const double* b, double tolerance, int64_t num_iterations,
                 double* x /* in: initial guess, out: solution */, int* converged_out) {
    double* r = dalloc(m); double* s = dalloc(n); double* p = dalloc(n); double* ap = dalloc(m); double* sn = dalloc(n);
    orc_gemm_nn(a, m, m, n, x, n, 1, ap, m);
    for (int64_t i = 0; i < m; ++i) r[i] = b[i] - ap[i];                                  /* :30 */
    orc_gemm_tn(a, m, m, n, r, m, 1, s, n);                                               /* :31 */
    memcpy(p, s, (size_t)n * sizeof(double));                                             /* :32 */
    double norm_s = 0.0; for (int64_t i = 0; i < n; ++i) norm_s += s[i] * s[i];           /* :33 */
    int converged = 0; int64_t it = 0;
    for (it = 0; it < num_iterations; ++it) {                                             /* :35 */
        orc_gemm_nn(a, m, m, n, p, n, 1, ap, m);                                          /* :36 */
        double apap = 0.0; for (int64_t i = 0; i < m; ++i) apap += ap[i] * ap[i];
        const double alpha = norm_s / apap;                                               /* :37 */
        for (int64_t i = 0; i < n; ++i) x[i] += alpha * p[i];                             /* :38 */
        for (int64_t i = 0; i < m; ++i) r[i] -= alpha * ap[i];                            /* :39 */
        orc_gemm_tn(a, m, m, n, r, m, 1, sn, n);                                          /* :40 */
        double nn = 0.0; for (int64_t i = 0; i < n; ++i) nn += sn[i] * sn[i];             /* :41 */
        if (sqrt(nn) < tolerance) { converged = 1; ++it; break; }                         /* :44-48 */
        const double beta = nn / norm_s;                                                  /* :50 */
        norm_s = nn;                                                                      /* :51 */
        for (int64_t i = 0; i < n; ++i) p[i] = sn[i] + beta * p[i];                       /* :52 */
    }
    if (converged_out) *converged_out = converged;
    free(r); free(s); free(p); free(ap); free(sn);
    return it;
}

/* src/sketch_and_precondition.rs:26-59; the sketch operator is this build's (kind 0 dense / 1 SASO / 2 block SASO, the
 * reference draws a dense Gaussian from its own stream); everything after the sketch follows the
 * reference literally, including the dense product A R^-1 (:56).  Returns 0, or 4/1 for the validation errors (:29-48),
 * or 6 if R is singular (the reference unwraps :55). */
int orc_blendenpik(const double* A, int64_t m, int64_t n, const double* b, double epsilon, int64_t l, double sampling_factor,
                   int kind, int dist_or_width, int zeta, uint64_t seed, double* x, int64_t* iters_out, int* converged_out) {
    if (m < n) return 4;                                                                  /* :29-33 */
    if (sampling_factor < 1.0 || epsilon <= 0.0 || l <= 0) return 1;                      /* :34-48 */
    const int64_t d = orc_sketch_dim(m, n, sampling_factor, 0);                           /* :49 */
    double* Ask = dalloc(d * n); double* bsk = dalloc(d);
    if (kind == 0) { orc_sketch_apply_dense(dist_or_width, seed, d, A, m, n, Ask); orc_sketch_apply_dense(dist_or_width, seed, d, b, m, 1, bsk); }
    else if (kind == 1) { orc_sketch_apply_saso(seed, d, zeta, A, m, n, Ask); orc_sketch_apply_saso(seed, d, zeta, b, m, 1, bsk); }
    else {
        const int w = dist_or_width ? dist_or_width : (zeta < 4 ? zeta : 4);
        orc_sketch_apply_saso_block(seed, d, zeta, w, A, m, n, 0, Ask); orc_sketch_apply_saso_block(seed, d, zeta, w, b, m, 1, 0, bsk);
    }
    double* Q = dalloc(d * n); double* R = dalloc(n * n);
    orc_qr(Ask, d, n, Q, R);                                                              /* :53 */
    double* z = dalloc(n);
    orc_gemm_tn(Q, d, d, n, bsk, d, 1, z, n);                                             /* :54 */
    double* Rinv = dalloc(n * n);                                                         /* :55 solve_upper_triangular(I) */
    int singular = 0;
    for (int64_t j = 0; j < n && !singular; ++j) {
        for (int64_t i = j; i >= 0; --i) {
            double sum = (i == j) ? 1.0 : 0.0;
            for (int64_t k = i + 1; k <= j; ++k) sum -= AT(R, n, i, k) * AT(Rinv, n, k, j);
            const double dii = AT(R, n, i, i);
            if (dii == 0.0) { singular = 1; break; }
            AT(Rinv, n, i, j) = sum / dii;
        }
    }
    int rc = 0;
    if (singular) rc = 6;
    else {
        double* Ap = dalloc(m * n);
        orc_gemm_nn(A, m, m, n, Rinv, n, n, Ap, m);                                       /* :56 */
        int conv = 0;
        const int64_t it = orc_cgls(Ap, m, n, b, epsilon, l, z, &conv);                   /* :57 */
        orc_gemm_nn(Rinv, n, n, n, z, n, 1, x, n);                                        /* :58 */
        if (iters_out) *iters_out = it;
        if (converged_out) *converged_out = conv;
        free(Ap);
    }
    free(Ask); free(bsk); free(Q); free(R); free(z); free(Rinv);
    return rc;
}

/* src/sketch_and_precondition.rs:82-119 (sketch operator as in orc_blendenpik).  Returns 0, 4/1 for the validation errors
 * (:85-104), 7 if the SVD fails. */
int orc_lsrn(const double* A, int64_t m, int64_t n, const double* b, double epsilon, int64_t l, double sampling_factor,
             int kind, int dist_or_width, int zeta, uint64_t seed, double* x, int64_t* iters_out, int* converged_out) {
    if (m < n) return 4;                                                                  /* :85-89 */
    if (sampling_factor < 1.0 || epsilon <= 0.0 || l <= 0) return 1;                      /* :90-104 */
    const int64_t d = orc_sketch_dim(m, n, sampling_factor, 0);                           /* :105 */
    double* Ask = dalloc(d * n);
    if (kind == 0) orc_sketch_apply_dense(dist_or_width, seed, d, A, m, n, Ask);          /* :106-107 */
    else if (kind == 1) orc_sketch_apply_saso(seed, d, zeta, A, m, n, Ask);
    else orc_sketch_apply_saso_block(seed, d, zeta, dist_or_width ? dist_or_width : (zeta < 4 ? zeta : 4), A, m, n, 0, Ask);
    const int64_t r = imin(d, n);
    double* U = dalloc(d * r); double* sg = dalloc(r); double* Vt = dalloc(r * n);
    int rc = 0;
    if (orc_svd(Ask, d, n, U, sg, Vt) != 0) rc = 7;                                       /* :109 svd(false, true) */
    else {
        double* N = dalloc(n * r);                                                        /* :110-113 n = v * sigma_inv */
        for (int64_t j = 0; j < r; ++j)
            for (int64_t i = 0; i < n; ++i) AT(N, n, i, j) = sg[j] != 0.0 ? AT(Vt, r, j, i) / sg[j] : 0.0;
        double* Ap = dalloc(m * r);
        orc_gemm_nn(A, m, m, n, N, n, r, Ap, m);                                          /* :114 */
        double* y = dalloc(r);                                                            /* :115 zeros */
        int conv = 0;
        const int64_t it = orc_cgls(Ap, m, r, b, epsilon, l, y, &conv);                   /* :116 */
        orc_gemm_nn(N, n, n, r, y, r, 1, x, n);                                           /* :117 */
        if (iters_out) *iters_out = it;
        if (converged_out) *converged_out = conv;
        free(N); free(Ap); free(y);
    }
    free(Ask); free(U); free(sg); free(Vt);
    return rc;
}
