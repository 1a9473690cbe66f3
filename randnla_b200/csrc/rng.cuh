// Counter-based sketch entries: Philox4x32-10 + exactly-reproducible transforms.
//
// The block function is bit-identical to rust-random123 `philox_4x32`
// (reference: rust-random123/src/philox.rs:149-154 round, :173-176 key bump, :211-223 ten rounds,
// constants :24-27; KAT :268-279).  On top of it this build defines its OWN entry map, because the
// reference's ThreeFry+ziggurat stream (src/sketch.rs:112-127) is order-dependent and cannot be
// evaluated per tile inside a kernel:
//
//   omega(r, c) = T_dist( philox4x32_10( ctr = (q_lo, q_hi, c, stream), key = (seed_lo, seed_hi) )[r & 3] ),
//   q = r >> 2   (four consecutive ROWS of one column share a Philox block)
//
// T_gauss is an inverse-CDF transform evaluated in FP32 with only IEEE-exact operations
// (cvt.rn, fma.rn, mul.rn, add.rn, sqrt.rn and an in-house polynomial log), so a CPU restatement
// reproduces every entry bit-for-bit; the result is widened to f64.  FP32 keeps the generator off
// the FP64 pipe that the DMMA main loop saturates.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define RNLA_HD __host__ __device__ __forceinline__
#else
#define RNLA_HD inline
#endif

namespace rnla {

enum Dist : int { DIST_GAUSSIAN = 0, DIST_UNIFORM = 1, DIST_RADEMACHER = 2 };

struct u32x4 { uint32_t x, y, z, w; };

RNLA_HD void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

RNLA_HD u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return u32x4{c0, c1, c2, c3};
}

// ---- exactly reproducible FP32 helpers ------------------------------------------------------
RNLA_HD float f_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}
RNLA_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
RNLA_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
RNLA_HD float f_sqrt(float a) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(a);
#else
    return __builtin_sqrtf(a);
#endif
}
RNLA_HD uint32_t f_bits(float a) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(a);
#else
    uint32_t u; __builtin_memcpy(&u, &a, 4); return u;
#endif
}
RNLA_HD float bits_f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float a; __builtin_memcpy(&a, &u, 4); return a;
#endif
}

// natural log of a positive normal float, |error| ~ 3e-8 + half an ulp of the result
RNLA_HD float log_pos(float t) {
    const uint32_t b = f_bits(t);
    int e = (int)(b >> 23) - 127;
    float m = bits_f((b & 0x007fffffu) | 0x3f800000u);          // [1,2)
    const bool big = m > 1.41421354f;                            // branch-free: fold into [sqrt(.5), sqrt(2))
    m = big ? f_mul(m, 0.5f) : m;
    e += big ? 1 : 0;
    const float f = f_sub(m, 1.0f);
    float p = -0x1.4237fep-4f;
    p = f_fma(p, f, 0x1.0696e4p-3f);
    p = f_fma(p, f, -0x1.0c524cp-3f);
    p = f_fma(p, f, 0x1.22973ap-3f);
    p = f_fma(p, f, -0x1.548882p-3f);
    p = f_fma(p, f, 0x1.99a3ecp-3f);
    p = f_fma(p, f, -0x1.000206p-2f);
    p = f_fma(p, f, 0x1.55554ep-2f);
    p = f_fma(p, f, -0x1.fffffep-2f);
    const float lm = f_fma(f_mul(f, f), p, f);                   // log(m)
    return f_fma((float)e, 0.693147182f, lm);
}

// standard normal from one 32-bit word: sign = top bit, tail mass v = (j + 1/2) 2^-31,
// |x| = 1 - v, z = sqrt(2) erfinv(|x|) with Giles' single-precision erfinv polynomials in w = -log(1 - x^2).
RNLA_HD float gauss_from_u32(uint32_t k) {
    const uint32_t j = k & 0x7fffffffu;
    const float v = f_fma((float)j, 0x1p-31f, 0x1p-32f);         // (0,1]
    const float t = f_mul(v, f_sub(2.0f, v));                    // 1 - x^2
    const float x = f_sub(1.0f, v);
    float w = -log_pos(t);
    float p;
    if (w < 5.0f) {
        w = f_sub(w, 2.5f);
        p = 2.81022636e-08f;
        p = f_fma(p, w, 3.43273939e-07f);
        p = f_fma(p, w, -3.5233877e-06f);
        p = f_fma(p, w, -4.39150654e-06f);
        p = f_fma(p, w, 0.00021858087f);
        p = f_fma(p, w, -0.00125372503f);
        p = f_fma(p, w, -0.00417768164f);
        p = f_fma(p, w, 0.246640727f);
        p = f_fma(p, w, 1.50140941f);
    } else {
        w = f_sub(f_sqrt(w), 3.0f);
        p = -0.000200214257f;
        p = f_fma(p, w, 0.000100950558f);
        p = f_fma(p, w, 0.00134934322f);
        p = f_fma(p, w, -0.00367342844f);
        p = f_fma(p, w, 0.00573950773f);
        p = f_fma(p, w, -0.0076224613f);
        p = f_fma(p, w, 0.00943887047f);
        p = f_fma(p, w, 1.00167406f);
        p = f_fma(p, w, 2.83297682f);
    }
    const float z = f_mul(f_mul(p, x), 1.41421354f);
    return (k >> 31) ? -z : z;
}

#if defined(__CUDACC__)
// Four samples of one Philox block at once, for the in-kernel generator: the common path of all four is straight-line
// code (so the independent polynomial chains interleave), and the rare |z| > ~3.1 tail polynomial runs under ONE
// warp-uniform vote instead of a divergent per-sample branch.  Bit-identical to gauss_from_u32 sample by sample.
// Must be called by all 32 lanes of a converged warp.
__device__ __forceinline__ void gauss_block4(const u32x4& b, float z[4]) {
    const uint32_t k[4] = {b.x, b.y, b.z, b.w};
    float w[4], x[4], pc[4];
    bool tail = false;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint32_t j = k[e] & 0x7fffffffu;
        const float v = f_fma((float)j, 0x1p-31f, 0x1p-32f);
        const float t = f_mul(v, f_sub(2.0f, v));
        x[e] = f_sub(1.0f, v);
        w[e] = -log_pos(t);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float wc = f_sub(w[e], 2.5f);
        float p = 2.81022636e-08f;
        p = f_fma(p, wc, 3.43273939e-07f);
        p = f_fma(p, wc, -3.5233877e-06f);
        p = f_fma(p, wc, -4.39150654e-06f);
        p = f_fma(p, wc, 0.00021858087f);
        p = f_fma(p, wc, -0.00125372503f);
        p = f_fma(p, wc, -0.00417768164f);
        p = f_fma(p, wc, 0.246640727f);
        p = f_fma(p, wc, 1.50140941f);
        pc[e] = p;
        tail |= !(w[e] < 5.0f);
    }
    if (__any_sync(0xffffffffu, tail)) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (!(w[e] < 5.0f)) {
                const float ws = f_sub(f_sqrt(w[e]), 3.0f);
                float p = -0.000200214257f;
                p = f_fma(p, ws, 0.000100950558f);
                p = f_fma(p, ws, 0.00134934322f);
                p = f_fma(p, ws, -0.00367342844f);
                p = f_fma(p, ws, 0.00573950773f);
                p = f_fma(p, ws, -0.0076224613f);
                p = f_fma(p, ws, 0.00943887047f);
                p = f_fma(p, ws, 1.00167406f);
                p = f_fma(p, ws, 2.83297682f);
                pc[e] = p;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float zz = f_mul(f_mul(pc[e], x[e]), 1.41421354f);
        z[e] = (k[e] >> 31) ? -zz : zz;
    }
}
#endif

// Integer-only widening f32 -> f64 (exact; the Gaussian transform never yields denormals or non-finite values).
// Keeps the generator off the FP64 pipe, which the DMMA main loop saturates (F2F.F64.F32 runs there).
RNLA_HD uint64_t f32_bits_to_f64_bits(uint32_t b) {
    const uint32_t e = (b >> 23) & 0xffu;
    if (e == 0) return (uint64_t)(b & 0x80000000u) << 32;                       // +-0
    return ((uint64_t)(b & 0x80000000u) << 32) | ((uint64_t)(e + 896u) << 52) | ((uint64_t)(b & 0x007fffffu) << 29);
}
RNLA_HD double bits_d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double a; __builtin_memcpy(&a, &u, 8); return a;
#endif
}
RNLA_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

// Uniform(-1,1): (2k+1) 2^-32 - 1, exact in f64 (a 33-bit odd integer times 2^-32), built with integer ops only.
// Rademacher: +1 iff top bit clear (mirrors `Bernoulli(0.5)`: true iff u < 2^63, src/sketch.rs:124-125).
RNLA_HD double sample_from_u32(int dist, uint32_t k) {
    if (dist == DIST_GAUSSIAN) return bits_d(f32_bits_to_f64_bits(f_bits(gauss_from_u32(k))));
    if (dist == DIST_UNIFORM) {
        const int64_t nsig = 2 * (int64_t)k + 1 - ((int64_t)1 << 32);           // odd, |n| < 2^32
        const uint64_t mag = (uint64_t)(nsig < 0 ? -nsig : nsig);
        const int p = 63 - clz64(mag);                                          // position of the leading one
        const uint64_t mant = (mag << (52 - p)) & 0x000fffffffffffffull;
        return bits_d((nsig < 0 ? 0x8000000000000000ull : 0ull) | ((uint64_t)(p - 32 + 1023) << 52) | mant);
    }
    return bits_d((k >> 31) ? 0xBFF0000000000000ull : 0x3FF0000000000000ull);
}

// the four entries rows 4q..4q+3 of column c
RNLA_HD u32x4 omega_block(uint64_t seed, uint32_t stream, uint64_t q, uint32_t c) {
    return philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), c, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
}

RNLA_HD double omega_entry(int dist, uint64_t seed, uint32_t stream, uint64_t r, uint32_t c) {
    const u32x4 b = omega_block(seed, stream, r >> 2, c);
    const uint32_t lane = (uint32_t)(r & 3);
    const uint32_t k = lane == 0 ? b.x : lane == 1 ? b.y : lane == 2 ? b.z : b.w;
    return sample_from_u32(dist, k);
}

// ---- ThreeFry2x64-20, for the reference-compatible Uniform / Rademacher streams -------------
// (reference: rust-random123/src/threefry.rs:30-93; seeding = rand_core 0.6.4 `seed_from_u64` PCG32 expansion.)
RNLA_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

RNLA_HD void threefry2x64_20(uint64_t c0, uint64_t c1, uint64_t k0, uint64_t k1, uint64_t& o0, uint64_t& o1) {
    const uint64_t ks[3] = {k0, k1, 0x1BD11BDAA9FC1A22ull ^ k0 ^ k1};
    uint64_t x0 = c0 + ks[0], x1 = c1 + ks[1];
    const int R[8] = {16, 42, 12, 31, 16, 32, 24, 21};
#pragma unroll
    for (int blk = 0; blk < 5; ++blk) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x0 += x1; x1 = rotl64(x1, R[(blk & 1) * 4 + i]); x1 ^= x0;
        }
        const int s = blk + 1;
        x0 += ks[s % 3]; x1 += ks[(s + 1) % 3] + (uint64_t)s;
    }
    o0 = x0; o1 = x1;
}

}  // namespace rnla
