/*
 * oracle/det_math.h -- TEST INFRASTRUCTURE (CPU oracle).  Never linked into the product.
 *
 * "detmath v1": the deterministic elementary functions of the canonical arithmetic.
 *
 * Why it exists: the reference (pf::resamplers::mn_resampler, restated in
 * SURVEY.md A.3; in-tree twin /root/reference/include/ssme/liu_west_filter.h:97-101)
 * forms resampling weights as exp(lw - max).  glibc's exp and CUDA's libdevice exp
 * differ in the last bit, and one flipped bit in a weight can flip an ancestor index.
 * To make ancestors bit-exact between this oracle and the sm_100a kernel at ANY size,
 * both sides evaluate the same polynomials with the same IEEE-754 operation sequence
 * (every rounding step is spelled out below; fma = one rounding).  The CUDA twin is
 * ssme_b200/csrc/det_math.cuh, written independently from this spec.
 * Coefficients are derived by tools/gen_coeffs.py (Chebyshev-node interpolation).
 *
 * Compile with -ffp-contract=off so the compiler performs exactly these operations.
 */
#ifndef SSME_ORACLE_DET_MATH_H
#define SSME_ORACLE_DET_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint64_t dm_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double dm_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
static inline uint32_t dm_f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static inline float dm_u2f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }

#define DM_LOG2E 0x1.71547652b82fep+0
#define DM_LN2_HI 0x1.62e4200000000p-1
#define DM_LN2_LO 0x1.fdf473de6af28p-22
#define DM_SHIFT 0x1.8p52
#define DM_HALF_LOG_2PI 0x1.d67f1c864beb5p-1
#define DM_SQRT2 0x1.6a09e667f3bcdp+0

static const double DM_EXP_Q[10] = {
    0x1.0000000000001p-1, 0x1.5555555555556p-3, 0x1.5555555553d63p-5, 0x1.11111111109b3p-7,
    0x1.6c16c1788bd90p-10, 0x1.a01a01a7c41d5p-13, 0x1.a019b90d2ae7ap-16, 0x1.71de0dae63bb3p-19,
    0x1.289185613a3d6p-22, 0x1.af38a9b0ec855p-26,
};

/* exp(x).  NaN -> NaN; x <= -708 -> +0 (no subnormal results); x > 709 -> +inf. */
static inline double dm_exp(double x)
{
    if (x != x) return x;
    if (x <= -708.0) return 0.0;
    if (x > 709.0) return INFINITY;
    double t = fma(x, DM_LOG2E, DM_SHIFT);
    int32_t k = (int32_t)(uint32_t)(dm_d2u(t) & 0xffffffffu);
    double kd = t - DM_SHIFT;
    double r = fma(kd, -DM_LN2_HI, x);
    r = fma(kd, -DM_LN2_LO, r);
    double q = DM_EXP_Q[9];
    for (int i = 8; i >= 0; --i) q = fma(q, r, DM_EXP_Q[i]);
    double p = fma(q, r, 1.0);
    p = fma(p, r, 1.0);
    double scale = dm_u2d((uint64_t)(uint32_t)(k + 1023) << 52);
    return p * scale;
}

static const double DM_LOG_R[8] = {
    0x1.5555555555555p-1, 0x1.9999999999a38p-2, 0x1.2492492476c42p-2, 0x1.c71c720168526p-3,
    0x1.745cf8fe328f9p-3, 0x1.3b1c36b445cebp-3, 0x1.0fbe71ad855c9p-3, 0x1.0c05166ec4148p-3,
};

/* log(x).  NaN or x<0 -> NaN; 0 -> -inf; +inf -> +inf; subnormals pre-scaled by 2^54. */
static inline double dm_log(double x)
{
    if (x != x || x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    int32_t e = 0;
    uint64_t b = dm_d2u(x);
    if ((b >> 52) == 0) { x = x * 0x1p54; b = dm_d2u(x); e = -54; }
    e += (int32_t)(b >> 52) - 1023;
    double m = dm_u2d((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
    if (m > DM_SQRT2) { m = m * 0.5; e += 1; }
    double ke = (double)e;
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double R = DM_LOG_R[7];
    for (int i = 6; i >= 0; --i) R = fma(R, z, DM_LOG_R[i]);
    double t1 = (s * z) * R;
    double lo = fma(ke, DM_LN2_LO, t1);
    double mid = fma(2.0, s, lo);
    return fma(ke, DM_LN2_HI, mid);
}

/* ---- float32 Box-Muller: two N(0,1) variates from two 32-bit words ------------------------- */
#define DM_LN2_F 0x1.62e430p-1f
#define DM_SQRT2_F 0x1.6a09e6p+0f
static const float DM_FLOG_P[9] = {
    -0x1.fffffep-2f, 0x1.555554p-2f, -0x1.00020cp-2f, 0x1.99a014p-3f, -0x1.548368p-3f,
    0x1.22ea2cp-3f, -0x1.0cdb1ep-3f, 0x1.04915ap-3f, -0x1.3a4fa2p-4f,
};
static const float DM_FSIN_S[5] = {
    0x1.921fb6p+0f, -0x1.4abbc4p-1f, 0x1.4668f0p-4f, -0x1.32531ep-8f, 0x1.3e1420p-13f,
};
static const float DM_FCOS_C[6] = {
    0x1.000000p+0f, -0x1.3bd3ccp+0f, 0x1.03c1e4p-2f, -0x1.55d074p-6f, 0x1.e126b0p-11f, -0x1.8fb3f4p-16f,
};

static inline void dm_box_muller(uint32_t a, uint32_t b, float* z0, float* z1)
{
    /* radius: u in (0,1] with 24 bits, r = sqrt(-2 ln u) */
    float u = (float)((a >> 8) + 1u) * 0x1p-24f;
    uint32_t ib = dm_f2u(u);
    int32_t e = (int32_t)(ib >> 23) - 127;
    float m = dm_u2f((ib & 0x007fffffu) | 0x3f800000u);
    if (m > DM_SQRT2_F) { m = m * 0.5f; e += 1; }
    float f = m - 1.0f;
    float P = DM_FLOG_P[8];
    for (int i = 7; i >= 0; --i) P = fmaf(P, f, DM_FLOG_P[i]);
    float lnm = fmaf(f * f, P, f);
    float lnu = fmaf((float)e, DM_LN2_F, lnm);
    /* u = 1 (one radius word in 2^24) gives -2 ln u = -0: clamped to 2^-100, radius 2^-50 instead of 0 (the kernel's
     * branch-free square root is correctly rounded on [2^-100, 2^100]) */
    float v2 = -2.0f * lnu;
    float r = sqrtf(v2 > 0x1p-100f ? v2 : 0x1p-100f);
    /* angle: quadrant from the top two bits, 24-bit fraction of a quarter turn below them */
    uint32_t quad = b >> 30;
    float t = (float)((b >> 6) & 0x00ffffffu) * 0x1p-24f;
    float z = t * t;
    float S = DM_FSIN_S[4];
    for (int i = 3; i >= 0; --i) S = fmaf(S, z, DM_FSIN_S[i]);
    float C = DM_FCOS_C[5];
    for (int i = 4; i >= 0; --i) C = fmaf(C, z, DM_FCOS_C[i]);
    float sn = t * S, cs = C;
    float c2 = (quad & 1u) ? -sn : cs; /* rotate by quad * 90 degrees */
    float s2 = (quad & 1u) ? cs : sn;
    if (quad & 2u) { c2 = -c2; s2 = -s2; }
    *z0 = r * c2;
    *z1 = r * s2;
}

/* 53-bit uniform in [0,1) from two 32-bit words */
/* ---- float32 exp for the optional fp32 mode: Cody-Waite reduction with the 1.5*2^23 shift, degree-5 polynomial on
 * (e^r - 1 - r)/r^2 (Cephes expf coefficients), every operation a correctly rounded float operation in this order.
 * NaN -> NaN; x <= -87 -> +0; x > 88 -> +inf. */
static inline float dm_fexp(float x)
{
    if (x != x) return x;
    if (x <= -87.0f) return 0.0f;
    if (x > 88.0f) return INFINITY;
    float t = fmaf(x, 0x1.715476p+0f, 0x1.8p23f);
    int32_t k = (int32_t)dm_f2u(t) - 0x4B400000;
    float kd = t - 0x1.8p23f;
    float r = fmaf(kd, -0x1.62e400p-1f, x);
    r = fmaf(kd, -0x1.7f7d1cp-20f, r);
    float p = 0x1.a0d2cep-13f;
    p = fmaf(p, r, 0x1.6e879cp-10f);
    p = fmaf(p, r, 0x1.1112fap-7f);
    p = fmaf(p, r, 0x1.555502p-5f);
    p = fmaf(p, r, 0x1.555550p-3f);
    p = fmaf(p, r, 0x1.000000p-1f);
    float v = fmaf(r * r, p, r) + 1.0f;
    return v * dm_u2f((uint32_t)(k + 127) << 23);
}
/* 24-bit uniform in [0,1) */
static inline float dm_uniform24(uint32_t w) { return (float)(w >> 8) * 0x1p-24f; }

static inline double dm_uniform53(uint32_t hi, uint32_t lo)
{
    uint64_t v = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return (double)v * 0x1p-53;
}

/* 32-bit uniform in [0,1): word * 2^-32, exact ("detmath v2": the i.i.d. multinomial resampling targets, 4 per block) */
static inline double dm_uniform32(uint32_t w) { return (double)w * 0x1p-32; }

/* ---- Philox4x32-R (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3") ------
 * The filters use R = DM_PHILOX_ROUNDS = 7 ("detmath v2"), the smallest round count of Philox4x32 that passes BigCrush in
 * that paper; the round function is pinned to Random123's known answers at R = 10 (tests/test_oracle.py). */
#define DM_PHILOX_ROUNDS 7
static inline void dm_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < rounds; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline void dm_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { dm_philox4x32_r(ctr, key, 10, out); }
static inline void dm_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { dm_philox4x32_r(ctr, key, DM_PHILOX_ROUNDS, out); }

#endif
