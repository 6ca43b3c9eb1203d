// Scalar math building blocks of the fp32 hot path.
//
// Design rule (DESIGN.md "pipe balance"): the SFU/XU pipe retires 16 lanes/clk/SM against 128
// for FMA, so sin/cos are evaluated as short FMA polynomials after an exact quadrant reduction
// (also ~4x more accurate than MUFU.SIN, which matters for the 1e-5 parity bound), and MUFU is
// kept for lg2 / ex2 / rcp / rsqrt / sqrt only.
//
// Every function is __host__ __device__ so tests/hostcheck can compile the SAME source for the
// CPU and check it against the oracle without a GPU (the approximate MUFU ops are replaced by
// libm there; results agree to a few ulp).  Product code only ever runs the device side.
#pragma once
#include <math.h>
#include <stdint.h>

#include "philox.cuh"

namespace rvmc {

// ---- MUFU wrappers --------------------------------------------------------------------------
RVMC_HD float fast_lg2(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return log2f(x);
#endif
}
RVMC_HD float fast_ex2(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return exp2f(x);
#endif
}
RVMC_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
}
RVMC_HD float fast_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
}
RVMC_HD float fast_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
}

// p^y for p > 0 with the two exponents the Sana et al. defaults produce special-cased
// (kind is uniform over a launch: 1 -> y == 1, 2 -> y == 2, 0 -> general lg2/ex2).
enum PowKind : int { POW_GENERAL = 0, POW_ONE = 1, POW_TWO = 2 };
RVMC_HD float pow_kind(float p, float y, int kind) {
    if (kind == POW_ONE) return p;
    if (kind == POW_TWO) return p * p;
    return fast_ex2(y * fast_lg2(p));
}

// ---- sin/cos on [-pi/4, pi/4] (Cephes single-precision minimax coefficients) ----------------
RVMC_HD void sincos_kernel(float r, float& s, float& c) {
    const float z = r * r;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(z, ps, -1.6666654611e-1f);
    s = fmaf(r * z, ps, r);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(z, pc, 4.166664568298827e-2f);
    c = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
}
RVMC_HD void quadrant_fix(int q, float ss, float cc, float& s, float& c) {
    // rotate (sin, cos) of the reduced angle by q quarter turns
    const float a = (q & 1) ? cc : ss;
    const float b = (q & 1) ? ss : cc;
    s = (q & 2) ? -a : a;
    c = ((q + 1) & 2) ? -b : b;
}
RVMC_HD uint32_t float_bits(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    union { float f; uint32_t u; } v;
    v.f = x;
    return v.u;
#endif
}
// sin/cos of an angle in [0, pi] (slight excursions past either end are fine): the eccentric anomaly
// of a mean anomaly folded into [0, pi].  Folded at pi/2 -- pi - E is exact in fp32 for E > pi/2
// (Sterbenz) and the low word of pi restores the last bit -- then odd/even minimax polynomials on
// [0, pi/2] (degree 9 / 10; 6.6e-9 relative / 2.7e-10 absolute before rounding, 1.3e-7 / 1.0e-7 in
// fp32): sin keeps its RELATIVE accuracy towards both ends, where an error in E is amplified most
// (periastron).  16 FMA-pipe instructions, nothing on the XU pipe, against 25 + FRND + F2I for the
// general quadrant reduction.
RVMC_HD void sincos_0pi(float E, float& s, float& c) {
    const bool hi = E > 1.5707963267948966f;
    const float y = hi ? (3.14159274101257324f - E) + -8.74227765734758577e-08f : E;
    const float z = y * y;
    float ps = fmaf(z, 2.6043765322843951e-06f, -1.9808937104531256e-04f);
    ps = fmaf(z, ps, 8.3330565112898092e-03f);
    ps = fmaf(z, ps, -1.6666659127902242e-01f);
    s = fmaf(y * z, ps, y);
    float pc = fmaf(z, -2.6063337837564222e-07f, 2.4761072372806247e-05f);
    pc = fmaf(z, pc, -1.3888386773697650e-03f);
    pc = fmaf(z, pc, 4.1666639349409541e-02f);
    pc = fmaf(z, pc, -4.9999999513042248e-01f);
    const float cy = fmaf(z, pc, 1.0f);
    c = hi ? -cy : cy;
}
// sin/cos of 2*pi*t for a "turn" t (|t| small, e.g. a uniform draw): the quadrant split
// t - k/4 is exact in fp32, so the reduced angle carries one rounding only.
RVMC_HD void sincos_turn(float t, float& s, float& c) {
    // round-to-nearest of 4t by the 1.5 * 2^23 shift: the quadrant lands in the low mantissa bits, so the
    // rounding and the float->int conversion cost one FFMA + one FADD and no XU-pipe instruction
    // (FRND / F2I share the 16-lane XU pipe with MUFU); valid for |4t| < 2^22
    const float kb = fmaf(4.0f, t, 12582912.0f);
    const float k = kb - 12582912.0f;
    const float r = fmaf(-0.25f, k, t) * 6.283185307179586f;
    float ss, cc;
    sincos_kernel(r, ss, cc);
    quadrant_fix((int)float_bits(kb), ss, cc, s, c);
}
// sin/cos of the orientation angle omega = 2 pi t, t a uniform draw in [0, 1): MUFU.SIN / MUFU.COS on the
// angle folded into [-pi, pi) (t - 1 is exact for t >= 0.5) -- 5 instructions against 26 for the polynomial
// pair.  The unit's absolute error there is < 5e-7 (+ 2e-7 from rounding 2 pi t), and the RV is LINEAR in
// (sin omega, cos omega), so this costs < 1e-6 of K; measured on 1.2e6 orbits per population the largest
// |RV - oracle| / max(|RV|, K) moves from 1.3-1.6e-6 to 1.3-1.8e-6 (tools/parity_margin.py), against a bound
// of 1e-5, for +2 % on the multi-epoch kernel and +4 % on the fused sigma_1D kernels.  The eccentric anomaly
// keeps its polynomial pair (periastron amplifies an error in E, not one in omega).  -DRVMC_OMEGA_POLY restores
// the polynomial.
RVMC_HD void sincos_omega(float t, float& s, float& c) {
#if defined(__CUDA_ARCH__) && !defined(RVMC_OMEGA_POLY)
    const float th = (t - (t >= 0.5f ? 1.0f : 0.0f)) * 6.283185307179586f;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
#else
    sincos_turn(t, s, c);
#endif
}
// sin/cos of x radians for |x| up to a few pi (two-constant Cody-Waite reduction).
RVMC_HD void sincos_rad(float x, float& s, float& c) {
    const float k = rintf(x * 0.6366197723675814f);
    float r = fmaf(-k, 1.5707962512969971f, x);      // pi/2 high part (24 bits)
    r = fmaf(-k, 7.549789415861596e-8f, r);          // pi/2 low part
    float ss, cc;
    sincos_kernel(r, ss, cc);
    quadrant_fix((int)k, ss, cc, s, c);
}

// fp64 sin/cos for |x| up to a few hundred radians (guard path: E in [0, pi]).  fdlibm's
// __kernel_sin/__kernel_cos minimax polynomials on [-pi/4, pi/4] after a three-constant
// Cody-Waite reduction; ~1 ulp.  Unlike sin()/cos() it carries no large-argument slow path, so
// the hot kernels need no stack frame and no call.
RVMC_HD void sincos_f64_bounded(double x, double& s, double& c) {
    const double k = rint(x * 0.63661977236758134308);
    double r = fma(-k, 1.57079632673412561417e+00, x);     // pi/2, first 33 bits
    r = fma(-k, 6.07710050650619224932e-11, r);            // next 33 bits
    r = fma(-k, 2.02226624879595063154e-21, r);            // tail
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double ss = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double cc = fma(z * z, pc, fma(z, -0.5, 1.0));
    const int q = (int)k;
    const double a = (q & 1) ? cc : ss;
    const double b = (q & 1) ? ss : cc;
    s = (q & 2) ? -a : a;
    c = ((q + 1) & 2) ? -b : b;
}

// 2^t in fp64 for moderate |t| (|t| < ~1000), relative error < 5e-14: t = n + f with |f| <= 0.5,
// 2^f by an 11th-degree Taylor polynomial in f ln 2, scaled by 2^n through the exponent field.
RVMC_HD double exp2_fast_f64(double t) {
    const double n = rint(t);
    const double g = (t - n) * 0.69314718055994530942;    // f ln 2, |g| <= 0.3466
    double p = 2.50521083854417187751e-08;                // 1/11!
    p = fma(p, g, 2.75573192239858906526e-07);            // 1/10!
    p = fma(p, g, 2.75573192239858906526e-06);            // 1/9!
    p = fma(p, g, 2.48015873015873015873e-05);            // 1/8!
    p = fma(p, g, 1.98412698412698412698e-04);            // 1/7!
    p = fma(p, g, 1.38888888888888888889e-03);            // 1/6!
    p = fma(p, g, 8.33333333333333333333e-03);            // 1/5!
    p = fma(p, g, 4.16666666666666666667e-02);            // 1/4!
    p = fma(p, g, 1.66666666666666666667e-01);            // 1/3!
    p = fma(p, g, 0.5);
    p = fma(p, g, 1.0);
    p = fma(p, g, 1.0);
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(p) + ((int)n << 20), __double2loint(p));
#else
    return ldexp(p, (int)n);
#endif
}
// 10^y, ~18 instructions against ~50 for exp10(): the multi-epoch kernel needs 1/P to ~1e-10.
RVMC_HD double exp10_fast_f64(double y) { return exp2_fast_f64(y * 3.3219280948873623479); }   // log2(10)

// p^y in fp64 for a positive NORMAL p (the argument of a power-law inverse CDF, bounded away from 0
// and infinity by construction) and moderate |y log2 p|; relative error < 1e-13 |y|.  log2 p =
// k + 2 atanh((m-1)/(m+1)) log2 e with m in [sqrt(1/2), sqrt 2); the reciprocal is two Newton steps
// from MUFU.RCP, so there is no division, no libm call and no stack frame: the multi-epoch kernel
// keeps its call-free instantiation for a general period-law exponent (RVMC_bias_corr.py:282-288
// with pi not in {0, -0.5}).  ~55 instructions against a pow() call.
RVMC_HD double pow_bounded_f64(double p, double y) {
#if defined(__CUDA_ARCH__)
    int hi = __double2hiint(p);
    int k = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                    // m in [1, 2)
    if (hi >= 0x3ff6a09e) {                                 // m >= ~sqrt 2: halve it
        hi -= 0x00100000;
        k += 1;
    }
    const double m = __hiloint2double(hi, __double2loint(p));
    const double d = m + 1.0;
    double r = (double)fast_rcp((float)d);
    r = r * fma(-d, r, 2.0);
    r = r * fma(-d, r, 2.0);
    const double f = (m - 1.0) * r;
    const double s = f * f;
    double q = 1.0 / 19.0;
    q = fma(q, s, 1.0 / 17.0);
    q = fma(q, s, 1.0 / 15.0);
    q = fma(q, s, 1.0 / 13.0);
    q = fma(q, s, 1.0 / 11.0);
    q = fma(q, s, 1.0 / 9.0);
    q = fma(q, s, 1.0 / 7.0);
    q = fma(q, s, 1.0 / 5.0);
    q = fma(q, s, 1.0 / 3.0);
    q = fma(q, s, 1.0);
    const double l2 = fma(f * q, 2.8853900817779268147, (double)k);     // 2 log2(e) f q + k
    return exp2_fast_f64(y * l2);
#else
    return pow(p, y);
#endif
}

// ---- normals ---------------------------------------------------------------------------------
// Box-Muller from two Philox words: z0 = r cos(theta), z1 = r sin(theta); replaces
// np.random.normal (legacy polar method) at every noise site of the reference.
RVMC_HD void box_muller_f32(uint32_t w1, uint32_t w2, float& z0, float& z1) {
    const float u1 = u01_open_f32(w1);
    const float r = fast_sqrt(-1.3862943611198906f * fast_lg2(u1));   // sqrt(-2 ln u1)
    float s, c;
    sincos_turn(u01_f32(w2), s, c);
    z0 = r * c;
    z1 = r * s;
}
// The measurement / cluster noise normals of the fused kernels: sin and cos of the uniform angle come
// from MUFU.SIN / MUFU.COS (absolute error < 5e-7, i.e. < 3e-6 of the noise sigma -- six orders below
// the Monte-Carlo scatter of any statistic of the noise), 4 instructions against 26 for the
// polynomial pair.  The orbital angles keep the polynomial paths: they carry the 1e-5 parity bound.
RVMC_HD void box_muller_noise_f32(uint32_t w1, uint32_t w2, float& z0, float& z1) {
#if defined(__CUDA_ARCH__)
    const float u1 = u01_open_f32(w1);
    const float r = fast_sqrt(-1.3862943611198906f * fast_lg2(u1));   // sqrt(-2 ln u1)
    const float th = (float)(w2 >> 8) * 3.7450702829239286e-07f;      // 2 pi u, u = (w2 >> 8) 2^-24
    float s, c;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
    z0 = r * c;
    z1 = r * s;
#else
    box_muller_f32(w1, w2, z0, z1);
#endif
}
// Six normals from ONE Philox block: three Box-Muller pairs, each from a 22-bit radius uniform in
// (0, 1] (largest deviate 5.5 sigma against 5.8 for 24 bits) and a 20-bit angle (6e-6 rad): 126 of the
// block's 128 bits.  The measurement noise of the multi-epoch path draws from it -- six epochs per
// block instead of four -- which takes a whole Philox block (20 IMAD.WIDE, ~65 instructions) off every
// binary of a six-epoch campaign.  u[], a[]: the integer fields, shared by both precisions.
RVMC_HD void noise6_fields(const u32x4& w, uint32_t u[3], uint32_t a[3]) {
    u[0] = w.x >> 10;
    a[0] = ((w.x << 10) | (w.y >> 22)) & 0xfffffu;
    u[1] = w.y & 0x3fffffu;
    a[1] = w.z >> 12;
    u[2] = ((w.z << 10) | (w.w >> 22)) & 0x3fffffu;
    a[2] = (w.w >> 2) & 0xfffffu;
}
RVMC_HD void noise6_f32(const u32x4& w, float z[6]) {
    uint32_t u[3], a[3];
    noise6_fields(w, u, a);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 0; p < 3; ++p) {
        const float u1 = (float)(u[p] + 1u) * 2.384185791015625e-07f;                   // 2^-22
        const float r = fast_sqrt(-1.3862943611198906f * fast_lg2(u1));                 // sqrt(-2 ln u1)
        const float th = (float)a[p] * 5.9921124526782858e-06f;                         // 2 pi 2^-20
        float s, c;
#if defined(__CUDA_ARCH__)
        asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
        asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
#else
        s = sinf(th);
        c = cosf(th);
#endif
        z[2 * p] = r * c;
        z[2 * p + 1] = r * s;
    }
}
RVMC_HD void noise6_f64(const u32x4& w, double z[6]) {
    uint32_t u[3], a[3];
    noise6_fields(w, u, a);
    for (int p = 0; p < 3; ++p) {
        const double r = sqrt(-2.0 * log((double)(u[p] + 1u) * 2.384185791015625e-07));
        const double th = (double)a[p] * 5.9921124526782858e-06;
        z[2 * p] = r * cos(th);
        z[2 * p + 1] = r * sin(th);
    }
}
RVMC_HD void box_muller_f64(uint32_t w1, uint32_t w2, double& z0, double& z1) {
    const double u1 = u01_open_f64(w1);
    const double r = sqrt(-2.0 * log(u1));
    const double th = RVMC_TWO_PI * u01_f64(w2);
    z0 = r * cos(th);
    z1 = r * sin(th);
}

}  // namespace rvmc
