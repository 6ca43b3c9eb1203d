// Per-orbit device functions: inverse-CDF sampling, Kepler solve, radial velocity.
//
// Two precisions share the same Philox draws:
//   * OrbitF32 / *_f32 : the fused hot path.  Works in the variables the formulas need
//     (log2 M1, x = log10 P/day, orbital phase in turns, cos i) and never materialises P, M1, a.
//   * *_f64            : the materialised drop-in arrays and the fp64 replay entries; same maps
//     written in the reference's own units (kg, s, rad).
//
// Reference code replaced (file:line in /root/reference):
//   sample_powerlaw RVMC.py:103-112   sample_masses :121-126   sample_period :128-134
//   sample_mass_ratio :136-140        sample_eccentricity :148-161   sample_inclination :114-119
//   sample_orbit_rotation :142-146    sample_time :163-168
//   calculate_eccentric_anomaly :170-180 (SciPy secant -> closed-form starter + Halley)
//   calculate_semi_major_axis :182-187, max_orbital_velocity :189-198, binary_radial_velocity :200-212
#pragma once
#include "rvmc_constants.h"
#include "rvmc_math.cuh"

namespace rvmc {

// ---------------------------------------------------------------------------------------------
// Derived constants of one population (host-computed in double by derive_pop(), abi.cu)
// ---------------------------------------------------------------------------------------------
// Power-law inverse CDF  x = min * (a + b*u)^inv,  a = pmax = (max/min)^(power+1), b = 1 - pmax.
template <typename T>
struct PowerLaw {
    T a, b, inv, minv;
    int kind;   // PowKind of `inv`
};

template <typename T>
struct PopT {
    PowerLaw<T> mass;     // solar masses (2e30 kg)
    PowerLaw<T> logp;     // x = log10(P/day)
    PowerLaw<T> q;
    PowerLaw<T> e_long;   // P > p_mid      (e < 0.9)
    PowerLaw<T> e_mid;    // p_circ < P <= p_mid (e < 0.5)
    T x_circ, x_mid;      // log10(4), log10(6): thresholds in x
    T p_circ_s, p_mid_s;  // the same thresholds in seconds (fp64 path compares like the reference)
    T log2_mass_min;      // log2(min_mass)
    T log2_c0;            // log2 of C0 = (2 pi G Msun / day)^(1/3) * 1e-3  [km/s]
    T sigma_dyn;
    T guard_amp;          // fp64 guard: refine when sqrt(1-e^2)/(1-e cos E)^2 exceeds this
};
typedef PopT<float> PopF32;
typedef PopT<double> PopF64;

// ---------------------------------------------------------------------------------------------
// fp32 hot path
// ---------------------------------------------------------------------------------------------
struct OrbitF32 {
    float log2m;   // log2(M1 / Msun_ref)
    float x;       // log10(P / day)
    float phase;   // orbital phase in turns, wrapped to [-0.5, 0.5]  (M = 2 pi phase)
    float uphase;  // the raw draw in [0,1): time = uphase * P
    float q;
    float e;
    float cosi, sini;
    float wturn;   // omega / 2 pi
};

RVMC_HD float powerlaw_f32(const PowerLaw<float>& pl, float u) {
    return pl.minv * pow_kind(fmaf(pl.b, u, pl.a), pl.inv, pl.kind);
}

// Orbit of `item` from Philox blocks ORBIT_A/ORBIT_B (7 of the 8 words are used).
RVMC_HD OrbitF32 sample_orbit_f32(const PopF32& P, const u32x4& wa, const u32x4& wb) {
    OrbitF32 o;
    // mass: only log2 m is needed downstream
    {
        const float p = fmaf(P.mass.b, u01_f32(wa.x), P.mass.a);
        o.log2m = (P.mass.kind == POW_ONE) ? fast_lg2(P.mass.minv * p)
                                           : fmaf(P.mass.inv, fast_lg2(p), P.log2_mass_min);
    }
    o.x = powerlaw_f32(P.logp, u01_f32(wa.y));
    o.uphase = u01_f32(wa.z);
    o.phase = o.uphase - rintf(o.uphase);
    o.q = powerlaw_f32(P.q, u01_f32(wa.w));
    {
        const float ue = u01_f32(wb.x);
        const bool is_long = o.x > P.x_mid;
        const PowerLaw<float>& pl = is_long ? P.e_long : P.e_mid;
        const float e = powerlaw_f32(pl, ue);
        o.e = (o.x > P.x_circ) ? e : 0.0f;
    }
    o.cosi = u01_f32(wb.y);
    o.sini = fast_sqrt((1.0f - o.cosi) * (1.0f + o.cosi));
    o.wturn = u01_f32(wb.z);
    return o;
}

// Semi-amplitude K = q/(1+q) * (2 pi G M1 (1+q) / P)^(1/3) * sin i / sqrt(1-e^2)  [km/s]
// == RVMC.py:187,197-198,209 rearranged; one ex2 on a sum of log2 terms.
RVMC_HD float semi_amplitude_f32(const PopF32& P, const OrbitF32& o) {
    const float third = 0.3333333333333333f;
    float l = fmaf(third, o.log2m, P.log2_c0);
    l = fmaf(-1.1073093649624542f, o.x, l);                      // -(log2 10)/3 * x
    l = fmaf(-0.6666666666666666f, fast_lg2(1.0f + o.q), l);
    const float ome2 = (1.0f - o.e) * (1.0f + o.e);
    return fast_ex2(l) * o.q * o.sini * fast_rsqrt(ome2);
}

// Kepler's equation E - e sin E = M for 0 <= M <= pi (callers fold the sign of M out).
// Markley (1995) cubic starter (|error| < 5e-4 rad for every e < 1) followed by `iters` Halley
// steps; the iteration count is a launch-uniform constant, so a warp never diverges here.
// Returns E and its sin/cos (angle-addition update of the last evaluated pair).
template <int ITERS = 0>
RVMC_HD float kepler_f32(float M, float e, int iters_rt, float& sE, float& cE, float& resid) {
    const int iters = ITERS > 0 ? ITERS : iters_rt;      // ITERS > 0: fully unrolled, no loop overhead
    // alpha = (3 pi^2 + 1.6 pi (pi - M)/(1+e)) / (pi^2 - 6) = a0 - a1 M.  Everything that depends on e
    // alone (rc, a0, a1, ome2, ome3) is loop-invariant for a caller that solves several epochs of one
    // orbit: the unrolled epoch loop of the multi-epoch kernel shares it.
    const float ome = 1.0f - e;
    const float rc = fast_rcp(1.0f + e);
    const float a1 = 1.2989824604108398f * rc;                        // 1.6 pi / (pi^2 - 6)
    const float a0 = fmaf(4.080873754768689f, rc, 7.651638290191292f);   // 1.6 pi^2/(pi^2-6), 3 pi^2/(pi^2-6)
    const float ome2 = ome + ome, ome3 = 3.0f * ome;
    const float alpha = fmaf(-a1, M, a0);
    const float d = fmaf(alpha, e, ome3);
    const float ad = alpha * d;
    const float M2 = M * M;
    const float q = fmaf(ad, ome2, -M2);
    const float g = fmaf(3.0f, ad * fmaf(alpha, e, ome2), M2);       // r / M, with d - ome = 2 ome + alpha e
    const float r = M * g;                                           // >= 0 for M >= 0
    float w = r + fast_sqrt(fmaf(q * q, q, r * r));
    w = fast_ex2(0.6666666666666666f * fast_lg2(w));                 // w^(2/3)
    const float den = fmaf(w, w + q, q * q);
    // E0 = (2 r w / den + M) / d  with a single reciprocal.  At M = 0: r = 0, w = q > 0, den = 3 q^2 > 0
    // for every e < 1, so E0 = 0 exactly and no 0/0 guard is needed.
    float E = fmaf(r + r, w, M * den) * fast_rcp(den * d);
    float s, c, f = 0.0f;
    sincos_0pi(E, s, c);
#pragma unroll(ITERS > 0 ? ITERS : 1)
    for (int it = 0; it < iters; ++it) {
        f = fmaf(-e, s, E) - M;
        const float fp = fmaf(-e, c, 1.0f);
        const float fpp = e * s;
        const float dE = -f * fp * fast_rcp(fmaf(fp, fp, -0.5f * f * fpp));
        E += dE;
        if (it + 1 < iters) {
            sincos_0pi(E, s, c);
        } else {
            // sin/cos(E + dE) from sin/cos(E): |dE| < 1e-3 after the starter, so cos dE = 1 - dE^2/2 and
            // sin dE = dE to 4e-14 and 2e-10 absolute
            const float cd = fmaf(dE * dE, -0.5f, 1.0f);
            const float s1 = fmaf(s, cd, c * dE);
            c = fmaf(c, cd, -s * dE);
            s = s1;
        }
    }
    sE = s;
    cE = c;
    resid = fmaf(-e, s, E) - M;
    return E;
}

// fp64 guard: given the fp32 solution, polish E in double and return sin/cos in double.
RVMC_HD void kepler_refine_f64(double M, double e, double& E, double& sE, double& cE) {
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
        sincos_f64_bounded(E, sE, cE);
        E -= (E - e * sE - M) / (1.0 - e * cE);
    }
    sincos_f64_bounded(E, sE, cE);
}

// RV / K = cos(nu + omega) + e cos(omega) with the true anomaly taken in closed form
// (cos nu = (cos E - e)/(1 - e cos E), sin nu = sqrt(1-e^2) sin E/(1 - e cos E)); equal to the
// 2*atan(sqrt((1+e)/(1-e)) tan(E/2)) form of RVMC.py:211-212.  rome2 = sqrt(1 - e^2).
RVMC_HD float rv_shape_f32(float e, float rome2, float sE, float cE, float sw, float cw) {
    // shape = [ (cE - e) cw - sqrt(1-e^2) sE sw ] / (1 - e cE) + e cw; the products of per-orbit
    // quantities (rome2 sw, e cw) are shared by all epochs of an orbit
    const float rd = fast_rcp(fmaf(-e, cE, 1.0f));
    const float B = rome2 * sw, C = e * cw;
    return fmaf(rd, fmaf(-B, sE, fmaf(cE, cw, -C)), C);
}

// Kepler solve + shape for a mean anomaly of magnitude M in [0, pi] and sign `neg`
// (M64 is the same magnitude in double, used by the guard only).  The fp64 guard engages when
// the amplification of an error in M into the true anomaly, sqrt(1-e^2)/(1-e cos E)^2, exceeds
// guard_amp: only for e > ~0.9, i.e. never with the reference's hard-coded bounds (RVMC.py:158).
// `guard` / `bad` report the guard path and a residual above 1e-5 (SciPy's RuntimeWarning).
// GUARD = false compiles the guard out: the host proves it unreachable when the largest
// amplification the eccentricity bounds allow, sqrt(1-e^2)/(1-e)^2 at e_max, stays below guard_amp
// (true for the reference's bounds: 43.6 at e = 0.9 against the default threshold 64).
// `aresid` returns |E - e sin E - M| of the fp32 solve (0 on the guard path): callers that evaluate several
// epochs of one orbit keep a running maximum and test it once (one FMNMX per epoch instead of a compare
// and a predicated add).
template <int ITERS = 0, bool GUARD = true>
RVMC_HD float kepler_shape_resid_f32(float M, double M64, bool neg, float e, float rome2, float sw, float cw,
                                     int iters, float guard_amp, int& guard, float& aresid) {
    float sE, cE, resid;
    const float E = kepler_f32<ITERS>(M, e, iters, sE, cE, resid);
    const float dd = fmaf(-e, cE, 1.0f);
    if (GUARD && dd * dd * guard_amp < rome2) {
        double E64 = (double)E, s64, c64;
        const double e64 = (double)e;
        kepler_refine_f64(M64, e64, E64, s64, c64);
        if (neg) s64 = -s64;
        const double rd = 1.0 / (1.0 - e64 * c64);
        const double cnu = (c64 - e64) * rd;
        const double snu = sqrt((1.0 - e64) * (1.0 + e64)) * s64 * rd;
        guard = 1;
        aresid = 0.0f;
        return (float)(cnu * (double)cw - snu * (double)sw + e64 * (double)cw);
    }
    guard = 0;
    aresid = fabsf(resid);
    return rv_shape_f32(e, rome2, neg ? -sE : sE, cE, sw, cw);
}
template <int ITERS = 0, bool GUARD = true>
RVMC_HD float kepler_shape_f32(float M, double M64, bool neg, float e, float rome2, float sw, float cw,
                               int iters, float guard_amp, int& guard, int& bad) {
    float aresid;
    const float shape = kepler_shape_resid_f32<ITERS, GUARD>(M, M64, neg, e, rome2, sw, cw, iters, guard_amp, guard,
                                                             aresid);
    bad = (aresid > 1e-5f) ? 1 : 0;
    return shape;
}

// Full single-epoch orbital RV of one sampled orbit [km/s]; the sign of the phase folds M into
// [0, pi].
template <int ITERS = 0, bool GUARD = true>
RVMC_HD float orbit_rv_f32(const PopF32& P, const OrbitF32& o, int iters, float* K_out, int& guard,
                           int& bad) {
    const float K = semi_amplitude_f32(P, o);
    const float aphase = fabsf(o.phase);
    float sw, cw;
    sincos_omega(o.wturn, sw, cw);
    const float rome2 = fast_sqrt((1.0f - o.e) * (1.0f + o.e));
    const float shape = kepler_shape_f32<ITERS, GUARD>(6.283185307179586f * aphase, RVMC_TWO_PI * (double)aphase,
                                         o.phase < 0.0f, o.e, rome2, sw, cw, iters, P.guard_amp, guard, bad);
    if (K_out) *K_out = K;
    // a rounded product, never contracted into the caller's accumulation: every kernel that adds this
    // RV to a noise term then produces the same bits (sigma1d_fused_kernel / sigma1d_fused_fbins_kernel)
#if defined(__CUDA_ARCH__)
    return __fmul_rn(K, shape);
#else
    return K * shape;
#endif
}

// ---------------------------------------------------------------------------------------------
// fp64 path: the same draws in the reference's units
// ---------------------------------------------------------------------------------------------
struct OrbitF64 {
    double primary_mass;    // kg
    double period;          // s
    double time;            // s
    double mass_ratio;
    double eccentricity;
    double inclination;     // rad
    double orbit_rotation;  // rad
};

RVMC_HD double powerlaw_f64(const PowerLaw<double>& pl, double u) {
    // np.random.uniform(low, high) is low + (high-low)*u with separate roundings: no FMA
    // contraction here, so the materialised arrays equal NumPy's bit for bit for exponents 1 and 2
#if defined(__CUDA_ARCH__)
    const double p = __dadd_rn(pl.a, __dmul_rn(pl.b, u));
#else
    const double p = pl.a + pl.b * u;
#endif
    // exponents 1 and 2 (kappa = 0, pi = e_power = -0.5: the Sana et al. defaults) need no pow()
    const double y = (pl.kind == POW_ONE) ? p : (pl.kind == POW_TWO) ? p * p : pow(p, pl.inv);
    return y * pl.minv;
}

// RVMC.py:148-161 with one uniform slot per orbit
RVMC_HD double eccentricity_for_period_f64(const PopF64& P, double period, double ue) {
    if (period > P.p_mid_s) return powerlaw_f64(P.e_long, ue);
    if (period > P.p_circ_s) return powerlaw_f64(P.e_mid, ue);
    return 0.0;
}

RVMC_HD OrbitF64 sample_orbit_f64(const PopF64& P, const u32x4& wa, const u32x4& wb) {
    OrbitF64 o;
    o.primary_mass = powerlaw_f64(P.mass, u01_f64(wa.x)) * RVMC_MSUN_KG;
    o.period = RVMC_DAY_S * pow(10.0, powerlaw_f64(P.logp, u01_f64(wa.y)));
    o.time = u01_f64(wa.z) * o.period;
    o.mass_ratio = powerlaw_f64(P.q, u01_f64(wa.w));
    o.eccentricity = eccentricity_for_period_f64(P, o.period, u01_f64(wb.x));
    o.inclination = acos(u01_f64(wb.y));
    o.orbit_rotation = u01_f64(wb.z) * RVMC_TWO_PI;
    return o;
}

// Root of E - e sin E = M for any real M, returned on the branch the reference reaches from
// x0 = M (E in [0, 2 pi) for M in [0, 2 pi)).  Converged to the last bit or two; `resid` lets
// the caller count failures (SciPy's RuntimeWarning equivalent).
RVMC_HD double kepler_f64(double M, double e, double& resid) {
    const double pi = 3.141592653589793;
    const double turns = floor(M / RVMC_TWO_PI + 0.5);
    const double Mr = M - turns * RVMC_TWO_PI;            // [-pi, pi]
    const double a = fabs(Mr);
    const double ome = 1.0 - e;
    const double alpha = (3.0 * pi * pi + 1.6 * pi * (pi - a) / (1.0 + e)) / (pi * pi - 6.0);
    const double d = 3.0 * ome + alpha * e;
    const double q = 2.0 * alpha * d * ome - a * a;
    const double r = 3.0 * alpha * d * (d - ome) * a + a * a * a;
    double w = fabs(r) + sqrt(q * q * q + r * r);
    w = cbrt(w * w);
    double E = (a == 0.0) ? 0.0 : (2.0 * r * w / (w * w + w * q + q * q) + a) / d;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const double s = sin(E), c = cos(E);
        const double f = E - e * s - a;
        const double fp = 1.0 - e * c;
        E -= f * fp / (fp * fp - 0.5 * f * e * s);
    }
    resid = E - e * sin(E) - a;
    E = (Mr < 0.0) ? -E : E;
    return E + turns * RVMC_TWO_PI;
}

// RVMC.py:182-212 evaluated as written (semi-major axis, periastron speed, arctan/tan form).
RVMC_HD double semi_major_axis_f64(double primary_mass, double period, double q) {
    return cbrt(RVMC_G_VALUE * (1.0 + q) * primary_mass * period * period / (4.0 * 9.869604401089358));
}
RVMC_HD double max_orbital_velocity_f64(double primary_mass, double q, double e, double a) {
    return q * sqrt(RVMC_G_VALUE * primary_mass * (1.0 + e) / ((1.0 + q) * (1.0 - e) * a)) * 0.001;
}
RVMC_HD double semi_amplitude_f64(double primary_mass, double period, double q, double e, double inc) {
    const double a = semi_major_axis_f64(primary_mass, period, q);
    return max_orbital_velocity_f64(primary_mass, q, e, a) / (1.0 + e) * sin(inc);
}
RVMC_HD double rv_from_K_f64(double K, double e, double omega, double E) {
    const double nu = 2.0 * atan(sqrt((1.0 + e) / (1.0 - e)) * tan(0.5 * E));
    return K * (e * cos(omega) + cos(nu + omega));
}
RVMC_HD double orbit_rv_f64(double primary_mass, double period, double q, double e, double inc,
                            double omega, double E, double* K_out) {
    const double K = semi_amplitude_f64(primary_mass, period, q, e, inc);
    if (K_out) *K_out = K;
    return rv_from_K_f64(K, e, omega, E);
}

// Multi-epoch mean anomaly.  quirk != 0 reproduces RVMC_bias_corr.py:357, whose operator
// precedence yields ((2 pi t) mod P) / P in [0,1) used as radians; quirk == 0
// is the physical 2 pi frac(t / P).  Needs fp64: 2 pi t reaches 2e8 s against P down to 1e5 s.
RVMC_HD double mean_anomaly_epoch_f64(double t, double period, int quirk) {
    if (quirk) {
        const double a = RVMC_TWO_PI * t;
        double m = fmod(a, period);
        if (m < 0.0) m += period;          // np.mod takes the sign of the divisor
        return m / period;
    }
    const double ph = t / period;
    return RVMC_TWO_PI * (ph - floor(ph));
}

}  // namespace rvmc
