// Host-side derivation of the per-population constants (double), shared by the C ABI and by
// tests/hostcheck.  Mirrors the scalar prologue of RVMC.sample_powerlaw (RVMC.py:110):
// pmax = (max/min)**(power+1); the draw p = pmax + (1-pmax) u; value = min * p**(1/(power+1)).
#pragma once
#include <math.h>

#include "../../include/rvmc_b200.h"
#include "orbit.cuh"

namespace rvmc {

inline int pow_kind_of(double inv) {
    if (inv == 1.0) return POW_ONE;
    if (inv == 2.0) return POW_TWO;
    return POW_GENERAL;
}

// fp64: the reference's expression as written.  fp32: the same map renormalised so that the
// base p = a + b*u always lies in (0, 1] -- when pmax > 1 we use p/pmax and fold pmax**inv into
// the prefactor (min * pmax**inv == max).  lg2.approx has a small ABSOLUTE error near 1, so the
// end of the distribution that sits at p == 1 (e -> e_max for the default-sign indices, where
// K ~ 1/sqrt(1-e^2) is most sensitive) is computed most accurately.
template <typename T>
inline PowerLaw<T> make_powerlaw(double minv, double maxv, double power) {
    PowerLaw<T> pl;
    const double pmax = pow(maxv / minv, power + 1.0);
    const double inv = 1.0 / (power + 1.0);
    if (sizeof(T) == sizeof(float) && pmax > 1.0) {
        pl.a = (T)1.0;
        pl.b = (T)((1.0 - pmax) / pmax);
        pl.minv = (T)(minv * pow(pmax, inv));
    } else {
        pl.a = (T)pmax;
        pl.b = (T)(1.0 - pmax);
        pl.minv = (T)minv;
    }
    pl.inv = (T)inv;
    pl.kind = pow_kind_of(inv);
    return pl;
}

// returns 0, or RVMC_ERR_ZERODIV when an index is exactly -1 (RVMC.py:112 divides by power+1)
template <typename T>
inline int derive_pop(const rvmc_pop_params& p, PopT<T>& d) {
    if (p.mass_power == -1.0 || p.period_pi == -1.0 || p.mass_ratio_kappa == -1.0 || p.e_power == -1.0)
        return RVMC_ERR_ZERODIV;
    d.mass = make_powerlaw<T>(p.min_mass, p.max_mass, p.mass_power);
    d.logp = make_powerlaw<T>(log10(p.min_period), log10(p.max_period), p.period_pi);
    d.q = make_powerlaw<T>(p.min_q, p.max_q, p.mass_ratio_kappa);
    d.e_long = make_powerlaw<T>(p.e_min, p.e_max_long, p.e_power);
    d.e_mid = make_powerlaw<T>(p.e_min, p.e_max_mid, p.e_power);
    d.x_circ = (T)log10(p.p_circ_days);
    d.x_mid = (T)log10(p.p_mid_days);
    d.p_circ_s = (T)(p.p_circ_days * RVMC_DAY_S);
    d.p_mid_s = (T)(p.p_mid_days * RVMC_DAY_S);
    d.log2_mass_min = (T)log2((double)d.mass.minv);
    const double c0 = cbrt(RVMC_TWO_PI * RVMC_G_VALUE * RVMC_MSUN_KG / RVMC_DAY_S) * 1e-3;
    d.log2_c0 = (T)log2(c0);
    d.sigma_dyn = (T)p.sigma_dyn;
    d.guard_amp = (T)(p.guard_amp > 0 ? p.guard_amp : 64.0);
    return 0;
}

// Can the fp64 guard of kepler_shape_f32 ever engage for this population?  The amplification
// sqrt(1-e^2)/(1-e cos E)^2 is largest at E = 0 and grows with e, and sampled eccentricities never
// leave [e_min, max(e_max_mid, e_max_long)] (1 % margin for the fp32 evaluation).
inline bool guard_reachable(const rvmc_pop_params& p) {
    double e = fmax(fmax(p.e_max_mid, p.e_max_long), p.e_min);
    if (!(e < 1.0)) return true;
    const double amp = sqrt(1.0 - e * e) / ((1.0 - e) * (1.0 - e));
    const double thr = p.guard_amp > 0 ? p.guard_amp : 64.0;
    return amp * 1.01 >= thr;
}

inline void pop_defaults(rvmc_pop_params* p) {
    p->min_mass = 6.0;
    p->max_mass = 20.0;
    p->mass_power = -2.35;
    p->min_period = pow(10.0, 0.15);
    p->max_period = pow(10.0, 3.5);
    p->period_pi = -0.5;
    p->min_q = 0.1;
    p->max_q = 1.0;
    p->mass_ratio_kappa = 0.0;
    p->e_power = -0.5;
    p->e_min = 1e-5;
    p->e_max_mid = 0.5;
    p->e_max_long = 0.9;
    p->p_circ_days = 4.0;
    p->p_mid_days = 6.0;
    p->sigma_dyn = 2.0;
    p->fbin = 0.7;
    p->guard_amp = 64.0;
    p->kepler_iters = 1;
    p->substream = 0;
}

}  // namespace rvmc
