// TEST INFRASTRUCTURE (never loaded by rvmc_b200/): compiles the __host__ __device__ math of
// rvmc_b200/csrc/{philox,rvmc_math,orbit}.cuh for the CPU with g++, so `-m "not gpu"` tests can
// check the exact source the kernels run (Philox known answers, samplers, Kepler, RV) against
// the oracle without a GPU.  MUFU approximations are libm calls here (a few ulp apart).
#include <stdint.h>
#include <string.h>

#include "../../rvmc_b200/csrc/pop_derive.h"

using namespace rvmc;

extern "C" {

void hc_philox(uint64_t seed, uint64_t item, uint32_t stream, uint32_t sub, uint32_t* out) {
    u32x4 r = philox_block(seed, item, stream, sub);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

void hc_philox_raw(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    u32x4 c = {ctr[0], ctr[1], ctr[2], ctr[3]};
    u32x4 r = philox4x32<10>(c, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

void hc_pop_defaults(rvmc_pop_params* p) { pop_defaults(p); }

// fp32 fused-path math for items [first, first+n): outputs 9 float arrays + flags
int hc_orbits_f32(const rvmc_pop_params* p, uint64_t seed, uint64_t first, int64_t n, float* log2m,
                  float* x, float* uphase, float* q, float* e, float* cosi, float* wturn, float* K,
                  float* rv, int32_t* guard, int32_t* bad) {
    PopF32 P;
    int rc = derive_pop<float>(*p, P);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        u32x4 wa = philox_block(seed, first + i, RVMC_STREAM_ORBIT_A, 0);
        u32x4 wb = philox_block(seed, first + i, RVMC_STREAM_ORBIT_B, 0);
        OrbitF32 o = sample_orbit_f32(P, wa, wb);
        int g, b;
        float Kv;
        rv[i] = orbit_rv_f32(P, o, p->kepler_iters, &Kv, g, b);
        log2m[i] = o.log2m; x[i] = o.x; uphase[i] = o.uphase; q[i] = o.q; e[i] = o.e;
        cosi[i] = o.cosi; wturn[i] = o.wturn; K[i] = Kv; guard[i] = g; bad[i] = b;
    }
    return 0;
}

// fp64 materialised path for items [first, first+n)
int hc_orbits_f64(const rvmc_pop_params* p, uint64_t seed, uint64_t first, int64_t n, double* m1,
                  double* period, double* time, double* q, double* e, double* inc, double* om,
                  double* E, double* rv, double* K) {
    PopF64 P;
    int rc = derive_pop<double>(*p, P);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        u32x4 wa = philox_block(seed, first + i, RVMC_STREAM_ORBIT_A, 0);
        u32x4 wb = philox_block(seed, first + i, RVMC_STREAM_ORBIT_B, 0);
        OrbitF64 o = sample_orbit_f64(P, wa, wb);
        double resid;
        E[i] = kepler_f64(RVMC_TWO_PI * o.time / o.period, o.eccentricity, resid);
        rv[i] = orbit_rv_f64(o.primary_mass, o.period, o.mass_ratio, o.eccentricity, o.inclination,
                             o.orbit_rotation, E[i], &K[i]);
        m1[i] = o.primary_mass; period[i] = o.period; time[i] = o.time; q[i] = o.mass_ratio;
        e[i] = o.eccentricity; inc[i] = o.inclination; om[i] = o.orbit_rotation;
    }
    return 0;
}

// direct Kepler + shape checks on caller-chosen (phase, e, wturn)
void hc_kepler_shape_f32(int64_t n, const float* phase, const float* e, const float* wturn, int iters,
                         float guard_amp, float* shape, float* E_out, int32_t* guard) {
    PopF32 P;
    memset(&P, 0, sizeof(P));
    P.guard_amp = guard_amp;
    P.logp.a = 1;   // unused
    for (int64_t i = 0; i < n; ++i) {
        OrbitF32 o;
        memset(&o, 0, sizeof(o));
        o.phase = phase[i]; o.e = e[i]; o.wturn = wturn[i]; o.q = 1.0f; o.sini = 1.0f; o.cosi = 0.0f;
        int g, b;
        float Kv;
        // K multiplies the shape; divide it back out (K is finite and non-zero here)
        float rvv = orbit_rv_f32(P, o, iters, &Kv, g, b);
        shape[i] = rvv / Kv;
        float sE, cE, resid;
        E_out[i] = kepler_f32(6.283185307179586f * fabsf(o.phase), o.e, iters, sE, cE, resid);
        guard[i] = g;
    }
}

void hc_kepler_f64(int64_t n, const double* M, const double* e, double* E) {
    for (int64_t i = 0; i < n; ++i) {
        double r;
        E[i] = kepler_f64(M[i], e[i], r);
    }
}

void hc_mean_anomaly_epoch(int64_t n, const double* t, const double* period, int quirk, double* M) {
    for (int64_t i = 0; i < n; ++i) M[i] = mean_anomaly_epoch_f64(t[i], period[i], quirk);
}

void hc_box_muller(int64_t n, const uint32_t* w1, const uint32_t* w2, float* z0, float* z1, double* d0,
                   double* d1) {
    for (int64_t i = 0; i < n; ++i) {
        box_muller_f32(w1[i], w2[i], z0[i], z1[i]);
        box_muller_f64(w1[i], w2[i], d0[i], d1[i]);
    }
}

// six normals per Philox block: w is [n][4] words, z32 / z64 are [n][6]
void hc_noise6(int64_t n, const uint32_t* w, float* z32, double* z64) {
    for (int64_t i = 0; i < n; ++i) {
        u32x4 b;
        b.x = w[4 * i];
        b.y = w[4 * i + 1];
        b.z = w[4 * i + 2];
        b.w = w[4 * i + 3];
        noise6_f32(b, z32 + 6 * i);
        noise6_f64(b, z64 + 6 * i);
    }
}

void hc_exp10_fast(int64_t n, const double* y, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = exp10_fast_f64(y[i]);
}

void hc_sincos_f64(int64_t n, const double* x, double* s, double* c) {
    for (int64_t i = 0; i < n; ++i) sincos_f64_bounded(x[i], s[i], c[i]);
}

// turn: 1 sincos_turn (angle in turns), 0 sincos_rad, 2 sincos_0pi (eccentric anomaly in [0, pi])
void hc_sincos(int64_t n, const float* t, int turn, float* s, float* c) {
    for (int64_t i = 0; i < n; ++i) {
        if (turn == 1) sincos_turn(t[i], s[i], c[i]);
        else if (turn == 2) sincos_0pi(t[i], s[i], c[i]);
        else sincos_rad(t[i], s[i], c[i]);
    }
}
}
