// Materialised (fp64) population entries and the finite-pool bootstrap.
//
// These back the public attribute arrays of the drop-in RVMC class and the replay/parity entries.
// Reference code replaced (file:line under /root/reference):
//   RVMC.__init__ samplers                     RVMC.py:79-95 (via :103-168)
//   calculate_eccentric_anomaly                RVMC.py:170-180
//   calculate_semi_major_axis / max_orbital_velocity / binary_radial_velocity   RVMC.py:182-212
//   calculate_radial_velocity (pool mixing)    RVMC.py:228-236
//   subsample (bootstrap + dispersion)         RVMC.py:238-273
#include "common.cuh"
#include "cluster_reduce.cuh"
#include "pop_derive.h"

namespace rvmc {

// ---------------------------------------------------------------------------------------------
// K1+K2 materialised
// ---------------------------------------------------------------------------------------------
__global__ void sample_orbits_kernel(PopF64 P, uint64_t seed, uint64_t first_item, int64_t n, uint32_t given,
                                     rvmc_orbit_soa out, uint32_t* __restrict__ nonconverged) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t item = first_item + (uint64_t)i;
        const u32x4 wa = philox_block(seed, item, RVMC_STREAM_ORBIT_A, 0);
        const u32x4 wb = philox_block(seed, item, RVMC_STREAM_ORBIT_B, 0);
        OrbitF64 o = sample_orbit_f64(P, wa, wb);
        if (given & RVMC_GIVEN_MASS) o.primary_mass = out.primary_mass[i];
        if (given & RVMC_GIVEN_PERIOD) {
            // period_dist (RVMC.py:84-87): time (:89) and eccentricity (:91) follow the given period
            o.period = out.period[i];
            o.time = u01_f64(wa.z) * o.period;
            o.eccentricity = eccentricity_for_period_f64(P, o.period, u01_f64(wb.x));
        }
        if (given & RVMC_GIVEN_TIME) o.time = out.time[i];
        if (given & RVMC_GIVEN_ECC) o.eccentricity = out.eccentricity[i];
        if (out.primary_mass && !(given & RVMC_GIVEN_MASS)) out.primary_mass[i] = o.primary_mass;
        if (out.period && !(given & RVMC_GIVEN_PERIOD)) out.period[i] = o.period;
        if (out.time && !(given & RVMC_GIVEN_TIME)) out.time[i] = o.time;
        if (out.mass_ratio) out.mass_ratio[i] = o.mass_ratio;
        if (out.eccentricity && !(given & RVMC_GIVEN_ECC)) out.eccentricity[i] = o.eccentricity;
        if (out.inclination) out.inclination[i] = o.inclination;
        if (out.orbit_rotation) out.orbit_rotation[i] = o.orbit_rotation;
        if (out.eccentric_anomaly) {
            double resid;
            out.eccentric_anomaly[i] = kepler_f64(RVMC_TWO_PI * o.time / o.period, o.eccentricity, resid);
            bad += (fabs(resid) <= 1e-9) ? 0u : 1u;      // NaN counts as a failure
        }
    }
    if (nonconverged && bad) atomicAdd(nonconverged, bad);
}

__global__ void orbit_scales_f64_kernel(int64_t n, rvmc_orbit_soa in, const double* __restrict__ a_in,
                                        double* __restrict__ a_out, double* __restrict__ vmax_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double q = in.mass_ratio[i], m1 = in.primary_mass[i], P = in.period[i];
        const double a = a_in ? a_in[i] : semi_major_axis_f64(m1, P, q);
        if (a_out) a_out[i] = a;
        if (vmax_out) vmax_out[i] = max_orbital_velocity_f64(m1, q, in.eccentricity[i], a);
    }
}

__global__ void rv_from_vmax_f64_kernel(int64_t n, rvmc_orbit_soa in, const double* __restrict__ v_max,
                                        double* __restrict__ rv_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = in.eccentricity[i];
        const double K = v_max[i] / (1.0 + e) * sin(in.inclination[i]);
        rv_out[i] = rv_from_K_f64(K, e, in.orbit_rotation[i], in.eccentric_anomaly[i]);
    }
}

__global__ void kepler_f64_kernel(int64_t n, const double* __restrict__ time, const double* __restrict__ period,
                                  const double* __restrict__ ecc, double* __restrict__ E_out,
                                  uint32_t* __restrict__ nonconverged) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double resid;
        E_out[i] = kepler_f64(RVMC_TWO_PI * time[i] / period[i], ecc[i], resid);
        bad += (fabs(resid) <= 1e-9) ? 0u : 1u;      // NaN counts as a failure
    }
    if (nonconverged && bad) atomicAdd(nonconverged, bad);
}

__global__ void rv_replay_f64_kernel(int64_t n, rvmc_orbit_soa in, double* __restrict__ rv_out,
                                     double* __restrict__ K_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double K;
        const double rv = orbit_rv_f64(in.primary_mass[i], in.period[i], in.mass_ratio[i], in.eccentricity[i],
                                       in.inclination[i], in.orbit_rotation[i], in.eccentric_anomaly[i], &K);
        rv_out[i] = rv;
        if (K_out) K_out[i] = K;
    }
}

// The fused path's fp32 arithmetic on caller-supplied fp64 arrays: the reference's units are
// converted to the fused path's variables in double (the conversion is not part of what is
// being checked), then semi_amplitude_f32 / kepler_f32 / rv_shape_f32 run exactly as in K4/K5.
__global__ void rv_replay_f32_kernel(int64_t n, rvmc_orbit_soa in, PopF32 P, int iters, float* __restrict__ rv_out,
                                     float* __restrict__ E_out, uint32_t* __restrict__ counters) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t n_guard = 0, n_bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        OrbitF32 o;
        o.log2m = (float)log2(in.primary_mass[i] / RVMC_MSUN_KG);
        o.x = (float)log10(in.period[i] / RVMC_DAY_S);
        const double ph = in.time[i] / in.period[i];
        const double fr = ph - floor(ph);
        o.uphase = (float)fr;
        o.phase = (float)(fr - rint(fr));
        o.q = (float)in.mass_ratio[i];
        o.e = (float)in.eccentricity[i];
        o.cosi = (float)cos(in.inclination[i]);
        o.sini = (float)sin(in.inclination[i]);
        const double wt = in.orbit_rotation[i] / RVMC_TWO_PI;
        o.wturn = (float)(wt - floor(wt));
        int g, b;
        rv_out[i] = orbit_rv_f32(P, o, iters, nullptr, g, b);
        if (E_out) {
            float sE, cE, resid;
            const float E = kepler_f32(6.283185307179586f * fabsf(o.phase), o.e, iters, sE, cE, resid);
            E_out[i] = (o.phase < 0.0f) ? 6.283185307179586f - E : E;
        }
        n_guard += g;
        n_bad += b;
    }
    if (counters) {
        if (n_guard) atomicAdd(counters + 0, n_guard);
        if (n_bad) atomicAdd(counters + 1, n_bad);
    }
}

// ---------------------------------------------------------------------------------------------
// a15: pool mixing
// ---------------------------------------------------------------------------------------------
// Keyed bijection of [0, n): a 6-round balanced Feistel network on the smallest even-width
// power-of-two domain >= n, cycle-walked back into range.  perm(i), i < nb, is a selection of nb
// distinct indices == np.random.choice(rv_binary, nb, replace=False) at RVMC.py:234.
struct FeistelKey {
    uint32_t k[6];
    uint32_t half_bits;
};

__device__ __forceinline__ uint32_t feistel_round(uint32_t r, uint32_t k) {
    uint32_t h = (r ^ k) * 0x9E3779B1u;
    h ^= h >> 15;
    h *= 0x85EBCA77u;
    h ^= h >> 13;
    h *= 0xC2B2AE3Du;
    h ^= h >> 16;
    return h;
}

__device__ __forceinline__ uint64_t feistel_perm(uint64_t i, uint64_t n, const FeistelKey& key) {
    const uint32_t hb = key.half_bits;
    const uint32_t mask = (hb >= 32) ? 0xFFFFFFFFu : ((1u << hb) - 1u);
    uint64_t x = i;
    do {
        uint32_t L = (uint32_t)(x >> hb) & mask, R = (uint32_t)x & mask;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const uint32_t t = L ^ (feistel_round(R, key.k[r]) & mask);
            L = R;
            R = t;
        }
        x = ((uint64_t)L << hb) | R;
    } while (x >= n);
    return x;
}

__global__ void mix_pool_kernel(uint64_t seed, uint32_t call_id, int64_t n, int64_t nb, FeistelKey key,
                                const double* __restrict__ rv_binary, const double* __restrict__ cluster,
                                double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (i < nb) {
            out[i] = rv_binary[feistel_perm((uint64_t)i, (uint64_t)n, key)];
        } else {
            // with replacement: one 32-bit word per pick, index = floor(w * n / 2^32)
            const u32x4 w = philox_block(seed, (uint64_t)i >> 2, RVMC_STREAM_MIX, (call_id << 1) | 1u);
            const uint32_t ww = ((i & 3) == 0) ? w.x : ((i & 3) == 1) ? w.y : ((i & 3) == 2) ? w.z : w.w;
            out[i] = cluster[(int64_t)(((uint64_t)ww * (uint64_t)n) >> 32)];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a16: finite-pool bootstrap, fp64
// ---------------------------------------------------------------------------------------------
// One tile of R realizations per block.  Phase A: one Philox block per PAIR of star slots
// (slot = realization*S + j): two pool indices + one Box-Muller pair; values go to shared memory.
// Phase B: cluster_reduce.cuh.  Per-star values never reach HBM.
__global__ void __launch_bounds__(256) sigma1d_pool_kernel(uint64_t seed, uint32_t call_id,
                                                           const double* __restrict__ pool, uint32_t pool_n,
                                                           const double* __restrict__ meas_err, int S, int R,
                                                           int G, int weighted, int ddof, uint64_t first_real,
                                                           int64_t n_real, double* __restrict__ sigma_out) {
    extern __shared__ double smem_d[];
    double* v = smem_d;                 // [R*S]
    double* err = v + (size_t)R * S;    // [S]
    double* wts = err + S;              // [S]
    __shared__ double wsum_s;

    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const double e = meas_err[j];
        err[j] = e;
        wts[j] = 1.0 / (e * e);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int j = 0; j < S; ++j) s += wts[j];
        wsum_s = s;
    }
    const int64_t n_tiles = (n_real + R - 1) / R;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r0 = tile * R;
        const int nr = (int)min((int64_t)R, n_real - r0);
        const uint64_t slot0 = (first_real + (uint64_t)r0) * (uint64_t)S;
        const int nslots = nr * S;
        const uint64_t p_begin = slot0 >> 1;
        const int npairs = (int)(((slot0 + nslots + 1) >> 1) - p_begin);
        const float invS = 1.0f / (float)S;
        __syncthreads();   // previous tile's reduction has finished reading v
        for (int pi = threadIdx.x; pi < npairs; pi += blockDim.x) {
            const uint64_t pair = p_begin + pi;
            const u32x4 w = philox_block(seed, pair, RVMC_STREAM_BOOT, call_id);
            double z0, z1;
            box_muller_f64(w.z, w.w, z0, z1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t l = (int64_t)(2 * pair + h - slot0);
                if (l >= 0 && l < nslots) {
                    const int r = (int)(((float)l + 0.5f) * invS);
                    const int j = (int)l - r * S;
                    const uint32_t word = h ? w.y : w.x;
                    const uint32_t idx = (uint32_t)(((uint64_t)word * (uint64_t)pool_n) >> 32);
                    RVMC_DEV_ASSERT(idx < pool_n && r >= 0 && r < nr && j >= 0 && j < S);
                    v[l] = pool[idx] + (h ? z1 : z0) * err[j];
                }
            }
        }
        __syncthreads();
        reduce_tile_sigma<double, double>(v, nr, S, G, weighted ? wts : nullptr, wsum_s, ddof, sigma_out + r0);
    }
}

}  // namespace rvmc

using namespace rvmc;

static int orbit_inputs_present(const rvmc_orbit_soa& a, bool need_E) {
    return a.primary_mass && a.period && a.mass_ratio && a.eccentricity && a.inclination && a.orbit_rotation &&
           (!need_E || a.eccentric_anomaly);
}

extern "C" {

int rvmc_sample_orbits(const rvmc_pop_params* p, uint64_t seed, uint64_t first_item, int64_t n,
                       uint32_t given_mask, rvmc_orbit_soa out, uint32_t* nonconverged, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(p != nullptr, "population parameters are NULL");
    RVMC_REQUIRE(n >= 0, "number of stars must be >= 0 (got %lld)", (long long)n);
    PopF64 P;
    int rc = derive_pop<double>(*p, P);
    if (rc) {
        set_error("float division by zero");   // the message Python gives at RVMC.py:112
        return rc;
    }
    if (n == 0) return RVMC_OK;
    RVMC_REQUIRE(!(given_mask & RVMC_GIVEN_MASS) || out.primary_mass, "given primary_mass is NULL");
    RVMC_REQUIRE(!(given_mask & RVMC_GIVEN_PERIOD) || out.period, "given period is NULL");
    RVMC_REQUIRE(!(given_mask & RVMC_GIVEN_TIME) || out.time, "given time is NULL");
    RVMC_REQUIRE(!(given_mask & RVMC_GIVEN_ECC) || out.eccentricity, "given eccentricity is NULL");
    const int threads = 128;
    sample_orbits_kernel<<<grid_stride_blocks(n, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        P, seed, first_item, n, given_mask, out, nonconverged);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_kepler_f64(int64_t n, const double* time, const double* period, const double* eccentricity,
                    double* E_out, uint32_t* nonconverged, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return RVMC_OK;
    RVMC_REQUIRE(time && period && eccentricity && E_out, "time, period, eccentricity and E_out are required");
    const int threads = 128;
    kepler_f64_kernel<<<grid_stride_blocks(n, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        n, time, period, eccentricity, E_out, nonconverged);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_rv_replay_f64(int64_t n, rvmc_orbit_soa in, double* rv_out, double* K_out, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return RVMC_OK;
    RVMC_REQUIRE(orbit_inputs_present(in, true) && rv_out,
                 "primary_mass, period, mass_ratio, eccentricity, inclination, orbit_rotation, "
                 "eccentric_anomaly and rv_out are required");
    const int threads = 128;
    rv_replay_f64_kernel<<<grid_stride_blocks(n, threads), threads, 0, (cudaStream_t)cuda_stream>>>(n, in, rv_out,
                                                                                                  K_out);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_orbit_scales_f64(int64_t n, rvmc_orbit_soa in, const double* a_in, double* a_out, double* vmax_out,
                          void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return RVMC_OK;
    RVMC_REQUIRE(in.primary_mass && in.period && in.mass_ratio, "primary_mass, period and mass_ratio are required");
    RVMC_REQUIRE(!vmax_out || in.eccentricity, "eccentricity is required for v_max");
    const int threads = 128;
    orbit_scales_f64_kernel<<<grid_stride_blocks(n, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        n, in, a_in, a_out, vmax_out);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_rv_from_vmax_f64(int64_t n, rvmc_orbit_soa in, const double* v_max, double* rv_out, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return RVMC_OK;
    RVMC_REQUIRE(in.eccentricity && in.inclination && in.orbit_rotation && in.eccentric_anomaly && v_max && rv_out,
                 "eccentricity, inclination, orbit_rotation, eccentric_anomaly, v_max and rv_out are required");
    const int threads = 128;
    rv_from_vmax_f64_kernel<<<grid_stride_blocks(n, threads), threads, 0, (cudaStream_t)cuda_stream>>>(n, in, v_max,
                                                                                                     rv_out);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_rv_replay_f32(int64_t n, rvmc_orbit_soa in, const rvmc_pop_params* p, float* rv_out, float* E_out,
                       uint32_t* counters, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n >= 0, "n must be >= 0");
    RVMC_REQUIRE(p != nullptr, "population parameters are NULL");
    if (n == 0) return RVMC_OK;
    RVMC_REQUIRE(orbit_inputs_present(in, false) && in.time && rv_out,
                 "primary_mass, period, time, mass_ratio, eccentricity, inclination, orbit_rotation and "
                 "rv_out are required");
    PopF32 P;
    int rc = derive_pop<float>(*p, P);
    if (rc) {
        set_error("float division by zero");
        return rc;
    }
    RVMC_REQUIRE(p->kepler_iters >= 1 && p->kepler_iters <= 8, "kepler_iters must be in [1,8]");
    const int threads = 128;
    rv_replay_f32_kernel<<<grid_stride_blocks(n, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        n, in, P, p->kepler_iters, rv_out, E_out, counters);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_mix_pool(uint64_t seed, uint32_t call_id, int64_t n, double fbin, const double* rv_binary,
                  const double* cluster_velocities, double* radial_velocities, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n >= 0, "n must be >= 0");
    RVMC_REQUIRE(call_id < (1u << 31), "call_id must be < 2^31");
    RVMC_REQUIRE(fbin >= 0.0 && fbin <= 1.0, "fbin must be in [0,1] (got %g)", fbin);
    if (n == 0) return RVMC_OK;
    RVMC_REQUIRE(rv_binary && cluster_velocities && radial_velocities, "array pointers are required");
    RVMC_REQUIRE(n < ((int64_t)1 << 32), "pool larger than 2^32 stars is not supported");
    const int64_t nb = (int64_t)((double)n * fbin);     // int(N * fbin), RVMC.py:233
    FeistelKey key;
    uint32_t bits = 2;
    while (((uint64_t)1 << bits) < (uint64_t)n) bits += 2;
    key.half_bits = bits / 2;
    // round keys: two Philox blocks of the call, evaluated on the host (same generator)
    const u32x4 a = philox_block(seed, 0, RVMC_STREAM_MIX, call_id << 1);
    const u32x4 b = philox_block(seed, 1, RVMC_STREAM_MIX, call_id << 1);
    key.k[0] = a.x; key.k[1] = a.y; key.k[2] = a.z; key.k[3] = a.w; key.k[4] = b.x; key.k[5] = b.y;
    const int threads = 256;
    mix_pool_kernel<<<grid_stride_blocks(n, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        seed, call_id, n, nb, key, rv_binary, cluster_velocities, radial_velocities);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_sigma1d_pool(uint64_t seed, uint32_t call_id, const double* pool, int64_t pool_n,
                      const double* meas_err, int sample_size, int weighted, int ddof,
                      uint64_t first_real, int64_t n_real, double* sigma_out, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n_real >= 0, "number_of_samples must be >= 0");
    RVMC_REQUIRE(sample_size >= 1, "sample_size must be >= 1 (got %d)", sample_size);
    RVMC_REQUIRE(ddof == 0 || ddof == 1, "ddof must be 0 or 1");
    RVMC_REQUIRE(pool_n >= 1, "the pool is empty");
    if (sample_size > 4096) {
        set_error("sample_size %d exceeds the 4096 stars per cluster the bootstrap kernel is built for", sample_size);
        return RVMC_ERR_UNSUPPORTED;
    }
    if (pool_n >= ((int64_t)1 << 32)) {
        set_error("pool larger than 2^32 stars is not supported");
        return RVMC_ERR_UNSUPPORTED;
    }
    if (n_real == 0) return RVMC_OK;
    RVMC_REQUIRE(pool && meas_err && sigma_out, "pool, meas_err and sigma_out are required");
    const int S = sample_size;
    const int T = 2048;                                  // star slots per tile (16 KB of fp64)
    const int R = S >= T ? 1 : T / S;
    const int G = reduce_group_size(S, R, 256);
    const size_t smem = ((size_t)R * S + 2 * (size_t)S) * sizeof(double);
    if (smem > 48 * 1024)
        RVMC_CUDA(cudaFuncSetAttribute(sigma1d_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_tiles = (n_real + R - 1) / R;
    const int64_t cap = (int64_t)sm_count() * 8;
    const unsigned int blocks = (unsigned int)(n_tiles < cap ? n_tiles : cap);
    sigma1d_pool_kernel<<<blocks, 256, smem, (cudaStream_t)cuda_stream>>>(
        seed, call_id, pool, (uint32_t)pool_n, meas_err, S, R, G, weighted, ddof, first_real, n_real, sigma_out);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

}  // extern "C"
