// K5: multi-epoch Delta-RV and detection probability (BiasCorr), fused, plus the materialised
// fp64 entries that back the drop-in class's attribute arrays and replay workflow.
//
// Reference code replaced (file:line under /root/reference):
//   BiasCorr.__init__ samplers                      RVMC_bias_corr.py:175-243 (via :256-328)
//   calculate_semi_major_axis / max_orbital_velocity :330-346
//   calculate_eccentric_anomaly                      :348-358  (phase quirk Q1 at :357)
//   binary_radial_velocity                           :360-372
//   generate_rv_errors                               :395-404
//   calculate_radial_velocities                      :406-422
//   generate_single_delta_rv                         :424-442
//   get_detected_binaries                            :451-476
//
// Fused kernel: one thread per binary (star i, sample s); the star is block-uniform so the
// observing epochs, their errors and the pair thresholds of the significance test sit in shared
// memory.  Per binary: two Philox blocks -> orbit; the period and the per-epoch phase reduction
// run in fp64 (2 pi t / P reaches ~2e3 turns: fp32 would lose the phase), everything after the
// mean anomaly is fp32: closed-form starter + Halley, closed-form true anomaly, RV.  Noise is
// one Philox block per six epochs.  max/min and the pairwise test are taken in registers; only
// Delta-RV (4 B) and the detection flag (1 B) per binary reach HBM, and both are optional.
#include "common.cuh"
#include "pop_derive.h"

namespace rvmc {

#ifndef BC_THREADS
#define BC_THREADS 256
#endif

struct BcArgs {
    PopF32 P32;
    PopF64 P64;
    PhiloxKey key, noise_key;     // expanded on the host: round keys are constant-bank operands
    uint64_t seed;                // (kept for the fp64 mass draw)
    int32_t nstars, max_epochs;
    int32_t star0;                // first star of this launch (gridDim.y is limited to 65535)
    int32_t spt;                  // samples per thread
    int32_t first_star;           // global index of local star 0 (RNG counter only)
    const int32_t* n_epochs;      // [nstars]
    const double* t_obs;          // [nstars][max_epochs]  s
    const float* rv_err;          // [nstars][max_epochs]  km/s, or nullptr
    int32_t mass_mode;            // 0 IMF, 1 N(masses, mass_errors), 2 masses
    const double* masses;         // [nstars] kg
    const double* mass_errors;    // [nstars] kg
    int32_t quirk_phase;
    int32_t kepler_iters;
    float delta_rv_min, n_sigma;
    uint64_t first_sample;
    int64_t n_samples, sample_stride;
    float* delta_rv;              // [nstars][sample_stride]
    uint8_t* detected;            // [nstars][sample_stride]
    uint32_t* det_bits;           // [nstars][bits_stride]: the same flags, one bit per sample (ballot words)
    int64_t bits_stride;
    float* rv_out;                // [nstars][max_epochs][sample_stride]
    unsigned long long* counters; // [4]: binary-epoch RVs, detected, guard, non-converged
    uint32_t* det_per_star;       // [nstars]
};

__device__ __forceinline__ uint64_t bc_item(int star, uint64_t sample) {
    return ((uint64_t)star << 40) + sample;
}

// Primary mass of (star, sample) in kg for the three mass modes; `u` is the IMF uniform.
__device__ __forceinline__ double bc_mass_f64(const PopF64& P, int mode, const double* masses,
                                              const double* mass_errors, int star, uint64_t seed, uint64_t item,
                                              double u) {
    if (mode == 0) return powerlaw_f64(P.mass, u) * RVMC_MSUN_KG;                  // RVMC_bias_corr.py:205-207
    if (mode == 2) return masses[star];                                            // :190
    // :183 -- N(masses, mass_errors); one Philox block per pair of samples
    const u32x4 w = philox_block(seed, item >> 1, RVMC_STREAM_MASS, 0);
    double z0, z1;
    box_muller_f64(w.x, w.y, z0, z1);
    return masses[star] + mass_errors[star] * ((item & 1) ? z1 : z0);
}

// Noise normals of epochs 6g..6g+5 of one binary: one Philox block (noise6_*, rvmc_math.cuh).
__device__ __forceinline__ void bc_noise6_f32(const PhiloxKey& K, uint64_t item, uint32_t stream, int g, float z[6]) {
    noise6_f32(philox_block(K, item, stream, (uint32_t)g), z);
}
__device__ __forceinline__ void bc_noise6_f64(uint64_t seed, uint64_t item, uint32_t stream, int g, double z[6]) {
    noise6_f64(philox_block(seed, item, stream, (uint32_t)g), z);
}

// NE    : compile-time bound on the epochs per star (<= 16: epoch loop fully unrolled, RVs in registers)
// FAST  : every launch-uniform switch resolved at compile time -- phase convention QUIRK, measurement
//         errors given iff NOISY, per-epoch RVs written iff RVOUT, one Halley step, fp64 guard proven
//         unreachable by the host (guard_reachable()); a general period-law exponent goes through the
//         call-free pow_bounded_f64.  These instantiations contain no call and no stack frame.
//         !FAST keeps the switches as run-time values (same arithmetic, ~15 % more instructions,
//         run-time Halley count, fp64 guard, pow()).
// Each thread walks a.spt samples of its star (1..BC_MAX_SPT, chosen by the host so that the grid
// still holds >= 8 blocks per SM), so the per-block prologue (epoch tables, pair thresholds, two
// barriers) is amortised over BC_THREADS * spt binaries.  3 blocks/SM at <= 85 registers measured best.
// PERSISTENT blocks (a grid of k blocks per SM, each walking a contiguous run of such chunks and
// restaging the tables only when the star changes) were built and measured in round 2 (commit b0198e4):
// 3.20 / 3.19 / 3.12 ms for k = 3 / 4 / 6 against 3.08 ms for this one-chunk-per-block grid on 10^8 x 6 --
// the ~24 500 short-lived blocks are not the cost: freshly scheduled blocks keep the three resident blocks
// of an SM in different phases of the per-binary instruction mix, blocks that start together do not.
#ifndef BC_MAX_SPT
#define BC_MAX_SPT 16
#endif
#ifndef BC_MIN_BLOCKS
#define BC_MIN_BLOCKS 3
#endif
template <int NE, bool FAST, bool QUIRK, bool RVOUT, bool NOISY>
__global__ void __launch_bounds__(BC_THREADS, BC_MIN_BLOCKS) biascorr_fused_kernel(BcArgs a) {
    __shared__ double A_s[NE];            // 2 pi t_obs (quirk) or t_obs (physical)
    __shared__ float err_s[NE];
    __shared__ float thr_s[NE * NE];      // pair thresholds max(delta_RV, n_sigma sqrt(ea^2+eb^2))
    __shared__ int uniform_s;
    __shared__ unsigned int c_det, c_guard, c_bad;

    constexpr int ITERS = FAST ? 1 : 0;
    constexpr bool GUARD = !FAST;
    const int star = a.star0 + (int)blockIdx.y;
    const int ne = a.n_epochs[star];
    RVMC_DEV_ASSERT(star >= 0 && star < a.nstars && ne >= 0 && ne <= NE && ne <= a.max_epochs);
    const bool noisy = FAST ? NOISY : (a.rv_err != nullptr);
    const bool quirk = FAST ? QUIRK : (a.quirk_phase != 0);
    float* const rv_out = (FAST && !RVOUT) ? nullptr : a.rv_out;

    if (threadIdx.x < ne) {
        const double t = a.t_obs[(size_t)star * a.max_epochs + threadIdx.x];
        A_s[threadIdx.x] = quirk ? RVMC_TWO_PI * t : t;
        err_s[threadIdx.x] = noisy ? a.rv_err[(size_t)star * a.max_epochs + threadIdx.x] : 0.0f;
    }
    if (threadIdx.x == 0) {
        c_det = 0;
        c_guard = 0;
        c_bad = 0;
    }
    __syncthreads();
    if (noisy) {
        for (int p = threadIdx.x; p < ne * ne; p += blockDim.x) {
            const int x = p / ne, y = p - x * ne;
            const float ex = err_s[x], ey = err_s[y];
            thr_s[x * NE + y] = fmaxf(a.delta_rv_min, a.n_sigma * sqrtf(fmaf(ex, ex, ey * ey)));
        }
        if (threadIdx.x == 0) {
            int u = 1;
            for (int k = 1; k < ne; ++k) u &= (err_s[k] == err_s[0]);
            uniform_s = u;
        }
    }
    __syncthreads();

    unsigned int my_guard = 0, my_bad = 0, my_det = 0;
    const int lane = threadIdx.x & 31;
    // the block's samples are [block_base, block_base + limit); inside the loop the index is a 32-bit offset
    const int64_t block_base = (int64_t)blockIdx.x * (BC_THREADS * a.spt);
    const int limit = (int)min((int64_t)(BC_THREADS * a.spt), a.n_samples - block_base);
    // word of the packed flags that holds this warp's samples of the first pass; BC_THREADS / 32 words further
    // per pass.  Only lane 0 stores.
    uint32_t* bits_w = a.det_bits ? a.det_bits + (size_t)star * a.bits_stride + (block_base >> 5) + (threadIdx.x >> 5)
                                  : nullptr;
#pragma unroll 1
    for (int il = threadIdx.x; il < limit; il += BC_THREADS, bits_w += BC_THREADS / 32) {
        const int64_t i = block_base + il;
        const uint64_t sample = a.first_sample + (uint64_t)i;
        const uint64_t item = bc_item(a.first_star + star, sample);
        const u32x4 wa = philox_block(a.key, item, RVMC_STREAM_ORBIT_A, 0);
        const u32x4 wb = philox_block(a.key, item, RVMC_STREAM_ORBIT_B, 0);

        // ---- per-binary quantities --------------------------------------------------------------
        OrbitF32 o;
        if (a.mass_mode == 0) {
            const float p = fmaf(a.P32.mass.b, u01_f32(wa.x), a.P32.mass.a);
            o.log2m = (a.P32.mass.kind == POW_ONE) ? fast_lg2(a.P32.mass.minv * p)
                                                   : fmaf(a.P32.mass.inv, fast_lg2(p), a.P32.log2_mass_min);
        } else {
            // masses at face value (:190) or N(masses, mass_errors) (:183); the normal is the fp32
            // Box-Muller of the SAME words rvmc_biascorr_sample uses (1e-7 relative apart): the hot
            // kernel stays free of fp64 transcendentals and of the function calls they bring
            double m = a.masses[star];
            if (a.mass_mode == 1) {
                const u32x4 w = philox_block(a.key, item >> 1, RVMC_STREAM_MASS, 0);
                float z0, z1;
                box_muller_f32(w.x, w.y, z0, z1);
                m = fma(a.mass_errors[star], (double)((item & 1) ? z1 : z0), m);
            }
            o.log2m = fast_lg2((float)(m * (1.0 / RVMC_MSUN_KG)));     // m <= 0 -> NaN like the reference (Q12)
        }
        // Period in fp64: the phase needs ~1e-10 relative on P (2 pi t / P reaches ~2e3 turns).  Only
        // 1/P is ever needed: time = u P (RVMC_bias_corr.py:328), so t / P = t_obs / P + u, and the
        // eccentricity regimes compare log10 P.
        double xd;
        if constexpr (FAST) {
            // exponent 1 or 2 (pi = 0 or the Sana et al. -0.5) directly, any other by the call-free
            // pow_bounded_f64: no pow(), no call, no stack frame
            const double p = __dadd_rn(a.P64.logp.a, __dmul_rn(a.P64.logp.b, u01_f64(wa.y)));
            const double pw = (a.P64.logp.kind == POW_ONE) ? p
                              : (a.P64.logp.kind == POW_TWO) ? p * p : pow_bounded_f64(p, a.P64.logp.inv);
            xd = pw * a.P64.logp.minv;
        } else {
            xd = powerlaw_f64(a.P64.logp, u01_f64(wa.y));
        }
        const double invP = exp10_fast_f64(-4.9365137424788932 - xd);   // 1 / (86400 * 10^x)  [1/s]
        o.x = (float)xd;
        const double uphase = u01_f64(wa.z);
        o.q = powerlaw_f32(a.P32.q, u01_f32(wa.w));
        {
            const float ue = u01_f32(wb.x);
            const bool is_long = xd > a.P64.x_mid;
            const float e = powerlaw_f32(is_long ? a.P32.e_long : a.P32.e_mid, ue);
            o.e = (xd > a.P64.x_circ) ? e : 0.0f;
        }
        o.cosi = u01_f32(wb.y);
        o.sini = fast_sqrt((1.0f - o.cosi) * (1.0f + o.cosi));
        o.wturn = u01_f32(wb.z);
        const float K = semi_amplitude_f32(a.P32, o);
        float sw, cw;
        sincos_omega(o.wturn, sw, cw);
        const float Ksw = K * sw, Kcw = K * cw;
        const float rome2 = fast_sqrt((1.0f - o.e) * (1.0f + o.e));
        const double c0 = quirk ? RVMC_TWO_PI * uphase : uphase;

        // ---- epochs -----------------------------------------------------------------------------
        float rvs[NE];
        float vmax = -INFINITY, vmin = INFINITY;
        float z[6];
        float res_max = 0.0f;             // largest Kepler residual over the epochs of this binary
        auto epoch = [&](int k) {
            if (noisy && (k % 6) == 0) bc_noise6_f32(a.noise_key, item, RVMC_STREAM_EPOCH, k / 6, z);
            float M;
            double M64;
            bool neg = false;
            // quirk (RVMC_bias_corr.py:357): ((2 pi t) mod P) / P == frac(2 pi t_obs / P + 2 pi u), in
            // [0,1) and used as radians; physical: 2 pi frac(t_obs / P + u)
            const double v = fma(A_s[k], invP, c0);
            const double fr = v - floor(v);
            if (quirk) {
                M64 = fr;
            } else {
                const double fold = fr - rint(fr);                  // [-0.5, 0.5]
                neg = fold < 0.0;
                M64 = RVMC_TWO_PI * fabs(fold);
            }
            M = (float)M64;
            int g;
            float ar;
            // K is folded into the orientation terms (the shape is linear in sin/cos omega)
            float rv = kepler_shape_resid_f32<ITERS, GUARD>(M, M64, neg, o.e, rome2, Ksw, Kcw, a.kepler_iters,
                                                            a.P32.guard_amp, g, ar);
            my_guard += g;
            res_max = fmaxf(res_max, ar);
            if (noisy) rv = fmaf(err_s[k], z[k % 6], rv);
            rvs[k] = rv;
            vmax = fmaxf(vmax, rv);
            vmin = fminf(vmin, rv);
            if (rv_out) rv_out[((size_t)star * a.max_epochs + k) * a.sample_stride + i] = rv;
        };
        if constexpr (NE <= 16) {
            // fully unrolled: rvs[] and z[] live in registers; a star with exactly NE epochs (the common
            // case: NE is chosen from max_epochs) takes the branch-free copy
            if (ne == NE) {
#pragma unroll
                for (int k = 0; k < NE; ++k) epoch(k);
            } else {
#pragma unroll
                for (int k = 0; k < NE; ++k)
                    if (k < ne) epoch(k);
            }
        } else {
#pragma unroll 1
            for (int k = 0; k < ne; ++k) epoch(k);
        }
        my_bad += (res_max > 1e-5f) ? 1u : 0u;
        // np.max/np.min propagate NaN (mass draw <= 0, Q12); fmaxf/fminf would drop it
        float drv = vmax - vmin;
        if (K != K) drv = K;
        if (a.delta_rv) a.delta_rv[(size_t)star * a.sample_stride + i] = drv;

        // ---- significance test --------------------------------------------------------------------
        bool det = false;
        if (noisy) {
            if (uniform_s) {
                // equal errors: every pair has the same threshold, so the largest difference decides
                det = (ne >= 2) && (drv > thr_s[1]);
            } else if constexpr (NE <= 16) {
#pragma unroll
                for (int x = 0; x < NE; ++x)
#pragma unroll
                    for (int y = x + 1; y < NE; ++y)
                        if (y < ne) det = det || (fabsf(rvs[x] - rvs[y]) > thr_s[x * NE + y]);
            } else {
                for (int x = 0; x < ne; ++x)
                    for (int y = x + 1; y < ne; ++y) det = det || (fabsf(rvs[x] - rvs[y]) > thr_s[x * NE + y]);
            }
        }
        if (a.detected) a.detected[(size_t)star * a.sample_stride + i] = det ? 1 : 0;
        if (a.det_bits) {
            // RVMC_bias_corr.py:476 returns one bool per binary: one BIT of it leaves the chip.  A warp holds
            // 32 consecutive samples; the lanes past the end of a ragged last warp have left the loop and are
            // not named in the mask, so their bits are 0.  (a.det_bits is tested, not bits_w: launch-uniform)
            unsigned int word;
            if ((il | 31) < limit) word = __ballot_sync(0xffffffffu, det);
            else word = __ballot_sync((1u << (limit & 31)) - 1u, det);    // a computed mask costs ~30 instructions: tail only
            if (lane == 0) *bits_w = word;
        }
        my_det += det ? 1u : 0u;
    }

    // ---- counters ---------------------------------------------------------------------------------
    if (a.counters || a.det_per_star) {
        for (int off = 16; off > 0; off >>= 1) my_det += __shfl_xor_sync(0xffffffffu, my_det, off);
        if ((threadIdx.x & 31) == 0 && my_det) atomicAdd(&c_det, my_det);
        if (my_guard) atomicAdd(&c_guard, my_guard);
        if (my_bad) atomicAdd(&c_bad, my_bad);
        __syncthreads();
        if (threadIdx.x == 0) {
            const int64_t first = (int64_t)blockIdx.x * (BC_THREADS * a.spt);
            const int64_t nvalid = min((int64_t)(BC_THREADS * a.spt), a.n_samples - first);
            if (a.counters) {
                atomicAdd(a.counters + 0, (unsigned long long)(nvalid * ne));
                if (c_det) atomicAdd(a.counters + 1, (unsigned long long)c_det);
                if (c_guard) atomicAdd(a.counters + 2, (unsigned long long)c_guard);
                if (c_bad) atomicAdd(a.counters + 3, (unsigned long long)c_bad);
            }
            if (a.det_per_star && c_det) atomicAdd(a.det_per_star + star, c_det);
        }
    }
}

template <int NE, bool FAST, bool QUIRK, bool RVOUT, bool NOISY>
static int bc_launch(BcArgs& a, int64_t blocks_per_star, cudaStream_t stream) {
    // gridDim.y carries the star (<= 65535 per launch)
    for (int s0 = 0; s0 < a.nstars; s0 += 65535) {
        a.star0 = s0;
        const int ny = a.nstars - s0 < 65535 ? a.nstars - s0 : 65535;
        biascorr_fused_kernel<NE, FAST, QUIRK, RVOUT, NOISY>
            <<<dim3((unsigned int)blocks_per_star, (unsigned int)ny), BC_THREADS, 0, stream>>>(a);
        RVMC_LAUNCH_CHECK();
    }
    return RVMC_OK;
}

template <int NE>
static int bc_dispatch(BcArgs& a, bool guard, int64_t blocks_per_star, cudaStream_t stream) {
    if constexpr (NE <= 16) {
        // launch-uniform switches -> template arguments (phase convention, RVs written, errors given)
        if (a.kepler_iters == 1 && !guard) {
            const int sel = (a.quirk_phase ? 4 : 0) | (a.rv_out ? 2 : 0) | (a.rv_err ? 1 : 0);
            switch (sel) {
                case 0: return bc_launch<NE, true, false, false, false>(a, blocks_per_star, stream);
                case 1: return bc_launch<NE, true, false, false, true>(a, blocks_per_star, stream);
                case 2: return bc_launch<NE, true, false, true, false>(a, blocks_per_star, stream);
                case 3: return bc_launch<NE, true, false, true, true>(a, blocks_per_star, stream);
                case 4: return bc_launch<NE, true, true, false, false>(a, blocks_per_star, stream);
                case 5: return bc_launch<NE, true, true, false, true>(a, blocks_per_star, stream);
                case 6: return bc_launch<NE, true, true, true, false>(a, blocks_per_star, stream);
                default: return bc_launch<NE, true, true, true, true>(a, blocks_per_star, stream);
            }
        }
    }
    return bc_launch<NE, false, false, true, true>(a, blocks_per_star, stream);
}

// ---------------------------------------------------------------------------------------------
// materialised fp64 entries
// ---------------------------------------------------------------------------------------------
__global__ void biascorr_sample_kernel(PopF64 P, uint64_t seed, int nstars, int mass_mode,
                                       const double* __restrict__ masses, const double* __restrict__ mass_errors,
                                       uint64_t first_sample, int64_t n_samples, uint32_t given,
                                       rvmc_orbit_soa out) {
    const int64_t total = (int64_t)nstars * n_samples;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int star = (int)(t / n_samples);
        const int64_t i = t - (int64_t)star * n_samples;
        const uint64_t item = bc_item(star, first_sample + (uint64_t)i);
        const u32x4 wa = philox_block(seed, item, RVMC_STREAM_ORBIT_A, 0);
        const u32x4 wb = philox_block(seed, item, RVMC_STREAM_ORBIT_B, 0);
        OrbitF64 o = sample_orbit_f64(P, wa, wb);
        if (mass_mode != 0) o.primary_mass = bc_mass_f64(P, mass_mode, masses, mass_errors, star, seed, item, 0.0);
        if (given & RVMC_GIVEN_PERIOD) {
            o.period = out.period[t];
            o.time = u01_f64(wa.z) * o.period;
            o.eccentricity = eccentricity_for_period_f64(P, o.period, u01_f64(wb.x));
        }
        if (out.primary_mass && !(given & RVMC_GIVEN_MASS)) out.primary_mass[t] = o.primary_mass;
        if (out.period && !(given & RVMC_GIVEN_PERIOD)) out.period[t] = o.period;
        if (out.time) out.time[t] = o.time;
        if (out.mass_ratio) out.mass_ratio[t] = o.mass_ratio;
        if (out.eccentricity) out.eccentricity[t] = o.eccentricity;
        if (out.inclination) out.inclination[t] = o.inclination;
        if (out.orbit_rotation) out.orbit_rotation[t] = o.orbit_rotation;
    }
}

// RVMC_bias_corr.py:406-422 in fp64 on caller-supplied arrays (possibly user-overwritten).
__global__ void biascorr_replay_f64_kernel(int nstars, int64_t n_samples, int max_epochs,
                                           const int32_t* __restrict__ n_epochs, const double* __restrict__ t_obs,
                                           rvmc_orbit_soa in, int64_t mass_cols, const float* __restrict__ rv_err,
                                           uint64_t seed, uint64_t first_sample, int quirk,
                                           double* __restrict__ rv_out, double* __restrict__ delta_rv) {
    const int64_t total = (int64_t)nstars * n_samples;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int star = (int)(t / n_samples);
        const int64_t i = t - (int64_t)star * n_samples;
        const double m1 = in.primary_mass[mass_cols == 1 ? star : t];
        const double P = in.period[t], e = in.eccentricity[t], w = in.orbit_rotation[t];
        const double K = semi_amplitude_f64(m1, P, in.mass_ratio[t], e, in.inclination[t]);
        const double time = in.time[t];
        const int ne = n_epochs[star];
        const uint64_t item = bc_item(star, first_sample + (uint64_t)i);
        double vmax = -INFINITY, vmin = INFINITY;
        bool any_nan = false;
        double z[6];
        for (int k = 0; k < ne; ++k) {
            if (rv_err && (k % 6) == 0) bc_noise6_f64(seed, item, RVMC_STREAM_EPOCH, k / 6, z);
            const double tt = t_obs[(size_t)star * max_epochs + k] + time;
            const double M = mean_anomaly_epoch_f64(tt, P, quirk);
            double resid;
            const double E = kepler_f64(M, e, resid);
            double rv = rv_from_K_f64(K, e, w, E);
            if (rv_err) rv += (double)rv_err[(size_t)star * max_epochs + k] * z[k % 6];
            if (rv_out) rv_out[((size_t)star * max_epochs + k) * n_samples + i] = rv;
            any_nan |= (rv != rv);
            vmax = fmax(vmax, rv);
            vmin = fmin(vmin, rv);
        }
        if (delta_rv) delta_rv[t] = any_nan ? nan("") : vmax - vmin;
    }
}

// calculate_eccentric_anomaly(time_deltas) for one star (RVMC_bias_corr.py:348-358)
__global__ void kepler_epochs_f64_kernel(int n_epochs, int64_t n_samples, const double* __restrict__ td,
                                         const double* __restrict__ period, const double* __restrict__ ecc, int quirk,
                                         double* __restrict__ E_out, uint32_t* __restrict__ nonconverged) {
    const int64_t total = (int64_t)n_epochs * n_samples;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t bad = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t i = t % n_samples;
        double resid;
        E_out[t] = kepler_f64(mean_anomaly_epoch_f64(td[t], period[i], quirk), ecc[i], resid);
        bad += (fabs(resid) <= 1e-9) ? 0u : 1u;      // NaN counts as a failure
    }
    if (nonconverged && bad) atomicAdd(nonconverged, bad);
}

// binary_radial_velocity(eccentric_anomalies_i) for one star (RVMC_bias_corr.py:360-372)
__global__ void rv_epochs_f64_kernel(int n_epochs, int64_t n_samples, const double* __restrict__ E,
                                     const double* __restrict__ vmax, const double* __restrict__ ecc,
                                     const double* __restrict__ incl, const double* __restrict__ omega,
                                     double* __restrict__ rv_out) {
    const int64_t total = (int64_t)n_epochs * n_samples;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t i = t % n_samples;
        const double e = ecc[i];
        rv_out[t] = rv_from_K_f64(vmax[i] / (1.0 + e) * sin(incl[i]), e, omega[i], E[t]);
    }
}

// generate_single_delta_rv (RVMC_bias_corr.py:424-442)
__global__ void biascorr_singles_kernel(uint64_t seed, int nstars, int max_epochs,
                                        const int32_t* __restrict__ n_epochs, const float* __restrict__ rv_err,
                                        uint64_t first_single, int64_t n_singles, double* __restrict__ out) {
    const int64_t total = (int64_t)nstars * n_singles;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int star = (int)(t / n_singles);
        const int64_t i = t - (int64_t)star * n_singles;
        const uint64_t item = bc_item(star, first_single + (uint64_t)i);
        const int ne = n_epochs[star];
        double vmax = -INFINITY, vmin = INFINITY;
        double z[6];
        for (int k = 0; k < ne; ++k) {
            if ((k % 6) == 0) bc_noise6_f64(seed, item, RVMC_STREAM_SINGLE, k / 6, z);
            const double rv = (double)rv_err[(size_t)star * max_epochs + k] * z[k % 6];
            vmax = fmax(vmax, rv);
            vmin = fmin(vmin, rv);
        }
        out[t] = vmax - vmin;
    }
}

// get_detected_binaries (RVMC_bias_corr.py:451-476) as written, fp64
__global__ void biascorr_detect_kernel(int nstars, int64_t n_samples, int max_epochs,
                                       const int32_t* __restrict__ n_epochs, const double* __restrict__ rv,
                                       const float* __restrict__ rv_err, double delta_rv_min, double n_sigma,
                                       uint8_t* __restrict__ detected) {
    const int64_t total = (int64_t)nstars * n_samples;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int star = (int)(t / n_samples);
        const int64_t i = t - (int64_t)star * n_samples;
        const int ne = n_epochs[star];
        const double* col = rv + (size_t)star * max_epochs * n_samples + i;
        const float* err = rv_err + (size_t)star * max_epochs;
        bool det = false;
        for (int x = 0; x < ne; ++x) {
            const double vx = col[(size_t)x * n_samples];
            const double ex = (double)err[x];
            for (int y = x + 1; y < ne; ++y) {
                const double d = fabs(vx - col[(size_t)y * n_samples]);
                const double ey = (double)err[y];
                det = det || ((d / sqrt(ex * ex + ey * ey) > n_sigma) && (d > delta_rv_min));
            }
        }
        detected[t] = det ? 1 : 0;
    }
}

}  // namespace rvmc

using namespace rvmc;

static int bc_check_common(const rvmc_pop_params* p, int nstars, int max_epochs, const int32_t* n_epochs) {
    RVMC_REQUIRE(p != nullptr, "population parameters are NULL");
    RVMC_REQUIRE(nstars >= 0 && nstars < (1 << 24), "nstars must be in [0, 2^24)");
    RVMC_REQUIRE(max_epochs >= 1, "max_epochs must be >= 1");
    RVMC_REQUIRE(nstars == 0 || n_epochs != nullptr, "n_epochs is required");
    return RVMC_OK;
}

extern "C" {

int rvmc_biascorr_fused(const rvmc_pop_params* p, uint64_t seed, uint64_t noise_seed, int first_star, int nstars,
                        int max_epochs,
                        const int32_t* n_epochs, const double* t_obs, const float* rv_err, int mass_mode,
                        const double* masses, const double* mass_errors, int quirk_phase, float delta_rv_min,
                        float n_sigma, uint64_t first_sample, int64_t n_samples, int64_t sample_stride,
                        float* delta_rv, uint8_t* detected, float* rv_out, uint64_t* counters,
                        uint32_t* det_per_star, void* cuda_stream) {
    return rvmc_biascorr_fused_packed(p, seed, noise_seed, first_star, nstars, max_epochs, n_epochs, t_obs, rv_err,
                                      mass_mode, masses, mass_errors, quirk_phase, delta_rv_min, n_sigma, first_sample,
                                      n_samples, sample_stride, delta_rv, detected, nullptr, 0, rv_out, counters,
                                      det_per_star, cuda_stream);
}

int rvmc_biascorr_fused_packed(const rvmc_pop_params* p, uint64_t seed, uint64_t noise_seed, int first_star,
                               int nstars, int max_epochs, const int32_t* n_epochs, const double* t_obs,
                               const float* rv_err, int mass_mode, const double* masses, const double* mass_errors,
                               int quirk_phase, float delta_rv_min, float n_sigma, uint64_t first_sample,
                               int64_t n_samples, int64_t sample_stride, float* delta_rv, uint8_t* detected,
                               uint32_t* detected_bits, int64_t bits_stride, float* rv_out, uint64_t* counters,
                               uint32_t* det_per_star, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    int rc = bc_check_common(p, nstars, max_epochs, n_epochs);
    if (rc) return rc;
    RVMC_REQUIRE(n_samples >= 0, "number_of_samples must be >= 0");
    RVMC_REQUIRE(first_star >= 0 && (int64_t)first_star + nstars <= (1 << 24), "star index must stay below 2^24");
    RVMC_REQUIRE(sample_stride >= n_samples, "sample_stride must be >= n_samples");
    RVMC_REQUIRE(detected_bits == nullptr || bits_stride >= (n_samples + 31) / 32,
                 "bits_stride must be >= ceil(n_samples / 32) words");
    RVMC_REQUIRE(mass_mode >= 0 && mass_mode <= 2, "mass_mode must be 0, 1 or 2");
    RVMC_REQUIRE(mass_mode == 0 || masses != nullptr, "masses are required for mass_mode %d", mass_mode);
    RVMC_REQUIRE(mass_mode != 1 || mass_errors != nullptr, "mass_errors are required for mass_mode 1");
    RVMC_REQUIRE(first_sample + (uint64_t)n_samples <= ((uint64_t)1 << 40), "sample index must stay below 2^40");
    RVMC_REQUIRE(p->kepler_iters >= 1 && p->kepler_iters <= 8, "kepler_iters must be in [1,8]");
    if (max_epochs > 64) {
        set_error("%d epochs per star exceed the 64 the fused kernel is built for", max_epochs);
        return RVMC_ERR_UNSUPPORTED;
    }
    BcArgs a;
    rc = derive_pop<float>(*p, a.P32);
    if (!rc) rc = derive_pop<double>(*p, a.P64);
    if (rc) {
        set_error("float division by zero");
        return rc;
    }
    if (nstars == 0 || n_samples == 0) return RVMC_OK;
    RVMC_REQUIRE(t_obs != nullptr, "t_obs is required");
    a.seed = seed;
    a.key = philox_key(seed);
    a.noise_key = philox_key(noise_seed);
    a.nstars = nstars;
    a.max_epochs = max_epochs;
    a.star0 = 0;
    a.first_star = first_star;
    a.n_epochs = n_epochs;
    a.t_obs = t_obs;
    a.rv_err = rv_err;
    a.mass_mode = mass_mode;
    a.masses = masses;
    a.mass_errors = mass_errors;
    a.quirk_phase = quirk_phase;
    a.kepler_iters = p->kepler_iters;
    a.delta_rv_min = delta_rv_min;
    a.n_sigma = n_sigma;
    a.first_sample = first_sample;
    a.n_samples = n_samples;
    a.sample_stride = sample_stride;
    a.delta_rv = delta_rv;
    a.detected = detected;
    a.det_bits = detected_bits;
    a.bits_stride = bits_stride;
    a.rv_out = rv_out;
    a.counters = (unsigned long long*)counters;
    a.det_per_star = det_per_star;
    // samples per thread: as many as keep >= 8 blocks per SM in flight
    int64_t spt = ((int64_t)nstars * n_samples) / ((int64_t)BC_THREADS * sm_count() * 8);
    spt = spt < 1 ? 1 : (spt > BC_MAX_SPT ? BC_MAX_SPT : spt);
    a.spt = (int32_t)spt;
    const int64_t blocks_per_star = (n_samples + BC_THREADS * spt - 1) / (BC_THREADS * spt);
    if (blocks_per_star > 0x7fffffffLL) {
        set_error("too many samples in one launch (%lld); shard them", (long long)n_samples);
        return RVMC_ERR_UNSUPPORTED;
    }
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const bool guard = guard_reachable(*p);
    if (max_epochs <= 6) return bc_dispatch<6>(a, guard, blocks_per_star, stream);
    if (max_epochs <= 8) return bc_dispatch<8>(a, guard, blocks_per_star, stream);
    if (max_epochs <= 16) return bc_dispatch<16>(a, guard, blocks_per_star, stream);
    return bc_dispatch<64>(a, guard, blocks_per_star, stream);      // 17..64 epochs: run-time loops, RVs in local memory
}

int rvmc_biascorr_sample(const rvmc_pop_params* p, uint64_t seed, int nstars, int mass_mode, const double* masses,
                         const double* mass_errors, uint64_t first_sample, int64_t n_samples, uint32_t given_mask,
                         rvmc_orbit_soa out, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(p != nullptr, "population parameters are NULL");
    RVMC_REQUIRE(nstars >= 0 && nstars < (1 << 24), "nstars must be in [0, 2^24)");
    RVMC_REQUIRE(n_samples >= 0, "number_of_samples must be >= 0");
    RVMC_REQUIRE(mass_mode >= 0 && mass_mode <= 2, "mass_mode must be 0, 1 or 2");
    RVMC_REQUIRE(mass_mode == 0 || masses != nullptr, "masses are required for mass_mode %d", mass_mode);
    RVMC_REQUIRE(mass_mode != 1 || mass_errors != nullptr, "mass_errors are required for mass_mode 1");
    RVMC_REQUIRE(first_sample + (uint64_t)n_samples <= ((uint64_t)1 << 40), "sample index must stay below 2^40");
    PopF64 P;
    int rc = derive_pop<double>(*p, P);
    if (rc) {
        set_error("float division by zero");
        return rc;
    }
    const int64_t total = (int64_t)nstars * n_samples;
    if (total == 0) return RVMC_OK;
    RVMC_REQUIRE(!(given_mask & RVMC_GIVEN_PERIOD) || out.period, "given period is NULL");
    RVMC_REQUIRE(!(given_mask & (RVMC_GIVEN_TIME | RVMC_GIVEN_ECC)), "time / eccentricity cannot be given here");
    const int threads = 128;
    biascorr_sample_kernel<<<grid_stride_blocks(total, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        P, seed, nstars, mass_mode, masses, mass_errors, first_sample, n_samples, given_mask, out);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_kepler_epochs_f64(int n_epochs, int64_t n_samples, const double* time_deltas, const double* period_row,
                           const double* ecc_row, int quirk_phase, double* E_out, uint32_t* nonconverged,
                           void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n_epochs >= 0 && n_samples >= 0, "bad shape");
    const int64_t total = (int64_t)n_epochs * n_samples;
    if (total == 0) return RVMC_OK;
    RVMC_REQUIRE(time_deltas && period_row && ecc_row && E_out, "array pointers are required");
    const int threads = 128;
    kepler_epochs_f64_kernel<<<grid_stride_blocks(total, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        n_epochs, n_samples, time_deltas, period_row, ecc_row, quirk_phase, E_out, nonconverged);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_rv_epochs_f64(int n_epochs, int64_t n_samples, const double* E, const double* vmax_row,
                       const double* ecc_row, const double* incl_row, const double* omega_row, double* rv_out,
                       void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n_epochs >= 0 && n_samples >= 0, "bad shape");
    const int64_t total = (int64_t)n_epochs * n_samples;
    if (total == 0) return RVMC_OK;
    RVMC_REQUIRE(E && vmax_row && ecc_row && incl_row && omega_row && rv_out, "array pointers are required");
    const int threads = 128;
    rv_epochs_f64_kernel<<<grid_stride_blocks(total, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        n_epochs, n_samples, E, vmax_row, ecc_row, incl_row, omega_row, rv_out);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_biascorr_replay_f64(int nstars, int64_t n_samples, int max_epochs, const int32_t* n_epochs,
                             const double* t_obs, rvmc_orbit_soa in, int64_t mass_cols, const float* rv_err,
                             uint64_t seed, uint64_t first_sample, int quirk_phase, double* rv_out,
                             double* delta_rv, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(nstars >= 0 && nstars < (1 << 24), "nstars must be in [0, 2^24)");
    RVMC_REQUIRE(n_samples >= 0, "number_of_samples must be >= 0");
    RVMC_REQUIRE(max_epochs >= 1, "max_epochs must be >= 1");
    const int64_t total = (int64_t)nstars * n_samples;
    if (total == 0) return RVMC_OK;
    RVMC_REQUIRE(n_epochs && t_obs, "n_epochs and t_obs are required");
    RVMC_REQUIRE(in.primary_mass && in.period && in.time && in.mass_ratio && in.eccentricity && in.inclination &&
                     in.orbit_rotation,
                 "primary_mass, period, time, mass_ratio, eccentricity, inclination and orbit_rotation are required");
    RVMC_REQUIRE(mass_cols == 1 || mass_cols == n_samples, "primary_mass must have 1 or n_samples columns");
    const int threads = 128;
    biascorr_replay_f64_kernel<<<grid_stride_blocks(total, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        nstars, n_samples, max_epochs, n_epochs, t_obs, in, mass_cols, rv_err, seed, first_sample, quirk_phase,
        rv_out, delta_rv);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_biascorr_singles(uint64_t seed, int nstars, int max_epochs, const int32_t* n_epochs, const float* rv_err,
                          uint64_t first_single, int64_t n_singles, double* delta_out, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(nstars >= 0 && nstars < (1 << 24), "nstars must be in [0, 2^24)");
    RVMC_REQUIRE(n_singles >= 0, "n_singles must be >= 0");
    RVMC_REQUIRE(max_epochs >= 1, "max_epochs must be >= 1");
    const int64_t total = (int64_t)nstars * n_singles;
    if (total == 0) return RVMC_OK;
    RVMC_REQUIRE(n_epochs && rv_err && delta_out, "n_epochs, rv_err and delta_out are required");
    const int threads = 128;
    biascorr_singles_kernel<<<grid_stride_blocks(total, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        seed, nstars, max_epochs, n_epochs, rv_err, first_single, n_singles, delta_out);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_biascorr_detect(int nstars, int64_t n_samples, int max_epochs, const int32_t* n_epochs, const double* rv,
                         const float* rv_err, double delta_rv_min, double n_sigma, uint8_t* detected,
                         void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(nstars >= 0 && n_samples >= 0 && max_epochs >= 1, "bad shape");
    const int64_t total = (int64_t)nstars * n_samples;
    if (total == 0) return RVMC_OK;
    RVMC_REQUIRE(n_epochs && rv && rv_err && detected, "n_epochs, rv, rv_err and detected are required");
    const int threads = 128;
    biascorr_detect_kernel<<<grid_stride_blocks(total, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        nstars, n_samples, max_epochs, n_epochs, rv, rv_err, delta_rv_min, n_sigma, detected);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

}  // extern "C"
