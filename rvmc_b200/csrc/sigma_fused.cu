// K1-K4 fused: sample + Kepler + RV + per-cluster dispersion in one pass, and K6 (grid search)
// built on it.
//
// Reference code replaced (file:line under /root/reference):
//   RVMC.__init__ samplers RVMC.py:79-101, calculate_binary_velocities :214-226,
//   calculate_radial_velocity :228-236 and subsample :238-273 in the infinite-pool limit
//   (SURVEY 7.2): every star of a cluster realization is a binary with probability fbin -- a fresh
//   orbit plus N(0, sigma_dyn) -- else a single N(0, sigma_dyn); plus N(0, err_j);
//   create_sig1D_distribution findBestDistributions.py:21-32 (ddof = 0 there);
//   the statistics of get_Pdistr_stats / get_fbinDist_stats :47-60 / :86-99.
//
// Kernel structure (one tile of R realizations x S stars per block, T = R*S star slots):
//   phase A  one thread per PAIR of slots: one Philox block -> two binary/single coins + one
//            Box-Muller pair; noise goes to v[slot] in shared memory, binary slots are appended to
//            a shared work queue (warp-aggregated atomics);
//   phase B  the queue is consumed 32 binaries per warp at a time: orbit sampling, Kepler solve
//            and RV run on fully populated warps whatever fbin is (no lane idles on a single
//            star); v[slot] += RV;
//   phase C  cluster_reduce.cuh: two-pass dispersion per realization by sub-warp groups.
// Every value is a function of (seed, global slot index) only, so the result does not depend on
// the tile size, the launch geometry or the GPU the tile ran on.  Per-star values never reach HBM.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "cluster_reduce.cuh"
#include "pop_derive.h"

namespace rvmc {

struct FusedPoint {
    PopF32 pop;
    uint32_t seed_lo, seed_hi;
    uint32_t coin_thresh;      // binary iff (word >> 8) < coin_thresh ; coin_thresh = round(fbin * 2^24)
    int32_t kepler_iters;
    uint32_t sub;              // Philox counter word 3 of this point's draws (rvmc_pop_params.substream)
};

struct FusedArgs {
    FusedPoint single;         // the population when points == nullptr (single-population launch)
    PhiloxKey key;             // expanded key schedule (constant-bank operands of the Philox rounds) of the
                               // single population, or of the seed ALL points of a grid launch share (KEYED)
    const FusedPoint* points;  // device, one per grid point, or nullptr
    const float* meas_err;     // device [S]
    float* sigma_out;          // [n_points][sigma_stride]
    unsigned long long* counters;   // [4] or nullptr: binary RVs, guard lanes, non-converged lanes
    uint32_t* fail;                 // [rows][2] or nullptr: per grid point {guard lanes, non-converged lanes}
    uint64_t first_real;
    int64_t n_real;            // realizations per point
    int64_t sigma_stride;
    int32_t S, R, G, tiles_per_point, weighted, ddof;
};

__device__ __forceinline__ uint64_t seed_of(const FusedPoint& p) {
    return ((uint64_t)p.seed_hi << 32) | p.seed_lo;
}

// The slot's noise scale: the two independent normals N(0, sigma_dyn) (RVMC.py:101,224) and
// N(0, err_j) (RVMC.py:265,271) are drawn as one normal of the summed variance.
__device__ __forceinline__ float slot_scale(float sigma_dyn, float err) {
    return sqrtf(fmaf(sigma_dyn, sigma_dyn, err * err));
}

// SINGLE: one population for the whole launch, read straight from the kernel parameters (constant
//         bank: no staging, no registers spent on it, pre-expanded Philox keys); otherwise one
//         descriptor per grid point, staged in shared memory.
// KEYED : every point of the launch has the same seed (the points differ in their substream, i.e. in
//         the Philox counter): the round keys are the launch-uniform a.key, as for SINGLE, instead of
//         two integer adds per round and Philox block for a per-point key schedule.
// ITERS : Halley steps as a compile-time constant (0 = run-time kepler_iters).
// GUARD : keep the fp64 guard path of the Kepler solve (false when proven unreachable on the host).
template <bool SINGLE, bool KEYED, int ITERS, bool GUARD>
__global__ void __launch_bounds__(256, SINGLE ? 3 : 4) sigma1d_fused_kernel(FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.S;
    const int T = a.R * S;
    float* v = reinterpret_cast<float*>(smem_raw);               // [T]
    float* scale = v + T;                                        // [S]
    float* wts = scale + S;                                      // [S]
    uint16_t* queue = reinterpret_cast<uint16_t*>(wts + S);      // [T]
    __shared__ FusedPoint pt_s;
    __shared__ int qn;
    __shared__ float wsum_s;
    __shared__ unsigned int cnt_guard, cnt_bad;

    const int64_t tile = blockIdx.x;
    const int64_t point = tile / a.tiles_per_point;
    const int64_t r0 = (tile - point * a.tiles_per_point) * a.R;
    const int nr = (int)min((int64_t)a.R, a.n_real - r0);
    const int nslots = nr * S;

    // stage the point descriptor and the per-star tables
    if (!SINGLE) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.points + point);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&pt_s);
        for (int i = threadIdx.x; i < (int)(sizeof(FusedPoint) / 4); i += blockDim.x) dst[i] = src[i];
    }
    const FusedPoint& pt = SINGLE ? a.single : pt_s;
    if (threadIdx.x == 0) {
        qn = 0;
        cnt_guard = 0;
        cnt_bad = 0;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const float e = a.meas_err[j];
        scale[j] = slot_scale(pt.pop.sigma_dyn, e);
        wts[j] = 1.0f / (e * e);
    }
    __syncthreads();
    if (a.weighted && threadIdx.x == 0) {
        float s = 0.0f;
        for (int j = 0; j < S; ++j) s += wts[j];
        wsum_s = s;
    }

    const uint64_t seed = seed_of(pt);
    const uint32_t sub = pt.sub;
    auto draw = [&](uint64_t item, uint32_t stream) {
        return (SINGLE || KEYED) ? philox_block(a.key, item, stream, sub) : philox_block(seed, item, stream, sub);
    };
    const uint64_t slot0 = (a.first_real + (uint64_t)r0) * (uint64_t)S;
    const uint64_t p_begin = slot0 >> 1;
    const int npairs = (int)(((slot0 + nslots + 1) >> 1) - p_begin);
    const float invS = 1.0f / (float)S;
    const uint32_t thresh = pt.coin_thresh;
    const int lane = threadIdx.x & 31;

    // ---- phase A: coins + noise ---------------------------------------------------------------
    const int npairs_pad = (npairs + 31) & ~31;      // whole warps stay in the loop for the ballots
    for (int pi = threadIdx.x; pi < npairs_pad; pi += blockDim.x) {
        bool bin0 = false, bin1 = false;
        int l0 = 0, l1 = 0;
        if (pi < npairs) {
            const uint64_t pair = p_begin + pi;
            const u32x4 w = draw(pair, RVMC_STREAM_SLOT);
            float z0, z1;
            box_muller_noise_f32(w.z, w.w, z0, z1);
            const int64_t la = (int64_t)(2 * pair - slot0);
            l0 = (int)la;
            l1 = l0 + 1;
            RVMC_DEV_ASSERT(l0 >= -1 && l0 < nslots && l1 <= nslots);
            if (l0 >= 0) {          // l0 < nslots always (pair range)
                const int r = (int)(((float)l0 + 0.5f) * invS);
                RVMC_DEV_ASSERT(r >= 0 && r < nr && l0 - r * S >= 0 && l0 - r * S < S);
                v[l0] = z0 * scale[l0 - r * S];
                bin0 = (w.x >> 8) < thresh;
            }
            if (l1 < nslots) {
                const int r = (int)(((float)l1 + 0.5f) * invS);
                RVMC_DEV_ASSERT(r >= 0 && r < nr && l1 - r * S >= 0 && l1 - r * S < S);
                v[l1] = z1 * scale[l1 - r * S];
                bin1 = (w.y >> 8) < thresh;
            }
        }
        // warp-aggregated append of the binary slots
        const unsigned m0 = __ballot_sync(0xffffffffu, bin0);
        const unsigned m1 = __ballot_sync(0xffffffffu, bin1);
        const int c0 = __popc(m0), c1 = __popc(m1);
        int base = 0;
        if (lane == 0 && (c0 + c1)) base = atomicAdd(&qn, c0 + c1);
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned below = (1u << lane) - 1u;
        RVMC_DEV_ASSERT(base >= 0 && base + c0 + c1 <= nslots);
        if (bin0) queue[base + __popc(m0 & below)] = (uint16_t)l0;
        if (bin1) queue[base + c0 + __popc(m1 & below)] = (uint16_t)l1;
    }
    __syncthreads();

    // ---- phase B: orbits of the queued binaries -------------------------------------------------
    const int nq = qn;
    RVMC_DEV_ASSERT(nq >= 0 && nq <= nslots);
    unsigned int my_guard = 0, my_bad = 0;
    for (int k = threadIdx.x; k < nq; k += blockDim.x) {
        const int l = queue[k];
        RVMC_DEV_ASSERT(l >= 0 && l < nslots);
        const uint64_t slot = slot0 + (uint64_t)l;
        const u32x4 wa = draw(slot, RVMC_STREAM_ORBIT_A);
        const u32x4 wb = draw(slot, RVMC_STREAM_ORBIT_B);
        const OrbitF32 o = sample_orbit_f32(pt.pop, wa, wb);
        int g, b;
        const float rv = orbit_rv_f32<ITERS, GUARD>(pt.pop, o, pt.kepler_iters, nullptr, g, b);
        v[l] += rv;
        my_guard += g;
        my_bad += b;
    }
    if (my_guard) atomicAdd(&cnt_guard, my_guard);
    if (my_bad) atomicAdd(&cnt_bad, my_bad);
    __syncthreads();

    // ---- phase C: dispersion per realization ----------------------------------------------------
    reduce_tile_sigma<float, float>(v, nr, S, a.G, a.weighted ? wts : nullptr, wsum_s, a.ddof,
                                    a.sigma_out + point * a.sigma_stride + r0);
    if (a.counters && threadIdx.x == 0) {
        atomicAdd(a.counters + 0, (unsigned long long)nq);
        if (cnt_guard) atomicAdd(a.counters + 1, (unsigned long long)cnt_guard);
        if (cnt_bad) atomicAdd(a.counters + 2, (unsigned long long)cnt_bad);
    }
    // per-point failure counts (rvmc_point_stats.n_guard / n_nonconverged): what scipy.optimize.newton
    // reports per call as a RuntimeWarning (RVMC.py:177-180)
    if (a.fail && threadIdx.x == 0 && (cnt_guard | cnt_bad)) {
        if (cnt_guard) atomicAdd(a.fail + 2 * point, cnt_guard);
        if (cnt_bad) atomicAdd(a.fail + 2 * point + 1, cnt_bad);
    }
}


// ---------------------------------------------------------------------------------------------
// Binary-fraction axis of a grid in ONE pass
// ---------------------------------------------------------------------------------------------
// Grid points that differ only in f_bin and share a seed consume identical draws: the same star
// slots, the same orbits, the same noise -- only the coin threshold differs.  This kernel evaluates
// every slot once (orbit solved iff its coin is below the LARGEST threshold) and then reduces the
// tile once per binary fraction, selecting per slot  coin < thresh_k ? noise + rv : noise.  The
// result for fraction k is BIT-IDENTICAL to sigma1d_fused_kernel run with fbin = fbins[k] and the
// same seed (same integer coin compare, same fp32 adds, same reduction order), so this is purely an
// execution strategy: an n_f-point axis costs about as much as its f_bin = max point alone.
#define FBINS_MAX 32
#define FB_THREADS 256          // block size as a compile-time constant: the loops' trip counts need no division
struct FbinArgs {
    FusedArgs base;                 // points: one descriptor per (other-axes) point; coin_thresh unused
    int32_t n_fbins;
    uint32_t thresh[FBINS_MAX];     // ascending or not: thresh_max is computed on the host
    uint32_t thresh_max;
};

template <bool KEYED, int ITERS, bool GUARD>
__global__ void __launch_bounds__(FB_THREADS, 4) sigma1d_fused_fbins_kernel(FbinArgs fa) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FusedArgs& a = fa.base;
    const int S = a.S;
    const int T = a.R * S;
    float* noise = reinterpret_cast<float*>(smem_raw);           // [T]
    float* rv = noise + T;                                       // [T]
    uint32_t* coin = reinterpret_cast<uint32_t*>(rv + T);        // [T] 24-bit coin of every slot
    float* scale = reinterpret_cast<float*>(coin + T);           // [S]
    float* wts = scale + S;                                      // [S]
    uint16_t* queue = reinterpret_cast<uint16_t*>(wts + S);      // [T]
    __shared__ FusedPoint pt;
    __shared__ int qn;
    __shared__ float wsum_s;
    __shared__ unsigned int cnt_guard, cnt_bad;
    __shared__ unsigned int fail_k[2 * FBINS_MAX];       // per fraction: guard lanes, non-converged lanes

    // 32-bit division: a launch holds < 2^31 tiles (checked on the host); the 64-bit one costs ~70 instructions
    const uint32_t tile = blockIdx.x;
    const uint32_t point_u = tile / (uint32_t)a.tiles_per_point;
    const int64_t point = point_u;
    const int64_t r0 = (int64_t)(tile - point_u * (uint32_t)a.tiles_per_point) * a.R;
    const int nr = (int)min((int64_t)a.R, a.n_real - r0);
    const int nslots = nr * S;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.points + point);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&pt);
        for (int i = threadIdx.x; i < (int)(sizeof(FusedPoint) / 4); i += FB_THREADS) dst[i] = src[i];
    }
    if (threadIdx.x < 2 * FBINS_MAX) fail_k[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        qn = 0;
        cnt_guard = 0;
        cnt_bad = 0;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < S; j += FB_THREADS) {
        const float e = a.meas_err[j];
        scale[j] = slot_scale(pt.pop.sigma_dyn, e);
        wts[j] = 1.0f / (e * e);
    }
    __syncthreads();
    if (a.weighted && threadIdx.x == 0) {
        float s = 0.0f;
        for (int j = 0; j < S; ++j) s += wts[j];
        wsum_s = s;
    }
    const uint64_t seed = seed_of(pt);
    const uint32_t sub = pt.sub;
    auto draw = [&](uint64_t item, uint32_t stream) {
        return KEYED ? philox_block(a.key, item, stream, sub) : philox_block(seed, item, stream, sub);
    };
    const uint64_t slot0 = (a.first_real + (uint64_t)r0) * (uint64_t)S;
    const uint64_t p_begin = slot0 >> 1;
    const int npairs = (int)(((slot0 + nslots + 1) >> 1) - p_begin);
    const float invS = 1.0f / (float)S;
    const uint32_t thresh = fa.thresh_max;
    const int lane = threadIdx.x & 31;
    // largest fraction = 1: every slot is a binary, the work queue is the identity and is skipped
    const bool all_binaries = thresh >= (1u << 24);

    // ---- phase A: coins + noise (as sigma1d_fused_kernel, the coin itself is kept) ---------------
    const int npairs_pad = (npairs + 31) & ~31;
    for (int pi = threadIdx.x; pi < npairs_pad; pi += FB_THREADS) {
        bool bin0 = false, bin1 = false;
        int l0 = 0, l1 = 0;
        if (pi < npairs) {
            const uint64_t pair = p_begin + pi;
            const u32x4 w = draw(pair, RVMC_STREAM_SLOT);
            float z0, z1;
            box_muller_noise_f32(w.z, w.w, z0, z1);
            l0 = (int)(int64_t)(2 * pair - slot0);
            l1 = l0 + 1;
            RVMC_DEV_ASSERT(l0 >= -1 && l0 < nslots && l1 <= nslots);
            if (l0 >= 0) {
                const int r = (int)(((float)l0 + 0.5f) * invS);
                noise[l0] = z0 * scale[l0 - r * S];
                coin[l0] = w.x >> 8;
                bin0 = (w.x >> 8) < thresh;
            }
            if (l1 < nslots) {
                const int r = (int)(((float)l1 + 0.5f) * invS);
                noise[l1] = z1 * scale[l1 - r * S];
                coin[l1] = w.y >> 8;
                bin1 = (w.y >> 8) < thresh;
            }
        }
        if (all_binaries) continue;
        const unsigned m0 = __ballot_sync(0xffffffffu, bin0);
        const unsigned m1 = __ballot_sync(0xffffffffu, bin1);
        const int c0 = __popc(m0), c1 = __popc(m1);
        int base = 0;
        if (lane == 0 && (c0 + c1)) base = atomicAdd(&qn, c0 + c1);
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned below = (1u << lane) - 1u;
        RVMC_DEV_ASSERT(base >= 0 && base + c0 + c1 <= nslots);
        if (bin0) queue[base + __popc(m0 & below)] = (uint16_t)l0;
        if (bin1) queue[base + c0 + __popc(m1 & below)] = (uint16_t)l1;
    }
    __syncthreads();

    // ---- phase B: one orbit per slot that is a binary under the largest fraction -------------------
    const int nq = all_binaries ? nslots : qn;
    unsigned int my_guard = 0, my_bad = 0;
    for (int k = threadIdx.x; k < nq; k += FB_THREADS) {
        const int l = all_binaries ? k : queue[k];
        RVMC_DEV_ASSERT(l >= 0 && l < nslots);
        const uint64_t slot = slot0 + (uint64_t)l;
        const u32x4 wa = draw(slot, RVMC_STREAM_ORBIT_A);
        const u32x4 wb = draw(slot, RVMC_STREAM_ORBIT_B);
        const OrbitF32 o = sample_orbit_f32(pt.pop, wa, wb);
        int g, b;
        rv[l] = orbit_rv_f32<ITERS, GUARD>(pt.pop, o, pt.kepler_iters, nullptr, g, b);
        my_guard += g;
        my_bad += b;
        if (g | b) {
            // (never with the reference's eccentricity bounds) the orbit counts for every fraction whose
            // coin threshold makes this slot a binary
            const uint32_t c = coin[l];
            for (int k = 0; k < fa.n_fbins; ++k)
                if (c < fa.thresh[k]) {
                    if (g) atomicAdd(&fail_k[2 * k], 1u);
                    if (b) atomicAdd(&fail_k[2 * k + 1], 1u);
                }
        }
    }
    if (my_guard) atomicAdd(&cnt_guard, my_guard);
    if (my_bad) atomicAdd(&cnt_bad, my_bad);
    __syncthreads();

    // ---- phase C: one reduction per binary fraction -------------------------------------------------
    if (a.G == 1 && !a.weighted) {
        // small clusters: a work item is (realization, group of eight fractions) -- one sweep over the row
        // with the eight accumulator pairs in registers; same operations in the same order as the generic
        // reducer.  Items are dealt group-major, so a warp holds 32 consecutive realizations of one group
        // (coalesced stores); the host sizes the tile so that the items fill the block's threads.
        // Everything around the sweep is kept off the per-item path (it was 11 % of the kernel's instructions,
        // profiles/r02_grid_kernels_hot_lines.txt): no integer division for (group, realization), thresholds
        // padded with zeros by the host instead of eight bounds tests, one output pointer walked by the row
        // stride instead of eight 64-bit index products, and unconditional stores for a full group.
        const int ngroups = (fa.n_fbins + 7) >> 3;
        const SigmaNorm<float> norm(S, a.ddof);
        const int nitems = nr * ngroups;
        float* const out_tile = a.sigma_out + (size_t)point * fa.n_fbins * a.sigma_stride + r0;
        for (int item = threadIdx.x; item < nitems; item += FB_THREADS) {
            int grp = 0, r = item;
            while (r >= nr) {                    // ngroups <= FBINS_MAX / 8 = 4
                r -= nr;
                ++grp;
            }
            const int k0 = grp << 3;
            const int base = r * S;
            float s1[8], s2[8], pivot[8];
            uint32_t thr[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) thr[k] = fa.thresh[k0 + k];      // 0 beyond n_fbins: never a binary, never stored
            {
                const float nz = noise[base], both = nz + rv[base];
                const uint32_t c = coin[base];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    pivot[k] = c < thr[k] ? both : nz;
                    s1[k] = 0.0f;
                    s2[k] = 0.0f;
                }
            }
            for (int j = 0; j < S; ++j) {
                const float nz = noise[base + j], both = nz + rv[base + j];
                const uint32_t c = coin[base + j];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float d = (c < thr[k] ? both : nz) - pivot[k];
                    s1[k] += d;
                    s2[k] = fmaf(d, d, s2[k]);
                }
            }
            float* o = out_tile + (size_t)k0 * a.sigma_stride + r;
            if (k0 + 8 <= fa.n_fbins) {
#pragma unroll
                for (int k = 0; k < 8; ++k, o += a.sigma_stride) *o = sigma_from_sums(s1[k], s2[k], norm);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k, o += a.sigma_stride)
                    if (k0 + k < fa.n_fbins) *o = sigma_from_sums(s1[k], s2[k], norm);
            }
        }
    } else
    for (int k = 0; k < fa.n_fbins; ++k) {
        const uint32_t thr = fa.thresh[k];
        reduce_tile_sigma_fn<float, float>(
            [=](int l) { return coin[l] < thr ? noise[l] + rv[l] : noise[l]; }, nr, S, a.G,
            a.weighted ? wts : nullptr, wsum_s, a.ddof, a.sigma_out + (point * fa.n_fbins + k) * a.sigma_stride + r0);
    }
    if (a.counters && threadIdx.x == 0) {
        atomicAdd(a.counters + 0, (unsigned long long)nq);
        if (cnt_guard) atomicAdd(a.counters + 1, (unsigned long long)cnt_guard);
        if (cnt_bad) atomicAdd(a.counters + 2, (unsigned long long)cnt_bad);
    }
    if (a.fail && (cnt_guard | cnt_bad) && threadIdx.x < 2 * fa.n_fbins && fail_k[threadIdx.x])
        atomicAdd(a.fail + 2 * point * fa.n_fbins + threadIdx.x, fail_k[threadIdx.x]);
}

// Per-slot dump of what the fused kernel computes (parity/debug entry): same device functions,
// same draws; the orbit is evaluated for every slot, `is_binary` says whether K4 would use it.
__global__ void fused_trace_kernel(FusedPoint pt, const float* __restrict__ meas_err, int S, uint64_t first_slot,
                                   int64_t n, rvmc_trace_soa out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint64_t seed = seed_of(pt);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t slot = first_slot + (uint64_t)i;
        const u32x4 ws = philox_block(seed, slot >> 1, RVMC_STREAM_SLOT, pt.sub);
        const u32x4 wa = philox_block(seed, slot, RVMC_STREAM_ORBIT_A, pt.sub);
        const u32x4 wb = philox_block(seed, slot, RVMC_STREAM_ORBIT_B, pt.sub);
        float z0, z1;
        box_muller_noise_f32(ws.z, ws.w, z0, z1);
        const int odd = (int)(slot & 1);
        const int j = (int)(slot % (uint64_t)S);
        const float noise = (odd ? z1 : z0) * slot_scale(pt.pop.sigma_dyn, meas_err[j]);
        const bool is_bin = ((odd ? ws.y : ws.x) >> 8) < pt.coin_thresh;
        const OrbitF32 o = sample_orbit_f32(pt.pop, wa, wb);
        int g, b;
        float K;
        const float rv = orbit_rv_f32(pt.pop, o, pt.kepler_iters, &K, g, b);
        if (out.words) {
            uint32_t* w = out.words + 12 * i;
            w[0] = wa.x; w[1] = wa.y; w[2] = wa.z; w[3] = wa.w;
            w[4] = wb.x; w[5] = wb.y; w[6] = wb.z; w[7] = wb.w;
            w[8] = ws.x; w[9] = ws.y; w[10] = ws.z; w[11] = ws.w;
        }
        if (out.log2m) out.log2m[i] = o.log2m;
        if (out.x) out.x[i] = o.x;
        if (out.uphase) out.uphase[i] = o.uphase;
        if (out.q) out.q[i] = o.q;
        if (out.e) out.e[i] = o.e;
        if (out.cosi) out.cosi[i] = o.cosi;
        if (out.wturn) out.wturn[i] = o.wturn;
        if (out.K) out.K[i] = K;
        if (out.rv) out.rv[i] = rv;
        if (out.noise) out.noise[i] = noise;
        if (out.is_binary) out.is_binary[i] = is_bin ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// K6: per-point statistics of the sigma_1D distribution (findBestDistributions.py:47-60)
// ---------------------------------------------------------------------------------------------
// One block per point.  Exact order statistics by a most-significant-digit radix select on the
// float bit patterns (sigma >= 0, so the unsigned order is the numeric order) in three passes of
// 12 + 10 + 10 bits over the point's sigma array (three streaming reads: a batch of rows is larger
// than the L2, see point_stats_cluster_kernel for the single-read form).  The ten ranks needed for the 5/16/50/84/95 percentiles (NumPy 'linear' rule: virtual
// index q/100*(n-1), neighbours floor and floor+1) are selected together:
//   pass 0  one 4096-bin histogram of the top 12 bits (run-length cached per thread: neighbouring
//           values share them) + the sum (fp64) and the maximum;
//   pass 1  elements whose top 12 bits carry a target (one byte-table lookup) go to that prefix
//           group's 1024-bin histogram of the next 10 bits; the `bin`-wide histogram for the mode
//           (first maximal bin, :49-52) is filled here, now that the maximum is known;
//   pass 2  a second table keyed by (group, middle digit) routes the few remaining candidates to
//           the histogram of the last 10 bits.
// ~45 instructions per element in total; targets that share a prefix share a histogram.
#define STATS_THREADS 512
#define STATS_TARGETS 10
#define STATS_MODE_BINS_MAX 4096
#define STATS_B0 12
#define STATS_B1 10
#define STATS_B2 10
#define STATS_SMEM_BYTES (((STATS_TARGETS << STATS_B1) + STATS_MODE_BINS_MAX) * 4 + (1 << STATS_B0) + (STATS_TARGETS << STATS_B1))

struct StatsArgs {
    const float* sigma;        // [n_points][sigma_stride]
    int64_t sigma_stride;
    int64_t n_real;
    double bin;
    int32_t hist_bins;         // row length of hist_out
    rvmc_point_stats* stats_out;
    uint32_t* hist_out;        // [n_points][hist_bins] or nullptr
    const uint32_t* fail;      // [n_points][2] {guard lanes, non-converged lanes} of the producer, or nullptr
    uint8_t* todo;             // [n_points] or nullptr.  point_stats_fast_kernel sets todo[row] = 1 for the rows it
                               // leaves to the general kernel, which (when todo is given) skips all others
};

__device__ __forceinline__ double block_sum_f64(double x, double* scratch) {
    // deterministic: warp shuffle tree, then thread 0 adds the warp totals in order
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) scratch[w] = x;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += scratch[i];
        scratch[0] = tot;
    }
    __syncthreads();
    tot = scratch[0];
    __syncthreads();
    return tot;
}

// Warp t < STATS_TARGETS (all 32 lanes): locate the digit holding rank r inside histogram
// h[0..nbins), nbins a multiple of 32 -- 32 bins per step: a shuffle scan of their counts, a ballot
// for the first bin whose running total exceeds r.  On return r is the rank inside that bin.
// (A single thread walking the bins one by one took ~90k cycles per block over the three passes.)
__device__ __forceinline__ unsigned int locate_digit(const unsigned int* h, int nbins, unsigned int& r) {
    const int lane = threadIdx.x & 31;
    unsigned int rr = r;
    for (int base = 0; base < nbins; base += 32) {
        const unsigned int c = h[base + lane];
        unsigned int s = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += t;
        }
        const unsigned int total = __shfl_sync(0xffffffffu, s, 31);
        if (rr < total) {
            const int l = __ffs(__ballot_sync(0xffffffffu, rr < s)) - 1;
            r = rr - __shfl_sync(0xffffffffu, s - c, l);
            RVMC_DEV_ASSERT(r < h[base + l]);
            return (unsigned int)(base + l);
        }
        rr -= total;
    }
    r = rr;                     // unreachable for a rank below the histogram's total
    return (unsigned int)nbins - 1u;
}

__global__ void __launch_bounds__(STATS_THREADS, 3) point_stats_kernel(StatsArgs a) {
    // dynamic shared memory (STATS_SMEM_BYTES, 70 KB => 3 blocks per SM); region A is the 4096-bin
    // histogram of pass 0, then the per-group histograms of passes 1 and 2
    extern __shared__ __align__(16) unsigned char stats_smem[];
    unsigned int* histA = reinterpret_cast<unsigned int*>(stats_smem);           // [10 << 10]  40 KB
    unsigned int* mode_hist = histA + (STATS_TARGETS << STATS_B1);               // [4096]      16 KB
    signed char* table0 = reinterpret_cast<signed char*>(mode_hist + STATS_MODE_BINS_MAX);   // top-12 prefix -> group / -1
    signed char* table1 = table0 + (1 << STATS_B0);                              // (group, middle digit) -> group2 / -1
    __shared__ double red[STATS_THREADS / 32];
    __shared__ unsigned int prefix[STATS_TARGETS];      // selected high bits so far, per target
    __shared__ unsigned int rank[STATS_TARGETS];        // rank still to find inside the prefix
    __shared__ int group_of[STATS_TARGETS];
    __shared__ int group0_of[STATS_TARGETS];
    __shared__ unsigned int max_bits;

    if (a.todo && !a.todo[blockIdx.x]) return;      // done by point_stats_fast_kernel (block-uniform)
    const int64_t n = a.n_real;
    const float* __restrict__ sig = a.sigma + (size_t)blockIdx.x * a.sigma_stride;
    rvmc_point_stats* st = a.stats_out + blockIdx.x;
    const double qs[5] = {5.0, 16.0, 50.0, 84.0, 95.0};

    if (threadIdx.x < STATS_TARGETS) {
        const int t = threadIdx.x;
        const double pos = (qs[t >> 1] / 100.0) * (double)(n - 1);
        const int64_t lo = (int64_t)floor(pos);
        rank[t] = (unsigned int)((t & 1) ? min(lo + 1, n - 1) : lo);
        prefix[t] = 0;
    }
    if (threadIdx.x == 0) max_bits = 0;
    for (int i = threadIdx.x; i < (1 << STATS_B0); i += blockDim.x) {
        histA[i] = 0;
        table0[i] = -1;
    }
    for (int i = threadIdx.x; i < STATS_MODE_BINS_MAX; i += blockDim.x) mode_hist[i] = 0;
    for (int i = threadIdx.x; i < (STATS_TARGETS << STATS_B1); i += blockDim.x) table1[i] = -1;
    __syncthreads();

    // four independent loads in flight per thread
    auto for_each = [&](auto visit) {
        int64_t i = threadIdx.x;
        for (; i + 3 * (int64_t)blockDim.x < n; i += 4 * (int64_t)blockDim.x) {
            const float x0 = __ldg(sig + i), x1 = __ldg(sig + i + blockDim.x), x2 = __ldg(sig + i + 2 * blockDim.x),
                        x3 = __ldg(sig + i + 3 * blockDim.x);
            visit(x0);
            visit(x1);
            visit(x2);
            visit(x3);
        }
        for (; i < n; i += blockDim.x) visit(__ldg(sig + i));
    };

    // ---- pass 0: top 12 bits, sum, max ----------------------------------------------------------------
    double sum = 0.0;
    {
        unsigned int mx = 0, run_digit = 0xffffffffu, run_count = 0;
        for_each([&](float x) {
            const unsigned int bits = __float_as_uint(x);
            const unsigned int digit = bits >> (32 - STATS_B0);
            sum += (double)x;
            mx = max(mx, bits);
            if (digit != run_digit) {
                if (run_count) atomicAdd(&histA[run_digit], run_count);
                run_digit = digit;
                run_count = 0;
            }
            ++run_count;
        });
        if (run_count) atomicAdd(&histA[run_digit], run_count);
        atomicMax(&max_bits, mx);
    }
    sum = block_sum_f64(sum, red);       // contains the barriers
    const double smax = (double)__uint_as_float(max_bits);
    // len(np.arange(0, max+bin, bin)) - 1 bins
    const double nbd = ceil((smax + a.bin) / a.bin) - 1.0;
    const int nb_mode = (nbd >= 1.0) ? (nbd < 2.0e9 ? (int)nbd : 2000000000) : 0;   // NaN sigma -> no histogram
    if ((threadIdx.x >> 5) < STATS_TARGETS) {
        const int t = threadIdx.x >> 5;
        unsigned int r = rank[t];
        const unsigned int d = locate_digit(histA, 1 << STATS_B0, r);
        if ((threadIdx.x & 31) == 0) {
            prefix[t] = d;
            rank[t] = r;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ng = 0;
        for (int t = 0; t < STATS_TARGETS; ++t) {
            if (table0[prefix[t]] < 0) table0[prefix[t]] = (signed char)ng++;
            group0_of[t] = table0[prefix[t]];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (STATS_TARGETS << STATS_B1); i += blockDim.x) histA[i] = 0;
    __syncthreads();

    // ---- pass 1: middle 10 bits of the candidates + the mode histogram -----------------------------------
    {
        const bool do_mode = nb_mode > 0 && nb_mode <= STATS_MODE_BINS_MAX;
        const double last_edge = (double)nb_mode * a.bin;
        const float inv_bin = (float)(1.0 / a.bin);
        // a bin width that is a power of two (the reference's 0.5, findBestDistributions.py:47) makes
        // x / bin exact in fp32: floor(x / bin) is the bin and only the closed right end of the last bin
        // (np.histogram) needs a clamp -- no fp64 edge comparison per element
        int bin_exp;
        const bool pow2_bin = frexp(a.bin, &bin_exp) == 0.5 && bin_exp > -100 && bin_exp < 100;
        if (do_mode && pow2_bin) {
            const int last = nb_mode - 1;
            for_each([&](float x) {
                const unsigned int bits = __float_as_uint(x);
                const int g = table0[bits >> (32 - STATS_B0)];
                if (g >= 0) atomicAdd(&histA[(g << STATS_B1) + ((bits >> STATS_B2) & ((1u << STATS_B1) - 1u))], 1u);
                const int k = min((int)(x * inv_bin), last);     // NaN -> 0 here, but then nb_mode == 0 above
                atomicAdd(&mode_hist[k], 1u);
            });
        } else
        for_each([&](float x) {
            const unsigned int bits = __float_as_uint(x);
            const int g = table0[bits >> (32 - STATS_B0)];
            if (g >= 0) atomicAdd(&histA[(g << STATS_B1) + ((bits >> STATS_B2) & ((1u << STATS_B1) - 1u))], 1u);
            if (do_mode) {
                // edges are k*bin (np.arange(0, max+bin, bin)); bin k = [k*bin, (k+1)*bin), the last
                // bin is closed on the right (np.histogram).  fp32 guess, exact fp64 correction.
                const double xd = (double)x;
                int k = (int)(x * inv_bin);
                if ((double)k * a.bin > xd) --k;
                else if ((double)(k + 1) * a.bin <= xd) ++k;
                if (k >= nb_mode && xd <= last_edge) k = nb_mode - 1;
                RVMC_DEV_ASSERT(k >= -1 && k <= nb_mode);
                if (k >= 0 && k < nb_mode) atomicAdd(&mode_hist[k], 1u);
            }
        });
    }
    __syncthreads();
    if ((threadIdx.x >> 5) < STATS_TARGETS) {
        const int t = threadIdx.x >> 5;
        unsigned int r = rank[t];
        const unsigned int d = locate_digit(histA + (group0_of[t] << STATS_B1), 1 << STATS_B1, r);
        if ((threadIdx.x & 31) == 0) {
            rank[t] = r;
            prefix[t] = (prefix[t] << STATS_B1) | d;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ng = 0;
        for (int t = 0; t < STATS_TARGETS; ++t) {
            const int slot = (group0_of[t] << STATS_B1) + (int)(prefix[t] & ((1u << STATS_B1) - 1u));
            if (table1[slot] < 0) table1[slot] = (signed char)ng++;
            group_of[t] = table1[slot];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (STATS_TARGETS << STATS_B2); i += blockDim.x) histA[i] = 0;
    __syncthreads();

    // ---- pass 2: last 10 bits of the few remaining candidates ------------------------------------------
    for_each([&](float x) {
        const unsigned int bits = __float_as_uint(x);
        const int g0 = table0[bits >> (32 - STATS_B0)];
        if (g0 >= 0) {
            const int g = table1[(g0 << STATS_B1) + ((bits >> STATS_B2) & ((1u << STATS_B1) - 1u))];
            if (g >= 0) atomicAdd(&histA[(g << STATS_B2) + (bits & ((1u << STATS_B2) - 1u))], 1u);
        }
    });
    __syncthreads();
    if ((threadIdx.x >> 5) < STATS_TARGETS) {
        const int t = threadIdx.x >> 5;
        unsigned int r = rank[t];
        const unsigned int d = locate_digit(histA + (group_of[t] << STATS_B2), 1 << STATS_B2, r);
        if ((threadIdx.x & 31) == 0) prefix[t] = (prefix[t] << STATS_B2) | d;
    }
    __syncthreads();

    // mode: first maximal bin of the `bin`-wide histogram (:49-52), by warp 0 -- each lane keeps the
    // first maximum of its strided bins, the shuffle tree prefers the larger count, then the lower bin
    int mode_arg = 0;
    if (threadIdx.x < 32 && nb_mode > 0 && nb_mode <= STATS_MODE_BINS_MAX) {
        unsigned int best = 0;
        int arg = 0x7fffffff;
        for (int k = threadIdx.x; k < nb_mode; k += 32) {
            const unsigned int c = mode_hist[k];
            if (c > best) {
                best = c;
                arg = k;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned int ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
            if (ob > best || (ob == best && oa < arg)) {
                best = ob;
                arg = oa;
            }
        }
        mode_arg = (best > 0) ? arg : 0;        // an all-empty histogram: bin 0, as the serial scan gave
    }
    if (threadIdx.x == 0) {
        double pct[5];
        for (int k = 0; k < 5; ++k) {
            // NumPy's _lerp: a + (b-a)*t, evaluated from b when t >= 0.5
            const double pos = (qs[k] / 100.0) * (double)(n - 1);
            const double t = pos - floor(pos);
            const double lo = (double)__uint_as_float(prefix[2 * k]);
            const double hi = (double)__uint_as_float(prefix[2 * k + 1]);
            const double diff = hi - lo;
            // separate roundings, as NumPy's ufuncs do: a contracted FMA differs in the last bit now and then
            pct[k] = (t >= 0.5) ? __dsub_rn(hi, __dmul_rn(diff, 1.0 - t)) : __dadd_rn(lo, __dmul_rn(diff, t));
        }
        st->p05 = pct[0];
        st->p16 = pct[1];
        st->median = pct[2];
        st->p84 = pct[3];
        st->p95 = pct[4];
        st->mean = sum / (double)n;
        st->sigma_max = smax;
        st->n_hist_bins = (uint32_t)nb_mode;
        st->n_guard = a.fail ? a.fail[2 * blockIdx.x] : 0u;
        st->n_nonconverged = a.fail ? a.fail[2 * blockIdx.x + 1] : 0u;
        st->reserved = 0;
        if (nb_mode > 0 && nb_mode <= STATS_MODE_BINS_MAX) {
            st->mode = __dadd_rn(__dmul_rn((double)mode_arg, a.bin), a.bin / 2.0);    // bins[i] + width / 2, rounded like NumPy
        } else {
            // no bin at all (every sigma is 0: the reference's np.max of an empty histogram raises)
            // or more bins than the kernel's table holds (flagged)
            st->mode = nan("");
            st->reserved = nb_mode > STATS_MODE_BINS_MAX ? 1 : 0;
        }
    }
    if (a.hist_out) {
        uint32_t* ho = a.hist_out + (size_t)blockIdx.x * a.hist_bins;
        for (int k = threadIdx.x; k < a.hist_bins; k += blockDim.x)
            ho[k] = (k < nb_mode && k < STATS_MODE_BINS_MAX) ? mode_hist[k] : 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// K6, fast form: TWO reads of the row instead of three, ~3x fewer instructions per element
// ---------------------------------------------------------------------------------------------
// A velocity dispersion in km/s lives in a narrow window of the float range: every value in [2^-10, 2^22)
// has its top 16 bits (sign 0, exponent, 7 mantissa bits) inside a run of 4096 consecutive prefixes.  So ONE
// histogram of 4096 bins resolves 16 bits at once (the general kernel above spends a pass on 12), and the
// bins that hold the ten target ranks are so thin (0.8 % wide: ~10^3 of 10^5 values each) that the second
// read can simply COLLECT their members into shared memory, where two 8-bit steps per target finish the
// selection.  Exact order statistics as before (the same ranks, NumPy's interpolation rule), bit-equal results.
//   pass A  16-bit-prefix histogram + fp64 sum + maximum + the `bin`-wide histogram for the mode;
//           block-wide exclusive scan of the histogram, one binary search per target rank;
//   pass B  a byte table maps a prefix to its candidate list (or -1: 1 load + 1 compare per element);
//           then one warp per target: two 256-bin histograms over its list.
// One block of 1024 threads per row and ONE block per SM (110 KB of shared memory): 148 rows = 59 MB are in
// flight at a time, so the second read of a row finds it in the 126 MB L2 -- a row crosses the HBM interface
// once.  Rows this form cannot take -- a target rank among values outside the window (zeros, denormals,
// > 4e6 km/s), NaNs, more candidates than the list holds -- are flagged in `todo` and done by the general
// kernel, which is launched right behind and returns at once for every other row.
#define FS_THREADS 1024
#define FS_BINS 4096
#define FS_LO16 (117u << 7)                 // bits >> 16 of 2^-10
#define FS_CAND_MAX 16384
#define FS_MODE_BINS (STATS_MODE_BINS_MAX + 1)
#define FS_SMEM_BYTES ((FS_BINS + FS_MODE_BINS + FS_CAND_MAX + STATS_TARGETS * 256) * 4 + FS_BINS)

__global__ void __launch_bounds__(FS_THREADS, 1) point_stats_fast_kernel(StatsArgs a) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    unsigned int* hist = reinterpret_cast<unsigned int*>(fs_smem);                 // [4096] 16-bit prefixes
    unsigned int* mode_hist = hist + FS_BINS;                                      // [4097]
    unsigned int* cand = mode_hist + FS_MODE_BINS;                                 // [16384]; first the scan of hist
    unsigned int* sub = cand + FS_CAND_MAX;                                        // [10][256]
    signed char* table = reinterpret_cast<signed char*>(sub + STATS_TARGETS * 256);   // prefix -> list / -1
    __shared__ double red[FS_THREADS / 32];
    __shared__ unsigned int warp_tot[FS_THREADS / 32];
    __shared__ unsigned int rank[STATS_TARGETS], tbin[STATS_TARGETS], value_bits[STATS_TARGETS];
    __shared__ int grp[STATS_TARGETS];
    __shared__ unsigned int seg_off[STATS_TARGETS], seg_cnt[STATS_TARGETS], seg_fill[STATS_TARGETS];
    __shared__ unsigned int max_bits, n_below, n_above;
    __shared__ int give_up;

    const int64_t n = a.n_real;
    const float* __restrict__ sig = a.sigma + (size_t)blockIdx.x * a.sigma_stride;
    rvmc_point_stats* st = a.stats_out + blockIdx.x;
    const double qs[5] = {5.0, 16.0, 50.0, 84.0, 95.0};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x < STATS_TARGETS) {
        const int t = threadIdx.x;
        const double pos = (qs[t >> 1] / 100.0) * (double)(n - 1);
        const int64_t lo = (int64_t)floor(pos);
        rank[t] = (unsigned int)((t & 1) ? min(lo + 1, n - 1) : lo);
        seg_fill[t] = 0;
    }
    if (threadIdx.x == 0) {
        max_bits = 0;
        n_below = 0;
        n_above = 0;
        give_up = 0;
    }
    for (int i = threadIdx.x; i < FS_BINS; i += FS_THREADS) {
        hist[i] = 0;
        table[i] = -1;
    }
    for (int i = threadIdx.x; i < FS_MODE_BINS; i += FS_THREADS) mode_hist[i] = 0;
    for (int i = threadIdx.x; i < STATS_TARGETS * 256; i += FS_THREADS) sub[i] = 0;
    __syncthreads();

    // 16-byte loads, two in flight per thread; scalar head/tail for rows that do not start on 16 bytes
    const int head = (int)min((int64_t)(((16u - (unsigned int)((uintptr_t)sig & 15u)) & 15u) >> 2), n);
    const float4* __restrict__ sig4 = reinterpret_cast<const float4*>(sig + head);
    const int64_t n4 = (n - head) >> 2;
    const int64_t tail0 = head + 4 * n4;
    auto for_each = [&](auto visit) {
        if (threadIdx.x < head) visit(__ldg(sig + threadIdx.x));
        int64_t j = threadIdx.x;
        for (; j + FS_THREADS < n4; j += 2 * FS_THREADS) {
            const float4 u = __ldg(sig4 + j), v = __ldg(sig4 + j + FS_THREADS);
            visit(u.x); visit(u.y); visit(u.z); visit(u.w);
            visit(v.x); visit(v.y); visit(v.z); visit(v.w);
        }
        if (j < n4) {
            const float4 u = __ldg(sig4 + j);
            visit(u.x); visit(u.y); visit(u.z); visit(u.w);
        }
        if (tail0 + threadIdx.x < n) visit(__ldg(sig + tail0 + threadIdx.x));
    };

    // ---- pass A ------------------------------------------------------------------------------------------
    double sum = 0.0;
    {
        unsigned int mx = 0, below = 0, above = 0;
        const float inv_bin = (float)(1.0 / a.bin);
        int bin_exp;
        const bool pow2_bin = frexp(a.bin, &bin_exp) == 0.5 && bin_exp > -100 && bin_exp < 100;
        auto common = [&](float x, unsigned int bits) {
            const unsigned int w = (bits >> 16) - FS_LO16;
            if (w < FS_BINS) atomicAdd(&hist[w], 1u);
            else if (bits < (FS_LO16 << 16)) ++below;
            else ++above;
            sum += (double)x;
            mx = max(mx, bits);
        };
        if (pow2_bin) {
            // x / bin is exact in fp32 for a power-of-two bin (the reference's 0.5): floor is the bin
            for_each([&](float x) {
                const unsigned int bits = __float_as_uint(x);
                common(x, bits);
                atomicAdd(&mode_hist[min((int)(x * inv_bin), STATS_MODE_BINS_MAX)], 1u);
            });
        } else {
            for_each([&](float x) {
                const unsigned int bits = __float_as_uint(x);
                common(x, bits);
                // edges are k*bin (np.arange(0, max+bin, bin)): fp32 guess, exact fp64 correction
                const double xd = (double)x;
                int k = (int)(x * inv_bin);
                if ((double)k * a.bin > xd) --k;
                else if ((double)(k + 1) * a.bin <= xd) ++k;
                atomicAdd(&mode_hist[min(max(k, 0), STATS_MODE_BINS_MAX)], 1u);
            });
        }
        atomicMax(&max_bits, mx);
        for (int off = 16; off > 0; off >>= 1) {
            below += __shfl_xor_sync(0xffffffffu, below, off);
            above += __shfl_xor_sync(0xffffffffu, above, off);
        }
        if (lane == 0) {
            if (below) atomicAdd(&n_below, below);
            if (above) atomicAdd(&n_above, above);
        }
    }
    sum = block_sum_f64(sum, red);       // contains the barriers
    const double smax = (double)__uint_as_float(max_bits);
    const double nbd = ceil((smax + a.bin) / a.bin) - 1.0;           // len(np.arange(0, max+bin, bin)) - 1
    const int nb_mode = (nbd >= 1.0) ? (nbd < 2.0e9 ? (int)nbd : 2000000000) : 0;

    // exclusive scan of the prefix histogram -> cand[0..4096): 4 bins per thread, warp scan, warp totals
    {
        const int b0 = threadIdx.x * 4;
        const unsigned int c0 = hist[b0], c1 = hist[b0 + 1], c2 = hist[b0 + 2], c3 = hist[b0 + 3];
        const unsigned int mine = c0 + c1 + c2 + c3;
        unsigned int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned int wt = warp_tot[lane];
            unsigned int winc = wt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, winc, off);
                if (lane >= off) winc += t;
            }
            warp_tot[lane] = winc - wt;          // exclusive offset of each warp
        }
        __syncthreads();
        const unsigned int ex = warp_tot[warp] + incl - mine;
        cand[b0] = ex;
        cand[b0 + 1] = ex + c0;
        cand[b0 + 2] = ex + c0 + c1;
        cand[b0 + 3] = ex + c0 + c1 + c2;
    }
    __syncthreads();
    // one binary search per target rank: the last bin whose exclusive count is <= r
    if (threadIdx.x < STATS_TARGETS) {
        const int t = threadIdx.x;
        const unsigned int r = rank[t];
        const bool nan_present = max_bits > 0x7f800000u;
        if (nan_present || r < n_below || (int64_t)r >= n - (int64_t)n_above) {
            give_up = 1;
        } else {
            const unsigned int rr = r - n_below;
            int lo = 0, hi = FS_BINS - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (cand[mid] <= rr) lo = mid;
                else hi = mid - 1;
            }
            RVMC_DEV_ASSERT(hist[lo] > 0 && rr - cand[lo] < hist[lo]);
            tbin[t] = (unsigned int)lo;
            rank[t] = rr - cand[lo];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && !give_up) {
        int ng = 0;
        unsigned int total = 0;
        for (int t = 0; t < STATS_TARGETS; ++t) {
            const unsigned int b = tbin[t];
            if (table[b] < 0) {
                table[b] = (signed char)ng;
                seg_off[ng] = total;
                seg_cnt[ng] = hist[b];
                total += hist[b];
                ++ng;
            }
            grp[t] = table[b];
        }
        if (total > FS_CAND_MAX) give_up = 1;
    }
    __syncthreads();
    if (give_up) {                           // block-uniform: the general kernel takes this row
        if (threadIdx.x == 0) a.todo[blockIdx.x] = 1;
        return;
    }

    // ---- pass B: collect the members of the target bins ----------------------------------------------------
    for_each([&](float x) {
        const unsigned int bits = __float_as_uint(x);
        const unsigned int w = (bits >> 16) - FS_LO16;
        if (w < FS_BINS) {
            const int g = table[w];
            if (g >= 0) {
                const unsigned int pos = atomicAdd(&seg_fill[g], 1u);
                RVMC_DEV_ASSERT(pos < seg_cnt[g]);
                cand[seg_off[g] + pos] = bits;
            }
        }
    });
    __syncthreads();
    // one warp per target: two 8-bit steps over its list (the order inside a list does not matter)
    if (warp < STATS_TARGETS) {
        const int t = warp, g = grp[t];
        const unsigned int* list = cand + seg_off[g];
        const unsigned int m = seg_cnt[g];
        unsigned int* h = sub + t * 256;
        unsigned int r = rank[t];
        for (unsigned int i = lane; i < m; i += 32) atomicAdd(&h[(list[i] >> 8) & 255u], 1u);
        __syncwarp();
        const unsigned int d1 = locate_digit(h, 256, r);
        __syncwarp();
        for (int i = lane; i < 256; i += 32) h[i] = 0;
        __syncwarp();
        for (unsigned int i = lane; i < m; i += 32) {
            const unsigned int v = list[i];
            if (((v >> 8) & 255u) == d1) atomicAdd(&h[v & 255u], 1u);
        }
        __syncwarp();
        const unsigned int d0 = locate_digit(h, 256, r);
        if (lane == 0) value_bits[t] = ((tbin[t] + FS_LO16) << 16) | (d1 << 8) | d0;
    }
    // the closed right end of the last bin (np.histogram): a value equal to the last edge counts in the last bin
    if (threadIdx.x == 0 && nb_mode > 0 && nb_mode <= STATS_MODE_BINS_MAX) {
        mode_hist[nb_mode - 1] += mode_hist[nb_mode];
        mode_hist[nb_mode] = 0;
    }
    __syncthreads();

    // mode: first maximal bin (findBestDistributions.py:49-52), as in the general kernel
    int mode_arg = 0;
    if (threadIdx.x < 32 && nb_mode > 0 && nb_mode <= STATS_MODE_BINS_MAX) {
        unsigned int best = 0;
        int arg = 0x7fffffff;
        for (int k = threadIdx.x; k < nb_mode; k += 32) {
            const unsigned int c = mode_hist[k];
            if (c > best) {
                best = c;
                arg = k;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned int ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
            if (ob > best || (ob == best && oa < arg)) {
                best = ob;
                arg = oa;
            }
        }
        mode_arg = (best > 0) ? arg : 0;
    }
    if (threadIdx.x == 0) {
        double pct[5];
        for (int k = 0; k < 5; ++k) {
            // NumPy's _lerp: a + (b-a)*t, evaluated from b when t >= 0.5
            const double pos = (qs[k] / 100.0) * (double)(n - 1);
            const double t = pos - floor(pos);
            const double lo = (double)__uint_as_float(value_bits[2 * k]);
            const double hi = (double)__uint_as_float(value_bits[2 * k + 1]);
            const double diff = hi - lo;
            // separate roundings, as NumPy's ufuncs do: a contracted FMA differs in the last bit now and then
            pct[k] = (t >= 0.5) ? __dsub_rn(hi, __dmul_rn(diff, 1.0 - t)) : __dadd_rn(lo, __dmul_rn(diff, t));
        }
        st->p05 = pct[0];
        st->p16 = pct[1];
        st->median = pct[2];
        st->p84 = pct[3];
        st->p95 = pct[4];
        st->mean = sum / (double)n;
        st->sigma_max = smax;
        st->n_hist_bins = (uint32_t)nb_mode;
        st->n_guard = a.fail ? a.fail[2 * blockIdx.x] : 0u;
        st->n_nonconverged = a.fail ? a.fail[2 * blockIdx.x + 1] : 0u;
        st->reserved = 0;
        if (nb_mode > 0 && nb_mode <= STATS_MODE_BINS_MAX) {
            st->mode = __dadd_rn(__dmul_rn((double)mode_arg, a.bin), a.bin / 2.0);    // bins[i] + width / 2, rounded like NumPy
        } else {
            st->mode = nan("");
            st->reserved = nb_mode > STATS_MODE_BINS_MAX ? 1 : 0;
        }
        a.todo[blockIdx.x] = 0;
    }
    if (a.hist_out) {
        uint32_t* ho = a.hist_out + (size_t)blockIdx.x * a.hist_bins;
        for (int k = threadIdx.x; k < a.hist_bins; k += blockDim.x)
            ho[k] = (k < nb_mode && k < STATS_MODE_BINS_MAX) ? mode_hist[k] : 0u;
    }
}

// Both statistics kernels behind one call: the fast form for every row, the general form for what it left.
static int launch_point_stats(StatsArgs sa, int64_t rows, uint8_t* todo, cudaStream_t stream) {
    const bool general_only = getenv("RVMC_STATS_GENERAL") != nullptr;       // A/B and tests (read per call)
    sa.todo = nullptr;
    if (!general_only && todo) {
        sa.todo = todo;
        point_stats_fast_kernel<<<(unsigned int)rows, FS_THREADS, FS_SMEM_BYTES, stream>>>(sa);
        int rc = cuda_status(cudaGetLastError(), "point_stats_fast_kernel launch");
        if (rc) return rc;
    }
    point_stats_kernel<<<(unsigned int)rows, STATS_THREADS, STATS_SMEM_BYTES, stream>>>(sa);
    return cuda_status(cudaGetLastError(), "point_stats_kernel launch");
}

// Per host thread and device: the side stream the statistics kernels run on and the three events that
// order it against the caller's stream, created once and reused by every grid call of the thread (a grid
// fit makes many calls; creating and destroying them costs more than a batch of a small grid), and the
// one-time opt-in of the two statistics kernels to > 48 KB of shared memory.
struct GridAux {
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fused = nullptr, ev_stats[2] = {nullptr, nullptr};
};
static int stats_kernels_setup();
static int grid_aux(GridAux** out) {
    static thread_local std::vector<GridAux> per_device;
    int dev = 0;
    int rc = cuda_status(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    if ((int)per_device.size() <= dev) per_device.resize(dev + 1);
    GridAux& a = per_device[dev];
    if (!a.side) {
        rc = stats_kernels_setup();
        if (!rc) rc = cuda_status(cudaStreamCreateWithFlags(&a.side, cudaStreamNonBlocking), "cudaStreamCreate");
        if (!rc) rc = cuda_status(cudaEventCreateWithFlags(&a.ev_fused, cudaEventDisableTiming), "cudaEventCreate");
        for (int k = 0; k < 2 && !rc; ++k)
            rc = cuda_status(cudaEventCreateWithFlags(&a.ev_stats[k], cudaEventDisableTiming), "cudaEventCreate");
        if (rc) {
            a.side = nullptr;        // try again on the next call
            return rc;
        }
    }
    *out = &a;
    return RVMC_OK;
}

static int stats_kernels_setup() {
    int rc = cuda_status(cudaFuncSetAttribute(point_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              STATS_SMEM_BYTES), "point_stats_kernel smem");
    if (!rc) rc = cuda_status(cudaFuncSetAttribute(point_stats_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   FS_SMEM_BYTES), "point_stats_fast_kernel smem");
    return rc;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int make_fused_point(const rvmc_pop_params& p, uint64_t seed, FusedPoint* out) {
    int rc = derive_pop<float>(p, out->pop);
    if (rc) {
        set_error("float division by zero");     // RVMC.py:112 with power == -1
        return rc;
    }
    RVMC_REQUIRE(p.fbin >= 0.0 && p.fbin <= 1.0, "fbin must be in [0,1] (got %g)", p.fbin);
    RVMC_REQUIRE(p.kepler_iters >= 1 && p.kepler_iters <= 8, "kepler_iters must be in [1,8] (got %d)",
                 p.kepler_iters);
    // the power law is in log10(P/day) (RVMC.py:134): periods <= 1 d give log10 <= 0 and NaNs in the
    // reference.  min > max is not an error there (the law simply runs backwards).
    RVMC_REQUIRE(p.min_period > 1.0 && p.max_period > 1.0, "min_period and max_period must be > 1 day (RVMC.py:134)");
    RVMC_REQUIRE(p.min_mass > 0 && p.max_mass > 0 && p.min_q > 0 && p.max_q > 0,
                 "mass and mass-ratio bounds must be > 0");
    out->seed_lo = (uint32_t)seed;
    out->seed_hi = (uint32_t)(seed >> 32);
    out->coin_thresh = (uint32_t)llrint(p.fbin * 16777216.0);
    out->kepler_iters = p.kepler_iters;
    out->sub = p.substream;
    return RVMC_OK;
}

// The seed every point of a grid call shares (the points then differ in their substream), if any.
static bool common_seed_of(const std::vector<FusedPoint>& pts, uint64_t* seed) {
    if (pts.empty()) return false;
    for (size_t i = 1; i < pts.size(); ++i)
        if (pts[i].seed_lo != pts[0].seed_lo || pts[i].seed_hi != pts[0].seed_hi) return false;
    *seed = ((uint64_t)pts[0].seed_hi << 32) | pts[0].seed_lo;
    return true;
}

struct TileShape {
    int R, G;
    size_t smem;
};

static int fused_tile_shape(int S, TileShape* ts) {
    if (S < 1) {
        set_error("sample_size must be >= 1 (got %d)", S);
        return RVMC_ERR_INVALID;
    }
    if (S > 8192) {
        set_error("sample_size %d exceeds the 8192 stars per cluster the fused kernel is built for", S);
        return RVMC_ERR_UNSUPPORTED;
    }
    const int T = 4096;
    ts->R = S >= T ? 1 : T / S;
    ts->G = reduce_group_size(S, ts->R, 256);
    const size_t slots = (size_t)ts->R * S;
    ts->smem = slots * sizeof(float) + 2 * (size_t)S * sizeof(float) + slots * sizeof(uint16_t) + 16;
    return RVMC_OK;
}

static int launch_fused(const FusedPoint* single, const FusedPoint* d_points, int64_t n_points, const float* meas_err, int S, int weighted,
                        int ddof, uint64_t first_real, int64_t n_real, float* sigma_out, int64_t sigma_stride,
                        uint64_t* counters, bool iters1, bool guard, cudaStream_t stream,
                        const uint64_t* common_seed = nullptr, uint32_t* fail = nullptr) {
    TileShape ts;
    int rc = fused_tile_shape(S, &ts);
    if (rc) return rc;
    FusedArgs a;
    if (single) {
        a.single = *single;
        a.key = philox_key(((uint64_t)single->seed_hi << 32) | single->seed_lo);
    }
    if (!single && common_seed) a.key = philox_key(*common_seed);      // KEYED: one key for every point
    a.points = d_points;
    a.meas_err = meas_err;
    a.sigma_out = sigma_out;
    a.counters = (unsigned long long*)counters;
    a.fail = fail;
    a.first_real = first_real;
    a.n_real = n_real;
    a.sigma_stride = sigma_stride;
    a.S = S;
    a.R = ts.R;
    a.G = ts.G;
    a.weighted = weighted;
    a.ddof = ddof;
    const int64_t tiles_per_point = (n_real + ts.R - 1) / ts.R;
    const int64_t total = tiles_per_point * n_points;
    if (tiles_per_point > 0x7fffffff || total > 0x7fffffff) {
        set_error("too many tiles in one launch (%lld); split the call", (long long)total);
        return RVMC_ERR_UNSUPPORTED;
    }
    a.tiles_per_point = (int32_t)tiles_per_point;
    // kepler_iters is launch-uniform for a single population; a grid takes the run-time count unless
    // the caller guarantees 1 everywhere (iters1)
    auto go = [&](auto kernel) -> int {
        if (ts.smem > 48 * 1024)
            RVMC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts.smem));
        kernel<<<(unsigned int)total, 256, ts.smem, stream>>>(a);
        RVMC_LAUNCH_CHECK();
        return RVMC_OK;
    };
    const bool fast = (single ? single->kepler_iters == 1 : iters1) && !guard;
    if (single)
        return fast ? go(sigma1d_fused_kernel<true, true, 1, false>) : go(sigma1d_fused_kernel<true, true, 0, true>);
    if (common_seed)
        return fast ? go(sigma1d_fused_kernel<false, true, 1, false>) : go(sigma1d_fused_kernel<false, true, 0, true>);
    return fast ? go(sigma1d_fused_kernel<false, false, 1, false>) : go(sigma1d_fused_kernel<false, false, 0, true>);
}

}  // namespace rvmc

using namespace rvmc;

extern "C" {

int rvmc_sigma1d_fused(const rvmc_pop_params* p, uint64_t seed, const float* meas_err, int sample_size,
                       int weighted, int ddof, uint64_t first_real, int64_t n_real, float* sigma_out,
                       uint64_t* counters, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(p != nullptr, "population parameters are NULL");
    RVMC_REQUIRE(n_real >= 0, "number_of_samples must be >= 0");
    RVMC_REQUIRE(ddof == 0 || ddof == 1, "ddof must be 0 or 1");
    FusedPoint fp;
    int rc = make_fused_point(*p, seed, &fp);
    if (rc) return rc;
    TileShape ts;
    rc = fused_tile_shape(sample_size, &ts);
    if (rc) return rc;
    if (n_real == 0) return RVMC_OK;
    RVMC_REQUIRE(meas_err && sigma_out, "meas_err and sigma_out are required");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    // one launch covers at most 2^31 tiles
    const int64_t max_real = (int64_t)ts.R * 0x7fff0000LL;
    for (int64_t done = 0; done < n_real; done += max_real) {
        const int64_t chunk = (n_real - done < max_real) ? n_real - done : max_real;
        rc = launch_fused(&fp, nullptr, 1, meas_err, sample_size, weighted, ddof, first_real + done, chunk,
                          sigma_out + done, 0, counters, false, guard_reachable(*p), stream);
        if (rc) return rc;
    }
    return RVMC_OK;
}

int rvmc_fused_trace(const rvmc_pop_params* p, uint64_t seed, const float* meas_err, int sample_size,
                     uint64_t first_slot, int64_t n, rvmc_trace_soa out, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(p != nullptr, "population parameters are NULL");
    RVMC_REQUIRE(n >= 0, "n must be >= 0");
    RVMC_REQUIRE(sample_size >= 1, "sample_size must be >= 1");
    FusedPoint fp;
    int rc = make_fused_point(*p, seed, &fp);
    if (rc) return rc;
    if (n == 0) return RVMC_OK;
    RVMC_REQUIRE(meas_err != nullptr, "meas_err is required");
    const int threads = 128;
    fused_trace_kernel<<<grid_stride_blocks(n, threads), threads, 0, (cudaStream_t)cuda_stream>>>(
        fp, meas_err, sample_size, first_slot, n, out);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_grid_scratch_slots(int device) {
    // rows of sigma_1D the scratch holds; it is used as two halves (the statistics kernel of one
    // batch overlaps the fused kernel of the next), so one statistics launch covers slots/2 rows:
    // 592 blocks = 4 per SM.  RVMC_GRID_SCRATCH_SLOTS overrides it (tuning).
    (void)device;
    static int slots = 0;
    if (slots == 0) {
        const char* env = getenv("RVMC_GRID_SCRATCH_SLOTS");
        int v = env ? atoi(env) : 0;
        slots = (v >= 64 && v <= (1 << 20)) ? (v & ~1) : 1184;
    }
    return slots;
}

int rvmc_grid_run(const rvmc_grid_point* points, int64_t n_points, const float* meas_err, int sample_size,
                  int weighted, int ddof, int64_t n_real, double bin, int hist_bins,
                  rvmc_point_stats* stats_out, float* sigma_out, uint32_t* hist_out, float* scratch,
                  uint64_t* counters, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n_points >= 0, "n_points must be >= 0");
    RVMC_REQUIRE(n_real >= 1, "number_of_samples must be >= 1");
    RVMC_REQUIRE(n_real < ((int64_t)1 << 31), "number_of_samples per grid point must be < 2^31");
    RVMC_REQUIRE(ddof == 0 || ddof == 1, "ddof must be 0 or 1");
    RVMC_REQUIRE(bin > 0.0, "bin must be > 0");
    RVMC_REQUIRE(hist_out == nullptr || hist_bins >= 1, "hist_bins must be >= 1 when hist_out is given");
    TileShape ts;
    int rc = fused_tile_shape(sample_size, &ts);
    if (rc) return rc;
    if (n_points == 0) return RVMC_OK;
    RVMC_REQUIRE(points && meas_err && stats_out, "points, meas_err and stats_out are required");
    RVMC_REQUIRE(sigma_out || scratch, "either sigma_out or scratch must be given");
    cudaStream_t stream = (cudaStream_t)cuda_stream;

    std::vector<FusedPoint> host(n_points);
    bool iters1 = true, guard = false;
    for (int64_t i = 0; i < n_points; ++i) {
        rc = make_fused_point(points[i].pop, points[i].seed, &host[i]);
        if (rc) return rc;
        iters1 = iters1 && host[i].kepler_iters == 1;
        guard = guard || guard_reachable(points[i].pop);
    }
    uint64_t shared_seed = 0;
    const bool keyed = common_seed_of(host, &shared_seed);
    // stream-ordered scratch for the descriptors: no library-global state, so concurrent calls on
    // different streams do not interact
    // [descriptors | per-point failure counters]
    void* ws = nullptr;
    const size_t desc_bytes = sizeof(FusedPoint) * (size_t)n_points;
    const size_t fail_bytes = 2 * sizeof(uint32_t) * (size_t)n_points;
    const size_t todo_bytes = (size_t)n_points;
    RVMC_CUDA(cudaMallocAsync(&ws, desc_bytes + fail_bytes + todo_bytes, stream));
    // the copy from pageable memory has been staged by the time the call returns; `host` may die
    rc = cuda_status(cudaMemcpyAsync(ws, host.data(), desc_bytes, cudaMemcpyHostToDevice, stream),
                     "cudaMemcpyAsync(points)");
    if (!rc) rc = cuda_status(cudaMemsetAsync((char*)ws + desc_bytes, 0, fail_bytes, stream), "cudaMemsetAsync");
    if (rc) {
        cudaFreeAsync(ws, stream);
        return rc;
    }
    const FusedPoint* d_points = (const FusedPoint*)ws;
    uint32_t* d_fail = (uint32_t*)((char*)ws + desc_bytes);
    uint8_t* d_todo = (uint8_t*)ws + desc_bytes + fail_bytes;

    const int64_t tiles_per_point = (n_real + ts.R - 1) / ts.R;
    int64_t batch = rvmc_grid_scratch_slots(0) / 2;               // the scratch is used as two halves
    if (sigma_out) batch = 0x7fff0000LL / tiles_per_point;       // no scratch limit: as many as one launch takes
    if (batch < 1) batch = 1;
    // The statistics kernel of batch b runs on a side stream while the fused kernel of batch b+1
    // runs on the caller's stream (scratch halves alternate), so its latency-bound passes hide
    // behind the compute-bound Monte Carlo.
    GridAux* aux = nullptr;
    rc = grid_aux(&aux);
    if (rc) {
        cudaFreeAsync(ws, stream);
        return rc;
    }
    cudaStream_t side = aux->side;
    cudaEvent_t ev_fused = aux->ev_fused, *ev_stats = aux->ev_stats;
    int64_t nb = 0;
    for (int64_t p0 = 0; p0 < n_points && !rc; p0 += batch, ++nb) {
        const int64_t np = (n_points - p0 < batch) ? n_points - p0 : batch;
        const int half = (int)(nb & 1);
        float* sig = sigma_out ? sigma_out + (size_t)p0 * n_real : scratch + (size_t)half * batch * n_real;
        // this scratch half was last read by the statistics kernel of batch nb-2
        if (!sigma_out && nb >= 2) rc = cuda_status(cudaStreamWaitEvent(stream, ev_stats[half], 0), "cudaStreamWaitEvent");
        if (rc) break;
        rc = launch_fused(nullptr, d_points + p0, np, meas_err, sample_size, weighted, ddof, 0, n_real, sig, n_real,
                          counters, iters1, guard, stream, keyed ? &shared_seed : nullptr, d_fail + 2 * p0);
        if (rc) break;
        rc = cuda_status(cudaEventRecord(ev_fused, stream), "cudaEventRecord");
        if (!rc) rc = cuda_status(cudaStreamWaitEvent(side, ev_fused, 0), "cudaStreamWaitEvent");
        if (rc) break;
        StatsArgs sa;
        sa.sigma = sig;
        sa.sigma_stride = n_real;
        sa.n_real = n_real;
        sa.bin = bin;
        sa.hist_bins = hist_bins;
        sa.stats_out = stats_out + p0;
        sa.hist_out = hist_out ? hist_out + (size_t)p0 * hist_bins : nullptr;
        sa.fail = d_fail + 2 * p0;
        rc = launch_point_stats(sa, np, d_todo + p0, side);
        if (!rc) rc = cuda_status(cudaEventRecord(ev_stats[half], side), "cudaEventRecord");
    }
    // the caller's stream continues only after every statistics kernel has finished
    for (int k = 0; k < 2; ++k)
        if (ev_stats[k] && nb > k) {
            int rc2 = cuda_status(cudaStreamWaitEvent(stream, ev_stats[k], 0), "cudaStreamWaitEvent");
            if (!rc) rc = rc2;
        }
    cudaFreeAsync(ws, stream);
    return rc;
}

int rvmc_grid_run_fbins(const rvmc_grid_point* points, int64_t n_points, const double* fbins, int n_fbins,
                        const float* meas_err, int sample_size, int weighted, int ddof, int64_t n_real, double bin,
                        int hist_bins, rvmc_point_stats* stats_out, float* sigma_out, uint32_t* hist_out,
                        float* scratch, uint64_t* counters, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n_points >= 0, "n_points must be >= 0");
    RVMC_REQUIRE(n_fbins >= 1 && n_fbins <= FBINS_MAX, "n_fbins must be in [1, %d] (got %d)", FBINS_MAX, n_fbins);
    RVMC_REQUIRE(fbins != nullptr, "fbins is NULL");
    RVMC_REQUIRE(n_real >= 1 && n_real < ((int64_t)1 << 31), "number_of_samples per grid point must be in [1, 2^31)");
    RVMC_REQUIRE(ddof == 0 || ddof == 1, "ddof must be 0 or 1");
    RVMC_REQUIRE(bin > 0.0, "bin must be > 0");
    RVMC_REQUIRE(hist_out == nullptr || hist_bins >= 1, "hist_bins must be >= 1 when hist_out is given");
    if (sample_size < 1) {
        set_error("sample_size must be >= 1 (got %d)", sample_size);
        return RVMC_ERR_INVALID;
    }
    if (sample_size > 2048) {
        set_error("sample_size %d exceeds the 2048 stars per cluster the f_bin-fused kernel is built for", sample_size);
        return RVMC_ERR_UNSUPPORTED;
    }
    FbinArgs fa;
    fa.n_fbins = n_fbins;
    fa.thresh_max = 0;
    for (int k = 0; k < FBINS_MAX; ++k) fa.thresh[k] = 0;       // the kernel reads whole groups of eight
    for (int k = 0; k < n_fbins; ++k) {
        RVMC_REQUIRE(fbins[k] >= 0.0 && fbins[k] <= 1.0, "fbin must be in [0,1] (got %g)", fbins[k]);
        fa.thresh[k] = (uint32_t)llrint(fbins[k] * 16777216.0);
        fa.thresh_max = fa.thresh[k] > fa.thresh_max ? fa.thresh[k] : fa.thresh_max;
    }
    if (n_points == 0) return RVMC_OK;
    RVMC_REQUIRE(points && meas_err && stats_out, "points, meas_err and stats_out are required");
    RVMC_REQUIRE(sigma_out || scratch, "either sigma_out or scratch must be given");
    cudaStream_t stream = (cudaStream_t)cuda_stream;

    std::vector<FusedPoint> host(n_points);
    bool iters1 = true, guard = false;
    int rc = RVMC_OK;
    for (int64_t i = 0; i < n_points; ++i) {
        rc = make_fused_point(points[i].pop, points[i].seed, &host[i]);
        if (rc) return rc;
        iters1 = iters1 && host[i].kepler_iters == 1;
        guard = guard || guard_reachable(points[i].pop);
    }
    uint64_t shared_seed = 0;
    const bool keyed = common_seed_of(host, &shared_seed);
    if (keyed) fa.base.key = philox_key(shared_seed);
    void* ws = nullptr;
    const size_t desc_bytes = sizeof(FusedPoint) * (size_t)n_points;
    const size_t fail_bytes = 2 * sizeof(uint32_t) * (size_t)n_points * (size_t)n_fbins;
    const size_t todo_bytes = (size_t)n_points * (size_t)n_fbins;
    RVMC_CUDA(cudaMallocAsync(&ws, desc_bytes + fail_bytes + todo_bytes, stream));
    rc = cuda_status(cudaMemcpyAsync(ws, host.data(), desc_bytes, cudaMemcpyHostToDevice, stream),
                     "cudaMemcpyAsync(points)");
    if (!rc) rc = cuda_status(cudaMemsetAsync((char*)ws + desc_bytes, 0, fail_bytes, stream), "cudaMemsetAsync");
    if (rc) {
        cudaFreeAsync(ws, stream);
        return rc;
    }
    const FusedPoint* d_points = (const FusedPoint*)ws;
    uint32_t* d_fail = (uint32_t*)((char*)ws + desc_bytes);
    uint8_t* d_todo = (uint8_t*)ws + desc_bytes + fail_bytes;

    const int S = sample_size;
    // Tile height: up to 3072 slots, and -- for the small clusters whose phase C runs one thread per
    // (realization, group of eight fractions) -- a multiple of what makes those items fill whole
    // passes of the 256 threads (S = 12, 16 fractions: R = 256, 512 items, two full passes).
    int R = 3072 / S;
    {
        const int ngroups = (n_fbins + 7) / 8;
        int step = 256;
        for (int g = ngroups; (g & 1) == 0 && step > 1; g >>= 1) step >>= 1;      // 256 / gcd(256, ngroups)
        if (R >= step) R -= R % step;
        if (R < 1) R = 1;
    }
    const size_t slots = (size_t)R * S;
    const size_t smem = slots * (2 * sizeof(float) + sizeof(uint32_t) + sizeof(uint16_t)) + 2 * (size_t)S * sizeof(float) + 16;
    const int64_t tiles_per_point = (n_real + R - 1) / R;
    // rows of sigma per batch: half the scratch (two halves alternate, as in rvmc_grid_run)
    int64_t batch = (rvmc_grid_scratch_slots(0) / 2) / n_fbins;
    if (sigma_out) batch = 0x7fff0000LL / tiles_per_point;
    if (batch < 1) batch = 1;
    if (batch * tiles_per_point > 0x7fff0000LL) batch = 0x7fff0000LL / tiles_per_point;
    if (!sigma_out && (int64_t)n_fbins > rvmc_grid_scratch_slots(0) / 2) {
        cudaFreeAsync(ws, stream);
        set_error("scratch too small for %d binary fractions", n_fbins);
        return RVMC_ERR_UNSUPPORTED;
    }
    const bool fast = iters1 && !guard;
    GridAux* aux = nullptr;
    rc = grid_aux(&aux);
    if (rc) {
        cudaFreeAsync(ws, stream);
        return rc;
    }
    cudaStream_t side = aux->side;
    cudaEvent_t ev_fused = aux->ev_fused, *ev_stats = aux->ev_stats;
    auto launch_fbins = [&](auto kernel, unsigned int total) -> int {
        if (smem > 48 * 1024) {
            int rc2 = cuda_status(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                  "smem attr");
            if (rc2) return rc2;
        }
        kernel<<<total, FB_THREADS, smem, stream>>>(fa);
        return cuda_status(cudaGetLastError(), "sigma1d_fused_fbins_kernel launch");
    };
    int64_t nb = 0;
    for (int64_t p0 = 0; p0 < n_points && !rc; p0 += batch, ++nb) {
        const int64_t np = (n_points - p0 < batch) ? n_points - p0 : batch;
        const int half = (int)(nb & 1);
        const int64_t rows = np * n_fbins;
        float* sig = sigma_out ? sigma_out + (size_t)p0 * n_fbins * n_real
                               : scratch + (size_t)half * (rvmc_grid_scratch_slots(0) / 2) * n_real;
        if (!sigma_out && nb >= 2) rc = cuda_status(cudaStreamWaitEvent(stream, ev_stats[half], 0), "cudaStreamWaitEvent");
        if (rc) break;
        fa.base.points = d_points + p0;
        fa.base.meas_err = meas_err;
        fa.base.sigma_out = sig;
        fa.base.counters = (unsigned long long*)counters;
        fa.base.fail = d_fail + 2 * p0 * n_fbins;
        fa.base.first_real = 0;
        fa.base.n_real = n_real;
        fa.base.sigma_stride = n_real;
        fa.base.S = S;
        fa.base.R = R;
        // the group size of rvmc_grid_run for this S (its tiles hold 4096 slots): same summation
        // grouping => bit-identical sigmas
        fa.base.G = reduce_group_size(S, S >= 4096 ? 1 : 4096 / S, 256);
        fa.base.tiles_per_point = (int32_t)tiles_per_point;
        fa.base.weighted = weighted;
        fa.base.ddof = ddof;
        const unsigned int total = (unsigned int)(tiles_per_point * np);
        if (keyed)
            rc = fast ? launch_fbins(sigma1d_fused_fbins_kernel<true, 1, false>, total)
                      : launch_fbins(sigma1d_fused_fbins_kernel<true, 0, true>, total);
        else
            rc = fast ? launch_fbins(sigma1d_fused_fbins_kernel<false, 1, false>, total)
                      : launch_fbins(sigma1d_fused_fbins_kernel<false, 0, true>, total);
        if (!rc) rc = cuda_status(cudaEventRecord(ev_fused, stream), "cudaEventRecord");
        if (!rc) rc = cuda_status(cudaStreamWaitEvent(side, ev_fused, 0), "cudaStreamWaitEvent");
        if (rc) break;
        StatsArgs sa;
        sa.sigma = sig;
        sa.sigma_stride = n_real;
        sa.n_real = n_real;
        sa.bin = bin;
        sa.hist_bins = hist_bins;
        sa.stats_out = stats_out + p0 * n_fbins;
        sa.hist_out = hist_out ? hist_out + (size_t)p0 * n_fbins * hist_bins : nullptr;
        sa.fail = d_fail + 2 * p0 * n_fbins;
        rc = launch_point_stats(sa, rows, d_todo + p0 * n_fbins, side);
        if (!rc) rc = cuda_status(cudaEventRecord(ev_stats[half], side), "cudaEventRecord");
    }
    for (int k = 0; k < 2; ++k)
        if (ev_stats[k] && nb > k) {
            int rc2 = cuda_status(cudaStreamWaitEvent(stream, ev_stats[k], 0), "cudaStreamWaitEvent");
            if (!rc) rc = rc2;
        }
    cudaFreeAsync(ws, stream);
    return rc;
}

}  // extern "C"
