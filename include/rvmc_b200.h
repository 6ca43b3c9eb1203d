/* rvmc_b200 -- C ABI of the B200-native RVMC Monte-Carlo hot path (librvmc_b200.so).
 *
 * The reference (FPABacks/RVMC) is pure Python and has no FFI of its own; this boundary sits
 * INSIDE the bodies of the methods it replaces.  Each entry point cites the reference code it
 * stands in for (file:line under /root/reference).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every array pointer is a DEVICE pointer owned by the caller unless the
 *     name ends in `_host` (then it is a host pointer and the call copies synchronously);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     stream-ordered and never synchronise unless documented;
 *   - return value: 0 = RVMC_OK, < 0 = argument error (RVMC_ERR_*), > 0 = a cudaError_t;
 *     rvmc_last_error() gives the message for the calling thread;
 *   - RNG: stateless Philox4x32-10.  Every value is a function of (seed, global item index,
 *     draw site) only, so results do not depend on launch geometry, chunking or GPU count.
 *     `first_*` arguments give the global index of the first item of a shard.
 */
#ifndef RVMC_B200_H
#define RVMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVMC_ABI_VERSION 2

/* the library is built with -fvisibility=hidden: only the entry points below are exported */
#if defined(__GNUC__)
#define RVMC_API __attribute__((visibility("default")))
#else
#define RVMC_API
#endif

enum {
    RVMC_OK = 0,
    RVMC_ERR_INVALID = -1,      /* bad argument (message says which)                         */
    RVMC_ERR_ZERODIV = -2,      /* a power-law index of -1: the reference raises              */
                                /* ZeroDivisionError at RVMC.py:112                           */
    RVMC_ERR_UNSUPPORTED = -3,  /* size outside what the kernels were built for               */
    RVMC_ERR_NO_DEVICE = -4     /* no CUDA device / wrong architecture: there is no fallback  */
};

/* Population description == the constructor arguments of RVMC.__init__ (RVMC.py:28-43) and
 * BiasCorr.__init__ (RVMC_bias_corr.py:57-75), plus the constants the reference hard-codes in
 * sample_eccentricity (RVMC.py:155-159).  Use rvmc_pop_defaults() to fill the defaults. */
typedef struct rvmc_pop_params {
    double min_mass, max_mass;          /* solar masses of 2e30 kg             RVMC.py:30-31,126 */
    double mass_power;                  /* IMF index (-2.35)                   RVMC.py:36        */
    double min_period, max_period;      /* days                                RVMC.py:32-33     */
    double period_pi;                   /* index of the law in log10 P         RVMC.py:34        */
    double min_q, max_q;                /*                                     RVMC.py:37-38     */
    double mass_ratio_kappa;            /*                                     RVMC.py:35        */
    double e_power;                     /*                                     RVMC.py:39        */
    double e_min, e_max_mid, e_max_long;/* 1e-5, 0.5, 0.9                      RVMC.py:158-159   */
    double p_circ_days, p_mid_days;     /* 4, 6                                RVMC.py:156-157   */
    double sigma_dyn;                   /* km/s                                RVMC.py:42        */
    double fbin;                        /*                                     RVMC.py:43        */
    double guard_amp;                   /* fp64 guard threshold on the M->nu error amplification
                                           sqrt(1-e^2)/(1-e cos E)^2 (default 64)               */
    int32_t kepler_iters;               /* Halley steps after the closed-form starter (default 1) */
    uint32_t substream;                 /* Philox counter word 3 of every draw of the fused sigma_1D
                                           kernels (default 0).  Grid points may share a seed and differ
                                           in substream: statistically independent streams under ONE
                                           key, which the grid kernels then read as launch-uniform
                                           operands instead of deriving ten round keys per point    */
} rvmc_pop_params;

/* SoA view of a materialised population: the public attribute arrays of RVMC.py:80-95.
 * Any pointer may be NULL (not produced / not needed). All float64, length n. */
typedef struct rvmc_orbit_soa {
    double* primary_mass;       /* kg   */
    double* period;             /* s    */
    double* time;               /* s    */
    double* mass_ratio;
    double* eccentricity;
    double* inclination;        /* rad  */
    double* orbit_rotation;     /* rad  */
    double* eccentric_anomaly;  /* rad  */
} rvmc_orbit_soa;

/* Per-slot dump of the fused fp32 path (parity/debug entry). All float32/uint32, length n. */
typedef struct rvmc_trace_soa {
    uint32_t* words;      /* 12 per slot: ORBIT_A[4], ORBIT_B[4], SLOT pair block[4]           */
    float* log2m;         /* log2(M1 / 2e30 kg)                                                */
    float* x;             /* log10(P / day)                                                    */
    float* uphase;        /* time / P in [0,1)                                                 */
    float* q;
    float* e;
    float* cosi;
    float* wturn;         /* omega / 2 pi                                                      */
    float* K;             /* semi-amplitude, km/s                                              */
    float* rv;            /* orbital RV, km/s (no noise)                                       */
    float* noise;         /* the slot's N(0, sqrt(sigma_dyn^2+err_j^2)) draw                    */
    uint8_t* is_binary;   /* the slot's coin                                                   */
} rvmc_trace_soa;

/* Per-point output of the grid search == the statistics of findBestDistributions.py:47-60. */
typedef struct rvmc_point_stats {
    double mode, mean, median, p05, p16, p84, p95;
    double sigma_max;           /* largest sigma_1D of the point (sets the histogram range)     */
    uint32_t n_guard;           /* orbits that took the fp64 guard path                         */
    uint32_t n_nonconverged;    /* orbits whose Kepler residual stayed above 1e-5               */
    uint32_t n_hist_bins;       /* bins of width `bin` the reference's arange(0,max+bin,bin) has */
    uint32_t reserved;
} rvmc_point_stats;

/* One grid point: a population, a cluster shape and its own seed. */
typedef struct rvmc_grid_point {
    rvmc_pop_params pop;
    uint64_t seed;
} rvmc_grid_point;

/* ---- library ---------------------------------------------------------------------------- */
RVMC_API int rvmc_abi_version(void);
RVMC_API const char* rvmc_last_error(void);
RVMC_API void rvmc_pop_defaults(rvmc_pop_params* p);               /* RVMC.py:28-43 defaults            */
/* Device probe: fails with RVMC_ERR_NO_DEVICE when there is no sm_100 device. */
RVMC_API int rvmc_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int* clock_khz);

/* Measurement aid (bench.py): sustained lane-operations per second of one SM pipe on the current
 * device -- kind 0 FFMA, 1 MUFU.EX2, 2 IMAD.WIDE.U32, 3 DFMA, 4 a mixed FFMA+ALU+MUFU issue stream.
 * These are the roofline denominators of a path that is bound by the instruction pipes, not by
 * HBM or the tensor cores.  Synchronises the device (default stream). */
RVMC_API int rvmc_pipe_peak(int kind, int iters, double* lane_ops_per_s, double* elapsed_ms);

/* ---- K1: Philox (known-answer / replay) ----------------------------------------------------
 * out[4*i..4*i+3] = Philox4x32-10 block of (item = first_item+i, stream, sub) under `seed`.  */
RVMC_API int rvmc_philox_blocks(uint64_t seed, uint64_t first_item, int64_t n, uint32_t stream, uint32_t sub,
                       uint32_t* out, void* cuda_stream);

/* ---- K1+K2: materialised population, fp64 (RVMC.__init__ RVMC.py:79-95; BiasCorr twins
 * RVMC_bias_corr.py:207-243).  Fills the non-NULL arrays of `out` for items
 * [first_item, first_item+n); eccentric_anomaly is the converged root for M = 2 pi time/period
 * (RVMC.py:170-180).  `given_mask` marks arrays that are INPUTS rather than outputs: bit 0 =
 * primary_mass (mass_dist, RVMC.py:79-82), bit 1 = period (period_dist, :84-87; time and
 * eccentricity then follow the given periods as at :89,:91), bit 2 = time, bit 3 = eccentricity
 * (bits 2-3 serve a standalone calculate_eccentric_anomaly on user-edited arrays).
 * *nonconverged (device uint32, may be NULL) is incremented per failure. */
#define RVMC_GIVEN_MASS 1u
#define RVMC_GIVEN_PERIOD 2u
#define RVMC_GIVEN_TIME 4u
#define RVMC_GIVEN_ECC 8u
RVMC_API int rvmc_sample_orbits(const rvmc_pop_params* p, uint64_t seed, uint64_t first_item, int64_t n,
                       uint32_t given_mask, rvmc_orbit_soa out, uint32_t* nonconverged,
                       void* cuda_stream);

/* out[i] (+)= mean + sigma * N(0,1) from stream `stream`/`sub`, item first_item+i.
 * Replaces np.random.normal at RVMC.py:101 (cluster_velocities) and :224 (sigma_dyn). */
RVMC_API int rvmc_normal_f64(uint64_t seed, uint64_t first_item, int64_t n, uint32_t stream, uint32_t sub,
                    double mean, double sigma, int accumulate, double* out, void* cuda_stream);

/* ---- K3: Kepler + RV on caller-supplied arrays (the replay/parity entries) ------------------
 * rvmc_kepler_f64: E = root of E - e sin E = 2 pi time/period      (RVMC.py:170-180).
 * rvmc_rv_replay_f64: RVMC.py:182-212 evaluated in fp64 as written.
 * rvmc_rv_replay_f32: the fused path's fp32 arithmetic (closed-form K, starter + Halley, fp64
 *   guard) on the same arrays; eccentric_anomaly is recomputed from time/period, E_out optional. */
RVMC_API int rvmc_kepler_f64(int64_t n, const double* time, const double* period, const double* eccentricity,
                    double* E_out, uint32_t* nonconverged, void* cuda_stream);
RVMC_API int rvmc_rv_replay_f64(int64_t n, rvmc_orbit_soa in, double* rv_out, double* K_out, void* cuda_stream);
/* The two intermediate public methods of the class, for scripts that call them one by one:
 * semi_major_axis [m] (RVMC.py:182-187) and max_orbital_velocity [km/s] (:189-198) from the
 * arrays of `in` (either output may be NULL; vmax uses a_in when given, else its own a), and
 * binary_radial_velocity (:200-212) for a caller-supplied v_max array. */
RVMC_API int rvmc_orbit_scales_f64(int64_t n, rvmc_orbit_soa in, const double* a_in, double* a_out,
                          double* vmax_out, void* cuda_stream);
RVMC_API int rvmc_rv_from_vmax_f64(int64_t n, rvmc_orbit_soa in, const double* v_max, double* rv_out,
                          void* cuda_stream);
RVMC_API int rvmc_rv_replay_f32(int64_t n, rvmc_orbit_soa in, const rvmc_pop_params* p, float* rv_out,
                       float* E_out, uint32_t* counters /* [guard, nonconverged] or NULL */,
                       void* cuda_stream);

/* ---- a15: pool mixing (RVMC.calculate_radial_velocity, RVMC.py:228-236) ---------------------
 * radial_velocities[0:nb) = a without-replacement selection of rv_binary (a keyed Feistel
 * permutation of [0,n)), [nb:n) = with-replacement picks of cluster_velocities; nb=int(n*fbin). */
RVMC_API int rvmc_mix_pool(uint64_t seed, uint32_t call_id, int64_t n, double fbin, const double* rv_binary,
                  const double* cluster_velocities, double* radial_velocities, void* cuda_stream);

/* ---- a16: finite-pool bootstrap (RVMC.subsample, RVMC.py:262-271) ---------------------------
 * sigma_out[r] for realizations [first_real, first_real+n_real): sample_size with-replacement
 * picks from pool[pool_n] + N(0, meas_err[j]); unweighted std with `ddof` (1 for RVMC.py:271,
 * 0 for findBestDistributions.py:29) or the weighted formula of RVMC.py:263-268 (ddof = 0; with
 * ddof = 1 the Bessel-corrected variant left commented out at RVMC.py:267).
 * meas_err: device float64[sample_size]. */
RVMC_API int rvmc_sigma1d_pool(uint64_t seed, uint32_t call_id, const double* pool, int64_t pool_n,
                      const double* meas_err, int sample_size, int weighted, int ddof,
                      uint64_t first_real, int64_t n_real, double* sigma_out, void* cuda_stream);

/* ---- K1-K4 fused: infinite-pool sigma_1D (sample + Kepler + RV + dispersion in one pass) ----
 * The N -> infinity limit of RVMC.py:228-273: each of the sample_size stars of a realization
 * is a binary with probability fbin (fresh orbit, RVMC.py:79-95,214-226) else a single, plus
 * N(0, sigma_dyn) and N(0, meas_err[j]).  Per-star values never reach HBM.
 * meas_err: device float32[sample_size].  counters: device uint64[4] =
 * {binary RVs evaluated, fp64-guard lanes, non-converged lanes, reserved} (accumulated) or NULL. */
RVMC_API int rvmc_sigma1d_fused(const rvmc_pop_params* p, uint64_t seed, const float* meas_err,
                       int sample_size, int weighted, int ddof, uint64_t first_real, int64_t n_real,
                       float* sigma_out, uint64_t* counters, void* cuda_stream);

/* Per-slot dump of exactly what rvmc_sigma1d_fused computes for star slots
 * [first_slot, first_slot+n) (slot = realization*sample_size + j). */
RVMC_API int rvmc_fused_trace(const rvmc_pop_params* p, uint64_t seed, const float* meas_err, int sample_size,
                     uint64_t first_slot, int64_t n, rvmc_trace_soa out, void* cuda_stream);

/* ---- K5 fused: multi-epoch Delta-RV + significance (BiasCorr, RVMC_bias_corr.py:406-476) ----
 * Binary (star, sample), samples [first_sample, first_sample+n_samples) of every star; `seed`
 * keys the orbit draws (the same numbers rvmc_biascorr_sample materialises), `noise_seed` the
 * measurement errors (a fresh one per calculate_radial_velocities() call).  `first_star` is the
 * global index of the first star of this call: every array below is indexed from it (star
 * s of the call is global star first_star + s), so stars can be sharded over calls, streams or
 * GPUs without changing any value.  t_obs: device float64[nstars*max_epochs] (row i holds n_epochs[i] valid entries, s);
 * rv_err: device float32[nstars*max_epochs] per-epoch errors in km/s or NULL (no noise, no
 * detection); mass_mode 0 = IMF (RVMC_bias_corr.py:205-207), 1 = N(masses, mass_errors) (:183),
 * 2 = masses at face value (:190); masses/mass_errors: device float64[nstars], kg.
 * quirk_phase != 0 reproduces RVMC_bias_corr.py:357.
 * Outputs (each may be NULL): delta_rv float32[nstars*n_samples] (:422), detected
 * uint8[nstars*n_samples] (:451-476), rv_out float32[nstars*max_epochs*n_samples] (the noisy
 * radial_velocities, :419-421), counters uint64[4] = {binary-epoch RVs, detected binaries, fp64-guard
 * epochs, binaries with a Kepler residual above 1e-5 at any epoch} (accumulated), det_per_star uint32[nstars] (accumulated). */
RVMC_API int rvmc_biascorr_fused(const rvmc_pop_params* p, uint64_t seed, uint64_t noise_seed, int first_star,
                        int nstars, int max_epochs,
                        const int32_t* n_epochs, const double* t_obs, const float* rv_err,
                        int mass_mode, const double* masses, const double* mass_errors,
                        int quirk_phase, float delta_rv_min, float n_sigma,
                        uint64_t first_sample, int64_t n_samples, int64_t sample_stride,
                        float* delta_rv, uint8_t* detected, float* rv_out, uint64_t* counters,
                        uint32_t* det_per_star, void* cuda_stream);

/* rvmc_biascorr_fused with the detection flags BIT-PACKED: detected_bits uint32[nstars*bits_stride],
 * bit j of word w of a star's row = sample 32 w + j of this call (np.unpackbits bitorder="little" on a
 * little-endian host); one __ballot_sync per warp, 1/8 of the bytes of `detected`.  The `bool
 * (n_stars, n_samples)` array of RVMC_bias_corr.py:451-476 is 1 bit of information per binary: this
 * is the form it should cross PCIe in (rvmc_unpack_bits_host widens it on the host).  A caller that
 * shards the samples passes the word holding ITS first sample, so `first` columns of a shard must be
 * multiples of 32.  bits_stride >= ceil(n_samples / 32) words.  Either flag output may be NULL. */
RVMC_API int rvmc_biascorr_fused_packed(const rvmc_pop_params* p, uint64_t seed, uint64_t noise_seed, int first_star,
                               int nstars, int max_epochs,
                               const int32_t* n_epochs, const double* t_obs, const float* rv_err,
                               int mass_mode, const double* masses, const double* mass_errors,
                               int quirk_phase, float delta_rv_min, float n_sigma,
                               uint64_t first_sample, int64_t n_samples, int64_t sample_stride,
                               float* delta_rv, uint8_t* detected, uint32_t* detected_bits, int64_t bits_stride,
                               float* rv_out, uint64_t* counters, uint32_t* det_per_star, void* cuda_stream);

/* HOST helpers for the packed flags (no GPU work; every pointer is a HOST pointer).
 * rvmc_unpack_bits_host: out_host[r*out_stride + c] = bit c of row r (0/1 bytes == NumPy bool) for
 *   r < n_rows, c < n_cols, with up to n_threads threads of a persistent pool (the calling thread is
 *   one of them).  bits_host / out_host may point INTO larger arrays (sub-blocks of a sharded pass).
 * rvmc_count_bits_host: counts_host[r] = number of set flags of row r (detection counts per star
 *   without ever widening the flags). */
RVMC_API int rvmc_unpack_bits_host(const uint32_t* bits_host, int64_t bits_stride_words, int64_t n_rows,
                          int64_t n_cols, uint8_t* out_host, int64_t out_stride, int n_threads);
RVMC_API int rvmc_count_bits_host(const uint32_t* bits_host, int64_t bits_stride_words, int64_t n_rows,
                         int64_t n_cols, int64_t* counts_host, int n_threads);
/* The device-to-host leg of the packed flags in one call.  bits_dev: DEVICE uint32[n_rows][bits_stride_words]
 * (the detected_bits of rvmc_biascorr_fused_packed); packed_host: PINNED host mirror of the same shape;
 * out_host: the bool array, out_stride >= n_cols bytes per row.  blocks[k] = rows [row0,row1) x columns
 * [col0,col1) (col0 a multiple of 32) with the CUDA event recorded behind the kernel launch that produced the
 * block (or NULL: ready now).  copy_stream: a stream of its own for the copies.  Blocks are copied in the
 * order given as soon as they are ready and widened as they land, while later blocks are still being
 * computed; returns when out_host is complete.  Replaces the tail of get_detected_binaries
 * (RVMC_bias_corr.py:451-476 returns `detected`). */
typedef struct rvmc_flag_block {
    int64_t row0, row1, col0, col1;
    void* ready_event;          /* cudaEvent_t or NULL */
} rvmc_flag_block;
RVMC_API int rvmc_fetch_flags_host(const uint32_t* bits_dev, int64_t bits_stride_words, int64_t n_cols,
                          const rvmc_flag_block* blocks, int n_blocks, void* copy_stream, uint32_t* packed_host,
                          uint8_t* out_host, int64_t out_stride, int n_threads);
/* memcpy with the same thread pool: moves a pinned staging block into pageable memory at DMA speed
 * (results too large to pin are staged through two fixed pinned buffers, rvmc_b200/_device.py). */
RVMC_API int rvmc_copy_host(void* dst_host, const void* src_host, int64_t nbytes, int n_threads);

/* Materialised per-binary attribute arrays of BiasCorr (RVMC_bias_corr.py:175-243) for the
 * same (seed, star, sample) indexing as rvmc_biascorr_fused; fp64, layout [nstars][n_samples]. */
RVMC_API int rvmc_biascorr_sample(const rvmc_pop_params* p, uint64_t seed, int nstars, int mass_mode,
                         const double* masses, const double* mass_errors, uint64_t first_sample,
                         int64_t n_samples, uint32_t given_mask, rvmc_orbit_soa out, void* cuda_stream);

/* The per-star building blocks of the loop at RVMC_bias_corr.py:414-419 for scripts that call
 * them one by one.  Row arrays (period, eccentricity, ... of the current star) have n_samples
 * elements and broadcast over the n_epochs rows of time_deltas / E, as in the reference.
 * rvmc_kepler_epochs_f64: calculate_eccentric_anomaly(time_deltas) (:348-358);
 * rvmc_rv_epochs_f64: binary_radial_velocity(eccentric_anomalies_i) (:360-372). */
RVMC_API int rvmc_kepler_epochs_f64(int n_epochs, int64_t n_samples, const double* time_deltas,
                           const double* period_row, const double* ecc_row, int quirk_phase,
                           double* E_out, uint32_t* nonconverged, void* cuda_stream);
RVMC_API int rvmc_rv_epochs_f64(int n_epochs, int64_t n_samples, const double* E, const double* vmax_row,
                       const double* ecc_row, const double* incl_row, const double* omega_row,
                       double* rv_out, void* cuda_stream);

/* fp64 replay of RVMC_bias_corr.py:406-422 on caller-supplied (possibly user-overwritten)
 * attribute arrays [nstars][n_samples]; primary_mass may have row length 1 (mass_cols = 1,
 * RVMC_bias_corr.py:190).  noise (seed, rv_err as above) is optional. */
RVMC_API int rvmc_biascorr_replay_f64(int nstars, int64_t n_samples, int max_epochs, const int32_t* n_epochs,
                             const double* t_obs, rvmc_orbit_soa in, int64_t mass_cols,
                             const float* rv_err, uint64_t seed, uint64_t first_sample,
                             int quirk_phase, double* rv_out, double* delta_rv, void* cuda_stream);

/* generate_single_delta_rv (RVMC_bias_corr.py:424-442): noise-only max-min per single star.
 * delta_out float64[nstars*n_singles]. */
RVMC_API int rvmc_biascorr_singles(uint64_t seed, int nstars, int max_epochs, const int32_t* n_epochs,
                          const float* rv_err, uint64_t first_single, int64_t n_singles,
                          double* delta_out, void* cuda_stream);

/* get_detected_binaries (RVMC_bias_corr.py:451-476) on materialised noisy RVs
 * rv[nstars][max_epochs][n_samples] (fp64). */
RVMC_API int rvmc_biascorr_detect(int nstars, int64_t n_samples, int max_epochs, const int32_t* n_epochs,
                         const double* rv, const float* rv_err, double delta_rv_min, double n_sigma,
                         uint8_t* detected, void* cuda_stream);

/* ---- K6: grid search (findBestDistributions.py:21-32,35-110,457-489) ------------------------
 * For each of n_points points: n_real infinite-pool realizations of sample_size stars, np.std
 * with `ddof` (0 at findBestDistributions.py:29-30), then mode (histogram of width `bin` from 0,
 * first maximum), mean, exact median and 5/16/84/95 percentiles (NumPy linear interpolation).
 * points: HOST array; meas_err: device float32[sample_size]; stats_out: device
 * rvmc_point_stats[n_points]; sigma_out: device float32[n_points*n_real] or NULL;
 * hist_out: device uint32[n_points*hist_bins] (counts of the `bin`-wide histogram) or NULL;
 * scratch: device float32[scratch_slots*n_real], scratch_slots = rvmc_grid_scratch_slots().
 * Points are independent: shard them over GPUs by passing each rank its own sub-array. */
RVMC_API int rvmc_grid_scratch_slots(int device);
RVMC_API int rvmc_grid_run(const rvmc_grid_point* points, int64_t n_points, const float* meas_err,
                  int sample_size, int weighted, int ddof, int64_t n_real, double bin, int hist_bins,
                  rvmc_point_stats* stats_out, float* sigma_out, uint32_t* hist_out, float* scratch,
                  uint64_t* counters, void* cuda_stream);

/* The binary-fraction axis of a grid in one pass.  Point i is evaluated for every fbins[k],
 * k < n_fbins <= 32, with the point's own seed shared along the axis (pop.fbin is ignored): each star
 * slot is sampled and solved once and the cluster is reduced once per fraction.  Results are
 * BIT-IDENTICAL to rvmc_grid_run on n_points*n_fbins points {pop_i with fbin = fbins[k], seed_i} --
 * this is an execution strategy, not a different estimator -- at roughly the cost of the largest
 * fraction alone.  Layouts: stats_out [n_points][n_fbins], sigma_out [n_points][n_fbins][n_real],
 * hist_out [n_points][n_fbins][hist_bins]; fbins is a HOST array; scratch as for rvmc_grid_run.
 * sample_size <= 2048. */
RVMC_API int rvmc_grid_run_fbins(const rvmc_grid_point* points, int64_t n_points, const double* fbins, int n_fbins,
                        const float* meas_err, int sample_size, int weighted, int ddof, int64_t n_real,
                        double bin, int hist_bins, rvmc_point_stats* stats_out, float* sigma_out,
                        uint32_t* hist_out, float* scratch, uint64_t* counters, void* cuda_stream);

/* HOST helper of the grid fit (no GPU work; stats_host is a HOST pointer): ONE pass over a table of
 * n_points records in grid (C) order.
 *   best fit -- findBestDistributions.py:138-158: the FIRST strict minimum of |central - observed| below the
 *     initial distance of 100, central = median (use_mode == 0) or mode; a NaN is never selected; nothing
 *     below 100 leaves best_index 0 and best_distance 100.  Skipped when have_observed == 0.
 *   totals[4] -- sums of `reserved` (points whose histogram the statistics kernel could not hold),
 *     `n_nonconverged`, the number of points with n_nonconverged != 0 (SciPy's RuntimeWarning at
 *     RVMC.py:177-180 is per call) and `n_guard`.
 * Any output pointer may be NULL.  Up to n_threads threads of the library's host pool. */
RVMC_API int rvmc_grid_best_fit_host(const rvmc_point_stats* stats_host, int64_t n_points, int use_mode,
                            int have_observed, double observed, int64_t* best_index, double* best_distance,
                            uint64_t* totals, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* RVMC_B200_H */
