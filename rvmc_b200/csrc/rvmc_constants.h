// Physical constants of the RVMC hot path, pinned to the values the reference uses.
//   G      : astropy.constants.G.value for astropy >= 4.0 (CODATA 2018)  -- RVMC.py:8,187,197
//   MSUN   : the reference's solar mass, 2e30 kg                         -- RVMC.py:126
//   DAY    : 24*3600 s                                                   -- RVMC.py:134
#pragma once

#define RVMC_G_VALUE 6.6743e-11
#define RVMC_MSUN_KG 2.0e30
#define RVMC_DAY_S 86400.0
#define RVMC_TWO_PI 6.283185307179586476925286766559

// Philox draw sites ("streams", counter word 2).  One (item, stream, sub) triple names one
// Philox4x32-10 block, so any value can be recomputed from (seed, indices) alone: results do
// not depend on grid size, launch chunking or the number of GPUs.
#define RVMC_STREAM_ORBIT_A 0u   // words: primary mass, period, orbital phase, mass ratio
#define RVMC_STREAM_ORBIT_B 1u   // words: eccentricity, inclination, omega, (spare)
#define RVMC_STREAM_SIGMA_DYN 2u // N(0, sigma_dyn) added to rv_binary          (RVMC.py:224)
#define RVMC_STREAM_CLUSTER 3u   // cluster_velocities                          (RVMC.py:101)
#define RVMC_STREAM_MIX 4u       // pool mixing: permutation key + single picks (RVMC.py:234-235)
#define RVMC_STREAM_BOOT 5u      // bootstrap pick + measurement noise          (RVMC.py:264-271)
#define RVMC_STREAM_SLOT 6u      // fused path: binary/single coin + noise, one block per slot pair
#define RVMC_STREAM_EPOCH 7u     // multi-epoch measurement noise               (RVMC_bias_corr.py:404)
#define RVMC_STREAM_SINGLE 8u    // single-star noise                           (RVMC_bias_corr.py:437)
#define RVMC_STREAM_MASS 9u      // N(mass, mass_error)                         (RVMC_bias_corr.py:183)
