// Counter-based Philox4x32-10 (Salmon et al., SC'11), kept entirely in registers.
// Replaces every np.random.* call of the reference (RVMC.py:101,111,119,146,168,224,234,235,
// 264,265,270,271; RVMC_bias_corr.py:183,264,280,306,328,404).
//
// Counter layout: ctr = (item_lo, item_hi, stream, sub), key = (seed_lo, seed_hi).
// `item` is a GLOBAL index (star of the pool, star slot of a cluster realization, binary of the
// multi-epoch grid), so a value never depends on which thread, launch or GPU computes it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RVMC_HD __host__ __device__ __forceinline__
#else
#define RVMC_HD inline
#endif

// Rounds of the generator.  10 is the Random123 default (Philox4x32-10) and what every result, test
// vector and figure of this repository uses; Salmon et al. (SC'11) report 7 as the smallest count that
// passes BigCrush.  A build with -DRVMC_PHILOX_ROUNDS=7 exists for experiments only (the known-answer
// tests then fail by design).
#ifndef RVMC_PHILOX_ROUNDS
#define RVMC_PHILOX_ROUNDS 10
#endif

namespace rvmc {

struct u32x4 {
    uint32_t x, y, z, w;
};

RVMC_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    // one 32x32->64 multiply: a single IMAD.WIDE.U32 on sm_100a
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
}

template <int ROUNDS = 10>
RVMC_HD u32x4 philox4x32(u32x4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(M0, c.x, hi0, lo0);
        mulhilo32(M1, c.z, hi1, lo1);
        u32x4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += W0;
        k1 += W1;
    }
    return c;
}

// Expanded key schedule: the ten round keys of a seed.  Kernels take it as a launch parameter so
// that every round key is a constant-bank operand of the LOP3 that consumes it, instead of two
// integer adds per round per thread.
struct PhiloxKey {
    uint32_t k0[10], k1[10];
};

RVMC_HD PhiloxKey philox_key(uint64_t seed) {
    PhiloxKey K;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        K.k0[r] = a;
        K.k1[r] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return K;
}

RVMC_HD u32x4 philox_block(const PhiloxKey& K, uint64_t item, uint32_t stream, uint32_t sub) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    u32x4 c;
    c.x = (uint32_t)item;
    c.y = (uint32_t)(item >> 32);
    c.z = stream;
    c.w = sub;
#pragma unroll
    for (int r = 0; r < RVMC_PHILOX_ROUNDS; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(M0, c.x, hi0, lo0);
        mulhilo32(M1, c.z, hi1, lo1);
        u32x4 n;
        n.x = hi1 ^ c.y ^ K.k0[r];
        n.y = lo1;
        n.z = hi0 ^ c.w ^ K.k1[r];
        n.w = lo0;
        c = n;
    }
    return c;
}

// One block for (item, stream, sub) under a 64-bit seed.
RVMC_HD u32x4 philox_block(uint64_t seed, uint64_t item, uint32_t stream, uint32_t sub) {
    u32x4 c;
    c.x = (uint32_t)item;
    c.y = (uint32_t)(item >> 32);
    c.z = stream;
    c.w = sub;
    return philox4x32<RVMC_PHILOX_ROUNDS>(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// 24-bit uniforms.  Both precisions consume the SAME value u = (w >> 8) * 2^-24 (exact in fp32),
// so the fp32 fused path and the fp64 materialised path see identical draws.
#if defined(RVMC_U01_MAGIC) && defined(__CUDA_ARCH__)
// Experiment (not the default): the same 24-bit value without an I2F on the XU pipe -- the top 23 bits through
// the 2^23 exponent trick, the last bit added by an FFMA.  6 ALU/FMA instructions against SHF + I2F + FMUL;
// measured SLOWER on B200 (the kernels are issue-bound, not XU-bound).
RVMC_HD float u01_f32(uint32_t w) {
    const uint32_t x = w >> 8;
    const float hi = __uint_as_float(0x4b000000u | (x >> 1)) - 8388608.0f;          // (x >> 1), exact
    return fmaf(hi, 1.1920928955078125e-7f, (x & 1u) ? 5.9604644775390625e-8f : 0.0f);
}
#else
RVMC_HD float u01_f32(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }        // [0,1)
#endif
RVMC_HD double u01_f64(uint32_t w) { return (double)(w >> 8) * 5.9604644775390625e-8; }
RVMC_HD float u01_open_f32(uint32_t w) { return (float)((w >> 8) + 1u) * 5.9604644775390625e-8f; }  // (0,1]
RVMC_HD double u01_open_f64(uint32_t w) { return (double)((w >> 8) + 1u) * 5.9604644775390625e-8; }

}  // namespace rvmc
