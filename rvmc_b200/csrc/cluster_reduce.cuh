// Per-realization dispersion of a tile of star velocities held in shared memory (K4).
//
// Replaces np.std(..., axis=1, ddof=1) at RVMC.py:270-271, np.std(..., axis=1) (ddof=0) at
// findBestDistributions.py:29-30 and the weighted formula of RVMC.py:262-268.
//
// A tile holds `nr` realizations of S stars, row-major.  Each realization is reduced by an
// aligned group of G lanes, G a power of two chosen so that the tile keeps every thread of the
// block busy: G = 1 (one thread sums its whole row, no shuffles) for the small clusters that
// dominate the workloads (S <= 16), up to a full warp for S >= ~170.  One pass: the sums of
// d = v - pivot and d^2 are accumulated with the row's first star as pivot, which keeps the
// fp32 cancellation in  var = (sum d^2 - (sum d)^2 / S) / (S - ddof)  harmless (the pivot is a
// sample of the distribution, so |mean - pivot| ~ sigma).  The summation order is fixed =>
// results are bit-reproducible across launches and GPUs.
#pragma once
#include <stdint.h>

#ifndef RVMC_DEV_ASSERT
#define RVMC_DEV_ASSERT(cond) ((void)0)
#endif

namespace rvmc {

// threads: block size; R: realizations per tile.  Smallest power of two G with R*G >= threads.
inline int reduce_group_size(int S, int R, int threads) {
    int g = 1;
    while (g < 32 && (int64_t)R * g < threads && g * 2 <= S) g *= 2;
    return g;
}

template <typename T>
__device__ __forceinline__ T group_sum(T x, int G) {
    for (int off = G >> 1; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}

// sigma from the pivot-shifted sums of one realization (shared by every reducer so that they all
// round alike).  The divisors enter as reciprocals computed once per thread (inv_S = 1/S,
// inv_dof = 1/(S - ddof); S == ddof -> 0 * inf = NaN like NumPy's 0/0): fp64 -- the finite-pool path
// that follows the reference to 1e-12 -- keeps true divisions and the IEEE square root, fp32 uses
// the reciprocals and MUFU.SQRT (2 ulp; the fused path's sigma carries 1e-5 anyway).
template <typename T>
struct SigmaNorm {
    T S, dof, inv_S, inv_dof;
    __device__ __forceinline__ SigmaNorm(int S_, int ddof) : S((T)S_), dof((T)(S_ - ddof)) {
        inv_S = (T)1 / S;
        inv_dof = (T)1 / dof;
    }
};
__device__ __forceinline__ double sigma_from_sums(double s1, double s2, const SigmaNorm<double>& n) {
    double var = (s2 - s1 * s1 / n.S) / n.dof;
    var = var < 0.0 ? 0.0 : var;         // rounding can leave -0.0...: keep sqrt real (NaN stays NaN)
    return sqrt(var);
}
__device__ __forceinline__ float sigma_from_sums(float s1, float s2, const SigmaNorm<float>& n) {
    float var = fmaf(-s1 * n.inv_S, s1, s2) * n.inv_dof;
    var = var < 0.0f ? 0.0f : var;
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(var));
    return r;
}

// value(l): velocity of tile slot l = r*S + j (a functor over shared memory); wts: [S] weights
// 1/err^2 in shared memory or nullptr (unweighted); wsum: sum of the weights; out[r] receives sigma
// of realization r.  All threads of the block must call this (it contains warp shuffles);
// blockDim.x must be a multiple of 32.
template <typename T, typename OutT, typename Value>
__device__ __forceinline__ void reduce_tile_sigma_fn(Value value, int nr, int S, int G, const T* wts, T wsum,
                                                     int ddof, OutT* out) {
    const int ng = blockDim.x / G;          // groups per block
    const int gi = threadIdx.x / G;
    const int gl = threadIdx.x & (G - 1);
    const int iters = (nr + ng - 1) / ng;
    const SigmaNorm<T> norm(S, ddof);
    RVMC_DEV_ASSERT(G >= 1 && G <= 32 && (G & (G - 1)) == 0 && gl < G);
    for (int it = 0; it < iters; ++it) {
        const int r = it * ng + gi;
        const bool valid = r < nr;
        const int base = (valid ? r : 0) * S;
        const T pivot = value(base);
        T s1 = 0, s2 = 0;
        if (wts) {
            for (int j = gl; j < S; j += G) {
                const T d = value(base + j) - pivot;
                const T wd = wts[j] * d;
                s1 += wd;
                s2 = fma(wd, d, s2);
            }
        } else {
            for (int j = gl; j < S; j += G) {
                const T d = value(base + j) - pivot;
                s1 += d;
                s2 = fma(d, d, s2);
            }
        }
        if (G > 1) {
            s1 = group_sum(s1, G);
            s2 = group_sum(s2, G);
        }
        if (valid && gl == 0) {
            if (!wts) {
                out[r] = (OutT)sigma_from_sums(s1, s2, norm);
                continue;
            }
            // sum w (v - mu)^2 / sum w  with  mu = pivot + s1 / wsum            (RVMC.py:263-268);
            // ddof = 1 applies the Bessel-type factor S/(S-1) of the variant the reference keeps
            // commented out at RVMC.py:267
            const T m = s1 / wsum;
            T var = s2 / wsum - m * m;
            if (ddof) var = var * (T)S / (T)(S - ddof);
            var = var < (T)0 ? (T)0 : var;       // rounding can leave -0.0...: keep sqrt real (NaN stays NaN)
            out[r] = (OutT)sqrt(var);
        }
    }
}

// v: [nr][S] in shared memory, row-major.
template <typename T, typename OutT>
__device__ __forceinline__ void reduce_tile_sigma(const T* v, int nr, int S, int G, const T* wts, T wsum,
                                                  int ddof, OutT* out) {
    reduce_tile_sigma_fn<T, OutT>([v](int l) { return v[l]; }, nr, S, G, wts, wsum, ddof, out);
}

}  // namespace rvmc
