// Per-realization dispersion of a tile of star velocities held in shared memory (K4).
//
// Replaces np.std(..., axis=1, ddof=1) at RVMC.py:270-271, np.std(..., axis=1) (ddof=0) at
// findBestDistributions.py:29-30 and the weighted formula of RVMC.py:262-268.
//
// A tile holds `nr` realizations of S stars, row-major.  Each realization is reduced by an
// aligned sub-warp group of G lanes (G = largest power of two <= min(S, 32)) with xor-shuffles,
// in two passes (mean, then squared deviations): the values already sit in shared memory, so the
// second pass costs one more LDS per star and avoids the cancellation of a one-pass formula.
// The summation order is fixed => results are bit-reproducible across launches and GPUs.
#pragma once
#include <stdint.h>

namespace rvmc {

inline int reduce_group_size(int S) {
    int g = 1;
    while (g * 2 <= S && g * 2 <= 32) g *= 2;
    return g;
}

template <typename T>
__device__ __forceinline__ T group_sum(T x, int G) {
    for (int off = G >> 1; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}

// v: [nr][S] in shared memory; wts: [S] weights 1/err^2 in shared memory or nullptr (unweighted);
// wsum: sum of the weights; out[r] receives sigma of realization r.  All threads of the block
// must call this (it contains warp shuffles); blockDim.x must be a multiple of 32.
template <typename T, typename OutT>
__device__ __forceinline__ void reduce_tile_sigma(const T* v, int nr, int S, int G, const T* wts, T wsum,
                                                  int ddof, OutT* out) {
    const int ng = blockDim.x / G;          // groups per block
    const int gi = threadIdx.x / G;
    const int gl = threadIdx.x & (G - 1);
    const int iters = (nr + ng - 1) / ng;
    for (int it = 0; it < iters; ++it) {
        const int r = it * ng + gi;
        const bool valid = r < nr;
        const T* row = v + (size_t)(valid ? r : 0) * S;
        T acc = 0;
        if (wts) {
            for (int j = gl; j < S; j += G) acc += wts[j] * row[j];
        } else {
            for (int j = gl; j < S; j += G) acc += row[j];
        }
        acc = group_sum(acc, G);
        const T mean = wts ? acc / wsum : acc / (T)S;
        T acc2 = 0;
        if (wts) {
            for (int j = gl; j < S; j += G) {
                const T d = row[j] - mean;
                acc2 += wts[j] * d * d;
            }
        } else {
            for (int j = gl; j < S; j += G) {
                const T d = row[j] - mean;
                acc2 += d * d;
            }
        }
        acc2 = group_sum(acc2, G);
        if (valid && gl == 0) {
            const T var = wts ? acc2 / wsum : acc2 / (T)(S - ddof);   // S == ddof -> 0/0 = NaN like NumPy
            out[r] = (OutT)sqrt(var);
        }
    }
}

}  // namespace rvmc
