// Shared host-side plumbing of librvmc_b200.so: error reporting, device gate, launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rvmc_b200.h"

namespace rvmc {

// thread-local message returned by rvmc_last_error()
void set_error(const char* fmt, ...);
// maps a cudaError_t to the ABI's return convention (0 or > 0) and records the message
int cuda_status(cudaError_t e, const char* what);
// RVMC_OK when the CURRENT device is an sm_100 part; there is no fallback of any kind
int require_device();
// number of SMs of the current device (cached)
int sm_count();

inline unsigned int blocks_for(int64_t n, int threads) {
    return (unsigned int)((n + threads - 1) / threads);
}
// grid of a grid-stride elementwise kernel: enough blocks to fill the machine, no more
inline unsigned int grid_stride_blocks(int64_t n, int threads, int blocks_per_sm = 16) {
    const int64_t need = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
    return (unsigned int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace rvmc

// Device-side bounds checks for debug builds (nvcc -DRVMC_BOUNDS_CHECK): compute-sanitizer is not
// available on every pool, so the shared-memory indexing of the fused kernels can assert itself.
#if defined(RVMC_BOUNDS_CHECK)
#include <assert.h>
#define RVMC_DEV_ASSERT(cond) assert(cond)
#else
#define RVMC_DEV_ASSERT(cond) ((void)0)
#endif

#define RVMC_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            rvmc::set_error(__VA_ARGS__);    \
            return RVMC_ERR_INVALID;         \
        }                                    \
    } while (0)

#define RVMC_CUDA(call)                                      \
    do {                                                     \
        int _rc = rvmc::cuda_status((call), #call);          \
        if (_rc) return _rc;                                 \
    } while (0)

#define RVMC_DEVICE_GATE()                     \
    do {                                       \
        int _rc = rvmc::require_device();      \
        if (_rc) return _rc;                   \
    } while (0)

#define RVMC_LAUNCH_CHECK(name) RVMC_CUDA(cudaGetLastError())
