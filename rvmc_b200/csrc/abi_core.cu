// librvmc_b200.so -- library plumbing (errors, device gate) and the two elementary
// generator entries (K1): raw Philox blocks and fp64 normals.
//
// Reference code replaced: np.random.normal at RVMC.py:101 (cluster_velocities) and RVMC.py:224
// (sigma_dyn term).  Everything else in this file is new boundary plumbing; the reference is
// pure Python and has none.
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "pop_derive.h"

namespace rvmc {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

int cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return RVMC_OK;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    cudaGetLastError();   // clear the sticky-less error state
    return (int)e;
}

struct DeviceState {
    int checked = 0;      // 0 = not yet, 1 = ok, -1 = unusable
    int sms = 0;
};
static std::mutex g_mu;
static std::vector<DeviceState> g_dev;

static DeviceState* current_state(int* rc_out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0) {
        cudaGetLastError();
        set_error("no CUDA device is available (%s); rvmc_b200 has no CPU fallback",
                  e == cudaSuccess ? "no current device" : cudaGetErrorString(e));
        *rc_out = RVMC_ERR_NO_DEVICE;
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if ((int)g_dev.size() <= dev) g_dev.resize(dev + 1);
    DeviceState& s = g_dev[dev];
    if (s.checked == 0) {
        int major = 0, minor = 0, sms = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        s.sms = sms;
        s.checked = (major == 10 && minor == 0) ? 1 : -1;
        if (s.checked > 0) {
            // the stream-ordered allocations of the grid entries (descriptors, counters: ~1 MB per call) stay
            // in the device's default pool between calls instead of going back to the driver at every sync
            cudaMemPool_t pool = nullptr;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool) {
                unsigned long long keep = 64ull << 20;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
        }
        if (s.checked < 0)
            set_error("device %d is sm_%d%d; the kernels are built for sm_100a only and there is no "
                      "fallback path", dev, major, minor);
    }
    if (s.checked < 0) {
        if (g_error.empty()) set_error("device %d is not an sm_100 part", dev);
        *rc_out = RVMC_ERR_NO_DEVICE;
        return nullptr;
    }
    *rc_out = RVMC_OK;
    return &s;
}

int require_device() {
    int rc;
    current_state(&rc);
    return rc;
}

int sm_count() {
    int rc;
    DeviceState* s = current_state(&rc);
    return s ? s->sms : 0;
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void philox_blocks_kernel(uint64_t seed, uint64_t first_item, int64_t n, uint32_t stream,
                                     uint32_t sub, uint4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u32x4 w = philox_block(seed, first_item + (uint64_t)i, stream, sub);
        out[i] = make_uint4(w.x, w.y, w.z, w.w);
    }
}

// One Philox block per PAIR of items: item i uses the cos branch of pair i/2 when i is even, the
// sin branch when odd (so a value depends on (seed, item) only).
__global__ void normal_f64_kernel(uint64_t seed, uint64_t first_item, int64_t n, uint32_t stream,
                                  uint32_t sub, double mean, double sigma, int accumulate,
                                  double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t item = first_item + (uint64_t)i;
        const u32x4 w = philox_block(seed, item >> 1, stream, sub);
        double z0, z1;
        box_muller_f64(w.x, w.y, z0, z1);
        const double z = (item & 1) ? z1 : z0;
        const double v = mean + sigma * z;
        out[i] = accumulate ? out[i] + v : v;
    }
}

}  // namespace rvmc

using namespace rvmc;

extern "C" {

int rvmc_abi_version(void) { return RVMC_ABI_VERSION; }

const char* rvmc_last_error(void) { return g_error.c_str(); }

void rvmc_pop_defaults(rvmc_pop_params* p) {
    if (p) pop_defaults(p);
}

int rvmc_device_info(int device, int* sms, int* cc_major, int* cc_minor, int* clock_khz) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device is available (%s); rvmc_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return RVMC_ERR_NO_DEVICE;
    }
    RVMC_REQUIRE(device >= 0 && device < count, "device %d out of range [0,%d)", device, count);
    int major = 0, minor = 0, n = 0, khz = 0;
    RVMC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    RVMC_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    RVMC_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
    RVMC_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    if (sms) *sms = n;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    if (clock_khz) *clock_khz = khz;
    if (!(major == 10 && minor == 0)) {
        set_error("device %d is sm_%d%d; the kernels are built for sm_100a only", device, major, minor);
        return RVMC_ERR_NO_DEVICE;
    }
    return RVMC_OK;
}

int rvmc_philox_blocks(uint64_t seed, uint64_t first_item, int64_t n, uint32_t stream, uint32_t sub,
                       uint32_t* out, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n >= 0, "n must be >= 0 (got %lld)", (long long)n);
    if (n == 0) return RVMC_OK;
    RVMC_REQUIRE(out != nullptr, "out is NULL");
    const int threads = 256;
    const unsigned int blocks = grid_stride_blocks(n, threads);
    philox_blocks_kernel<<<blocks, threads, 0, (cudaStream_t)cuda_stream>>>(seed, first_item, n, stream, sub,
                                                                           (uint4*)out);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

int rvmc_normal_f64(uint64_t seed, uint64_t first_item, int64_t n, uint32_t stream, uint32_t sub,
                    double mean, double sigma, int accumulate, double* out, void* cuda_stream) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(n >= 0, "n must be >= 0 (got %lld)", (long long)n);
    if (n == 0) return RVMC_OK;
    RVMC_REQUIRE(out != nullptr, "out is NULL");
    const int threads = 256;
    const unsigned int blocks = grid_stride_blocks(n, threads);
    normal_f64_kernel<<<blocks, threads, 0, (cudaStream_t)cuda_stream>>>(seed, first_item, n, stream, sub, mean,
                                                                        sigma, accumulate, out);
    RVMC_LAUNCH_CHECK();
    return RVMC_OK;
}

}  // extern "C"
