// Pipe-throughput microbenchmarks: the roofline denominators of the RVMC kernels.
//
// The hot path is a transcendental/FMA/integer pipeline with no dense contraction and almost no
// HBM traffic, so its roofline is the SM's instruction pipes, which
// MEASURED_PEAKS.json does not cover.  These kernels measure, on the device the bench runs on
// and at the clocks it actually holds, the sustained rate of
//   kind 0: FFMA        (fp32 fused multiply-add, 2 flop per lane-op)
//   kind 1: MUFU.EX2    (special-function unit)
//   kind 2: IMAD.WIDE   (32x32->64 integer multiply, the Philox round primitive)
//   kind 3: DFMA        (fp64 fused multiply-add)
//   kind 4: mixed issue (independent FFMA + LOP3 + MUFU streams; an issue-slot estimate)
// Each thread runs `iters` rounds of 8 independent dependency chains; the grid fills every SM
// with 2048 resident threads.  The result is lane-operations per second.
#include "common.cuh"

namespace rvmc {

template <int KIND>
__global__ void __launch_bounds__(256) pipe_kernel(int iters, float seedf, unsigned long long* sink) {
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (KIND == 0) {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = seedf + (float)(tid + j);
        const float a = 1.0000001f, b = seedf;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], a, b);
        }
        float s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += x[j];
        if (s == 12345.678f) sink[0] = tid;
    } else if (KIND == 1) {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = seedf * (float)(j + 1) * 1e-3f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        }
        float s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += x[j];
        if (s == 12345.678f) sink[0] = tid;
    } else if (KIND == 2) {
        unsigned long long x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = tid * 2654435761u + j;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned int lo = (unsigned int)x[j], hi = (unsigned int)(x[j] >> 32);
                    x[j] = (unsigned long long)(lo ^ hi) * 0xD2511F53ull;     // LOP3 + IMAD.WIDE.U32
                }
        }
        unsigned long long s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) s ^= x[j];
        if (s == 0x1234567ull) sink[0] = tid;
    } else if (KIND == 3) {
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (double)seedf + (double)(tid + j);
        const double a = 1.0000000001, b = (double)seedf;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
        }
        double s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += x[j];
        if (s == 12345.678) sink[0] = tid;
    } else {
        float x[4], m[2];
        unsigned int u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[j] = seedf + (float)(tid + j);
            u[j] = tid * 2654435761u + j;
        }
        m[0] = seedf * 1e-3f;
        m[1] = seedf * 2e-3f;
        const float a = 1.0000001f, b = seedf;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    x[j] = fmaf(x[j], a, b);
                    u[j] = (u[j] ^ (u[j] >> 7)) + 0x9E3779B9u;
                }
                if ((r & 3) == 0) {
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[0]));
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[1]));
                }
            }
        }
        float s = m[0] + m[1];
        unsigned int t = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s += x[j];
            t ^= u[j];
        }
        if (s == 12345.678f && t == 77u) sink[0] = tid;
    }
}

}  // namespace rvmc

using namespace rvmc;

extern "C" int rvmc_pipe_peak(int kind, int iters, double* lane_ops_per_s, double* elapsed_ms) {
    RVMC_DEVICE_GATE();
    RVMC_REQUIRE(kind >= 0 && kind <= 4, "kind must be in [0,4]");
    RVMC_REQUIRE(iters >= 1 && iters <= (1 << 24), "iters must be in [1, 2^24]");
    RVMC_REQUIRE(lane_ops_per_s != nullptr, "lane_ops_per_s is NULL");
    unsigned long long* sink = nullptr;
    RVMC_CUDA(cudaMalloc(&sink, 8));
    const unsigned int blocks = (unsigned int)sm_count() * 8;
    cudaEvent_t e0, e1;
    RVMC_CUDA(cudaEventCreate(&e0));
    RVMC_CUDA(cudaEventCreate(&e1));
    float ms = 0.f;
    for (int rep = 0; rep < 2; ++rep) {          // first repetition warms up
        cudaEventRecord(e0, 0);
        switch (kind) {
            case 0: pipe_kernel<0><<<blocks, 256>>>(iters, 1.5f, sink); break;
            case 1: pipe_kernel<1><<<blocks, 256>>>(iters, 1.5f, sink); break;
            case 2: pipe_kernel<2><<<blocks, 256>>>(iters, 1.5f, sink); break;
            case 3: pipe_kernel<3><<<blocks, 256>>>(iters, 1.5f, sink); break;
            default: pipe_kernel<4><<<blocks, 256>>>(iters, 1.5f, sink); break;
        }
        cudaEventRecord(e1, 0);
        int rc = cuda_status(cudaEventSynchronize(e1), "pipe_kernel");
        if (rc) {
            cudaFree(sink);
            return rc;
        }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    // lane-ops per thread per iteration: kinds 0-3: 64 ; kind 4: 8*(4 FFMA + 4*3 int) + 4 MUFU
    const double per_iter = (kind == 4) ? (8.0 * (4 + 12) + 4.0) : 64.0;
    const double total = (double)blocks * 256.0 * (double)iters * per_iter;
    *lane_ops_per_s = total / ((double)ms * 1e-3);
    if (elapsed_ms) *elapsed_ms = (double)ms;
    return RVMC_OK;
}
