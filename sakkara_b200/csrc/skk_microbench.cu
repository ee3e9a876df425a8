// ALU peak micro-benchmarks (the chain-per-lane path of configuration 5 is bound by the fp64
// pipe, and MEASURED_PEAKS.json has no fp64 / fp32 / MUFU entry): dependent-chain-free FMA and MUFU.EX2
// loops sized to fill every SM, timed with CUDA events.  Exposed through the C ABI as skk_alu_peak().
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/skk_b200.h"

namespace {

constexpr int kChains = 8;     // independent accumulators per thread (covers the pipe latency at 8+ warps per scheduler)
constexpr int kThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads) fma_peak_kernel(T *out, int iters, T a, T b) {
    T acc[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) acc[k] = (T)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < kChains; ++k) acc[k] = acc[k] * a + b;   // one FMA each
    T s = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += acc[k];
    if (s == (T)-1.2345) out[blockIdx.x * kThreads + threadIdx.x] = s;   // keeps the loop alive, never taken
}

__global__ void __launch_bounds__(kThreads) mufu_peak_kernel(float *out, int iters) {
    float acc[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) acc[k] = -1.0f - 0.01f * (float)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < kChains; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[k]));
    float s = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += acc[k];
    if (s == -1.2345f) out[blockIdx.x * kThreads + threadIdx.x] = s;
}

}  // namespace

extern "C" int skk_alu_peak(int32_t device, int32_t kind, double *ops_per_second) {
    if (!ops_per_second || kind < 0 || kind > 2) return SKK_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return SKK_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SKK_ERR_CUDA;
    const int grid = prop.multiProcessorCount * 8;   // 2048 threads per SM
    void *out = nullptr;
    if (cudaMalloc(&out, (size_t)grid * kThreads * sizeof(double)) != cudaSuccess) return SKK_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = kind == 0 ? 4096 : 16384;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0) fma_peak_kernel<double><<<grid, kThreads>>>(static_cast<double *>(out), iters, 1.0000001, 1e-9);
        else if (kind == 1) fma_peak_kernel<float><<<grid, kThreads>>>(static_cast<float *>(out), iters, 1.0000001f, 1e-9f);
        else mufu_peak_kernel<<<grid, kThreads>>>(static_cast<float *>(out), iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) {
            cudaFree(out);
            return SKK_ERR_CUDA;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)grid * kThreads * kChains * (double)iters / (ms * 1e-3);   // FMA (or MUFU) instructions per second
        if (rep > 0 && ops > best) best = ops;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *ops_per_second = best;
    return SKK_OK;
}
