// sb_probe.cu — measured denominators for the rasterizer's roofline ("FP32 and shared-memory peaks are
// nominal").  Two microbenchmarks, run live on the device bench.py runs on (MEASURED_PEAKS.json has HBM and bf16 only):
//   sb_probe_fp32_peak: FP32 instruction issue — independent FFMA chains, every lane busy; reported in lane-ops/s with an
//                       FMA counted ONCE (the unit of bench.py's rasterizer roofline: SURVEY 8d counts 24 lane-ops per pair);
//   sb_probe_smem_peak: shared-memory read bandwidth — conflict-free LDS.128 by every lane, in bytes/s.
// Both are timed with CUDA events on the caller's stream after a warm-up launch.
#include "sb_internal.h"

namespace sb {
namespace {

constexpr int kProbeThreads = 1024;
constexpr int kFmaChains = 8;
constexpr int kFmaInner = 64;

__global__ void __launch_bounds__(kProbeThreads) fp32_probe_kernel(float* out, int iters, float a, float b) {
    float acc[kFmaChains];
#pragma unroll
    for (int c = 0; c < kFmaChains; c++) acc[c] = (float)(threadIdx.x + c);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < kFmaInner; k++) {
#pragma unroll
            for (int c = 0; c < kFmaChains; c++) acc[c] = __fmaf_rn(acc[c], a, b);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < kFmaChains; c++) s += acc[c];
    if (s == 123.456f) out[0] = s;  // never true for the constants used; keeps the chains alive
}

constexpr int kLdsInner = 32;

__global__ void __launch_bounds__(kProbeThreads) smem_probe_kernel(uint32_t* out, int iters) {
    __shared__ uint4 buf[kProbeThreads];
    buf[threadIdx.x] = make_uint4(threadIdx.x, 1u, 2u, 3u);
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(buf);
    uint32_t x = 0, y = 0, z = 0, w = 0;
    uint32_t idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < kLdsInner; k++) {
            uint32_t a, b, c, d;
            // lane l of a warp reads 16 bytes at 16*l (+ a warp-uniform rotation): 4 conflict-free 128-byte wavefronts per LDS.128
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + ((idx + 32u * k) & (kProbeThreads - 1)) * 16u));
            x ^= a; y ^= b; z ^= c; w ^= d;
        }
    }
    if ((x ^ y ^ z ^ w) == 0xdeadbeefu) out[0] = x;
}

template <typename F>
cudaError_t time_launch(F&& launch, cudaStream_t stream, float* ms) {
    cudaEvent_t e0, e1;
    cudaError_t e = cudaEventCreate(&e0);
    if (e != cudaSuccess) return e;
    e = cudaEventCreate(&e1);
    if (e != cudaSuccess) return e;
    launch();  // warm-up
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0, stream);
        launch();
        cudaEventRecord(e1, stream);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float t = 0.0f;
        cudaEventElapsedTime(&t, e0, e1);
        best = t < best ? t : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms = best;
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace

cudaError_t probe_fp32_peak(int num_sms, cudaStream_t stream, double* lane_ops_per_s) {
    float* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 4);
    if (e != cudaSuccess) return e;
    const int iters = 2000, grid = num_sms * 2;
    float ms = 0.0f;
    e = time_launch([&] { fp32_probe_kernel<<<grid, kProbeThreads, 0, stream>>>(d, iters, 0.999f, 0.001f); }, stream, &ms);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    *lane_ops_per_s = (double)grid * kProbeThreads * (double)iters * kFmaInner * kFmaChains / (ms * 1e-3);
    return cudaSuccess;
}

cudaError_t probe_smem_peak(int num_sms, cudaStream_t stream, double* bytes_per_s) {
    uint32_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 4);
    if (e != cudaSuccess) return e;
    const int iters = 4000, grid = num_sms * 2;
    float ms = 0.0f;
    e = time_launch([&] { smem_probe_kernel<<<grid, kProbeThreads, 0, stream>>>(d, iters); }, stream, &ms);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    *bytes_per_s = (double)grid * kProbeThreads * (double)iters * kLdsInner * 16.0 / (ms * 1e-3);
    return cudaSuccess;
}

}  // namespace sb
