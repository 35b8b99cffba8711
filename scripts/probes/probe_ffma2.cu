// Issue/pipe throughput of packed FP32 (FFMA2) against scalar FFMA on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_ffma2 probe_ffma2.cu && ./probe_ffma2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int MODE>  // 0: scalar FFMA x8 chains, 1: FFMA2 x8 chains, 2: 4 FFMA2 + 4 FFMA interleaved, 3: FFMA2 + IADD interleaved
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float a[8];
    f32x2 p[8];
    uint32_t n[8];
    for (int i = 0; i < 8; i++) {
        a[i] = threadIdx.x * 1e-3f + i;
        p[i] = ((f32x2)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f);
        n[i] = threadIdx.x + i;
    }
    const f32x2 s2 = ((f32x2)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(s));
            if (MODE == 1) p[i] = fma2(p[i], s2, s2);
            if (MODE == 2) {
                if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(s));
                else p[i] = fma2(p[i], s2, s2);
            }
            if (MODE == 3) {
                if (i & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(n[i]) : "r"(n[i ^ 1]));
                else p[i] = fma2(p[i], s2, s2);
            }
        }
    }
    float acc = 0;
    for (int i = 0; i < 8; i++) acc += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32)) + n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, float* out, int sms) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, 100, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, iters, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // warp-instructions of the inner body per SM sub-partition per second
    const double winst = (double)sms * 8 * 8 /*warps*/ * iters * 8.0;
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %.3f warp-inst/clk/SMSP (at %d MHz nominal)\n", name, ms, winst / (ms * 1e-3) / (sms * 4.0) / (clk * 1e3), clk / 1000);
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, sms * 8 * 256 * sizeof(float));
    run<0>("FFMA", out, sms);
    run<1>("FFMA2", out, sms);
    run<2>("FFMA2+FFMA 1:1", out, sms);
    run<3>("FFMA2+IADD 1:1", out, sms);
    return 0;
}
