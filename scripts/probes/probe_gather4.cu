// Probe: semantics of cp.async.bulk.tensor.2d tile::gather4 on sm_100a (no PTX docs in this image).
// Usage: probe_gather4 <box_rows> ; prints the 4 gathered rows or "timeout".
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int* status, uint32_t expect_bytes) {
    __shared__ __align__(128) float dst[4 * 12 * 4];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), dst_a = (uint32_t)__cvta_generic_to_shared(dst);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4 * 12 * 4; i++) dst[i] = -1.0f;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(expect_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
                "r"(dst_a), "l"(&tmap), "r"(bar_a), "r"(0), "r"(5), "r"(1), "r"(7), "r"(3)
            : "memory");
        int ok = 0;
        for (int tries = 0; tries < 2000000 && !ok; tries++) {
            uint32_t p;
            asm volatile("{ .reg .pred q; mbarrier.test_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }" : "=r"(p) : "r"(bar_a) : "memory");
            ok = p;
        }
        *status = ok;
        for (int i = 0; i < 4 * 12 * 4; i++) out[i] = dst[i];
    }
}

int main(int argc, char** argv) {
    const uint32_t box_rows = argc > 1 ? atoi(argv[1]) : 1;
    const uint32_t expect = argc > 2 ? atoi(argv[2]) : 4 * 48;
    const int N = 16;
    std::vector<float> h(N * 12);
    for (int r = 0; r < N; r++) for (int c = 0; c < 12; c++) h[r * 12 + c] = r * 100 + c;
    float *d, *out; int* st;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 4 * 12 * 4 * 4); cudaMalloc(&st, 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {12, (cuuint64_t)N}, gstride[1] = {48};
    cuuint32_t box[2] = {12, box_rows}, estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode box_rows=%u -> %d\n", box_rows, (int)r);
    if (r != CUDA_SUCCESS) return 1;
    probe<<<1, 32>>>(tmap, out, st, expect);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<float> o(4 * 12 * 4); int s = 0;
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost);
    printf("barrier complete: %d (expect_tx %u)\n", s, expect);
    for (int r2 = 0; r2 < 8; r2++) { for (int c = 0; c < 12; c++) printf("%6.0f", o[r2 * 12 + c]); printf("\n"); }
    return 0;
}
