// C++ restatement of the reference's e2e viewer test (tests/e2e/viewer.rs:70-95) through the
// typed host mirror: render ONE red Gaussian at 1024x1024 and count lit pixels per channel.
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "splat_b200.hpp"

int main() {
    using namespace splat_b200;
    try {
        Context ctx(0);
        SbGaussian g{};
        g.pos[2] = 1.0f;
        g.rot[3] = 1.0f;
        g.scale[0] = g.scale[1] = g.scale[2] = 1.0f;
        g.color[0] = 255;
        g.color[3] = 255;
        Viewer viewer(ctx, SB_TARGET_RGBA8_UNORM, {g});
        Camera cam(0.1f, 1e4f, 1.04719755f);  // tests/common/given.rs:6-12
        cam.yaw = 0.1f;
        cam.pitch = 0.1f;
        const uint32_t W = 1024, H = 1024;
        viewer.update_camera(cam, W, H);
        void* d_pixels = nullptr;
        if (cudaMalloc(&d_pixels, (size_t)W * H * 4) != cudaSuccess) return 2;
        SbTarget t{d_pixels, W * 4, W, H, SB_TARGET_RGBA8_UNORM, 0, 0};
        viewer.render(nullptr, t);
        std::vector<unsigned char> px((size_t)W * H * 4);
        if (cudaMemcpy(px.data(), d_pixels, px.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return 3;
        unsigned long long sum[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < px.size(); i++) sum[i & 3] += px[i] > 0;
        std::printf("lit pixels r=%llu g=%llu b=%llu a=%llu\n", sum[0], sum[1], sum[2], sum[3]);
        // tests/e2e/viewer.rs:89-95
        const bool ok = sum[0] > 1 && sum[1] < 1 && sum[2] < 1 && sum[3] > 1;
        // errors behave like the reference's Result types
        bool threw = false;
        try {
            viewer.update_gaussian_transform(1.0f, SB_MODE_SPLAT, 4, false, 3.0f);  // GaussianShDegree::new(4) -> None
        } catch (const Error&) {
            threw = true;
        }
        cudaFree(d_pixels);
        std::printf(ok && threw ? "PASS\n" : "FAIL\n");
        return ok && threw ? 0 : 1;
    } catch (const std::exception& e) {
        std::printf("exception: %s\n", e.what());
        return 4;
    }
}
