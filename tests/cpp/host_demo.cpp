// host_demo.cpp — C++ host code driving the renderer through the C-ABI, written the way the
// reference's GPU demo does (projects/softgraphics/demo/DefaultDemo.cpp:232-375: create images,
// bind framebuffer / vertex / index buffers and shaders, set state, clear(), drawElements()).
// Build: g++ -std=c++17 -Iinclude tests/cpp/host_demo.cpp -Lifrit-v2_b200/lib -lifrit_b200 -o host_demo
// Prints the number of covered pixels of KAT-2 (centre sampling) and exits 0 on success.
#include <cstdio>
#include <vector>
#include "ifrit_b200/TileRasterRendererCuda.hpp"

int main() {
    using namespace ifrit_b200;
    const int W = 64, H = 64;
    std::vector<float> color(W * H * 4, 0.f), depth(W * H, 0.f);
    // KAT-2 triangle, attrs: float4 pos | float4 colour
    std::vector<float> vb = {-0.5f, -0.5f, 0.5f, 1, 1, 0, 0, 1, 0.5f, -0.5f, 0.5f, 1, 0, 1, 0, 1, 0.0f, 0.5f, 0.5f, 1, 0, 0, 1, 1};
    std::vector<int> ib = {0, 1, 2};
    try {
        TileRasterRendererCuda renderer(0);
        renderer.Init();
        renderer.bindFrameBuffer(color.data(), depth.data(), W, H, /*useDoubleBuffer=*/false);
        renderer.bindVertexBuffer(vb.data(), 3, 2);
        renderer.bindIndexBuffer(ib);
        renderer.bindVertexShader("passthrough", 1);
        renderer.bindFragmentShader("varying0");
        renderer.SetDepthFunc(IF_COMPARE_OP_LESS);
        renderer.SetDepthTestEnable(true);
        renderer.SetCullMode(IF_CULL_MODE_BACK);
        renderer.setClearValues({{0, 0, 0, 0}}, 255.0f);
        renderer.clear();
        renderer.drawElements();
        renderer.synchronize();   // colour image copied back into `color` (CU:3899-3909)
        int covered = 0;
        for (int i = 0; i < W * H; i++) covered += color[4 * i + 3] > 0.f;
        const iftrcuStats s = renderer.stats();
        std::printf("host_demo: covered=%d assembled=%llu fragments=%llu\n", covered,
                    (unsigned long long)s.triangles_assembled, (unsigned long long)s.fragments_shaded);
        return (covered > 400 && covered < 600 && (unsigned long long)covered == s.fragments_shaded) ? 0 : 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "host_demo failed: %s\n", e.what());
        return 1;
    }
}
