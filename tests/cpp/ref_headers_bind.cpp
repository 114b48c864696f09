// ref_headers_bind.cpp — host code written against the REFERENCE's own classes (FrameBuffer, ImageF32, VertexBuffer,
// VaryingDescriptor from /root/reference/include, linked against the unmodified reference build oracle/_ref/libifrit_ref.so
// for their out-of-line members) that drives THIS renderer through the templated binders of
// include/ifrit_b200/TileRasterRendererCuda.hpp — the code a maintainer would have after switching
// TileRasterRendererCuda implementations (projects/softgraphics/demo/DefaultDemo.cpp:232-375).
// Built by oracle/Makefile (target ref_bind) only where /root/reference exists; the binary travels in oracle/_ref/.
//   ref_bind                 -> instantiation check only (no GPU)
//   ref_bind <scene> <out>   -> draws the scene file on cuda:0 and writes the colour image (raw float32) to <out>
// Scene file: int32 {W, H, nverts, nattr, nidx, fs}, float32 verts[nverts * nattr * 4], int32 idx[nidx], float32 uniform[32].
#include "ifrit/softgraphics/engine/base/FrameBuffer.h"
#include "ifrit/softgraphics/engine/base/VertexBuffer.h"
#include "ifrit/softgraphics/engine/base/VaryingDescriptor.h"
#include "ifrit_b200/TileRasterRendererCuda.hpp"
#include <cstdio>
#include <type_traits>
#include <vector>

using namespace Ifrit::Graphics::SoftGraphics;
using Renderer = ifrit_b200::TileRasterRendererCuda;

template void Renderer::bindFrameBuffer<FrameBuffer>(FrameBuffer&, bool);
template void Renderer::bindVertexBuffer<VertexBuffer>(const VertexBuffer&);
template void Renderer::bindVertexShader<VaryingDescriptor>(int, VaryingDescriptor&);

int main(int argc, char** argv) {
    static_assert(std::is_convertible<decltype(std::declval<ImageF32&>().getData()), const float*>::value, "ImageF32::getData");
    static_assert(std::is_convertible<decltype(std::declval<const VertexBuffer&>().getVertexCount()), int>::value, "vertex count");
    static_assert(std::is_convertible<decltype(std::declval<VaryingDescriptor&>().getVaryingCounts()), int>::value, "varying count");
    static_assert(sizeof(ifrit_b200::Float4) == 16, "Vector4f layout");
    if (argc < 3) { std::printf("binders instantiate against the reference headers\n"); return 0; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int hd[6];
    if (std::fread(hd, 4, 6, f) != 6) return 3;
    const int W = hd[0], H = hd[1], nverts = hd[2], nattr = hd[3], nidx = hd[4], fs = hd[5];
    std::vector<float> verts((size_t)nverts * nattr * 4), uni(32);
    std::vector<int> idx(nidx);
    if (std::fread(verts.data(), 4, verts.size(), f) != verts.size() || std::fread(idx.data(), 4, idx.size(), f) != idx.size() ||
        std::fread(uni.data(), 4, 32, f) != 32) return 3;
    std::fclose(f);
    try {
        ImageF32 color(W, H, 4), depth(W, H, 1);
        FrameBuffer fb;
        fb.SetColorAttachments({&color});
        fb.SetDepthAttachment(depth);
        VertexBuffer vb;
        std::vector<TypeDescriptor> layout(nattr, TypeDescriptors.FLOAT4);
        vb.setLayout(layout);
        vb.allocateBuffer(nverts);
        for (int i = 0; i < nverts; i++)
            for (int a = 0; a < nattr; a++) {
                const float* v = &verts[((size_t)i * nattr + a) * 4];
                vb.setValue<Vector4f>(i, a, Vector4f(v[0], v[1], v[2], v[3]));
            }
        VaryingDescriptor vd;
        vd.setVaryingDescriptors(std::vector<TypeDescriptor>(nattr - 1, TypeDescriptors.FLOAT4));
        Renderer renderer(0);
        renderer.Init();
        renderer.bindFrameBuffer(fb, /*useDoubleBuffer=*/false);
        renderer.bindVertexBuffer(vb);
        renderer.bindIndexBuffer(idx);
        renderer.bindVertexShader(IFRIT_VS_MVP, vd);
        renderer.bindFragmentShader(fs);
        renderer.CreateBuffer(0, 128);
        renderer.copyHostBufferToBuffer(uni.data(), 0, 128);
        renderer.SetDepthFunc(IF_COMPARE_OP_LESS);
        renderer.SetDepthTestEnable(true);
        renderer.SetCullMode(IF_CULL_MODE_BACK);
        renderer.setClearValues({Vector4f(0.f, 0.f, 0.f, 0.f)}, 255.0f);
        renderer.clear();
        renderer.drawElements();
        renderer.synchronize();
        FILE* o = std::fopen(argv[2], "wb");
        if (!o) return 3;
        std::fwrite(color.getData(), 4, (size_t)W * H * 4, o);
        std::fclose(o);
        std::printf("ref_bind: drew %d triangles into a reference FrameBuffer\n", nidx / 3);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_bind failed: %s\n", e.what());
        return 1;
    }
}
