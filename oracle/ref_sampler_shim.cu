// ref_sampler_shim.cu — TEST INFRASTRUCTURE (never linked into the product).
//
// Compiles the REFERENCE's own texture sampling library
//   /root/reference/include/ifrit/softgraphics/engine/math/ShaderBuiltinCuda.cuh:97-272
// (TextureSampleImpl::textureImplLegacy / textureImplLod / textureImplAdaptiveLod /
// textureImplAnisotropicFilter / textureCubeImplLod — all IFRIT_DUAL, i.e. __host__ __device__) from where
// it lies and exposes its HOST instantiation through a flat C interface, so that the oracle's restatement
// of the sampler (oracle/ifrit_oracle.c: sample_level .. sample_cube) is pinned against the reference's own
// code (tests/test_oracle_pinned.py::test_sampler_restatement_vs_reference_source).
//
// The header is stale in this snapshot: it lives in namespace Ifrit::SoftRenderer while the
// types it names moved to Ifrit::Graphics::SoftGraphics.  The two using-directives below are the whole shim;
// no reference source is copied.  Built by oracle/Makefile into oracle/_ref/libifrit_ref_sampler.so with
// g++'s contraction switched on (-mfma -ffp-contract=fast), mimicking nvcc's default -fmad=true under which
// the reference's GPU build evaluates `a * b + c`.
#define IFRIT_FEATURE_CUDA 1
#include <cstdio>
#include <cstring>
#include "ifrit/softgraphics/engine/base/Structures.h"
#include "ifrit/softgraphics/engine/tilerastercuda/TileRasterConstantsCuda.h"
namespace Ifrit::SoftRenderer {
using namespace Ifrit::Graphics::SoftGraphics;
using namespace Ifrit::Graphics::SoftGraphics::TileRaster::CUDA;
}
#include "ifrit/softgraphics/engine/math/ShaderBuiltinCuda.cuh"

namespace SB = Ifrit::SoftRenderer::Math::ShaderOps::CUDA::TextureSampleImpl;
using Ifrit::Graphics::SoftGraphics::IfritSamplerT;

static IfritSamplerT make_sampler(const int* s) {   // {filter, addrU, addrV, border}
    IfritSamplerT t;
    t.filterMode = (Ifrit::Graphics::SoftGraphics::IfritFilter)s[0];
    t.addressModeU = (Ifrit::Graphics::SoftGraphics::IfritSamplerAddressMode)s[1];
    t.addressModeV = (Ifrit::Graphics::SoftGraphics::IfritSamplerAddressMode)s[2];
    t.addressModeW = t.addressModeU;
    t.borderColor = (Ifrit::Graphics::SoftGraphics::IfritBorderColor)s[3];
    return t;
}

extern "C" {

// kind 0: textureImplLegacy(uv, integer lod = (int)a[0], layer = (int)a[1])
// kind 1: textureImplLod(uv, lod = a[0])
// kind 2: textureImplAdaptiveLod(uv, dx = a[0..1], dy = a[2..3], bias = a[4])
// kind 3: textureImplAnisotropicFilter(maxAniso = (int)a[5], uv, du = a[0..1], dv = a[2..3], bias = a[4])
// kind 4: textureCubeImplLod(uvw = (uv[0], uv[1], a[1]), lod = a[0])      (uv holds x, y; a[1] holds z)
// args: n rows of 6 floats; uv: n rows of 2 floats; out: n rows of 4 floats
__attribute__((visibility("default")))
int refsb_sample(int kind, const int* sampler4, const float* tex, int w, int h, int layers, int n, const float* uv,
                 const float* args, float* out) {
    const IfritSamplerT s = make_sampler(sampler4);
    const float4* t = reinterpret_cast<const float4*>(tex);
    for (int i = 0; i < n; i++) {
        const float* a = args + 6 * (size_t)i;
        const float2 p = {uv[2 * (size_t)i], uv[2 * (size_t)i + 1]};
        float4 r = {0, 0, 0, 0};
        switch (kind) {
            case 0: r = SB::textureImplLegacy(s, t, w, h, layers, (int)a[1], p, int2{0, 0}, (int)a[0]); break;
            case 1: r = SB::textureImplLod(s, t, w, h, layers, p, int2{0, 0}, a[0]); break;
            case 2: r = SB::textureImplAdaptiveLod(s, t, w, h, layers, p, float2{a[0], a[1]}, float2{a[2], a[3]}, a[4], int2{0, 0}); break;
            case 3: r = SB::textureImplAnisotropicFilter(s, t, w, h, layers, (int)a[5], p, float2{a[0], a[1]}, float2{a[2], a[3]}, a[4], int2{0, 0}); break;
            case 4: r = SB::textureCubeImplLod(s, t, w, h, layers, float3{p.x, p.y, a[1]}, a[0]); break;
            default: return -1;
        }
        out[4 * (size_t)i] = r.x; out[4 * (size_t)i + 1] = r.y; out[4 * (size_t)i + 2] = r.z; out[4 * (size_t)i + 3] = r.w;
    }
    return 0;
}

}  // extern "C"
