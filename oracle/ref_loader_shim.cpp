// ref_loader_shim.cpp — C entry points around the reference's own OBJ loader (TEST INFRASTRUCTURE).
// Compiled by oracle/Makefile TOGETHER WITH projects/softgraphics/src/utility/loader/WavefrontLoader.cpp from where it
// lies under /root/reference into oracle/_ref/libifrit_ref_loader.so; tests/test_loaders.py pins ifrit_v2_b200.loaders
// (and the C++ twin include/ifrit_b200/WavefrontLoader.hpp) against it.
#include "ifrit/softgraphics/utility/loader/WavefrontLoader.h"
#include <cstring>
using namespace Ifrit::Graphics::SoftGraphics::Utility::Loader;

static std::vector<Vector3f> gV, gN;
static std::vector<Vector2f> gT;
static std::vector<uint32_t> gI;

extern "C" __attribute__((visibility("default"))) int refld_load(const char* path, int* counts) {
    gV.clear(); gN.clear(); gT.clear(); gI.clear();
    WavefrontLoader ld;
    ld.loadObject(path, gV, gN, gT, gI);
    counts[0] = (int)gV.size(); counts[1] = (int)gN.size(); counts[2] = (int)gT.size(); counts[3] = (int)gI.size();
    return 0;
}
extern "C" __attribute__((visibility("default"))) void refld_get(float* v, float* n, float* t, unsigned* idx, float* remappedNormals, float* remappedUVs) {
    for (size_t i = 0; i < gV.size(); i++) { v[3 * i] = gV[i].x; v[3 * i + 1] = gV[i].y; v[3 * i + 2] = gV[i].z; }
    for (size_t i = 0; i < gN.size(); i++) { n[3 * i] = gN[i].x; n[3 * i + 1] = gN[i].y; n[3 * i + 2] = gN[i].z; }
    for (size_t i = 0; i < gT.size(); i++) { t[2 * i] = gT[i].x; t[2 * i + 1] = gT[i].y; }
    if (!gI.empty()) memcpy(idx, gI.data(), gI.size() * 4);
    WavefrontLoader ld;
    if (remappedNormals && !gN.empty()) {
        auto rn = ld.RemapNormals(gN, gI, (int)gV.size());
        for (size_t i = 0; i < rn.size(); i++) { remappedNormals[3 * i] = rn[i].x; remappedNormals[3 * i + 1] = rn[i].y; remappedNormals[3 * i + 2] = rn[i].z; }
    }
    if (remappedUVs && !gT.empty()) {
        auto ru = ld.RemapUVs(gT, gI, (int)gV.size());
        for (size_t i = 0; i < ru.size(); i++) { remappedUVs[2 * i] = ru[i].x; remappedUVs[2 * i + 1] = ru[i].y; }
    }
}
