// TileRasterRendererCuda.hpp — header-only C++ mirror of the reference's
//   Ifrit::Graphics::SoftGraphics::TileRaster::CUDA::TileRasterRendererCuda
//   (include/ifrit/softgraphics/engine/tilerastercuda/TileRasterRendererCuda.h:81-115)
// over the C-ABI of ifrit_b200.h: same verb names, same argument meaning, same non-owning binding
// contract.  The bind* templates accept the reference's own FrameBuffer / VertexBuffer /
// VaryingDescriptor types (they only use GetColorAttachment / GetDepthAttachment / getData /
// GetBufferUnsafe / getVertexCount / getAttributeCount / getVaryingCounts — FrameBuffer.h:48-51,
// Image.h:141, VertexBuffer.h:42-86, VaryingDescriptor.h:47-52), so a reference caller switches
// renderers by changing one include and the shader-object arguments (registry ids instead of host
// virtual objects).  Errors throw std::runtime_error (the reference aborts).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "../ifrit_b200.h"

namespace ifrit_b200 {

// Four floats {x, y, z, w}.  The reference's math header defines Vector4f as a MACRO (ifrit/core/math/VectorDefs.h:211,
// CoreVec4<float>, same layout); when it has been included before this header that type is used as is.
#ifdef Vector4f
using Float4 = Vector4f;
#else
struct Vector4f { float x, y, z, w; };
using Float4 = Vector4f;
#endif

class TileRasterRendererCuda {
public:
    explicit TileRasterRendererCuda(int cudaDevice = 0) { ck(iftrcuCreateInstance(cudaDevice, &h_), "iftrcuCreateInstance"); }
    ~TileRasterRendererCuda() { if (h_) iftrcuDestroyInstance(h_); }
    TileRasterRendererCuda(const TileRasterRendererCuda&) = delete;
    TileRasterRendererCuda& operator=(const TileRasterRendererCuda&) = delete;

    void Init() { ck(iftrcuInit(h_)); }

    // raw-pointer forms (what the templates below reduce to)
    void bindFrameBuffer(float* colorRGBA32F, float* depthF32, int width, int height, bool useDoubleBuffer = true) {
        ck(iftrcuBindFrameBuffer(h_, colorRGBA32F, depthF32, width, height, useDoubleBuffer));
    }
    void bindVertexBuffer(const void* data, int vertexCount, int numFloat4Attributes) {
        std::vector<int> layout(numFloat4Attributes, IFTP_FLOAT4);
        ck(iftrcuBindVertexBuffer(h_, data, vertexCount, layout.data(), numFloat4Attributes));
    }
    void bindIndexBuffer(const std::vector<int>& indexBuffer) {
        ck(iftrcuBindIndexBuffer(h_, indexBuffer.data(), (int)indexBuffer.size()));
    }
    // reference-typed forms
    template <class FrameBufferT>
    void bindFrameBuffer(FrameBufferT& fb, bool useDoubleBuffer = true) {
        auto* c = fb.GetColorAttachment(0);
        auto* d = fb.GetDepthAttachment();
        bindFrameBuffer(c->getData(), d ? d->getData() : nullptr, (int)c->GetWidth(), (int)c->GetHeight(), useDoubleBuffer);
    }
    template <class VertexBufferT>
    void bindVertexBuffer(const VertexBufferT& vb) {
        bindVertexBuffer(vb.GetBufferUnsafe(), vb.getVertexCount(), vb.getAttributeCount());
    }
    // shaders: registry id (IFRIT_VS_* / IFRIT_FS_*) or name instead of VertexShader* / FragmentShader*
    void bindVertexShader(int vertexShaderId, int numVaryings) {
        std::vector<int> layout(numVaryings > 0 ? numVaryings : 1, IFTP_FLOAT4);
        ck(iftrcuBindVertexShader(h_, vertexShaderId, layout.data(), numVaryings));
    }
    template <class VaryingDescriptorT>
    void bindVertexShader(int vertexShaderId, VaryingDescriptorT& varyingDescriptor) {
        bindVertexShader(vertexShaderId, (int)varyingDescriptor.getVaryingCounts());
    }
    void bindVertexShader(const char* name, int numVaryings) { bindVertexShader(iftrcuLookupVertexShader(name), numVaryings); }
    void bindFragmentShader(int fragmentShaderId) { ck(iftrcuBindFragmentShader(h_, fragmentShaderId)); }
    void bindFragmentShader(const char* name) { bindFragmentShader(iftrcuLookupFragmentShader(name)); }
    // CUDA source of `__device__ float4 fs_main(const IfritFsInput& in)` compiled at run time (NVRTC) -> id for bindFragmentShader
    int createFragmentShaderFromSource(const char* name, const char* cudaSource, int numVaryings, bool needsDerivatives) {
        int id = -1;
        ck(iftrcuCreateFragmentShaderFromSource(h_, name, cudaSource, numVaryings, needsDerivatives ? 1 : 0, &id));
        return id;
    }
    void bindGeometryShader(int id) { ck(iftrcuBindGeometryShader(h_, id)); }
    // mesh / task front end: shaders are CUDA source of ms_main / ts_main compiled at run time (include/ifrit_b200.h)
    int createMeshShaderFromSource(const char* name, const char* cudaSource) {
        int id = -1; ck(iftrcuCreateMeshShaderFromSource(h_, name, cudaSource, &id)); return id;
    }
    int createTaskShaderFromSource(const char* name, const char* cudaSource) {
        int id = -1; ck(iftrcuCreateTaskShaderFromSource(h_, name, cudaSource, &id)); return id;
    }
    void bindMeshShader(int meshShaderId, int numVaryings, int lx, int ly, int lz) {
        std::vector<int> layout(numVaryings > 0 ? numVaryings : 1, IFTP_FLOAT4);
        ck(iftrcuBindMeshShader(h_, meshShaderId, layout.data(), numVaryings, lx, ly, lz));
    }
    template <class VaryingDescriptorT>
    void bindMeshShader(int meshShaderId, VaryingDescriptorT& varyingDescriptor, int lx, int ly, int lz) {
        bindMeshShader(meshShaderId, (int)varyingDescriptor.getVaryingCounts(), lx, ly, lz);
    }
    void bindTaskShader(int taskShaderId) { ck(iftrcuBindTaskShader(h_, taskShaderId, nullptr, 0)); }

    void CreateTexture(int slotId, const IfritImageCreateInfo& createInfo) { ck(iftrcuCreateTexture(h_, slotId, &createInfo)); }
    void CreateSampler(int slotId, const IfritSamplerT& samplerState) { ck(iftrcuCreateSampler(h_, slotId, &samplerState)); }
    void generateMipmap(int slotId, IfritFilter filter) { ck(iftrcuGenerateMipmap(h_, slotId, filter)); }
    void blitImage(int srcSlotId, int dstSlotId, const IfritImageBlit& region, IfritFilter filter) {
        ck(iftrcuBlitImage(h_, srcSlotId, dstSlotId, &region, filter));
    }
    void copyHostBufferToImage(void* srcBuffer, int dstSlot, const std::vector<IfritBufferImageCopy>& regions) {
        ck(iftrcuCopyHostBufferToImage(h_, srcBuffer, dstSlot, regions.data(), (int)regions.size()));
    }
    void CreateBuffer(int slotId, int bufSize) { ck(iftrcuCreateBuffer(h_, slotId, bufSize)); }
    void copyHostBufferToBuffer(const void* srcBuffer, int dstSlot, int size) { ck(iftrcuCopyHostBufferToBuffer(h_, srcBuffer, dstSlot, size)); }

    void SetScissors(const std::vector<Float4>& scissors) { ck(iftrcuSetScissors(h_, scissors.empty() ? nullptr : &scissors[0].x, (int)scissors.size())); }
    void setScissorTestEnable(bool option) { ck(iftrcuSetScissorTestEnable(h_, option)); }
    void setMsaaSamples(IfritSampleCountFlagBits msaaSamples) { ck(iftrcuSetMsaaSamples(h_, msaaSamples)); }
    void setRasterizerPolygonMode(IfritPolygonMode mode) { ck(iftrcuSetRasterizerPolygonMode(h_, mode)); }
    void setBlendFunc(IfritColorAttachmentBlendState state) { ck(iftrcuSetBlendFunc(h_, &state)); }
    void SetDepthFunc(IfritCompareOp depthFunc) { ck(iftrcuSetDepthFunc(h_, depthFunc)); }
    void SetDepthTestEnable(bool option) { ck(iftrcuSetDepthTestEnable(h_, option)); }
    void SetCullMode(IfritCullMode cullMode) { ck(iftrcuSetCullMode(h_, cullMode)); }
    void setClearValues(const std::vector<Float4>& clearColors, float clearDepth) {
        ck(iftrcuSetClearValues(h_, clearColors.empty() ? nullptr : &clearColors[0].x, (int)clearColors.size(), clearDepth));
    }

    void clear() { ck(iftrcuClear(h_)); }
    void drawElements() { ck(iftrcuDrawElements(h_)); }
    void DrawMeshTasks(int numWorkGroups, int firstWorkGroup) { ck(iftrcuDrawMeshTasks(h_, numWorkGroups, firstWorkGroup)); }

    // extensions
    void synchronize() { ck(iftrcuSynchronize(h_)); }
    void setTilePartition(int rank, int world) { ck(iftrcuSetTilePartition(h_, rank, world)); }
    void setSampleConvention(int c) { ck(iftrcuSetSampleConvention(h_, c)); }
    void setFramePipelining(bool on) { ck(iftrcuSetFramePipelining(h_, on)); }
    // packed RGBA8 presentation image, copied back instead of the RGBA32F one (hostRGBA8: width * height * 4 bytes or null)
    void bindPresentBuffer(void* hostRGBA8, bool enable = true) { ck(iftrcuBindPresentBuffer(h_, hostRGBA8, enable ? 1 : 0)); }
    void readPresent(void* dstRGBA8) { ck(iftrcuReadPresent(h_, dstRGBA8)); }
    iftrcuStats stats() { iftrcuStats s; ck(iftrcuGetStats(h_, &s)); return s; }
    iftrcuRenderer* handle() { return h_; }

private:
    void ck(int rc, const char* what = nullptr) {
        if (rc != IFRIT_OK)
            throw std::runtime_error(std::string("ifrit_b200 error ") + std::to_string(rc) + ": " +
                                     (h_ ? iftrcuGetLastError(h_) : (what ? what : "no CUDA device (no CPU fallback)")));
    }
    iftrcuRenderer* h_ = nullptr;
};

}  // namespace ifrit_b200
