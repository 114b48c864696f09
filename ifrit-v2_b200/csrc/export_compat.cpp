// export_compat.cpp — the reference's shared-object C ABI for its data objects (include/ifrit_b200_compat.h), host side only.
// The objects are plain containers with the reference's observable behaviour:
//   ImageF32      core/data/Image.h:30-249            W x H x C floats, zero-initialised, [(y*W + x)*C + c]
//   FrameBuffer   engine/base/FrameBuffer.cpp          non-owning colour attachment list + depth attachment; width / height of colour 0
//   VaryingDescriptor  engine/base/VaryingDescriptor.cpp   list of type descriptors (IFTP_UNDEFINED is stored as FLOAT4, EngineCommonExport.cpp:62-63)
//   VertexBuffer  engine/base/VertexBuffer.cpp:31-60   layout -> offsets / element size; allocateBuffer(vertexCount); interleaved bytes
//   TrivialBufferManager  engine/bufferman/BufferManager.cpp   byte buffers, bufferData = memcpy at an offset
// and the iftrcu*Handle binders unwrap them into the pointer / size form of include/ifrit_b200.h.
#include "../../include/ifrit_b200_compat.h"
#include <cstring>
#include <memory>
#include <new>
#include <vector>

struct ifcrImageF32 {
    size_t width, height, channel;
    std::vector<float> data;
};
struct iftrFrameBuffer {
    std::vector<ifcrImageF32*> color;
    ifcrImageF32* depth = nullptr;
    unsigned width = 0, height = 0;
};
struct iftrVaryingDescriptor {
    std::vector<int> desc;
};
struct iftrVertexBuffer {
    std::vector<uint8_t> buffer;
    std::vector<int> layout, offsets;
    int vertexCount = 0, elementSize = 0;
};
namespace {
struct BufferManager {};
struct Buffer {
    std::vector<uint8_t> bytes;
    std::weak_ptr<BufferManager> manager;
};
struct BufferManagerWrapper {
    std::shared_ptr<BufferManager> manager = std::make_shared<BufferManager>();
};
// TypeDescriptor sizes, engine/base/TypeDescriptor.h: N components x 4 bytes
int desc_size(int d) {
    switch (d) {
        case IFTP_FLOAT1: case IFTP_INT1: case IFTP_UINT1: return 4;
        case IFTP_FLOAT2: case IFTP_INT2: case IFTP_UINT2: return 8;
        case IFTP_FLOAT3: case IFTP_INT3: case IFTP_UINT3: return 12;
        case IFTP_FLOAT4: case IFTP_INT4: case IFTP_UINT4: return 16;
        default: return 0;
    }
}
int normalise_desc(int d) { return d == IFTP_UNDEFINED ? (int)IFTP_FLOAT4 : d; }   // EngineCommonExport.cpp:62-63, :108-109
}   // namespace

extern "C" {

ifcrImageF32* ifcrCreateImageFP32(size_t width, size_t height, size_t channel) {
    ifcrImageF32* im = new (std::nothrow) ifcrImageF32();
    if (!im) return nullptr;
    im->width = width; im->height = height; im->channel = channel;
    im->data.assign(width * height * channel, 0.0f);
    return im;
}
void ifcrDestroyImageFP32(ifcrImageF32* p) { delete p; }
float* ifcrGetImageRawDataFP32(ifcrImageF32* p) { return p ? p->data.data() : nullptr; }

iftrFrameBuffer* iftrCreateFrameBuffer(void) { return new (std::nothrow) iftrFrameBuffer(); }
void iftrDestroyFrameBuffer(iftrFrameBuffer* p) { delete p; }
void iftrSetFrameBufferColorAttachmentFP32(iftrFrameBuffer* p, ifcrImageF32* const* pImage, size_t nums) {
    if (!p) return;
    p->color.assign(pImage, pImage + nums);
    if (nums && pImage[0]) { p->width = (unsigned)pImage[0]->width; p->height = (unsigned)pImage[0]->height; }
}
void iftrSetFrameBufferDepthAttachmentFP32(iftrFrameBuffer* p, ifcrImageF32* pImage) {
    if (p) p->depth = pImage;
}

iftrVaryingDescriptor* iftrCreateVaryingDescriptor(void) { return new (std::nothrow) iftrVaryingDescriptor(); }
void iftrDestroyVaryingDescriptor(iftrVaryingDescriptor* p) { delete p; }
void iftrWriteVaryingDescriptor(iftrVaryingDescriptor* p, const iftrExportTypeDesc* pDesc, size_t num) {
    if (!p) return;
    p->desc.clear();
    for (size_t i = 0; i < num; i++)
        if (pDesc[i] >= IFTP_UNDEFINED && pDesc[i] <= IFTP_UINT4) p->desc.push_back(normalise_desc(pDesc[i]));
}

iftrVertexBuffer* iftrCreateVertexBuffer(void) { return new (std::nothrow) iftrVertexBuffer(); }
void iftrDestroyVertexBuffer(iftrVertexBuffer* p) { delete p; }
void iftrSetVertexBufferLayout(iftrVertexBuffer* p, const iftrExportTypeDesc* pDesc, size_t num) {
    if (!p) return;
    p->layout.clear(); p->offsets.clear();
    int off = 0;
    for (size_t i = 0; i < num; i++) {
        if (pDesc[i] < IFTP_UNDEFINED || pDesc[i] > IFTP_UINT4) continue;     // the reference drops unknown descriptors too
        const int d = normalise_desc(pDesc[i]);
        p->layout.push_back(d); p->offsets.push_back(off);
        off += desc_size(d);
    }
    p->elementSize = off;
}
void iftrSetVertexBufferSize(iftrVertexBuffer* p, size_t num) {
    if (p) p->vertexCount = (int)num;
}
void iftrAllocateVertexBuffer(iftrVertexBuffer* p) {
    if (p) p->buffer.assign((size_t)p->vertexCount * (size_t)p->elementSize, 0);
}
void iftrSetVertexBufferValueFloat4(iftrVertexBuffer* p, int index, int attr, void* value) {
    if (!p || !value || attr < 0 || attr >= (int)p->layout.size() || index < 0 || index >= p->vertexCount) return;
    const size_t off = (size_t)p->offsets[attr] + (size_t)index * (size_t)p->elementSize;
    if (off + 16 > p->buffer.size()) return;
    memcpy(p->buffer.data() + off, value, 16);
}

void* ifbufCreateBufferManager(void) { return new (std::nothrow) BufferManagerWrapper(); }
void ifbufDestroyBufferManager(void* p) { delete static_cast<BufferManagerWrapper*>(p); }
void* ifbufCreateBuffer(void* pManager, size_t bufSize) {
    if (!pManager) return nullptr;
    Buffer* b = new (std::nothrow) Buffer();
    if (!b) return nullptr;
    b->bytes.assign(bufSize, 0);
    b->manager = static_cast<BufferManagerWrapper*>(pManager)->manager;
    return b;
}
void ifbufDestroyBuffer(void* p) { delete static_cast<Buffer*>(p); }
void ifbufBufferData(void* pBuffer, const void* pData, size_t offset, size_t size) {
    Buffer* b = static_cast<Buffer*>(pBuffer);
    // a buffer whose manager is gone is dead (BufferManagerExport.cpp:62-70: the write is dropped)
    if (!b || !pData || b->manager.expired() || offset > b->bytes.size() || size > b->bytes.size() - offset) return;
    memcpy(b->bytes.data() + offset, pData, size);
}

int iftrcuBindFrameBufferHandle(iftrcuRenderer* h, iftrFrameBuffer* fb, int useDoubleBuffer) {
    if (!h || !fb || fb->color.empty() || !fb->color[0]) return IFRIT_ERR_INVALID;
    ifcrImageF32* c = fb->color[0];
    if (c->channel != 4 || c->width == 0 || c->height == 0) return IFRIT_ERR_INVALID;
    float* depth = nullptr;
    if (fb->depth) {
        if (fb->depth->width != c->width || fb->depth->height != c->height || fb->depth->channel != 1) return IFRIT_ERR_INVALID;
        depth = fb->depth->data.data();
    }
    return iftrcuBindFrameBuffer(h, c->data.data(), depth, (int)c->width, (int)c->height, useDoubleBuffer);
}
int iftrcuBindVertexBufferHandle(iftrcuRenderer* h, const iftrVertexBuffer* vb) {
    if (!h || !vb || vb->layout.empty() || vb->buffer.size() < (size_t)vb->vertexCount * (size_t)vb->elementSize) return IFRIT_ERR_INVALID;
    return iftrcuBindVertexBuffer(h, vb->buffer.data(), vb->vertexCount, vb->layout.data(), (int)vb->layout.size());
}
int iftrcuBindIndexBufferHandle(iftrcuRenderer* h, void* indexBuffer) {
    Buffer* b = static_cast<Buffer*>(indexBuffer);
    if (!h || !b || b->bytes.size() < 4) return IFRIT_ERR_INVALID;
    return iftrcuBindIndexBuffer(h, reinterpret_cast<const int*>(b->bytes.data()), (int)(b->bytes.size() / 4));
}
int iftrcuBindVertexShaderHandle(iftrcuRenderer* h, int vertexShaderId, const iftrVaryingDescriptor* d) {
    if (!h || !d) return IFRIT_ERR_INVALID;
    return iftrcuBindVertexShader(h, vertexShaderId, d->desc.data(), (int)d->desc.size());
}
int iftrcuCopyBufferHandleToBuffer(iftrcuRenderer* h, void* srcBuffer, int dstSlot) {
    Buffer* b = static_cast<Buffer*>(srcBuffer);
    if (!h || !b || b->bytes.empty()) return IFRIT_ERR_INVALID;
    return iftrcuCopyHostBufferToBuffer(h, b->bytes.data(), dstSlot, (int)b->bytes.size());
}

}   // extern "C"
