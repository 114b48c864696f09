/* ifrit_b200_compat.h — the reference's shared-object C ABI for its data objects, exported by libifrit_b200.so
 * under the SAME names and signatures, plus the iftrcu* binders that accept those handles.
 *
 * A C / C# caller of the reference (bindings/csharp: NativeMethods) creates images, frame buffers, vertex buffers,
 * varying descriptors and buffer-manager buffers through these entry points and passes the handles to the renderer
 * (iftrBindFrameBuffer, iftrBindVertexBuffer, iftrBindIndexBuffer, TileRasterRendererExport.h:33-76).  With this
 * header the same caller keeps every one of those calls and only swaps iftrCreateInstance / iftrBind* for
 * iftrcuCreateInstance / iftrcuBind*Handle.  The handles are opaque: the objects behind them are this library's own
 * (the reference's classes are C++ with private contexts, nothing of their layout crosses a C ABI).
 *
 * Replaces, one for one:
 *   include/ifrit/softgraphics/core/export/CommonExport.h:26-31            ifcrCreateImageFP32 / Destroy / GetRawData
 *   include/ifrit/softgraphics/engine/export/EngineCommonExport.h:33-70     iftr{Create,Destroy}FrameBuffer, iftrSetFrameBuffer*AttachmentFP32,
 *                                                                           iftr{Create,Destroy,Write}VaryingDescriptor,
 *                                                                           iftr{Create,Destroy}VertexBuffer, iftrSetVertexBufferLayout / Size,
 *                                                                           iftrAllocateVertexBuffer, iftrSetVertexBufferValueFloat4
 *   include/ifrit/softgraphics/engine/export/BufferManagerExport.h:25-35    ifbufCreateBufferManager / Destroy, ifbufCreateBuffer / Destroy, ifbufBufferData
 */
#ifndef IFRIT_B200_COMPAT_H
#define IFRIT_B200_COMPAT_H
#include "ifrit_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* opaque handles (reference: Core::Data::ImageF32*, FrameBuffer*, VaryingDescriptor*, VertexBuffer*, void*) */
typedef struct ifcrImageF32 ifcrImageF32;
typedef struct iftrFrameBuffer iftrFrameBuffer;
typedef struct iftrVaryingDescriptor iftrVaryingDescriptor;
typedef struct iftrVertexBuffer iftrVertexBuffer;
typedef int iftrExportTypeDesc; /* LibraryExport::ExportTypeDesc = int, values of IfritTypeDescriptorEnum (IFTP_*) */

/* ---- CommonExport.h:26-31 ---- */
IFRIT_B200_API ifcrImageF32* ifcrCreateImageFP32(size_t width, size_t height, size_t channel);
IFRIT_B200_API void ifcrDestroyImageFP32(ifcrImageF32* pInstance);
IFRIT_B200_API float* ifcrGetImageRawDataFP32(ifcrImageF32* pInstance); /* [(y*W + x)*C + c], Image.h:224 */

/* ---- EngineCommonExport.h:33-43 (FrameBuffer.h) ---- */
IFRIT_B200_API iftrFrameBuffer* iftrCreateFrameBuffer(void);
IFRIT_B200_API void iftrDestroyFrameBuffer(iftrFrameBuffer* pInstance);
IFRIT_B200_API void iftrSetFrameBufferColorAttachmentFP32(iftrFrameBuffer* pInstance, ifcrImageF32* const* pImage, size_t nums);
IFRIT_B200_API void iftrSetFrameBufferDepthAttachmentFP32(iftrFrameBuffer* pInstance, ifcrImageF32* pImage);

/* ---- EngineCommonExport.h:45-54 (VaryingDescriptor.h) ---- */
IFRIT_B200_API iftrVaryingDescriptor* iftrCreateVaryingDescriptor(void);
IFRIT_B200_API void iftrDestroyVaryingDescriptor(iftrVaryingDescriptor* pInstance);
IFRIT_B200_API void iftrWriteVaryingDescriptor(iftrVaryingDescriptor* pInstance, const iftrExportTypeDesc* pDesc, size_t num);

/* ---- EngineCommonExport.h:56-70 (VertexBuffer.h) ---- */
IFRIT_B200_API iftrVertexBuffer* iftrCreateVertexBuffer(void);
IFRIT_B200_API void iftrDestroyVertexBuffer(iftrVertexBuffer* pInstance);
IFRIT_B200_API void iftrSetVertexBufferLayout(iftrVertexBuffer* pInstance, const iftrExportTypeDesc* pDesc, size_t num);
IFRIT_B200_API void iftrSetVertexBufferSize(iftrVertexBuffer* pInstance, size_t num);
IFRIT_B200_API void iftrAllocateVertexBuffer(iftrVertexBuffer* pInstance);
IFRIT_B200_API void iftrSetVertexBufferValueFloat4(iftrVertexBuffer* pInstance, int index, int attr, void* value);

/* ---- BufferManagerExport.h:25-35 ---- */
IFRIT_B200_API void* ifbufCreateBufferManager(void);
IFRIT_B200_API void ifbufDestroyBufferManager(void* p);
IFRIT_B200_API void* ifbufCreateBuffer(void* pManager, size_t bufSize);
IFRIT_B200_API void ifbufDestroyBuffer(void* p);
IFRIT_B200_API void ifbufBufferData(void* pBuffer, const void* pData, size_t offset, size_t size);

/* ---- the renderer verbs that take those handles (TileRasterRendererExport.h:38-49: iftrBindFrameBuffer,
 *      iftrBindVertexBuffer, iftrBindIndexBuffer, iftrBindVertexShaderFunc's VaryingDescriptor) ---- */
/* colour attachment 0 must be W x H x 4, depth W x H x 1 (TileRasterRendererCuda.cpp bindFrameBuffer) */
IFRIT_B200_API int iftrcuBindFrameBufferHandle(iftrcuRenderer* h, iftrFrameBuffer* frameBuffer, int useDoubleBuffer);
IFRIT_B200_API int iftrcuBindVertexBufferHandle(iftrcuRenderer* h, const iftrVertexBuffer* vertexBuffer);
/* indexBuffer: an ifbufCreateBuffer handle holding int32 indices (iftrBindIndexBuffer takes the same void*) */
IFRIT_B200_API int iftrcuBindIndexBufferHandle(iftrcuRenderer* h, void* indexBuffer);
IFRIT_B200_API int iftrcuBindVertexShaderHandle(iftrcuRenderer* h, int vertexShaderId, const iftrVaryingDescriptor* vsOutDescriptors);
/* RC::copyHostBufferToBuffer from a buffer-manager buffer */
IFRIT_B200_API int iftrcuCopyBufferHandleToBuffer(iftrcuRenderer* h, void* srcBuffer, int dstSlot);

#ifdef __cplusplus
}
#endif
#endif
