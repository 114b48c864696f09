/* ifrit_b200.h — C-ABI of the B200-native tile rasterizer (drop-in boundary).
 *
 * The reference's CUDA renderer has NO C ABI of its own (only the CPU renderer is exported:
 * include/ifrit/softgraphics/engine/export/TileRasterRendererExport.h:33-76).  The interface
 * being replaced is the C++ class TileRasterRendererCuda
 * (include/ifrit/softgraphics/engine/tilerastercuda/TileRasterRendererCuda.h:81-115, "RC" below);
 * this header exports one `iftrcu*` function per public verb of that class, following the
 * naming / ownership conventions of the reference's `iftr*` exports:
 *   - opaque instance handle, create/destroy pair (TileRasterRendererExport.h:33-36);
 *   - bindings are NON-OWNING: the caller keeps host buffers alive
 *     (TileRasterContextCuda.h:42-49); device mirrors are owned by the instance;
 *   - POD structs / enums are bit-identical to engine/base/Structures.h:52-97 and
 *     engine/base/Constants.h:23-96.
 * Differences, all deliberate: every call returns a status code (the reference
 * aborts / ignores CUDA errors), state is per instance (the reference keeps it in process-wide
 * __constant__/__device__ globals, CU:39-58, one renderer per process), and shader objects are
 * selected from a compiled-in registry of device functors because host virtual objects /
 * host function pointers (Shaders.h:46-78, :117-120) cannot run on the device.
 *
 * No torch / C++ types appear in any signature: plain pointers, ints, floats, POD structs.
 * The library has NO CPU fallback: every entry point that needs the GPU fails with
 * IFRIT_ERR_CUDA when no CUDA device is usable.
 */
#ifndef IFRIT_B200_H
#define IFRIT_B200_H
#ifndef __CUDACC_RTC__ /* NVRTC (runtime-compiled fragment shaders) has no host headers; the kernels' common.cuh supplies the integer types */
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define IFRIT_B200_API __attribute__((visibility("default")))

/* ---- status codes ------------------------------------------------------------------- */
enum {
    IFRIT_OK = 0,
    IFRIT_ERR_INVALID = 1,     /* bad argument / handle */
    IFRIT_ERR_UNSUPPORTED = 2, /* verb of the reference API that is out of scope */
    IFRIT_ERR_CUDA = 3,        /* CUDA runtime failure; see iftrcuGetLastError */
    IFRIT_ERR_STATE = 4        /* draw issued with incomplete bindings */
};

/* ---- enums: engine/base/Constants.h:23-96 ------------------------------------------- */
typedef enum { IF_POLYGON_MODE_FILL = 0, IF_POLYGON_MODE_LINE = 1, IF_POLYGON_MODE_POINT = 2 } IfritPolygonMode;
typedef enum { IF_CULL_MODE_NONE = 0, IF_CULL_MODE_FRONT = 1, IF_CULL_MODE_BACK = 2,
               IF_CULL_MODE_FRONT_AND_BACK = 3 } IfritCullMode;
typedef enum { IF_SAMPLER_ADDRESS_MODE_REPEAT = 0, IF_SAMPLER_ADDRESS_MODE_MIRRORED_REPEAT = 1,
               IF_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE = 2, IF_SAMPLER_ADDRESS_MODE_CLAMP_TO_BORDER = 3,
               IF_SAMPLER_ADDRESS_MODE_MIRROR_CLAMP_TO_EDGE = 4 } IfritSamplerAddressMode;
typedef enum { IF_FILTER_NEAREST = 0, IF_FILTER_LINEAR = 1 } IfritFilter;
typedef enum { IF_BORDER_COLOR_BLACK = 0, IF_BORDER_COLOR_WHITE = 1 } IfritBorderColor;
typedef enum { IF_IMAGE_TILING_LINEAR = 1, IF_IMAGE_TILING_OPTIMAL = 2 } IfritImageTiling;
typedef enum { IF_FORMAT_R8G8B8A8_UINT = 0, IF_FORMAT_R32G32B32A32_SFLOAT = 1 } IfritFormat;
typedef enum { IF_BLEND_FACTOR_ZERO = 0, IF_BLEND_FACTOR_ONE = 1, IF_BLEND_FACTOR_SRC_COLOR = 2,
               IF_BLEND_FACTOR_ONE_MINUS_SRC_COLOR = 3, IF_BLEND_FACTOR_DST_COLOR = 4,
               IF_BLEND_FACTOR_ONE_MINUS_DST_COLOR = 5, IF_BLEND_FACTOR_SRC_ALPHA = 6,
               IF_BLEND_FACTOR_ONE_MINUS_SRC_ALPHA = 7, IF_BLEND_FACTOR_DST_ALPHA = 8,
               IF_BLEND_FACTOR_ONE_MINUS_DST_ALPHA = 9 } IfritBlendFactor;
typedef enum { IF_COMPARE_OP_NEVER = 0, IF_COMPARE_OP_LESS = 1, IF_COMPARE_OP_EQUAL = 2,
               IF_COMPARE_OP_LESS_OR_EQUAL = 3, IF_COMPARE_OP_GREATER = 4, IF_COMPARE_OP_NOT_EQUAL = 5,
               IF_COMPARE_OP_GREATER_OR_EQUAL = 6, IF_COMPARE_OP_ALWAYS = 7 } IfritCompareOp;
typedef enum { IF_SAMPLE_COUNT_1_BIT = 1, IF_SAMPLE_COUNT_2_BIT = 2, IF_SAMPLE_COUNT_4_BIT = 4,
               IF_SAMPLE_COUNT_8_BIT = 8, IF_SAMPLE_COUNT_16_BIT = 16 } IfritSampleCountFlagBits;
/* engine/base/TypeDescriptor.h:26-41 */
typedef enum { IFTP_UNDEFINED = 0, IFTP_FLOAT1, IFTP_FLOAT2, IFTP_FLOAT3, IFTP_FLOAT4, IFTP_INT1, IFTP_INT2,
               IFTP_INT3, IFTP_INT4, IFTP_UINT1, IFTP_UINT2, IFTP_UINT3, IFTP_UINT4 } IfritTypeDescriptorEnum;

/* ---- POD structs: engine/base/Structures.h:39-97 ------------------------------------ */
typedef struct { uint32_t width, height, depth; } IfritExtent3D;
typedef struct { int32_t x, y, z; } IfritOffset3D;
typedef struct {
    IfritFilter filterMode;
    IfritSamplerAddressMode addressModeU, addressModeV, addressModeW;
    IfritBorderColor borderColor;
    uint8_t anisotropyEnable; /* C++ `bool` */
    float maxAnisotropy;
} IfritSamplerT;
typedef struct {
    IfritExtent3D extent;
    IfritImageTiling tilingMode;
    uint32_t mipLevels;   /* levels BEYOND level 0 (CU:3561-3577) */
    uint32_t arrayLayers;
    IfritFormat format;
} IfritImageCreateInfo;
typedef struct { uint32_t mipLevel, baseArrayLayer; } IfritImageSubresourceLayers;
typedef struct {
    IfritImageSubresourceLayers srcSubresource;
    IfritExtent3D srcExtentSt, srcExtentEd;
    IfritImageSubresourceLayers dstSubresource;
    IfritExtent3D dstExtentSt, dstExtentEd;
} IfritImageBlit;
typedef struct {
    uint32_t bufferOffset;
    IfritImageSubresourceLayers imageSubresource;
    IfritOffset3D imageOffset;
    IfritExtent3D imageExtent;
} IfritBufferImageCopy;
typedef struct {
    uint8_t blendEnable; /* C++ `bool` */
    IfritBlendFactor srcColorBlendFactor, dstColorBlendFactor, srcAlphaBlendFactor, dstAlphaBlendFactor;
} IfritColorAttachmentBlendState;

/* ---- shader registry (replaces the VertexShader / FragmentShader objects of Shaders.h:46-78) --
 * Uniform block = buffer slot 0 (reference: GetBufferPtr(slot), ShaderBuiltinCuda.cuh:394):
 *   float[0..15] MVP row-major, [16..18] light direction, [20] lod bias / cube lod,
 *   [21] derivative-visualisation scale, [22] depth offset scale of IFRIT_FS_DEPTH_MOD. */
enum { IFRIT_VS_PASSTHROUGH = 0, /* pos = attr0, varying k = attr k+1 */
       IFRIT_VS_MVP = 1          /* pos = MVP * attr0, varying k = attr k+1 (DemoVertexShaderCuda) */ };
enum { IFRIT_FS_VARYING0 = 0,     /* colour = varying 0 */
       IFRIT_FS_NORMAL_COLOR = 1, /* varying0 * 0.5 + 0.5, a = 0.5 (DemoFragmentShaderCuda) */
       IFRIT_FS_PRIM_ID = 2,      /* (primitive id, fan index, 0, 1) */
       IFRIT_FS_TEXTURED = 3,     /* texture(0,0,vary,1,bias) * max(N.L, 0.2): adaptive-LOD or anisotropic */
       IFRIT_FS_TEXTURED_LOD0 = 4,/* textureLod(0,0,uv,0) * max(N.L, 0.2) */
       IFRIT_FS_SKYBOX = 5,       /* textureCubeLod(0,0,varying0.xyz,lod) */
       IFRIT_FS_DERIV_VIS = 6,    /* (ddx(v1).xy, ddy(v1).xy) * scale */
       IFRIT_FS_DEPTH_MOD = 7,    /* colour = varying 0, fragment depth += u[22] * varying0.x: a shader with
                                   * allowDepthModification (Shaders.h:61) -> ordered path, depth test before AND after
                                   * the shader (late Z, CU:1803-1805, CU:2517-2527); 1x only, as the reference */
       IFRIT_FS_USER = 1000       /* first id handed out by iftrcuCreateFragmentShaderFromSource */ };

/* coverage-sample convention: which of the reference's two renderers is reproduced */
enum { IFRIT_CONVENTION_CUDA = 0, /* TileRasterRendererCuda: pixel centre / MSAA table, clip at w=1e-3 */
       IFRIT_CONVENTION_CPU = 1   /* TileRasterRenderer (AVX2): pixel corner, clip at w=1, z reject */ };

typedef struct iftrcuRenderer iftrcuRenderer; /* opaque; = std::shared_ptr<TileRasterRendererCuda> in RC */

typedef struct {
    uint64_t triangles_submitted; /* input triangles of the last draw */
    uint64_t triangles_assembled; /* after clip / cull */
    uint64_t tile_pairs;          /* (triangle, tile) queue entries */
    uint64_t fragments_shaded;    /* FS invocations that wrote >= 1 sample */
    uint64_t kernel_launches;     /* kernels launched by this instance since creation */
    uint32_t queue_overflow;      /* 1 if the last draw had to be re-issued with a larger queue */
    uint32_t reserved;
} iftrcuStats;

/* one assembled triangle, for stage-level parity tests (debug read-back) */
typedef struct {
    int32_t prim, fan, clipped, slot;
    int32_t X[3], Y[3];
    float Z[3], IW[3];
    float bary[3][2];
} iftrcuAssembledTriangle;

#ifndef __CUDACC_RTC__ /* entry points: host side only (the header is also compiled by NVRTC for its enums and PODs) */
/* ---- lifetime (iftrCreateInstance / iftrDestroyInstance / iftrInit) ------------------ */
IFRIT_B200_API int iftrcuCreateInstance(int cudaDevice, iftrcuRenderer** out);
IFRIT_B200_API int iftrcuDestroyInstance(iftrcuRenderer* h);
IFRIT_B200_API int iftrcuInit(iftrcuRenderer* h);                                  /* RC::Init */
IFRIT_B200_API const char* iftrcuGetLastError(iftrcuRenderer* h);
/* run on a caller-provided cudaStream_t (e.g. torch's current stream).  NULL is the legacy default
 * stream; (void*)-1 returns to the instance's own non-blocking stream (the initial state). */
IFRIT_B200_API int iftrcuSetStream(iftrcuRenderer* h, void* cudaStream);

/* ---- bindings ------------------------------------------------------------------------- */
/* RC::bindFrameBuffer(FrameBuffer&, bool): colour = ImageF32 RGBA32F [(y*W+x)*4+c]
 * (core/data/Image.h:224), depth = ImageF32 F32; host pointers, non-owning; either may be NULL
 * (device-only rendering).  The colour image is copied back by iftrcuDrawElements (CU:3899-3909). */
IFRIT_B200_API int iftrcuBindFrameBuffer(iftrcuRenderer* h, float* hostColorRGBA32F, float* hostDepthF32,
                                         int width, int height, int useDoubleBuffer);
/* zero-copy variant: render straight into caller-owned DEVICE memory (no host copy-back) */
IFRIT_B200_API int iftrcuBindFrameBufferDevice(iftrcuRenderer* h, float* devColorRGBA32F, float* devDepthF32,
                                               int width, int height);
/* RC::bindVertexBuffer(const VertexBuffer&): interleaved attributes (VertexBuffer.h:28-90);
 * only IFTP_FLOAT4 attributes are supported (what every reference demo uses). */
IFRIT_B200_API int iftrcuBindVertexBuffer(iftrcuRenderer* h, const void* hostData, int vertexCount,
                                          const int* typeDescs, int numAttributes);
IFRIT_B200_API int iftrcuBindVertexBufferDevice(iftrcuRenderer* h, const void* devData, int vertexCount,
                                                const int* typeDescs, int numAttributes);
/* RC::bindIndexBuffer(const std::vector<int>&) */
IFRIT_B200_API int iftrcuBindIndexBuffer(iftrcuRenderer* h, const int* hostIndices, int indexCount);
IFRIT_B200_API int iftrcuBindIndexBufferDevice(iftrcuRenderer* h, const int* devIndices, int indexCount);
/* RC::bindVertexShader(VertexShader*, VaryingDescriptor&) / RC::bindFragmentShader(FragmentShader*) */
IFRIT_B200_API int iftrcuBindVertexShader(iftrcuRenderer* h, int vertexShaderId, const int* varyingTypeDescs,
                                          int numVaryings);
IFRIT_B200_API int iftrcuBindFragmentShader(iftrcuRenderer* h, int fragmentShaderId);
/* Fragment shader compiled at run time from CUDA source (the reference's counterpart is its shader JIT,
 * engine/base/ShaderRuntime.h:24-37 + ComLLVMRuntime).  `cudaSource` defines
 *     __device__ float4 fs_main(const IfritFsInput& in);
 * with  in.varying[k] / in.ddx[k] / in.ddy[k] (float4, k < numVaryings; derivatives only if needsDerivatives),
 * in.primitiveId, in.fanIndex, in.uniform[0..31] (buffer slot 0), and the sampling library
 *     ifritTexture(in, texSlot, sampSlot, float2 uv, float2 duvdx, float2 duvdy, bias)     texture()
 *     ifritTextureLod(in, texSlot, sampSlot, float2 uv, lod)                                 textureLod()
 *     ifritTextureCubeLod(in, texSlot, sampSlot, float3 dir, lod)                            textureCubeLod()
 * NVRTC (libnvrtc.so.12, loaded on first use) compiles the tile kernel around the function for sm_100a; the returned id
 * (>= IFRIT_FS_USER) goes to iftrcuBindFragmentShader.  A source error returns IFRIT_ERR_INVALID with the compiler log in
 * iftrcuGetLastError.  Not available on the ordered (blending) path. */
IFRIT_B200_API int iftrcuCreateFragmentShaderFromSource(iftrcuRenderer* h, const char* name, const char* cudaSource,
                                                        int numVaryings, int needsDerivatives, int* shaderIdOut);
/* compile only (no device, no instance): IFRIT_OK, IFRIT_ERR_INVALID + the compiler log, or IFRIT_ERR_UNSUPPORTED without libnvrtc */
IFRIT_B200_API int iftrcuCheckFragmentShaderSource(const char* cudaSource, int numVaryings, int needsDerivatives,
                                                   int msaaSamples, char* logOut, size_t logCap);
IFRIT_B200_API int iftrcuLookupVertexShader(const char* name);   /* -1 if unknown */
IFRIT_B200_API int iftrcuLookupFragmentShader(const char* name); /* -1 if unknown */
/* out of scope: always IFRIT_ERR_UNSUPPORTED */
IFRIT_B200_API int iftrcuBindGeometryShader(iftrcuRenderer* h, int shaderId);
/* Mesh / task front end (RC::bindMeshShader / bindTaskShader / DrawMeshTasks, RC:88-89, 115; meshShadingEntryPoint CU:375-500).
 * The shaders are CUDA source compiled at run time (NVRTC), with the argument lists of MeshShader::execute / TaskShader::execute:
 *     __device__ void ms_main(const IfritMeshInput& in, float4* outPos, float4* outVaryings, int* outIndices,
 *                             int& outNumVertices, int& outNumIndices);          // <= 768 vertices, <= 768 indices per work group
 *     __device__ void ts_main(int workGroupId, void* outPayload, int3* outMeshWorkGroups, int& outNumMeshWorkGroups);   // <= 32, 128-byte payload
 * in.localInvocation / in.workGroupId / in.payload / in.uniform[0..31]; ifritBuffer(in, slot) = buffer slot `slot` (float*);
 * ifritMvp(in, pos) = the registry's `mvp` vertex shader.  outPos is clip space; outVaryings[vertex * numVaryings + k].
 * iftrcuDrawMeshTasks runs [task ->] mesh work groups, compacts their output IN WORK-GROUP ORDER and draws it like drawElements. */
IFRIT_B200_API int iftrcuCreateMeshShaderFromSource(iftrcuRenderer* h, const char* name, const char* cudaSource, int* shaderIdOut);
IFRIT_B200_API int iftrcuCreateTaskShaderFromSource(iftrcuRenderer* h, const char* name, const char* cudaSource, int* shaderIdOut);
IFRIT_B200_API int iftrcuCheckMeshShaderSource(const char* cudaSource, int isTaskShader, char* logOut, size_t logCap); /* compile only, no device */
IFRIT_B200_API int iftrcuBindMeshShader(iftrcuRenderer* h, int shaderId, const int* varyingTypeDescs, int n,
                                        int lx, int ly, int lz);
IFRIT_B200_API int iftrcuBindTaskShader(iftrcuRenderer* h, int shaderId /* -1 unbinds */, const int* varyingTypeDescs, int n);

/* ---- resources ------------------------------------------------------------------------ */
IFRIT_B200_API int iftrcuCreateTexture(iftrcuRenderer* h, int slot, const IfritImageCreateInfo* ci); /* RC::CreateTexture */
IFRIT_B200_API int iftrcuCreateSampler(iftrcuRenderer* h, int slot, const IfritSamplerT* s);         /* RC::CreateSampler */
IFRIT_B200_API int iftrcuGenerateMipmap(iftrcuRenderer* h, int slot, IfritFilter filter);            /* RC::generateMipmap */
IFRIT_B200_API int iftrcuBlitImage(iftrcuRenderer* h, int srcSlot, int dstSlot, const IfritImageBlit* region,
                                   IfritFilter filter);                                               /* RC::blitImage */
IFRIT_B200_API int iftrcuCopyHostBufferToImage(iftrcuRenderer* h, const void* src, int dstSlot,
                                               const IfritBufferImageCopy* regions, int regionCount); /* RC::copyHostBufferToImage */
IFRIT_B200_API int iftrcuCopyImageToHostBuffer(iftrcuRenderer* h, int srcSlot, void* dst, size_t nbytes); /* test read-back of the whole chain */
IFRIT_B200_API int iftrcuCreateBuffer(iftrcuRenderer* h, int slot, int bufSize);                     /* RC::CreateBuffer */
IFRIT_B200_API int iftrcuCopyHostBufferToBuffer(iftrcuRenderer* h, const void* src, int dstSlot, int size); /* RC::copyHostBufferToBuffer */

/* ---- pipeline state ------------------------------------------------------------------- */
IFRIT_B200_API int iftrcuSetScissors(iftrcuRenderer* h, const float* xywh, int count); /* RC::SetScissors (Vector4f each) */
IFRIT_B200_API int iftrcuSetScissorTestEnable(iftrcuRenderer* h, int enable);
IFRIT_B200_API int iftrcuSetMsaaSamples(iftrcuRenderer* h, IfritSampleCountFlagBits samples);
IFRIT_B200_API int iftrcuSetRasterizerPolygonMode(iftrcuRenderer* h, IfritPolygonMode mode); /* FILL only */
/* Ordered alpha blending (pixelShadingAlphaBlendKernel CU:1649-1897): fragments of a pixel are depth-tested, shaded and
 * blended in submission order, depth is written by every passing fragment.  Factors: ZERO, ONE, SRC_ALPHA, DST_ALPHA,
 * ONE_MINUS_SRC_ALPHA, ONE_MINUS_DST_ALPHA with the reference's coefficient table (CU:3436-3537: its *_ALPHA rows
 * select 1 - a and its ONE_MINUS_* rows a — replicated); 1x MSAA only.  Tiles holding more than 4096 triangles are walked in rounds of submission-order ranges. */
IFRIT_B200_API int iftrcuSetBlendFunc(iftrcuRenderer* h, const IfritColorAttachmentBlendState* s);
IFRIT_B200_API int iftrcuSetDepthFunc(iftrcuRenderer* h, IfritCompareOp op);
IFRIT_B200_API int iftrcuSetDepthTestEnable(iftrcuRenderer* h, int enable);
IFRIT_B200_API int iftrcuSetCullMode(iftrcuRenderer* h, IfritCullMode mode);
IFRIT_B200_API int iftrcuSetClearValues(iftrcuRenderer* h, const float* clearColorsRGBA, int numColors,
                                        float clearDepth);
/* extensions with no reference counterpart */
IFRIT_B200_API int iftrcuSetSampleConvention(iftrcuRenderer* h, int convention);
/* sort-first partition: this instance renders only the 64-px super-tiles it owns */
IFRIT_B200_API int iftrcuSetTilePartition(iftrcuRenderer* h, int rank, int worldSize);
IFRIT_B200_API int iftrcuSetDebugPrimitiveIds(iftrcuRenderer* h, int enable);
/* vertex stage fused into triangle setup (default, 1) or run as the stand-alone kernel that mirrors
 * vertexProcessingKernel CU:330 (0); both produce identical frames */
IFRIT_B200_API int iftrcuSetVertexStageFused(iftrcuRenderer* h, int enable);
/* Sampler precision contract.  DEFAULT (0): four loads + fp32 lerps in software, the reference's arithmetic op for op
 * (textureImplLegacy / Lod / AnisotropicFilter, ShaderBuiltinCuda.cuh:106-272) — bit-identical to the CPU oracle up to
 * the last ulp of log2f, for texels of any range.  OPT-IN (1), "FAST": LINEAR 2x2 footprints are fetched and blended by
 * the texture unit (same texels, weights quantised to 8 bits) and sqrt / divide / log2 of the LOD and lighting math
 * use the SFU approximations.  Error <= ~2^-9 x (texel range): within 1/255 per channel for textures in [0, 1], NOT
 * for HDR-range texels (tests/test_gpu_parity.py::test_fast_sampler_precision_contract). */
IFRIT_B200_API int iftrcuSetSamplerHardwareFilter(iftrcuRenderer* h, int enable);
/* Framebuffer traffic of the tile kernel through TMA (cp.async.bulk.tensor: block stores, tile-sized clears, block loads
 * for draws that keep the previous contents).  DEFAULT on; results are bit-identical either way (A/B switch; the images
 * of the reference are written by per-thread stores, CU:2235-2241, CU:3363-3391). */
IFRIT_B200_API int iftrcuSetTmaFramebuffer(iftrcuRenderer* h, int enable);
/* Tile kernel grid = the exact CTA count measured by the draw that sized the queues for the bound inputs, instead of one
 * CTA per tile with the surplus exiting.  DEFAULT off: measured no gain (c3 tile 0.3323 vs 0.3327 ms, profiles/r02_ab.txt).  Inputs changed behind the library's back are caught
 * like a queue overflow: the draw is re-issued at the next synchronisation. */
IFRIT_B200_API int iftrcuSetTileGridHint(iftrcuRenderer* h, int enable);
/* Big-queue entries the tile kernel stages per bulk copy: 0 = automatic (32 for multisampled draws whose tiles hold few large
 * triangles — less shared memory, more L1 — else 64), or pinned to 32 / 64 (A/B switch). */
IFRIT_B200_API int iftrcuSetTileChunk(iftrcuRenderer* h, int entries);

/* ---- draw ----------------------------------------------------------------------------- */
IFRIT_B200_API int iftrcuClear(iftrcuRenderer* h);        /* RC::clear — fused into the next draw */
IFRIT_B200_API int iftrcuDrawElements(iftrcuRenderer* h); /* RC::drawElements — asynchronous */
IFRIT_B200_API int iftrcuDrawMeshTasks(iftrcuRenderer* h, int numWorkGroups, int firstWorkGroup); /* RC::DrawMeshTasks — synchronous */
/* wait for everything issued so far (incl. the host copy-back of the colour image) */
IFRIT_B200_API int iftrcuSynchronize(iftrcuRenderer* h);

/* ---- read-back / introspection --------------------------------------------------------- */
IFRIT_B200_API int iftrcuReadColor(iftrcuRenderer* h, float* hostDst);  /* W*H*4 floats, synchronous */
IFRIT_B200_API int iftrcuReadDepth(iftrcuRenderer* h, float* hostDst);  /* W*H floats */
IFRIT_B200_API int iftrcuReadPrimitiveIds(iftrcuRenderer* h, int32_t* hostDst); /* W*H*S, needs SetDebugPrimitiveIds */
IFRIT_B200_API int iftrcuGetDeviceColorPtr(iftrcuRenderer* h, void** out);
IFRIT_B200_API int iftrcuGetDeviceDepthPtr(iftrcuRenderer* h, void** out);
IFRIT_B200_API int iftrcuGetStats(iftrcuRenderer* h, iftrcuStats* out);
/* per-kernel device time: CUDA events recorded on the launching stream around each kernel of the draw
 * chain (vertex, setup+count, scan, fill, tile); averages over the draws issued since SetProfiling(1)
 * (ring of 256).  Reference analogue: the CU_PROFILER_* switches, KC:128-135. */
IFRIT_B200_API int iftrcuSetProfiling(iftrcuRenderer* h, int enable);
IFRIT_B200_API int iftrcuGetKernelTimes(iftrcuRenderer* h, float* msOut5, int* numDraws);
/* assembled triangles of the last draw, sorted by slot; returns count via *n (cap = capacity of dst) */
IFRIT_B200_API int iftrcuReadAssembled(iftrcuRenderer* h, iftrcuAssembledTriangle* dst, int cap, int* n);
/* pack / unpack the super-tiles owned by `rank` (sort-first gather plumbing, device pointers) */
/* Peer-memory sort-first (NVLink / NVSwitch, one process per GPU): the rank that owns the final frame exports
 * its device framebuffer (colour + depth, 2 x 64-byte CUDA IPC handles); every other rank binds it as its
 * render target, so that its tile kernel stores the tiles it owns straight into the owner's memory over
 * NVLink — no pack, no collective on the data path, no unpack.  The caller orders frame completion with a
 * barrier on the stream (e.g. a one-element NCCL all-reduce). */
#define IFRIT_IPC_HANDLE_BYTES 128
/* asynchronous colour read-back on the renderer's stream (the reference's double-buffered copy, CU:3899-3909);
 * hostDst should be pinned; ordered after every draw issued so far */
IFRIT_B200_API int iftrcuCopyColorToHostAsync(iftrcuRenderer* h, void* hostDst);
/* Packed presentation image (the step after the path — the reference copies RGBA32F over PCIe
 * and is bound by it, CU:3893-3914, docs/performance.md).  hostRGBA8: width*height*4 bytes, row-major, r in the low
 * byte, or NULL for a device-only image; format 1 = RGBA8 UNORM, 0 = off.  While bound, every draw / clear also
 * writes round(saturate(c) * 255) per channel next to the float colour (same kernel, no re-read) and the per-draw
 * copy-back carries this image INSTEAD of the RGBA32F one; useDoubleBuffer applies to it as to the float image. */
/* Frame pipelining (default on): the geometry stage of a draw (setup + count, scan, fill) runs on an internal stream
 * into one of two scratch sets, only the tile kernel runs on the caller's stream, so the geometry of frame k + 1
 * overlaps the tile kernel of frame k.  Results and their order on the caller's stream are unchanged.  The geometry
 * stage is ordered after the last bind / upload of its inputs (vertex, index, uniform buffers): a caller that
 * rewrites a DEVICE-bound buffer in place must re-bind it (or switch pipelining off) before the next draw. */
IFRIT_B200_API int iftrcuSetFramePipelining(iftrcuRenderer* h, int enable);
IFRIT_B200_API int iftrcuBindPresentBuffer(iftrcuRenderer* h, void* hostRGBA8, int format);
IFRIT_B200_API int iftrcuReadPresent(iftrcuRenderer* h, void* dstRGBA8);
IFRIT_B200_API int iftrcuExportFrameBuffer(iftrcuRenderer* h, void* handleOut128);
/* peer: the owner's COLOUR buffer becomes this instance's render target (its tiles are stored over NVLink); depth stays
 * in a local buffer — after a sort-first frame the owner holds the complete colour image and the depth of its own tiles */
IFRIT_B200_API int iftrcuBindPeerFrameBuffer(iftrcuRenderer* h, const void* handle128, int width, int height);
/* Sort-first frame flags (peer-memory mode; no reference counterpart — the reference is single-GPU).
 * The framebuffer owner's allocation carries a block of 32-bit flag words that every rank reaches through the same
 * CUDA IPC mapping as the framebuffer.  Per frame, in stream order and without NCCL or host round trips:
 *   every rank : iftrcuDrawElements; iftrcuFrameSignal(id)      "my tiles of frame id are stored"
 *   owner      : iftrcuFrameWait(id)                            polls until all `worldSize` ranks have signalled
 *                ... consume the frame ...; iftrcuFrameRelease(id)  "frame id may be overwritten"
 *   peers      : iftrcuFrameFence(id)                           the NEXT tile kernel first waits for the release of id
 * A wait gives up after 2 s (a later iftrcuSynchronize then returns IFRIT_ERR_STATE); the device never hangs. */
/* Distributed triangle setup (2..8 ranks): each rank classifies 1 / worldSize of the draw's triangles — which ranks own a tile of
 * the triangle's box, one mask byte per triangle into devMask[first ...] (caller-owned device array of one byte per triangle of
 * the whole draw) — the ranks all-gather their slices (NCCL), and iftrcuSetTriangleOwnerMask gives the complete array to the
 * following draws, whose setup pass then assembles only the triangles that name this rank.  Frames are unchanged. */
IFRIT_B200_API int iftrcuClassifyTriangles(iftrcuRenderer* h, int firstTriangle, int count, void* devMask);
IFRIT_B200_API int iftrcuSetTriangleOwnerMask(iftrcuRenderer* h, const void* devMask /* NULL = off */);
/* the stream the geometry stage (and iftrcuClassifyTriangles) runs on: the caller's all-gather of the masks must be ordered on it */
IFRIT_B200_API int iftrcuGetGeometryStream(iftrcuRenderer* h, void** cudaStreamOut);
/* Peer mode with two frames in flight: the owner's colour allocation holds `count` (1 or 2) frames back to back behind the same CUDA
 * IPC handle, with one flag block each; iftrcuSelectFrame picks the frame that the following clears / draws / read-backs and the
 * iftrcuFrame* verbs refer to.  Owner: iftrcuBindFrameBuffer -> iftrcuSetFrameRing -> iftrcuExportFrameBuffer; peers:
 * iftrcuBindPeerFrameBuffer -> iftrcuSetFrameRing.  With frame k in buffer (k - 1) & 1, a peer fences frame k + 1 on the release of frame
 * k - 1, so that the owner's wait / pre-clear / release of frame k overlaps the peers' tile kernels of frame k + 1. */
IFRIT_B200_API int iftrcuSetFrameRing(iftrcuRenderer* h, int count);
IFRIT_B200_API int iftrcuSelectFrame(iftrcuRenderer* h, int index);
IFRIT_B200_API int iftrcuFrameSignal(iftrcuRenderer* h, uint32_t frameId);
IFRIT_B200_API int iftrcuFrameWait(iftrcuRenderer* h, uint32_t frameId);
IFRIT_B200_API int iftrcuFrameRelease(iftrcuRenderer* h, uint32_t frameId);
IFRIT_B200_API int iftrcuFrameFence(iftrcuRenderer* h, uint32_t frameId);
/* Owner: FrameWait + (cleared ? FrameReleaseCleared : FrameRelease) on the renderer's exchange stream, behind what the draw stream holds now,
 * for frames nothing on the draw stream consumes; the next draw / clear into the same buffer of the ring is ordered behind it. */
IFRIT_B200_API int iftrcuFrameRetire(iftrcuRenderer* h, uint32_t frameId, int cleared);

/* The same pair with a promise: the NEXT draw into this framebuffer begins with iftrcuClear (same clear values on every
 * rank).  The owner clears the other ranks' tiles locally before raising the release; a peer that fences with the Cleared
 * variant then skips the NVLink stores of the colour of its empty tiles in that draw.  Frames stay bit-identical. */
IFRIT_B200_API int iftrcuFrameReleaseCleared(iftrcuRenderer* h, uint32_t frameId);
IFRIT_B200_API int iftrcuFrameFenceCleared(iftrcuRenderer* h, uint32_t frameId);
/* Sort-first distributed read-back (no reference counterpart — the reference copies the frame back over one PCIe link, CU:3899-3909):
 * stores the tiles this rank owns from its bound device framebuffer into the full-frame RGBA32F host image (row-major, W*H*16 bytes,
 * shared by the ranks), in stream order, through mapped pinned memory.  Not page-locked yet -> registered here. */
IFRIT_B200_API int iftrcuCopyOwnedTilesToHost(iftrcuRenderer* h, void* hostImage);
IFRIT_B200_API int iftrcuPackOwnedTiles(iftrcuRenderer* h, void* devDst, size_t dstBytes, size_t* usedBytes);
IFRIT_B200_API int iftrcuUnpackOwnedTiles(iftrcuRenderer* h, const void* devSrc, int rank, int worldSize);
IFRIT_B200_API size_t iftrcuOwnedTileBytes(int width, int height, int rank, int worldSize);

#endif /* !__CUDACC_RTC__ */
#ifdef __cplusplus
}
#endif
#endif /* IFRIT_B200_H */
