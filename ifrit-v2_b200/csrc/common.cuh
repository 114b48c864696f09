// common.cuh — internal types shared by the kernels and the C-ABI host code.
// B200 (sm_100a) only.  See DESIGN.md §3 for the arithmetic specification these kernels
// implement bit-for-bit (integer coverage / ids / depth) against oracle/ifrit_oracle.c.
#pragma once
#ifdef __CUDACC_RTC__
// Runtime compilation of user fragment shaders (capi.cu: iftrcuCreateFragmentShaderFromSource): NVRTC has the CUDA device
// API built in and no host headers; the few host-side names the shared structs mention get stand-ins.
typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;
typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
struct alignas(64) CUtensorMap { unsigned long long opaque[16]; };
#else
#include <cuda_runtime.h>
#include <cuda.h>      // CUtensorMap (type only: nothing links libcuda, the encoder is fetched with cudaGetDriverEntryPoint)
#include <stdint.h>
#endif

#define IFR_TILE 16          // tile edge in pixels (reference: CU_TILE_WIDTH, KC:36-42)
#define IFR_TILE_PIX 256
#define IFR_SUPER 64         // sort-first ownership granularity in pixels
#define IFR_SUBPIX 256       // 24.8 fixed point
#define IFR_MAX_TEX 8
#define IFR_MAX_BUF 8
#define IFR_MAX_VARY 4
#define IFR_MAX_SCISSOR 16
#define IFR_SLOT_NONE 0xffffffffu

// One assembled triangle = 64 bytes: three 16-byte snapped vertices + one word of triangle-level
// setup (geometryParamPostprocKernel CU:1549 keeps ~112 bytes per triangle).
struct TriVert { int32_t x, y; float z, iw; };
struct __align__(16) TriRec {
    TriVert v[3];
    float ar;            // 1 / float(d), d = twice the signed area in sub-pixel units (post-swap winding)
    uint32_t flags;
    uint32_t pad[2];
};
#define IFR_FLAG_CLIPPED 1   // clip barycentrics in clipBary[slot*6]
#define IFR_FLAG_SWAPPED 2   // cull-mode winding swap (v1 <-> v3)
#define IFR_FLAG_FITS32 4    // edge functions at every pixel corner that gets evaluated fit int32
#define IFR_FLAG_SMALL 8     // pixel bbox area <= IFR_SMALL_AREA: thread-per-triangle path of the tile kernel
#define IFR_FLAG_WIDE 16     // an edge delta exceeds 2^20 sub-pixels (4096 px): int64 coverage path
#define IFR_SMALL_AREA 16
// FAST textured shaders: per-triangle screen-space planes of the perspective numerators instead of per-fragment barycentric
// interpolation from the vertex data (-DIFR_PLANES=0: A/B switch, the FAST shaders interpolate like the exact ones)
#ifndef IFR_PLANES
#define IFR_PLANES 1
#endif
#define IFR_PLANE_QUADS 5

// Per (triangle, tile) raster setup, written by the fill pass into the tile's big-triangle queue and
// read with uniform 128-bit loads by every pixel thread of the tile: six 16-byte words.
struct __align__(16) TriS {
    int dX[3]; uint32_t box;     // box = x0 | x1<<8 | y0<<16 | y1<<24: tile-local pixel bbox that can hold covered samples
    int dY[3]; uint32_t slot;
    int T[3]; uint32_t flags;    // int32 thresholds: inside <=> dX*j - dY*i >= T (i, j in 1/16 px, tile local)
    long long E0[3];             // edge functions at the tile origin (sub-pixel units)
    float ar, z0, z1, z2;
    int pad[2];                  // pad[0]: coarse raster, bit per 8x4-pixel block of the tile: 0-7 may hold a covered
                                 // sample, 8-15 covered at every sample; pad[1]: Hi-Z bound (float bits), the most
                                 // optimistic depth of the triangle inside its tile-local box (bits 8-15 and
                                 // pad[1] are filled in by the tile kernel for deep queues, tri_setup.cuh)
};

#define IFR_MAX_MIPS 16
// one mip level: dimensions, texel offset of layer 0, texels per layer, reciprocals for address wrapping
struct TexLevel { int w, h; uint32_t off, layerStride; float invW, invH, inv2W, inv2H; float fw, fh, fw1, fh1; };   // fw1 = float(w - 1)
// hw[l]: CUDA texture object over layer 0 of level l (pitch-linear view of the same memory, clamp addressing,
// unnormalised coordinates, bilinear filter) or 0 when the level cannot be bound (alignment, 1-texel levels)
struct TexDesc { const float4* data; int w, h, mips, layers; TexLevel lvl[IFR_MAX_MIPS]; unsigned long long hw[IFR_MAX_MIPS]; };
struct SampDesc { int filter, addrU, addrV, addrW, border, aniso; float maxAniso; };

struct DrawParams {
    // framebuffer
    float4* color; float* depth; int32_t* primIds;
    uint32_t* present;   // optional packed RGBA8 presentation image (iftrcuBindPresentBuffer), written with the colour
    int W, H, tilesX, tilesY;
    int msaa, depthFunc, cullMode, convention;
    // visibility-key mode of depthFunc, precomputed by the host (kernels_tile.cu: KeyMode) so that the hot loops read
    // it from the constant bank instead of re-deriving it from the compare op
    uint32_t keyXhi, keyXlo; int keyOrdered, keyEqual, keyNever;
    int doClear; float clearColor[4]; float clearDepth;
    int scissorEnable, nScissor; float scissor[IFR_MAX_SCISSOR][4];
    int rank, world;
    // sort-first exchange: 1 = a peer whose owner has pre-cleared the frame (iftrcuFrameReleaseCleared) does not send the
    // colour of its EMPTY tiles over NVLink in a clearing draw; 2 (clear_kernel only) = clear the tiles of the OTHER ranks
    int skipEmptyStores;
    // distributed triangle setup (kernels_geometry.cu: classify_kernel): one owner-mask byte per input triangle, bit r = rank r owns
    // a tile of its box; null = every rank assembles every triangle
    const uint8_t* triMask;
    // ordered alpha blending (pixelShadingAlphaBlendKernel CU:1649-1897): colour src / colour dst / alpha src / alpha dst
    // coefficients {x1, x(1 - src.a), x(1 - dst.a), zero flag} (updateBlendStateKernel CU:3436-3537)
    int blendEnable; float blendCoef[4][4];
    // TMA descriptors of the framebuffer images (capi.cu: tma_maps): [0..2] colour / depth / present with the 8x4-pixel
    // block box a warp of the tile kernel owns, [3..5] the same images with a 16x16-pixel tile box (fused clear of empty
    // tiles).  useTma = 0 when an image cannot be described (row pitch not a multiple of 16 bytes, unaligned base).
    alignas(64) CUtensorMap tmap[6]; int useTma;
    int hwFilter;   // bilinear footprints through the TEX unit (8-bit weights) instead of four loads + fp32 lerps
    // coverage sample pattern of this draw, in 1/16 pixel (from cMsaaSampleSequenceAgg, CU:223-234)
    alignas(16) int32_t sampX[16]; alignas(16) int32_t sampY[16];   // 32-bit: read with wide constant-bank loads, no byte extraction in the sample loop
    int sampMinX, sampMaxX, sampMinY, sampMaxY;   // extents in sub-pixel (1/256) units
    // geometry
    const uint8_t* vb; int stride, nattr, nvary, nverts; int attrOff[IFR_MAX_VARY + 1];
    const int32_t* ib; int ntri;
    float4* pos; float4* vary;       // outputs of the stand-alone vertex stage (unused when it is fused into setup)
    // where the tile kernel reads varying k of vertex i: varyBase + i * varyStride + varyOff[k] (the vertex
    // buffer itself when the vertex stage is fused: the registry's vertex shaders pass attributes through)
    const uint8_t* varyBase; int varyStride; int varyOff[IFR_MAX_VARY];
    // assembled triangles: rec[slot*3+v]; slot = prim (fan 0) or ntri + prim (fan 1)
    TriRec* rec; float* clipBary;    // clipBary[slot*6]
    // attribute planes of the FAST textured shaders (kernels_geometry.cu: planes_kernel; tile_kernel.cuh: interp_planes):
    // IFR_PLANE_QUADS float4 per slot, null when the draw's shader does not use them
    const float4* planes;
    uint2* tileBox;                  // per slot: packed tile bbox, x = x0 | x1<<16, y = y0 | y1<<16 | small<<31; x0>x1 = none
    uint32_t* fan1List;              // slots of second fan triangles (clipped primitives only)
    uint32_t* heavyBits;             // one bit per first-fan slot: left to the heavy fill pass (multi-tile or big)
    // binning: two queues per tile.  Index 0 = small triangles (u32 slot per pair), 1 = big triangles
    // (one TriS per pair).  tileCount / tileCursor: [2][ntiles]; tileStart: [2][ntiles + 1]
    uint32_t* tileCount; uint32_t* tileStart; uint32_t* tileCursor;
    // tile kernel work list, written by the scan pass: tileOrder[0 .. nWork) = owned tiles with a non-empty queue,
    // tileOrder[ntiles - 1 - k], k < nEmpty = owned tiles with nothing binned (fused clear only); nWork = counters[2],
    // nEmpty = counters[7].  CTA b of the tile kernel takes tileOrder[b]; CTAs beyond nWork clear IFR_EMPTY_PER_CTA
    // empty tiles each.
    uint32_t* tileOrder;
    uint32_t* queue; uint32_t queueCap; TriS* bigQueue; uint32_t bigCap;
    // tile kernel grid: the exact CTA count (work tiles + groups of empty tiles) read back by the draw that sized the queues for
    // this geometry / state, 0 = not known (one CTA per tile, the surplus exits).  A grid that turns out too small (inputs changed
    // behind the library's back) raises counters[9]; the host then re-issues the draw like a queue overflow.
    uint32_t gridHint;
    int bigChunk;                    // big-queue entries per staged chunk of the tile kernel: 32 or 64 (tile_kernel.cuh)
    uint32_t epoch; unsigned long long* scanPartial;   // chained scan: (epoch << 32 | block total) per scan CTA
    unsigned long long* counters;    // [0] assembled (host sum), [1] small pairs, [2] nWork (device) / frags (host sum), [3] overflow flag,
                                     // [4] fan1 count, [5] big pairs, [6] blending: a tile held more than IFR_BLEND_CAP triangles,
                                     // [7] nEmpty, [8..11] scan tickets (one per scan class), [12] tile grid too small, [13..15] spare
    // resources
    TexDesc tex[IFR_MAX_TEX]; SampDesc samp[IFR_MAX_TEX]; const float* buf[IFR_MAX_BUF];
    // the uniform block of the shader registry = the first 128 bytes of buffer slot 0 as last written through
    // copyHostBufferToBuffer, shadowed by the host so that the kernels read it from the constant bank (the reference's
    // shaders read GetBufferPtr(slot) from global memory per invocation, SB:394)
    alignas(16) float u0[32];
};

// RGBA32F -> packed RGBA8 UNORM (r in the low byte): round-to-nearest-even of saturate(c) * 255
__device__ __forceinline__ uint32_t ifr_pack_rgba8(const float4& c) {
    return __float2uint_rn(__saturatef(c.x) * 255.0f) | (__float2uint_rn(__saturatef(c.y) * 255.0f) << 8) |
           (__float2uint_rn(__saturatef(c.z) * 255.0f) << 16) | (__float2uint_rn(__saturatef(c.w) * 255.0f) << 24);
}

// Statistics counters that every warp of a big grid bumps (assembled triangles, shaded fragments) are spread over
// IFR_SPREAD slots, one 32-byte sector each, behind the 16 scalar counters: a single address would serialise hundreds
// of thousands of reductions in one L2 slice (measured: 90 us of a 195 us tile kernel).  The host sums them.
#define IFR_SPREAD 256
#define IFR_COUNTER_SCALARS 16
#define IFR_COUNTER_BYTES (IFR_COUNTER_SCALARS * 8 + 2 * IFR_SPREAD * 32)
#define IFR_EMPTY_PER_CTA 8
__device__ __forceinline__ void ifr_count_spread(unsigned long long* counters, int which, unsigned int slotHint, unsigned long long v) {
    atomicAdd(counters + IFR_COUNTER_SCALARS + ((size_t)which * IFR_SPREAD + (slotHint & (IFR_SPREAD - 1))) * 4, v);
}

// sort-first ownership of a tile: 64-px super-tiles interleaved over ranks
__host__ __device__ inline int ifr_tile_owner(int tx, int ty, int world) {
    const int sx = tx / (IFR_SUPER / IFR_TILE), sy = ty / (IFR_SUPER / IFR_TILE);
    return (sx + 3 * sy) % world;
}

#ifndef __CUDACC_RTC__
#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) return r->fail(IFRIT_ERR_CUDA, #expr, cudaGetErrorString(_e)); \
    } while (0)

// kernel launchers (defined in the .cu files)
struct Launch { cudaStream_t stream; };
cudaError_t ifr_launch_vertex(const DrawParams& p, int vsId, cudaStream_t s);
cudaError_t ifr_launch_setup_count(const DrawParams& p, int vsMode, cudaStream_t s);   // vsMode -1: positions from vs_kernel
cudaError_t ifr_launch_scan(const DrawParams& p, cudaStream_t s);   // both queues
cudaError_t ifr_launch_fill(const DrawParams& p, cudaStream_t s);
cudaError_t ifr_launch_tiles(const DrawParams& p, int fsId, cudaStream_t s);
cudaError_t ifr_launch_tiles_blend(const DrawParams& p, int fsId, cudaStream_t s);
cudaError_t ifr_launch_clear(const DrawParams& p, cudaStream_t s);
cudaError_t ifr_launch_blit(const TexDesc& src, int srcLevel, int srcLayer, const TexDesc& dst, int dstLevel,
                            int dstLayer, const uint32_t srcRect[4], const uint32_t dstRect[4], int filter,
                            cudaStream_t s);
cudaError_t ifr_launch_pack(const DrawParams& p, float* dst, int rank, int world, int unpack, cudaStream_t s);
cudaError_t ifr_launch_owned_to_host(const float4* src, float4* hostMapped, int W, int H, int rank, int world, cudaStream_t s);
cudaError_t ifr_launch_flag_store(uint32_t* word, uint32_t value, cudaStream_t s);
cudaError_t ifr_launch_flag_wait(const uint32_t* words, int n, uint32_t value, uint32_t* timedOut, cudaStream_t s);
#define IFR_FLAG_WORDS 128    // frame-flag block at the tail of the depth allocation: [0..63] per-rank, [64] owner's release
cudaError_t ifr_launch_planes(const DrawParams& p, cudaStream_t s);
cudaError_t ifr_launch_classify(const DrawParams& p, int vsMode, int first, int count, uint8_t* mask, cudaStream_t s);
cudaError_t ifr_launch_prims(const DrawParams& p, int mode, int vsId, int fsId, unsigned long long* keys, float4* seg, cudaStream_t s);
cudaError_t ifr_launch_mesh_compact(const int* counts, int* starts, int nwg, const float4* pos, const float4* vary, const int* idx, int nvary,
                                    float4* vb, int* ib, int stage, cudaStream_t s);
cudaError_t ifr_launch_task_compact(const int3* sub, const int* num, int ntask, int* parentStart, int4* list, int* totals, int stage, cudaStream_t s);
int ifr_kernels_per_draw();
#endif   // !__CUDACC_RTC__
#ifndef IFR_FLAG_WORDS
#define IFR_FLAG_WORDS 128
#endif
