// kernels_image.cu — explicit clear, mip-chain blits, and sort-first super-tile pack / unpack.
//   imageResetFloat32{Mono,RGBA}Kernel CU:3337-3391  -> clear_kernel (128-bit stores; only used when a
//                                                       clear is not followed by a draw: normally fused)
//   blitImage{Bilinear,Nearest}Kernel  CUIMG:36-86   -> blit_kernel
//   (no reference counterpart)                       -> pack_kernel: gather plumbing for SURVEY §8e
#include "common.cuh"
#include "../../include/ifrit_b200.h"

__global__ void __launch_bounds__(256) clear_kernel(const __grid_constant__ DrawParams p) {
    const size_t n = (size_t)p.W * p.H;
    const float4 c = make_float4(p.clearColor[0], p.clearColor[1], p.clearColor[2], p.clearColor[3]);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (p.scissorEnable && p.nScissor) {
            const int x = (int)(i % p.W), y = (int)(i / p.W);
            bool pass = false;
            for (int k = 0; k < p.nScissor; k++) {
                const float* s = p.scissor[k];
                if ((float)x >= s[0] && (float)x <= s[0] + s[2] && (float)y >= s[1] && (float)y <= s[1] + s[3]) pass = true;
            }
            if (!pass) { if (p.present) p.present[i] = ifr_pack_rgba8(p.color[i]); continue; }
        }
        if (p.world > 1 && (ifr_tile_owner((int)(i % p.W) / IFR_TILE, (int)(i / p.W) / IFR_TILE, p.world) != p.rank) != (p.skipEmptyStores == 2)) continue;
        p.color[i] = c;
        if (p.skipEmptyStores == 2) continue;        // pre-clear of the other ranks' tiles: colour only (their depth stays with them)
        p.depth[i] = p.clearDepth;
        if (p.present) p.present[i] = ifr_pack_rgba8(c);
    }
}

cudaError_t ifr_launch_clear(const DrawParams& p, cudaStream_t s) {
    clear_kernel<<<148 * 8, 256, 0, s>>>(p);
    return cudaGetLastError();
}

struct BlitArgs {
    const float4* src; float4* dst;
    int srcW, srcH, dstW;
    uint32_t sx0, sy0, sx1, sy1, dx0, dy0, dx1, dy1;
    int filter;
};

__global__ void __launch_bounds__(256) blit_kernel(const BlitArgs a) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x + a.dx0 >= a.dx1 || y + a.dy0 >= a.dy1) return;
    const float px = (float)x / (float)(a.dx1 - a.dx0), py = (float)y / (float)(a.dy1 - a.dy0);
    const float sx = fmaf(px, (float)(a.sx1 - a.sx0), (float)a.sx0), sy = fmaf(py, (float)(a.sy1 - a.sy0), (float)a.sy0);
    float4 out;
    if (a.filter == IF_FILTER_LINEAR) {
        const int ix = min(max((int)sx, 0), a.srcW - 1), iy = min(max((int)sy, 0), a.srcH - 1);   // in range for validated rectangles
        const float fx = sx - (float)ix, fy = sy - (float)iy;
        const int ix1 = min(ix + 1, a.srcW - 1), iy1 = min(iy + 1, a.srcH - 1);
        const float4 c00 = __ldg(a.src + (size_t)iy * a.srcW + ix), c01 = __ldg(a.src + (size_t)iy * a.srcW + ix1);
        const float4 c10 = __ldg(a.src + (size_t)iy1 * a.srcW + ix), c11 = __ldg(a.src + (size_t)iy1 * a.srcW + ix1);
        const float4 p0 = make_float4(fmaf(c01.x - c00.x, fx, c00.x), fmaf(c01.y - c00.y, fx, c00.y),
                                      fmaf(c01.z - c00.z, fx, c00.z), fmaf(c01.w - c00.w, fx, c00.w));
        const float4 p1 = make_float4(fmaf(c11.x - c10.x, fx, c10.x), fmaf(c11.y - c10.y, fx, c10.y),
                                      fmaf(c11.z - c10.z, fx, c10.z), fmaf(c11.w - c10.w, fx, c10.w));
        out = make_float4(fmaf(p1.x - p0.x, fy, p0.x), fmaf(p1.y - p0.y, fy, p0.y), fmaf(p1.z - p0.z, fy, p0.z),
                          fmaf(p1.w - p0.w, fy, p0.w));
    } else {
        const int ix = min(max((int)roundf(sx), 0), a.srcW - 1), iy = min(max((int)roundf(sy), 0), a.srcH - 1);
        out = __ldg(a.src + (size_t)iy * a.srcW + ix);
    }
    a.dst[(size_t)(y + a.dy0) * a.dstW + (x + a.dx0)] = out;
}

static void level_of(const TexDesc& t, int level, int layer, int& w, int& h, long long& off) {
    long long o = 0;
    int lw = t.w, lh = t.h;
    for (int i = 0; i < level; i++) { o += (long long)lw * lh * t.layers; lw = (lw + 1) >> 1; lh = (lh + 1) >> 1; }
    w = lw; h = lh; off = o + (long long)lw * lh * layer;
}

cudaError_t ifr_launch_blit(const TexDesc& src, int srcLevel, int srcLayer, const TexDesc& dst, int dstLevel,
                            int dstLayer, const uint32_t sr[4], const uint32_t dr[4], int filter, cudaStream_t s) {
    BlitArgs a;
    int sw, sh, dw, dh;
    long long so, dof;
    level_of(src, srcLevel, srcLayer, sw, sh, so);
    level_of(dst, dstLevel, dstLayer, dw, dh, dof);
    a.src = src.data + so; a.dst = const_cast<float4*>(dst.data) + dof;
    a.srcW = sw; a.srcH = sh; a.dstW = dw;
    a.sx0 = sr[0]; a.sy0 = sr[1]; a.sx1 = sr[2]; a.sy1 = sr[3];
    a.dx0 = dr[0]; a.dy0 = dr[1]; a.dx1 = dr[2]; a.dy1 = dr[3];
    a.filter = filter;
    const int w = (int)(dr[2] - dr[0]), h = (int)(dr[3] - dr[1]);
    if (w <= 0 || h <= 0) return cudaSuccess;
    blit_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), 256, 0, s>>>(a);
    return cudaGetLastError();
}

// Sort-first gather plumbing: the super-tiles owned by `rank` are stored contiguously, in raster
// order of super-tiles, each as 64x64 float4 colours followed (after all colours) by nothing else:
// depth never leaves the GPU in the reference either (CU:3899-3909 copies colour only).
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ DrawParams p, float4* buf, int rank,
                                                   int world, int unpack) {
    const int sx = blockIdx.x, sy = blockIdx.y;
    const int stx = (p.W + IFR_SUPER - 1) / IFR_SUPER;
    if ((sx + 3 * sy) % world != rank) return;
    // packed index = number of owned super-tiles before (sx, sy) in raster order
    int idx = 0;
    for (int y = 0; y <= sy; y++) {
        const int xe = (y == sy) ? sx : stx;
        // count x in [0, xe) with (x + 3y) % world == rank
        const int r0 = ((rank - 3 * y) % world + world) % world;
        if (xe > r0) idx += (xe - r0 + world - 1) / world;
    }
    float4* tile = buf + (size_t)idx * IFR_SUPER * IFR_SUPER;
    for (int i = threadIdx.x; i < IFR_SUPER * IFR_SUPER; i += blockDim.x) {
        const int x = sx * IFR_SUPER + (i % IFR_SUPER), y = sy * IFR_SUPER + (i / IFR_SUPER);
        if (x >= p.W || y >= p.H) continue;
        if (unpack) p.color[(size_t)y * p.W + x] = tile[i];
        else tile[i] = p.color[(size_t)y * p.W + x];
    }
}

cudaError_t ifr_launch_pack(const DrawParams& p, float* dst, int rank, int world, int unpack, cudaStream_t s) {
    const dim3 grid((p.W + IFR_SUPER - 1) / IFR_SUPER, (p.H + IFR_SUPER - 1) / IFR_SUPER);
    pack_kernel<<<grid, 256, 0, s>>>(p, reinterpret_cast<float4*>(dst), rank, world, unpack);
    return cudaGetLastError();
}

// Sort-first distributed read-back (no reference counterpart; the single-GPU reference copies the whole frame back over one PCIe
// link, CU:3899-3909).  Every rank stores the super-tiles it owns from its LOCAL framebuffer straight into the full-frame, row-major
// host image through mapped pinned memory: the frame never crosses NVLink, and N ranks drive N PCIe links at once.  A CTA takes one
// owned super-tile; a warp writes 512 contiguous bytes per instruction (full PCIe write bursts).
__global__ void __launch_bounds__(256) owned_to_host_kernel(const float4* __restrict__ src, float4* __restrict__ host, int W, int H,
                                                            int rank, int world) {
    const int sx = blockIdx.x, sy = blockIdx.y;
    if ((sx + 3 * sy) % world != rank) return;
    const int x = sx * IFR_SUPER + (threadIdx.x & 63);
    if (x >= W) return;
#pragma unroll 4
    for (int r = threadIdx.x >> 6; r < IFR_SUPER; r += 4) {
        const int y = sy * IFR_SUPER + r;
        if (y >= H) break;
        host[(size_t)y * W + x] = __ldcs(src + (size_t)y * W + x);
    }
}
cudaError_t ifr_launch_owned_to_host(const float4* src, float4* hostMapped, int W, int H, int rank, int world, cudaStream_t s) {
    const dim3 grid((W + IFR_SUPER - 1) / IFR_SUPER, (H + IFR_SUPER - 1) / IFR_SUPER);
    owned_to_host_kernel<<<grid, 256, 0, s>>>(src, hostMapped, W, H, rank, world);
    return cudaGetLastError();
}

// ---- sort-first frame flags (peer-memory mode) ---------------------------------------------------------
// A block of 32-bit words at the tail of the framebuffer owner's depth allocation, mapped into every rank by
// the same CUDA IPC handle as the framebuffer: word r = last frame whose tiles rank r has stored, word 64 = last
// frame the owner has released for overwriting.  A store kernel after the tile kernel publishes, a one-warp
// kernel polls — no NCCL call and no host round trip per frame.
__global__ void flag_store_kernel(volatile uint32_t* word, uint32_t value) {
    __threadfence_system();     // the tile kernel's peer stores (previous kernel on this stream) before the flag
    *word = value;
    __threadfence_system();
}
__global__ void flag_wait_kernel(const volatile uint32_t* words, int n, uint32_t value, uint32_t* timedOut) {
    const int lane = threadIdx.x;
    if (lane >= n) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    // frame ids wrap: "reached" = not behind by half the range
    while ((int32_t)(words[lane] - value) < 0) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > 2000000000ull) { *timedOut = 1u; break; }   // 2 s: a peer is gone — never hang the device
        __nanosleep(200);
    }
    __threadfence_system();
}
cudaError_t ifr_launch_flag_store(uint32_t* word, uint32_t value, cudaStream_t s) {
    flag_store_kernel<<<1, 1, 0, s>>>(word, value);
    return cudaGetLastError();
}
cudaError_t ifr_launch_flag_wait(const uint32_t* words, int n, uint32_t value, uint32_t* timedOut, cudaStream_t s) {
    flag_wait_kernel<<<1, 64, 0, s>>>(words, n, value, timedOut);
    return cudaGetLastError();
}
