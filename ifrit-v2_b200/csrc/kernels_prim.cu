// kernels_prim.cu — point and line polygon modes.
//   pointRasterizationInsertKernel CU:2782 + Alloc / Place CU:2818-2838 + PointFragmentStage::pixelShadingKernel CU:2848
//   lineGeometryAssemblyKernel CU:3021 + devLineCullClip CU:2976 + bresenhamRasterizationKernel CU:3060 +
//   lineRasterAlloc / Place CU:3129-3162 + LineFragmentStage::pixelShadingKernel CU:3164
// The reference builds a per-pixel candidate list (three passes + atomics) that its shading kernel then walks for the
// smallest z.  Here the walk is folded into the insertion: every candidate does ONE 64-bit atomicMin of
// (monotone z, candidate id) on its pixel, the shading kernel reads the winner.  Arithmetic as restated in
// oracle/ifrit_oracle.c: orc_render_prims (same deviations: row = index / width, pixels past the frame edge skipped,
// ties keep the lowest candidate id).  1x, no blending, no derivatives — as the reference.
#include "common.cuh"
#include "tri_setup.cuh"
#include "sampler.cuh"

#define IFR_PRIM_FAR 1e9f      // localDepth = 1e9, CU:2861 / CU:3177

__device__ __forceinline__ float to_unit(float v) { return (float)((double)v * 0.5 + 0.5); }   // `v * 0.5 + 0.5`, double literals

__device__ __forceinline__ void key_min_global(unsigned long long* keys, size_t pix, float z, uint32_t id) {
    if (!(z < IFR_PRIM_FAR)) return;
    atomicMin(keys + pix, ((unsigned long long)mono(z) << 32) | id);
}

__global__ void __launch_bounds__(256) prim_reset_kernel(unsigned long long* keys, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) keys[i] = ~0ull;
}

template <int VS>
__device__ __forceinline__ float4 fetch_pos(const DrawParams& p, int vi) {
    return run_vs<VS>(p, __ldg(reinterpret_cast<const float4*>(p.vb + (size_t)vi * p.stride + p.attrOff[0])));
}

// one thread per index: the vertex is a point
template <int VS>
__global__ void __launch_bounds__(256) point_insert_kernel(const __grid_constant__ DrawParams p, unsigned long long* keys, int nidx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nidx) return;
    float4 v = fetch_pos<VS>(p, __ldg(p.ib + i));
    if (v.w < 1e-8f) return;                                  // CU_EPS
    v.x /= v.w; v.y /= v.w; v.z /= v.w;
    v.x = to_unit(v.x); v.y = to_unit(v.y);
    if (v.x < 0.0f || v.x > 1.0f || v.y < 0.0f || v.y > 1.0f) return;
    const int px = (int)__fmul_rn(v.x, (float)(p.W - 1)), py = (int)__fmul_rn(v.y, (float)(p.H - 1));
    key_min_global(keys, (size_t)py * p.W + px, v.z, (uint32_t)i);
}

__device__ __forceinline__ float lerp_ref(float a, float b, float t) { return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t)); }
__device__ __forceinline__ float line_percent(const DrawParams& p, float4 v1, float4 v2, int px, int py) {
    v1.x = __fmul_rn(v1.x, (float)p.W); v1.y = __fmul_rn(v1.y, (float)p.H); v2.x = __fmul_rn(v2.x, (float)p.W); v2.y = __fmul_rn(v2.y, (float)p.H);
    const float dx = fabsf(__fsub_rn(v2.x, v1.x)), dy = fabsf(__fsub_rn(v2.y, v1.y));
    if (dx >= dy) return __fdiv_rn(__fsub_rn((float)px, v1.x), __fsub_rn(v2.x, v1.x));
    return __fdiv_rn(__fsub_rn((float)py, v1.y), __fsub_rn(v2.y, v1.y));
}

// one thread per (triangle, edge): clip, map, Bresenham; seg[2l], seg[2l + 1] = the mapped ends (seg[2l].w = NaN marks a dropped segment)
template <int VS>
__global__ void __launch_bounds__(128) line_insert_kernel(const __grid_constant__ DrawParams p, unsigned long long* keys, float4* seg, int nlines) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const int t = l / 3, e = l % 3;
    float4 v1 = fetch_pos<VS>(p, __ldg(p.ib + 3 * t + e)), v2 = fetch_pos<VS>(p, __ldg(p.ib + 3 * t + (e + 1) % 3));
    bool ok = !(v1.w < 0 && v2.w < 0);
    if (ok) {
        if (v1.w < 0) {
            const float tt = __fdiv_rn(__fsub_rn(v2.w, 1.0f), __fsub_rn(v2.w, v1.w));
            v1 = make_float4(lerp_ref(v2.x, v1.x, tt), lerp_ref(v2.y, v1.y, tt), lerp_ref(v2.z, v1.z, tt), lerp_ref(v2.w, v1.w, tt));
        }
        if (v2.w < 0) {
            const float tt = __fdiv_rn(__fsub_rn(v1.w, 1.0f), __fsub_rn(v1.w, v2.w));
            v2 = make_float4(lerp_ref(v1.x, v2.x, tt), lerp_ref(v1.y, v2.y, tt), lerp_ref(v1.z, v2.z, tt), lerp_ref(v1.w, v2.w, tt));
        }
        v1.x = __fdiv_rn(v1.x, v1.w); v1.y = __fdiv_rn(v1.y, v1.w); v1.z = __fdiv_rn(v1.z, v1.w);
        v2.x = __fdiv_rn(v2.x, v2.w); v2.y = __fdiv_rn(v2.y, v2.w); v2.z = __fdiv_rn(v2.z, v2.w);
        v1.x = to_unit(v1.x); v1.y = to_unit(v1.y); v2.x = to_unit(v2.x); v2.y = to_unit(v2.y);
        ok = !(v1.x < 0.0f || v1.y < 0.0f || v1.x > 1.0f || v1.y > 1.0f) && !(v2.x < 0.0f || v2.y < 0.0f);
    }
    if (!ok) { seg[2 * l] = make_float4(0, 0, 0, __int_as_float(0x7fc00000)); return; }
    seg[2 * l] = v1; seg[2 * l + 1] = v2;
    int x1 = (int)__fmul_rn(v1.x, (float)(p.W - 1)), y1 = (int)__fmul_rn(v1.y, (float)(p.H - 1));
    int x2 = (int)__fmul_rn(v2.x, (float)(p.W - 1)), y2 = (int)__fmul_rn(v2.y, (float)(p.H - 1));
    const bool majorX = abs(x1 - x2) > abs(y1 - y2);
    if (majorX ? (x2 < x1) : (y2 < y1)) { int q = x1; x1 = x2; x2 = q; q = y1; y1 = y2; y2 = q; }
    int cx = x1, cy = y1;
    const int n = majorX ? (x2 - x1) : (y2 - y1);
    const float dslope = majorX ? fabsf(__fdiv_rn(__fmul_rn(1.0f, (float)(y2 - y1)), (float)(x2 - x1)))
                                : fabsf(__fdiv_rn(__fmul_rn(1.0f, (float)(x2 - x1)), (float)(y2 - y1)));
    const int step = majorX ? ((y2 - y1) > 0 ? 1 : -1) : ((x2 - x1) > 0 ? 1 : -1);
    float d = 0.0f;
    for (int i = 0; i <= n; i++) {
        if (cx >= 0 && cx < p.W && cy >= 0 && cy < p.H) {
            const float pc = line_percent(p, v1, v2, cx, cy);
            key_min_global(keys, (size_t)cy * p.W + cx, lerp_ref(v1.z, v2.z, pc), (uint32_t)l);
        }
        d = __fadd_rn(d, dslope);
        if (d > 0.5f) { if (majorX) cy += step; else cx += step; d = __fsub_rn(d, 1.0f); }
        if (majorX) cx++; else cy++;
    }
}

// one thread per pixel: the winner, if any, against the stored depth; colour replaces
template <int FS>
__global__ void __launch_bounds__(256) prim_shade_kernel(const __grid_constant__ DrawParams p, const unsigned long long* keys, const float4* seg, int mode) {
    constexpr int NV = FsTraits<FS>::nv;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int shaded = 0;
    if (pix < (size_t)p.W * p.H) {
        const unsigned long long k = keys[pix];
        int prim = -1;
        if (k != ~0ull) {
            const uint32_t zb = (uint32_t)(k >> 32);
            const uint32_t raw = (zb & 0x80000000u) ? (zb ^ 0x80000000u) : ~zb;      // inverse of mono()
            const float z = __uint_as_float(raw);
            const uint32_t id = (uint32_t)k;
            if (z < p.depth[pix]) {
                float4 v[NV > 0 ? NV : 1], zero[NV > 0 ? NV : 1];
#pragma unroll
                for (int j = 0; j < (NV > 0 ? NV : 1); j++) { v[j] = make_float4(0, 0, 0, 0); zero[j] = v[j]; }
                if (mode == IF_POLYGON_MODE_POINT) {
                    prim = __ldg(p.ib + id);
#pragma unroll
                    for (int j = 0; j < NV; j++)
                        if (j < p.nvary) v[j] = __ldg(reinterpret_cast<const float4*>(p.varyBase + p.varyOff[j] + (size_t)prim * p.varyStride));
                } else {
                    const int t = (int)id / 3, e = (int)id % 3;
                    const float4 v1 = seg[2 * id], v2 = seg[2 * id + 1];
                    const float pc = line_percent(p, v1, v2, (int)(pix % p.W), (int)(pix / p.W));
                    const float z1 = v1.w, z2 = v2.w;
                    const float zt = __fdiv_rn(__fmul_rn(z1, z2), __fadd_rn(z2, __fmul_rn(pc, __fsub_rn(z1, z2))));
                    const int i1 = __ldg(p.ib + 3 * t + e), i2 = __ldg(p.ib + 3 * t + (e + 1) % 3);
#pragma unroll
                    for (int j = 0; j < NV; j++) {
                        if (j >= p.nvary) continue;
                        float4 a1 = __ldg(reinterpret_cast<const float4*>(p.varyBase + p.varyOff[j] + (size_t)i1 * p.varyStride));
                        float4 a2 = __ldg(reinterpret_cast<const float4*>(p.varyBase + p.varyOff[j] + (size_t)i2 * p.varyStride));
                        a1.x = __fdiv_rn(a1.x, z1); a1.y = __fdiv_rn(a1.y, z1); a1.z = __fdiv_rn(a1.z, z1); a1.w = __fdiv_rn(a1.w, z1);
                        a2.x = __fdiv_rn(a2.x, z2); a2.y = __fdiv_rn(a2.y, z2); a2.z = __fdiv_rn(a2.z, z2); a2.w = __fdiv_rn(a2.w, z2);
                        v[j].x = __fmul_rn(__fadd_rn(__fmul_rn(pc, __fsub_rn(a2.x, a1.x)), a1.x), zt);
                        v[j].y = __fmul_rn(__fadd_rn(__fmul_rn(pc, __fsub_rn(a2.y, a1.y)), a1.y), zt);
                        v[j].z = __fmul_rn(__fadd_rn(__fmul_rn(pc, __fsub_rn(a2.z, a1.z)), a1.z), zt);
                        v[j].w = __fmul_rn(__fadd_rn(__fmul_rn(pc, __fsub_rn(a2.w, a1.w)), a1.w), zt);
                    }
                    prim = (int)id;
                }
                const float4 c = run_fs<FS>(p, v, zero, zero, mode == IF_POLYGON_MODE_POINT ? prim : (int)id / 3, 0);
                p.color[pix] = c;
                p.depth[pix] = z;
                if (p.present) p.present[pix] = ifr_pack_rgba8(c);
                shaded = 1;
            }
        }
        if (p.primIds) p.primIds[pix] = shaded ? prim : -1;
    }
    shaded = __reduce_add_sync(0xffffffffu, shaded);
    if ((threadIdx.x & 31) == 0 && shaded) ifr_count_spread(p.counters, 1, blockIdx.x * 8u + (threadIdx.x >> 5), (unsigned long long)shaded);
}

template <int FS>
static cudaError_t launch_shade(const DrawParams& p, const unsigned long long* keys, const float4* seg, int mode, cudaStream_t s) {
    const size_t n = (size_t)p.W * p.H;
    prim_shade_kernel<FS><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, keys, seg, mode);
    return cudaGetLastError();
}

// keys: W * H 64-bit words; seg: 2 float4 per (triangle, edge) in line mode
cudaError_t ifr_launch_prims(const DrawParams& p, int mode, int vsId, int fsId, unsigned long long* keys, float4* seg, cudaStream_t s) {
    const size_t n = (size_t)p.W * p.H;
    prim_reset_kernel<<<148 * 8, 256, 0, s>>>(keys, n);
    const int nidx = p.ntri * 3;
    if (nidx > 0) {
        if (mode == IF_POLYGON_MODE_POINT) {
            if (vsId == IFRIT_VS_MVP) point_insert_kernel<IFRIT_VS_MVP><<<(nidx + 255) / 256, 256, 0, s>>>(p, keys, nidx);
            else point_insert_kernel<IFRIT_VS_PASSTHROUGH><<<(nidx + 255) / 256, 256, 0, s>>>(p, keys, nidx);
        } else {
            if (vsId == IFRIT_VS_MVP) line_insert_kernel<IFRIT_VS_MVP><<<(nidx + 127) / 128, 128, 0, s>>>(p, keys, seg, nidx);
            else line_insert_kernel<IFRIT_VS_PASSTHROUGH><<<(nidx + 127) / 128, 128, 0, s>>>(p, keys, seg, nidx);
        }
    }
    switch (fsId) {
        case IFRIT_FS_VARYING0: case IFRIT_FS_DEPTH_MOD: return launch_shade<IFRIT_FS_VARYING0>(p, keys, seg, mode, s);
        case IFRIT_FS_NORMAL_COLOR: return launch_shade<IFRIT_FS_NORMAL_COLOR>(p, keys, seg, mode, s);
        case IFRIT_FS_PRIM_ID: return launch_shade<IFRIT_FS_PRIM_ID>(p, keys, seg, mode, s);
        case IFRIT_FS_TEXTURED: return launch_shade<IFRIT_FS_TEXTURED>(p, keys, seg, mode, s);
        case IFRIT_FS_TEXTURED_LOD0: return launch_shade<IFRIT_FS_TEXTURED_LOD0>(p, keys, seg, mode, s);
        case IFRIT_FS_SKYBOX: return launch_shade<IFRIT_FS_SKYBOX>(p, keys, seg, mode, s);
        case IFRIT_FS_DERIV_VIS: return launch_shade<IFRIT_FS_DERIV_VIS>(p, keys, seg, mode, s);
        default: return cudaErrorInvalidValue;
    }
}


// ---- mesh / task front end: ordered compaction (meshShadingCompactionKernel CU:384-401, taskShadingScan / CompactKernel CU:418-444) -----
// The reference allocates output ranges with atomics ("TODO: Ordering"); here a single-CTA scan keeps work-group order, so primitive ids —
// and with them depth-tie resolution — are those of the same geometry drawn through drawElements.
#define IFR_MS_MAX 768
__global__ void __launch_bounds__(1024) mesh_scan_kernel(const int* __restrict__ counts, int* __restrict__ starts, int nwg) {
    __shared__ int wsumV[32], wsumI[32], carryV, carryI;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { carryV = 0; carryI = 0; }
    __syncthreads();
    for (int base = 0; base < nwg; base += 1024) {
        const int w = base + tid;
        const int nv = w < nwg ? min(max(counts[2 * w], 0), IFR_MS_MAX) : 0, ni = w < nwg ? min(max(counts[2 * w + 1], 0), IFR_MS_MAX) : 0;
        int iv = nv, ii = ni;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, iv, o), b = __shfl_up_sync(0xffffffffu, ii, o);
            if (lane >= o) { iv += a; ii += b; }
        }
        if (lane == 31) { wsumV[wid] = iv; wsumI[wid] = ii; }
        __syncthreads();
        if (wid == 0) {
            int a = wsumV[lane], b = wsumI[lane];
            const int a0 = a, b0 = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(0xffffffffu, a, o), y = __shfl_up_sync(0xffffffffu, b, o);
                if (lane >= o) { a += x; b += y; }
            }
            wsumV[lane] = a - a0; wsumI[lane] = b - b0;
        }
        __syncthreads();
        const int cv = carryV, ci = carryI;
        if (w < nwg) { starts[2 * w] = cv + wsumV[wid] + iv - nv; starts[2 * w + 1] = ci + wsumI[wid] + ii - ni; }
        __syncthreads();
        if (tid == 1023) { carryV = cv + wsumV[31] + iv; carryI = ci + wsumI[31] + ii; }
        __syncthreads();
    }
    if (tid == 0) { starts[2 * nwg] = carryV; starts[2 * nwg + 1] = carryI; }
}
__global__ void __launch_bounds__(256) mesh_copy_kernel(const int* __restrict__ counts, const int* __restrict__ starts, const float4* __restrict__ pos,
                                                        const float4* __restrict__ vary, const int* __restrict__ idx, int nvary, float4* __restrict__ vb, int* __restrict__ ib) {
    const int w = blockIdx.x;
    const int nv = min(max(counts[2 * w], 0), IFR_MS_MAX), ni = min(max(counts[2 * w + 1], 0), IFR_MS_MAX);
    const int v0 = starts[2 * w], i0 = starts[2 * w + 1];
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        float4* dst = vb + (size_t)(v0 + v) * (1 + nvary);
        dst[0] = pos[(size_t)w * IFR_MS_MAX + v];
        for (int k = 0; k < nvary; k++) dst[1 + k] = vary[((size_t)w * IFR_MS_MAX + v) * nvary + k];
    }
    for (int i = threadIdx.x; i < ni; i += blockDim.x) ib[i0 + i] = idx[(size_t)w * IFR_MS_MAX + i] + v0;
}
cudaError_t ifr_launch_mesh_compact(const int* counts, int* starts, int nwg, const float4* pos, const float4* vary, const int* idx, int nvary,
                                    float4* vb, int* ib, int stage, cudaStream_t s) {
    if (stage == 0) mesh_scan_kernel<<<1, 1024, 0, s>>>(counts, starts, nwg);
    else mesh_copy_kernel<<<nwg, 256, 0, s>>>(counts, starts, pos, vary, idx, nvary, vb, ib);
    return cudaGetLastError();
}
// one CTA: ordered list of the mesh work groups the task shaders asked for
__global__ void __launch_bounds__(256) task_compact_kernel(const int3* __restrict__ sub, const int* __restrict__ num, int ntask, int* __restrict__ parentStart,
                                                           int4* __restrict__ list, int* __restrict__ totals) {
    __shared__ int run, mx, my, mz;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int w = 0; w < ntask; w++) { parentStart[w] = acc; acc += min(max(num[w], 0), 32); }
        run = acc; mx = 0; my = 0; mz = 0;
    }
    __syncthreads();
    for (int w = threadIdx.x; w < ntask; w += blockDim.x) {
        const int n = min(max(num[w], 0), 32), s0 = parentStart[w];
        for (int i = 0; i < n; i++) {
            const int3 d = sub[(size_t)w * 32 + i];
            list[s0 + i] = make_int4(d.x, d.y, d.z, w);
            atomicMax(&mx, d.x); atomicMax(&my, d.y); atomicMax(&mz, d.z);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { totals[0] = run; totals[1] = mx; totals[2] = my; totals[3] = mz; }
}
cudaError_t ifr_launch_task_compact(const int3* sub, const int* num, int ntask, int* parentStart, int4* list, int* totals, int, cudaStream_t s) {
    task_compact_kernel<<<1, 256, 0, s>>>(sub, num, ntask, parentStart, list, totals);
    return cudaGetLastError();
}
