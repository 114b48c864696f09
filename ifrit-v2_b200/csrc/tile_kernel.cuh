// tile_kernel.cuh — the per-tile raster + early-Z + quad shading + MSAA resolve kernel (template; instantiated per fragment
// shader in tile_inst.cu, one translation unit per shader id so that the build runs in parallel).
//
// One CTA (256 threads) owns one 16x16 screen tile from coverage through resolve; per-sample
// visibility keys live in shared memory, so HBM sees the framebuffer once (written once when the
// clear is fused).  Replaces, in one kernel:
//   secondBinnerWorklistRasterizerKernel CU:1254 + devFinerTilingRasterizationChildProcess CU:892
//       -> coverage: exact integer edge functions per sample (small triangles: one thread per
//          triangle + shared-memory atomicMin; large triangles: one thread per pixel)
//   pixelTaggingExecKernel CU:1901  -> order-independent 64-bit (depth, slot) keys in shared memory
//   pixelShadingExecKernel CU:2145  -> 2x2 quads = 4 consecutive lanes, ddx/ddy by __shfl_xor
//   pixelMultisampleCoverageModulationKernel CU:2319 -> resolve from the shared-memory keys
//   imageResetFloat32{Mono,RGBA}Kernel CU:3337/3363 -> fused clear
#pragma once
#include "common.cuh"
#include "tri_setup.cuh"
#include "sampler.cuh"

// ---- visibility keys (identical to oracle/ifrit_oracle.c key_*) --------------------------------
// LESS (mono(d), slot+1) / LESS_OR_EQUAL (mono(d), ~(slot+1)) / GREATER, GREATER_OR_EQUAL with ~mono /
// ALWAYS and EQUAL (slot+1, depth bits); the minimum key wins.  The compare op is folded into two xor
// masks so that the hot loops carry no switch.
struct KeyMode { uint32_t xhi, xlo; bool ordered, equal, never; };
__device__ __forceinline__ KeyMode key_mode(int func) {
    KeyMode m;
    m.ordered = func == IF_COMPARE_OP_LESS || func == IF_COMPARE_OP_LESS_OR_EQUAL || func == IF_COMPARE_OP_GREATER ||
                func == IF_COMPARE_OP_GREATER_OR_EQUAL;
    m.xhi = (func == IF_COMPARE_OP_GREATER || func == IF_COMPARE_OP_GREATER_OR_EQUAL) ? 0xffffffffu : 0u;
    m.xlo = (func == IF_COMPARE_OP_LESS_OR_EQUAL || func == IF_COMPARE_OP_GREATER_OR_EQUAL) ? 0xffffffffu : 0u;
    m.equal = func == IF_COMPARE_OP_EQUAL;
    m.never = !m.ordered && !m.equal && func != IF_COMPARE_OP_ALWAYS;
    return m;
}
__device__ __forceinline__ unsigned long long key_init(const KeyMode& m, float d0) {
    if (m.ordered) return ((unsigned long long)(mono(d0) ^ m.xhi) << 32) | m.xlo;
    return ((unsigned long long)0xffffffffu << 32) | __float_as_uint(d0 + 0.0f);
}
__device__ __forceinline__ bool key_make(const KeyMode& m, float d, float d0, uint32_t slot, unsigned long long& k) {
    if (m.ordered) { k = ((unsigned long long)(mono(d) ^ m.xhi) << 32) | ((slot + 1u) ^ m.xlo); return true; }
    if (m.never || (m.equal && !(d == d0))) return false;
    k = ((unsigned long long)(slot + 1u) << 32) | __float_as_uint(d + 0.0f);
    return true;
}
__device__ __forceinline__ uint32_t key_slot(const KeyMode& m, unsigned long long k) {   // IFR_SLOT_NONE if empty
    if (m.ordered) { const uint32_t s = (uint32_t)k ^ m.xlo; return s == 0 ? IFR_SLOT_NONE : s - 1; }
    const uint32_t s = (uint32_t)(k >> 32);
    return s == 0xffffffffu ? IFR_SLOT_NONE : s - 1;
}

__device__ __forceinline__ bool scissor_pass(const DrawParams& p, int x, int y) {
    if (!p.scissorEnable || p.nScissor == 0) return true;
    for (int i = 0; i < p.nScissor; i++) {
        const float* s = p.scissor[i];
        if ((float)x >= s[0] && (float)x <= s[0] + s[2] && (float)y >= s[1] && (float)y <= s[1] + s[3]) return true;
    }
    return false;
}

// Cold paths (triangles wider than 4096 px, edge values beyond int32) are written as separate functions so that they CAN
// be kept out of line (-DIFR_COLD=__noinline__: the tile kernel is ~90 KB of SASS and instruction fetch shows up in the
// stall samples).  Measured (profiles/r02_ab.txt): out of line c4 tile 0.523 -> 0.505 ms, but c3 0.332 -> 0.338 and c2
// 0.222 -> 0.248 (the call sites pin registers of the 40-register 1x kernels) — inlined by default.
#ifndef IFR_COLD
#define IFR_COLD __forceinline__
#endif
template <int S>
__device__ IFR_COLD uint32_t coverage_wide(const DrawParams& p, int dx0, int dx1, int dx2, int dy0, int dy1, int dy2,
                                           long long e0, long long e1, long long e2, int ci, int cj) {
    uint32_t mask = 0;
    for (int k = 0; k < S; k++) {
        const long long i = ci + p.sampX[k], j = cj + p.sampY[k];
        const bool in = (e0 + 16 * (dx0 * j - dy0 * i) > 0) && (e1 + 16 * (dx1 * j - dy1 * i) > 0) && (e2 + 16 * (dx2 * j - dy2 * i) > 0);
        mask |= (uint32_t)in << k;
    }
    return mask;
}
// coverage mask of the pixel whose origin is (ci, cj) = (16 lx, 16 ly) in 1/16 px, over the S samples
template <int S>
__device__ __forceinline__ uint32_t coverage(const DrawParams& p, const int dX[3], const int dY[3], const int T[3],
                                             const long long E0[3], int wide, int ci, int cj) {
    uint32_t mask = 0;
    if (wide) return coverage_wide<S>(p, dX[0], dX[1], dX[2], dY[0], dY[1], dY[2], E0[0], E0[1], E0[2], ci, cj);
#pragma unroll
    for (int k = 0; k < S; k++) {
        const int i = ci + p.sampX[k], j = cj + p.sampY[k];
        const bool in = (dX[0] * j - dY[0] * i >= T[0]) && (dX[1] * j - dY[1] * i >= T[1]) &&
                        (dX[2] * j - dY[2] * i >= T[2]);
        mask |= (uint32_t)in << k;
    }
    return mask;
}

// The same masks for S > 1, computed by the whole warp for one (non-wide) triangle: the sample term of every edge
// function, off[e][k] = dX[e] * sampY[k] - dY[e] * sampX[k], is the same for all 32 pixels of the warp's block — lanes
// 0 .. 3S-1 compute one each into the warp's scratch row, every lane then adds them (broadcast LDS.128) to its own
// pixel term base[e] = dX[e] * cj - dY[e] * ci - T[e]; inside <=> no sum is negative.  |base + off| < 2^31: non-wide
// edges keep |dX * j - dY * i| < 2^29 and T is clamped to +-2^30 (tri_setup.cuh).  Must be called by all 32 lanes.
template <int S>
__device__ __forceinline__ uint32_t coverage_coop(const DrawParams& p, int* __restrict__ scratch, const int dX[3], const int dY[3],
                                                  const int T[3], int ci, int cj, int lane) {
    __syncwarp();                                    // the previous triangle's offsets have been read by every lane
#pragma unroll
    for (int idx = lane; idx < 3 * S; idx += 32) {
        const int e = idx / S, k = idx % S;
        const int ex = e == 0 ? dX[0] : (e == 1 ? dX[1] : dX[2]), ey = e == 0 ? dY[0] : (e == 1 ? dY[1] : dY[2]);
        scratch[idx] = ex * p.sampY[k] - ey * p.sampX[k];
    }
    __syncwarp();
    const int b0 = dX[0] * cj - dY[0] * ci - T[0], b1 = dX[1] * cj - dY[1] * ci - T[1], b2 = dX[2] * cj - dY[2] * ci - T[2];
    uint32_t mask = 0;
#pragma unroll
    for (int k4 = 0; k4 < S; k4 += 4) {
        int o0[4], o1[4], o2[4];
        if (S >= 4) {
            const int4 a = *reinterpret_cast<const int4*>(scratch + k4), b = *reinterpret_cast<const int4*>(scratch + S + k4),
                       c = *reinterpret_cast<const int4*>(scratch + 2 * S + k4);
            o0[0] = a.x; o0[1] = a.y; o0[2] = a.z; o0[3] = a.w; o1[0] = b.x; o1[1] = b.y; o1[2] = b.z; o1[3] = b.w;
            o2[0] = c.x; o2[1] = c.y; o2[2] = c.z; o2[3] = c.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                o0[j] = j < S ? scratch[k4 + j] : 0; o1[j] = j < S ? scratch[S + k4 + j] : 0; o2[j] = j < S ? scratch[2 * S + k4 + j] : 0;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (k4 + j < S) {
                const uint32_t neg = (uint32_t)((b0 + o0[j]) | (b1 + o1[j]) | (b2 + o2[j])) >> 31;
                mask |= (neg ^ 1u) << (k4 + j);
            }
        }
    }
    return mask;
}

static __device__ IFR_COLD float depth_at64(int dx0, int dx1, int dx2, int dy0, int dy1, int dy2, long long e0, long long e1, long long e2,
                                     float ar, float z0, float z1, float z2, int lx, int ly) {
    const float b0 = __fmul_rn((float)(e0 + 256 * ((long long)dx0 * ly - (long long)dy0 * lx)), ar);
    const float b1 = __fmul_rn((float)(e1 + 256 * ((long long)dx1 * ly - (long long)dy1 * lx)), ar);
    const float b2 = __fmul_rn((float)(e2 + 256 * ((long long)dx2 * ly - (long long)dy2 * lx)), ar);
    return fmaf(z2, b2, fmaf(z1, b1, __fmul_rn(z0, b0)));
}
// depth at the integer pixel (lx, ly): beta_i = E_i(corner) / d, from exact integer edge values
__device__ __forceinline__ float depth_at(const int dX[3], const int dY[3], const long long E0[3], uint32_t flags, float ar,
                                          float z0, float z1, float z2, int lx, int ly) {
    if (!(flags & IFR_FLAG_FITS32)) return depth_at64(dX[0], dX[1], dX[2], dY[0], dY[1], dY[2], E0[0], E0[1], E0[2], ar, z0, z1, z2, lx, ly);
    float beta[3];
    // triangle-level bound: the edge values at this corner fit int32 -> wrapping 32-bit arithmetic, I2F.S32
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const uint32_t Ec = (uint32_t)E0[e] + (((uint32_t)dX[e] * (uint32_t)ly - (uint32_t)dY[e] * (uint32_t)lx) << 8);
        beta[e] = __fmul_rn((float)(int)Ec, ar);
    }
    return fmaf(z2, beta[2], fmaf(z1, beta[1], __fmul_rn(z0, beta[0])));
}

// tile-local pixel -> key index = index of the owning thread (warp = 8x4 block, 4 consecutive lanes = 2x2 quad)
__device__ __forceinline__ int pix_index(int lx, int ly) {
    return (((ly >> 2) * 2 + (lx >> 3)) << 5) + ((((ly >> 1) & 1) * 4 + ((lx >> 1) & 3)) << 2) + (ly & 1) * 2 + (lx & 1);
}

__device__ __forceinline__ void key_min_shared(unsigned long long* addr, unsigned long long k) {
    if (k < *reinterpret_cast<volatile unsigned long long*>(addr)) atomicMin(addr, k);
}


// ---- attribute interpolation at an integer pixel (pixelShadingExecKernel CU:2168-2215) ---------
// Perspective weights (u, v, 1 - (u + v)) of the clip-space barycentrics from the three edge values at a pixel
// corner; beta_i = float(E_i) * ar is also what the depth plane uses.
struct PixW { float u, v, w2; };
// FASTW (the FAST textured shaders only — colour within 1/255, depth is not computed from these): SFU reciprocal, and the
// unclipped barycentric mapping as plain selects instead of fma chains with 0 / 1 factors
template <bool FASTW>
__device__ __forceinline__ PixW persp_weights(const float beta[3], const float IW[3], uint32_t flags, const float* cb) {
    float b0 = __fmul_rn(beta[0], IW[0]), b1 = __fmul_rn(beta[1], IW[1]), b2 = __fmul_rn(beta[2], IW[2]);
    const float sum = __fadd_rn(__fadd_rn(b0, b1), b2);
    const float inv = FASTW ? rcp_fast(sum) : __frcp_rn(sum);
    b0 = __fmul_rn(b0, inv); b1 = __fmul_rn(b1, inv); b2 = __fmul_rn(b2, inv);
    PixW w;
    if (flags & IFR_FLAG_CLIPPED) {
        w.u = fmaf(b2, cb[4], fmaf(b1, cb[2], __fmul_rn(b0, cb[0])));
        w.v = fmaf(b2, cb[5], fmaf(b1, cb[3], __fmul_rn(b0, cb[1])));
    } else {
        // unclipped: clip barycentrics are those of (v2, v3, v1) of the input triangle, first and last
        // exchanged by a cull swap: B = {(0,1),(0,0),(1,0)} or {(1,0),(0,0),(0,1)}.  The fma chains with
        // exact 0 / 1 factors reduce to (x * 1 + 0 * y + 0) = x up to the sign of zero, which + 0.0f fixes.
        const bool sw = flags & IFR_FLAG_SWAPPED;
        if (FASTW) { w.u = sw ? b0 : b2; w.v = sw ? b2 : b0; }
        else {
            const float B00 = sw ? 1.0f : 0.0f, B01 = sw ? 0.0f : 1.0f;
            w.u = fmaf(b2, B01, fmaf(b1, 0.0f, __fmul_rn(b0, B00)));
            w.v = fmaf(b2, B00, fmaf(b1, 0.0f, __fmul_rn(b0, B01)));
        }
    }
    w.w2 = __fsub_rn(1.0f, __fadd_rn(w.u, w.v));
    return w;
}
__device__ __forceinline__ float4 mix3(const float4& v0, const float4& v1, const float4& v2, const PixW& w) {
    float4 o;
    o.x = fmaf(v2.x, w.w2, fmaf(v1.x, w.v, __fmul_rn(v0.x, w.u)));
    o.y = fmaf(v2.y, w.w2, fmaf(v1.y, w.v, __fmul_rn(v0.y, w.u)));
    o.z = fmaf(v2.z, w.w2, fmaf(v1.z, w.v, __fmul_rn(v0.z, w.u)));
    o.w = fmaf(v2.w, w.w2, fmaf(v1.w, w.v, __fmul_rn(v0.w, w.u)));
    return o;
}

struct Betas9 { float b[3], h[3], v[3]; };
static __device__ IFR_COLD Betas9 betas64(int x0, int x1, int x2, int y0, int y1, int y2, int sx, int sy, int hs, int vs, float ar) {
    const int X[3] = {x0, x1, x2}, Y[3] = {y0, y1, y2};
    Betas9 o;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int a = (i + 1) % 3, b = (i + 2) % 3;
        const long long dx = X[b] - X[a], dy = Y[b] - Y[a];
        const long long E = dx * (long long)(sy - Y[a]) - dy * (long long)(sx - X[a]);
        o.b[i] = __fmul_rn((float)E, ar);
        o.h[i] = __fmul_rn((float)(E - dy * hs), ar);
        o.v[i] = __fmul_rn((float)(E + dx * vs), ar);
    }
    return o;
}
// Interpolates slot's varyings at pixel (px, py).  DERIV: also at the two other pixels the 2x2 quad's coarse
// derivatives need — the horizontal and the vertical neighbour inside the screen-aligned quad (SB:36-76:
// ddx = right - left of the lane's row, ddy = bottom - top of its column).  The lane evaluates its neighbours'
// values itself from the same integer edge functions (E steps by -256 dY / +256 dX per pixel), which is
// bit-identical to what a helper lane would compute, so a pixel shades alone: no helper invocations, no
// quad-uniform control flow, every FS pass runs with all of its lanes useful.
template <int NV, bool DERIV, bool FASTW = false>
__device__ __forceinline__ void interp(const DrawParams& p, uint32_t slot, int px, int py, float4* out, float4* ddx,
                                       float4* ddy, float& depth, int& prim, int& fan) {
    const int4* r = reinterpret_cast<const int4*>(p.rec + slot);
    int X[3], Y[3];
    float Z[3], IW[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int4 v = __ldg(r + j);
        X[j] = v.x; Y[j] = v.y; Z[j] = __int_as_float(v.z); IW[j] = __int_as_float(v.w);
    }
    const int4 w3 = __ldg(r + 3);
    const float ar = __int_as_float(w3.x);
    const uint32_t flags = (uint32_t)w3.y;
    fan = slot >= (uint32_t)p.ntri;
    prim = fan ? (int)(slot - (uint32_t)p.ntri) : (int)slot;
    int idx[3];
    if (NV > 0) {
#pragma unroll
        for (int j = 0; j < 3; j++) idx[j] = __ldg(p.ib + 3 * (size_t)prim + j);
    }
    const int sx = px * IFR_SUBPIX, sy = py * IFR_SUBPIX;
    const int hs = (px & 1) ? -IFR_SUBPIX : IFR_SUBPIX, vs = (py & 1) ? -IFR_SUBPIX : IFR_SUBPIX;   // towards the quad neighbours
    float beta[3], betaH[3], betaV[3];
    if (flags & IFR_FLAG_FITS32) {
        // the edge values fit int32 (triangle-level bound, two pixels around any covered sample): wrapping
        // 32-bit arithmetic is exact
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int a = (i + 1) % 3, b = (i + 2) % 3;
            const uint32_t dx = (uint32_t)(X[b] - X[a]), dy = (uint32_t)(Y[b] - Y[a]);
            const uint32_t E = dx * (uint32_t)(sy - Y[a]) - dy * (uint32_t)(sx - X[a]);
            beta[i] = __fmul_rn((float)(int)E, ar);
            if (DERIV) {
                betaH[i] = __fmul_rn((float)(int)(E - dy * (uint32_t)hs), ar);
                betaV[i] = __fmul_rn((float)(int)(E + dx * (uint32_t)vs), ar);
            }
        }
    } else {
        const Betas9 o = betas64(X[0], X[1], X[2], Y[0], Y[1], Y[2], sx, sy, hs, vs, ar);
#pragma unroll
        for (int i = 0; i < 3; i++) { beta[i] = o.b[i]; betaH[i] = o.h[i]; betaV[i] = o.v[i]; }
    }
    depth = fmaf(Z[2], beta[2], fmaf(Z[1], beta[1], __fmul_rn(Z[0], beta[0])));
    if (NV == 0) return;
    float cb[6] = {0, 0, 0, 0, 0, 0};
    if (flags & IFR_FLAG_CLIPPED) {
        const float* g = p.clipBary + (size_t)slot * 6;
#pragma unroll
        for (int j = 0; j < 6; j++) cb[j] = __ldg(g + j);
    }
    const PixW w = persp_weights<FASTW>(beta, IW, flags, cb);
    PixW wh = w, wv = w;
    if (DERIV) { wh = persp_weights<FASTW>(betaH, IW, flags, cb); wv = persp_weights<FASTW>(betaV, IW, flags, cb); }
    const bool right = px & 1, bottom = py & 1;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        if (k >= p.nvary) {
            out[k] = make_float4(0, 0, 0, 0);
            if (DERIV) { ddx[k] = out[k]; ddy[k] = out[k]; }
            continue;
        }
        const uint8_t* vb = p.varyBase + p.varyOff[k];
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(vb + (size_t)idx[0] * p.varyStride)),
                     v1 = __ldg(reinterpret_cast<const float4*>(vb + (size_t)idx[1] * p.varyStride)),
                     v2 = __ldg(reinterpret_cast<const float4*>(vb + (size_t)idx[2] * p.varyStride));
        const float4 a = mix3(v0, v1, v2, w);
        out[k] = a;
        if (DERIV) {
            const float4 o1 = mix3(v0, v1, v2, wh), o2 = mix3(v0, v1, v2, wv);
            ddx[k] = right ? make_float4(a.x - o1.x, a.y - o1.y, a.z - o1.z, a.w - o1.w)
                           : make_float4(o1.x - a.x, o1.y - a.y, o1.z - a.z, o1.w - a.w);
            ddy[k] = bottom ? make_float4(a.x - o2.x, a.y - o2.y, a.z - o2.z, a.w - o2.w)
                            : make_float4(o2.x - a.x, o2.y - a.y, o2.z - a.z, o2.w - a.w);
        }
    }
}

// ---- FAST textured shaders: varyings from the per-triangle attribute planes (kernels_geometry.cu: planes_kernel) --------------
// v[0].xyz = the perspective numerator of the normal (direction only, `textured` normalises it), v[1].xy = uv; ddx / ddy of uv
// from the planes evaluated one pixel towards the quad neighbour (same definition as interp: right - left, bottom - top of the
// lane's 2x2 quad).  Five independent 16-byte loads, no triangle record, no indices, no vertex fetches.
template <bool DERIV>
__device__ __forceinline__ void interp_planes(const DrawParams& p, uint32_t slot, int px, int py, float4* out, float4* ddx, float4* ddy) {
    const float4* q = p.planes + (size_t)slot * IFR_PLANE_QUADS;
    const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4);
    const float fx = (float)px - q4.z, fy = (float)py - q4.w;
    const float D = fmaf(q0.z, fy, fmaf(q0.y, fx, q0.x));
    const float U = fmaf(q1.y, fy, fmaf(q1.x, fx, q0.w));
    const float V = fmaf(q2.x, fy, fmaf(q1.w, fx, q1.z));
    out[0] = make_float4(fmaf(q2.w, fy, fmaf(q2.z, fx, q2.y)), fmaf(q3.z, fy, fmaf(q3.y, fx, q3.x)), fmaf(q4.y, fy, fmaf(q4.x, fx, q3.w)), 0.0f);
    const float inv = rcp_fast(D);
    const float u = U * inv, v = V * inv;
    out[1] = make_float4(u, v, 0.0f, 0.0f);
    if (DERIV) {
        const float sh = (px & 1) ? -1.0f : 1.0f, sv = (py & 1) ? -1.0f : 1.0f;      // towards the quad neighbours
        const float invH = rcp_fast(fmaf(sh, q0.y, D)), invV = rcp_fast(fmaf(sv, q0.z, D));
        const float uH = fmaf(sh, q1.x, U) * invH, vH = fmaf(sh, q1.w, V) * invH;
        const float uV = fmaf(sv, q1.y, U) * invV, vV = fmaf(sv, q2.x, V) * invV;
        ddx[0] = make_float4(0, 0, 0, 0); ddy[0] = ddx[0];
        ddx[1] = make_float4((uH - u) * sh, (vH - v) * sh, 0.0f, 0.0f);
        ddy[1] = make_float4((uV - u) * sv, (vV - v) * sv, 0.0f, 0.0f);
    }
}
// the depth a visibility key was made from (key_make): the shading pass of the plane shaders takes it from the winning key
// instead of re-evaluating the depth plane from the triangle record — bit-identical, mono() is a bijection on canonical floats
__device__ __forceinline__ float depth_from_word(const KeyMode& m, uint32_t w) {     // w = high word (ordered) / low word (others)
    if (!m.ordered) return __uint_as_float(w);
    const uint32_t t = w ^ m.xhi;
    return __uint_as_float((t & 0x80000000u) ? (t ^ 0x80000000u) : ~t);
}
template <int FS> struct FsPlanes {
    static constexpr bool value = IFR_PLANES && (FS == IFRIT_FS_TEXTURED_FAST || FS == IFRIT_FS_TEXTURED_LOD0_FAST || FS == IFRIT_FS_TEXTURED_FAST_RA);
};

// ------------------------------------------------------------------------------------------------
#define IFR_CHUNK 256
// Big-queue entries staged per bulk copy, double buffered: DrawParams::bigChunk = 32 or 64, chosen per draw by the host (capi.cu:
// issue_draw).  The staging lives in dynamic shared memory behind the keys, so that a draw of shallow queues allocates 6 KB instead
// of 12: at 8x MSAA (16 KB of keys) that takes four resident CTAs from 117 to 93 KB and moves the L1 / shared split one step
// towards L1 — c3 tile 0.331 -> 0.319 ms; deep queues (c5) and the 1x kernels keep 64 (fewer chunk turns; profiles/r02_ab.txt).
#define IFR_BIG_CHUNK_MAX 64
#ifdef IFR_NO_STCS
#define IFR_FB_STORE(ptr, v) (*(ptr) = (v))
#else
#define IFR_FB_STORE(ptr, v) __stcs(ptr, v)
#endif

// ---- TMA bulk copy (global -> shared) completing on an mbarrier -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// ---- TMA tensor copies of framebuffer blocks (descriptors: DrawParams::tmap, built by capi.cu: tma_maps) ---------------
// Coordinates are in elements of the image (x innermost); boxes that hang over the image edge are clipped by the unit.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(x), "r"(y), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {      // the staged source may be reused / released afterwards
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// per-warp staging of one 8x4-pixel block: colour 512 B | depth 128 B | present 128 B; the eight rows together also hold
// one 16x16 tile (colour 4096 B | depth 1024 B | present 1024 B) for the clear of empty tiles
// IFR_TMA_EPILOGUE = 1: the blocks of covered tiles leave (and, for LOAD-op draws, arrive) through per-warp tensor copies
// too.  Measured SLOWER than per-thread stores (profiles/r02_tma_ab.txt: c3 tile 0.332 -> 0.384 ms) — every warp ends
// on a blocking wait for the unit to read its 512-byte staging row, which holds the CTA's registers and shared memory
// while nothing computes — so the default build keeps TMA for what it wins: whole-tile clears of empty tiles and the
// big-queue staging.
#ifndef IFR_TMA_EPILOGUE
#define IFR_TMA_EPILOGUE 0
#endif
#define IFR_STAGE_WORDS 192
#define IFR_STAGE_DEPTH 128
#define IFR_STAGE_PRESENT 160

// ---- big triangles: setups precomputed by the fill pass arrive by TMA bulk copy (double buffered); every thread
//      tests its own pixel against broadcast LDS.128 words, warps run independently.
// HIZ (deep queues, ordered compare ops): hierarchical early-Z one level above the pixel.  wfar = the worst (largest)
// key depth word any sample of this warp's 8x4 block currently holds; a triangle whose most optimistic key inside the
// tile (TriS pad[1], computed by setup_hiz on the staged entries) lies above it cannot win a single sample of the
// block and is skipped by the whole warp.  One level further up, entries that cannot win anywhere in the tile
// (bound above every warp's wfar) are dropped while the chunk is prepared: the warps only walk the compacted
// survivor list.  Results are unchanged (the reference has no Hi-Z: curTileHierZ is an unused field,
// TileRasterWorker.h:80).
struct HizShared { uint32_t warpFar[8]; uint32_t nSurv[2]; uint8_t surv[2][32]; };

template <int S>
__device__ __forceinline__ uint32_t block_far(const unsigned long long* keys, int tid, unsigned long long best, bool uni) {
    uint32_t mine;
    if (S == 1 || uni) mine = (uint32_t)(best >> 32);
    else {
        mine = 0;
#pragma unroll
        for (int sm = 0; sm < S; sm++) mine = max(mine, (uint32_t)(keys[sm * IFR_TILE_PIX + tid] >> 32));
    }
    return __reduce_max_sync(0xffffffffu, mine);
}

// one big-queue entry against this thread's pixel; returns true if a key of the pixel changed
template <int S, bool HIZ>
__device__ __forceinline__ bool big_test(const DrawParams& p, const KeyMode& km, unsigned long long* keys, int* covScratch, const int4* t4, uint32_t cz,
                                         int tid, int wrp, int lx, int ly, bool scis, float d0, unsigned long long& best, bool& uni) {
    const int ci = 16 * lx, cj = 16 * ly;
    const int4 A = t4[0];
    const int4 B = t4[1], C = t4[2];
    const int dX[3] = {A.x, A.y, A.z}, dY[3] = {B.x, B.y, B.z}, T[3] = {C.x, C.y, C.z};
    const uint32_t flags = (uint32_t)C.w;
    uint32_t m = 0;
    bool upd = false;
    if (HIZ && ((cz >> (8 + wrp)) & 1)) {
        if (scis) m = (1u << S) - 1u;          // trivial accept: every sample of the block
    } else if (!(flags & IFR_FLAG_WIDE)) {
        if (S > 1) {             // every branch up to here is warp-uniform (queue entry, block mask, Hi-Z): all 32 lanes arrive
            m = coverage_coop<S>(p, covScratch, dX, dY, T, ci, cj, tid & 31);
            if (!scis) m = 0;
        } else if (scis) m = coverage<S>(p, dX, dY, T, nullptr, 0, ci, cj);
    }
    if (m || (flags & IFR_FLAG_WIDE)) {
        const int4 D = t4[3], E = t4[4];
        const long long E0[3] = {(long long)(((unsigned long long)(uint32_t)D.y << 32) | (uint32_t)D.x),
                                 (long long)(((unsigned long long)(uint32_t)D.w << 32) | (uint32_t)D.z),
                                 (long long)(((unsigned long long)(uint32_t)E.y << 32) | (uint32_t)E.x)};
        if (!m && scis) m = coverage<S>(p, dX, dY, T, E0, 1, ci, cj);
        if (m) {
            const float2 Zr = *reinterpret_cast<const float2*>(t4 + 5);
            const float d = depth_at(dX, dY, E0, flags, __int_as_float(E.z), __int_as_float(E.w), Zr.x, Zr.y, lx, ly);
            unsigned long long k;
            if (key_make(km, d, d0, (uint32_t)B.w, k)) {
                if (S == 1) {
                    if (k < best) { best = k; upd = true; }
                } else if (uni) {
                    // every sample of the pixel still holds the same key (`best`, in a register: pixels inside one
                    // triangle never touch shared memory); the first partially covering winner materialises it
                    if (k < best) {
                        upd = true;
                        if (m == (1u << S) - 1u) best = k;
                        else {
#pragma unroll
                            for (int sm = 0; sm < S; sm++) keys[sm * IFR_TILE_PIX + tid] = ((m >> sm) & 1) ? k : best;
                            uni = false;
                        }
                    }
                } else {
#pragma unroll
                    for (int sm = 0; sm < S; sm++)
                        if (((m >> sm) & 1) && k < keys[sm * IFR_TILE_PIX + tid]) { keys[sm * IFR_TILE_PIX + tid] = k; upd = true; }
                }
            }
        }
    }
    return upd;
}

template <int S, bool HIZ>
__device__ __forceinline__ void big_loop(const DrawParams& p, const KeyMode& km, unsigned long long* keys, int* covScratch, TriS* bigBuf,
                                         uint64_t* mbar, HizShared& hz, const TriS* bigSrc, uint32_t nBig, int tid, int wrp, int lx,
                                         int ly, bool scis, float d0, unsigned long long& best, bool& uni) {
    const int lane = tid & 31;
    uint32_t wfar = 0xffffffffu;
    if (HIZ) {
        wfar = block_far<S>(keys, tid, best, uni);
        // published for the chunk preparation of warps 0 / 1, which may run concurrently: a stale value is larger, hence
        // conservative; the bound only ever shrinks, so the publication is an atomicMin and the readers use atomics too
        // (compute-sanitizer racecheck: no hazards)
        if (lane == 0) atomicMin(&hz.warpFar[wrp], wfar);
    }
    const uint32_t CH = (uint32_t)p.bigChunk;
    const uint32_t nchunks = (nBig + CH - 1) / CH;
    for (uint32_t c = 0; c < nchunks; c++) {
        if (tid == 0 && c + 1 < nchunks) {
            const uint32_t cnt1 = min(nBig - (c + 1) * CH, CH);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load(bigBuf + ((c + 1) & 1) * CH, bigSrc + (size_t)(c + 1) * CH, cnt1 * (uint32_t)sizeof(TriS), &mbar[(c + 1) & 1]);
        }
        mbar_wait(&mbar[c & 1], (c >> 1) & 1);
        const uint32_t cnt = min(nBig - c * CH, CH);
        const int4* chunk = reinterpret_cast<const int4*>(bigBuf + (c & 1) * CH);
        if (!HIZ) {
            const int4* t4 = chunk;
            for (uint32_t b = 0; b < cnt; b++, t4 += 6) {
                // warp-uniform coarse reject: the fill pass rasterised the triangle against the tile's 8x4 blocks
                const uint32_t cz = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(t4) + 88);
                if (!((cz >> wrp) & 1)) continue;
                big_test<S, false>(p, km, keys, covScratch, t4, cz, tid, wrp, lx, ly, scis, d0, best, uni);
            }
        } else {
            // chunk preparation, warps 0 and 1, one thread per staged entry: trivial-accept blocks + key bound written
            // back into the staged copy, entries that can still win somewhere in the tile compacted per warp
            if ((uint32_t)tid < CH) {
                bool alive = false;
                if ((uint32_t)tid < cnt) {
                    uint32_t tileFar = 0;
#pragma unroll
                    for (int w = 0; w < 8; w++) tileFar = max(tileFar, atomicOr(&hz.warpFar[w], 0u));
                    TriS* e = bigBuf + (c & 1) * CH + tid;
                    const uint32_t tb = setup_hiz(e, p.sampMinX, p.sampMaxX, p.sampMinY, p.sampMaxY, km.xhi);
                    alive = ((uint32_t)e->pad[0] & 0xffu) != 0 && tb <= tileFar;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, alive);
                if (alive) hz.surv[wrp][__popc(bal & ((1u << lane) - 1u))] = (uint8_t)tid;
                if (lane == 0) hz.nSurv[wrp] = __popc(bal);
            }
            __syncthreads();
            for (uint32_t half = 0; half < CH / 32; half++) {
                const uint32_t ns = hz.nSurv[half];
                for (uint32_t i = 0; i < ns; i++) {
                    const int4* t4 = chunk + 6 * (int)hz.surv[half][i];
                    const uint2 cz = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(t4) + 88);
                    if (!((cz.x >> wrp) & 1)) continue;
                    if (cz.y > wfar) continue;          // warp-uniform Hi-Z reject
                    const bool upd = big_test<S, true>(p, km, keys, covScratch, t4, cz.x, tid, wrp, lx, ly, scis, d0, best, uni);
                    if (__any_sync(0xffffffffu, upd)) {  // warp-convergent: every reject above is warp-uniform
                        wfar = block_far<S>(keys, tid, best, uni);
                        if (lane == 0) atomicMin(&hz.warpFar[wrp], wfar);
                    }
                }
            }
        }
        // the staged buffer (and, with HIZ, the survivor lists) are rewritten by the next iteration's preparation / the
        // bulk copy issued at the top of iteration c + 1
        if (HIZ ? (c + 1 < nchunks) : (c + 2 < nchunks)) __syncthreads();
    }
}

// A tile with nothing binned: the fused clear is all there is to do (imageResetFloat32* CU:3337/3363).  Thread = pixel,
// row-major inside the tile: a warp writes two full 256-byte colour rows.
template <int S>
__device__ __forceinline__ void clear_tile(const DrawParams& p, int tileX, int tileY, int tid) {
    const int px = tileX * IFR_TILE + (tid & 15), py = tileY * IFR_TILE + (tid >> 4);
    if (px >= p.W || py >= p.H) return;
    const size_t pidx = (size_t)py * p.W + px;
    const bool scis = scissor_pass(p, px, py);
    if (p.primIds) {      // per-draw debug ids: nothing of this draw lands here, cleared or not
#pragma unroll
        for (int s = 0; s < S; s++) p.primIds[pidx * S + s] = -1;
    }
    if (!p.doClear || !scis) {
        if (p.present) p.present[pidx] = ifr_pack_rgba8(p.color[pidx]);   // untouched pixel: present what is there
        if (!p.doClear) return;
    }
    if (scis) {
        const float4 cc = make_float4(p.clearColor[0], p.clearColor[1], p.clearColor[2], p.clearColor[3]);
        if (p.skipEmptyStores != 1 || p.scissorEnable) p.color[pidx] = cc;
        p.depth[pidx] = p.clearDepth;
        if (p.present) p.present[pidx] = ifr_pack_rgba8(cc);
    }
}

// The same for a group of empty tiles through TMA: the CTA builds one cleared tile in shared memory, one thread per tile
// stores it (2 - 3 tensor copies instead of 256 threads x 2 - 3 stores).  Whole-frame clears without scissors only.
template <int S>
__device__ __forceinline__ void clear_tiles_tma(const DrawParams& p, float* stage, uint32_t e0, uint32_t e1, int ntiles, int tid) {
    const float4 cc = make_float4(p.clearColor[0], p.clearColor[1], p.clearColor[2], p.clearColor[3]);
    reinterpret_cast<float4*>(stage)[tid] = cc;
    stage[1024 + tid] = p.clearDepth;
    reinterpret_cast<uint32_t*>(stage)[1280 + tid] = ifr_pack_rgba8(cc);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (e0 + (uint32_t)tid < e1) {
        const uint32_t t = p.tileOrder[ntiles - 1 - (int)(e0 + tid)];
        const int x0 = (int)(t & 0xffffu) * IFR_TILE, y0 = (int)(t >> 16) * IFR_TILE;
        if (p.skipEmptyStores != 1) tma_store_2d(&p.tmap[3], x0 * 4, y0, stage);     // 1: the frame's owner has cleared it already
        tma_store_2d(&p.tmap[4], x0, y0, stage + 1024);
        if (p.present) tma_store_2d(&p.tmap[5], x0, y0, stage + 1280);
        tma_store_commit_wait();
    }
    if (p.primIds) {
        for (uint32_t e = e0; e < e1; e++) {
            const uint32_t t = p.tileOrder[ntiles - 1 - (int)e];
            const int px = (int)(t & 0xffffu) * IFR_TILE + (tid & 15), py = (int)(t >> 16) * IFR_TILE + (tid >> 4);
            if (px >= p.W || py >= p.H) continue;
#pragma unroll
            for (int sm = 0; sm < S; sm++) p.primIds[((size_t)py * p.W + px) * S + sm] = -1;
        }
    }
}

#ifndef IFR_MINB1
#define IFR_MINB1 6
#endif
// resident CTAs per SM the register budget is cut for.  Measured (profiles/r01d_ab.txt): the 1x textured kernels are
// latency bound on the texture fetches — 6 CTAs (40 registers, some spills) beat 4 by 4 %; the plain-varying 1x
// kernels (small-triangle soup) lose 20 % with more than 4; the multisampled variants keep 64 registers.
#ifndef IFR_MINB8
#define IFR_MINB8 4
#endif
template <int FS, int S>
__global__ void __launch_bounds__(256, FsTraits<FS>::needsDerivs ? (S == 1 ? IFR_MINB1 : IFR_MINB8) : 4) tile_kernel(const __grid_constant__ DrawParams p) {
    extern __shared__ __align__(128) unsigned long long keys[];        // [S][256], index = owning thread; then the big-queue staging
    TriS* const bigBuf = reinterpret_cast<TriS*>(keys + S * IFR_TILE_PIX);   // [2][p.bigChunk]
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ HizShared hz;
    __shared__ __align__(16) int covOff[8][3 * (S > 1 ? S : 1)];   // coverage_coop: one row per warp
#if IFR_TMA_EPILOGUE
    __shared__ __align__(128) float stage[8][IFR_STAGE_WORDS];     // TMA staging, one row per warp (see IFR_STAGE_WORDS)
    __shared__ __align__(8) uint64_t stageBar[8];
#endif
    constexpr int NV = FsTraits<FS>::nv;

    // Work list (scan pass): CTA b < nWork takes a tile with a non-empty queue; the CTAs behind them share the owned tiles
    // with nothing binned, IFR_EMPTY_PER_CTA each (fused clear only); tiles of other ranks appear in neither list.
    const int ntiles = p.tilesX * p.tilesY;
    const int tid = threadIdx.x;
    const ulonglong2 hdr = *reinterpret_cast<const ulonglong2*>(p.counters + 2);   // {nWork, overflow flag}
    const uint32_t nWork = (uint32_t)hdr.x;
    const bool overflow = hdr.y != 0;
    // the grid may come from an earlier draw of the same inputs (DrawParams::gridHint): tell the host if it no longer fits
    if (blockIdx.x == gridDim.x - 1 && tid == 0 &&
        nWork + ((uint32_t)p.counters[7] + IFR_EMPTY_PER_CTA - 1) / IFR_EMPTY_PER_CTA > gridDim.x) p.counters[12] = 1ull;
    if (blockIdx.x >= nWork) {
        const uint32_t e0 = (blockIdx.x - nWork) * IFR_EMPTY_PER_CTA;
        const uint32_t nEmpty = (uint32_t)p.counters[7];

        if (p.useTma && p.doClear && !p.scissorEnable) {
            // staged in the big-queue buffer, which a CTA without a tile never fills (6 KB of its 12)
            if (e0 < nEmpty) clear_tiles_tma<S>(p, reinterpret_cast<float*>(bigBuf), e0, min(e0 + IFR_EMPTY_PER_CTA, nEmpty), ntiles, tid);
            return;
        }
        for (uint32_t e = e0; e < min(e0 + IFR_EMPTY_PER_CTA, nEmpty); e++) {
            const uint32_t t = p.tileOrder[ntiles - 1 - (int)e];
            clear_tile<S>(p, (int)(t & 0xffffu), (int)(t >> 16), tid);
        }
        return;
    }
    const uint32_t torder = p.tileOrder[blockIdx.x];
    const int tileX = (int)(torder & 0xffffu), tileY = (int)(torder >> 16);
    const int tile = tileY * p.tilesX + tileX;
    if (overflow) {    // a queue is too small: the host re-issues the draw with larger ones; until then the frame is cleared only
        clear_tile<S>(p, tileX, tileY, tid);
        return;
    }
    // a warp owns one 8x4-pixel block of the tile (block = warp: x = (w & 1) * 8, y = (w >> 1) * 4), a quad = 4 lanes
    const int wrp = tid >> 5, q = (tid >> 2) & 7, l = tid & 3;
    // under the 40-register budget of the 1x kernels the compiler re-derives lx / ly from the thread id (a dozen bit
    // operations) at every use inside the hot loops; an opaque packed copy makes it keep (or at worst reload) the value
    int lxy = ((wrp & 1) * 8 + (q & 3) * 2 + (l & 1)) | (((wrp >> 1) * 4 + (q >> 2) * 2 + (l >> 1)) << 16);
    asm volatile("" : "+r"(lxy));
    const int lx = lxy & 0xffff, ly = lxy >> 16;
    const int px = tileX * IFR_TILE + lx, py = tileY * IFR_TILE + ly;
    const bool inFrame = px < p.W && py < p.H;
    const bool scis = inFrame && scissor_pass(p, px, py);
    const uint32_t sStart = p.tileStart[tile], bStart = p.tileStart[ntiles + 1 + tile];
    const uint32_t nSmall = p.tileStart[tile + 1] - sStart;
    const uint32_t nBig = p.tileStart[ntiles + 1 + tile + 1] - bStart;
    const size_t pidx = (size_t)py * p.W + px;

    KeyMode km;     // precomputed by the host from p.depthFunc (same table as key_mode)
    km.xhi = p.keyXhi; km.xlo = p.keyXlo; km.ordered = p.keyOrdered != 0; km.equal = p.keyEqual != 0; km.never = p.keyNever != 0;
#if IFR_TMA_EPILOGUE
    // TMA path: the warp's 8x4 block leaves through one staged tensor store per image (epilogue).  A block that is not
    // cleared as a whole first fetches its previous depth and colour the same way, so that the store can always write
    // the whole block (pixels the draw leaves alone write back what was there; the tile is owned by this CTA).
    float* const stg = stage[wrp];
    const int bi = (ly & 3) * 8 + (lx & 7);
    const int bx0 = tileX * IFR_TILE + (wrp & 1) * 8, by0 = tileY * IFR_TILE + (wrp >> 1) * 4;
    bool prevLoaded = false;
    if (p.useTma) {
        prevLoaded = !__all_sync(0xffffffffu, !inFrame || (p.doClear && scis));
        if (prevLoaded) {
            if ((tid & 31) == 0) {
                mbar_init(&stageBar[wrp], 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&stageBar[wrp])), "r"(640u) : "memory");
                tma_load_2d(stg, &p.tmap[0], bx0 * 4, by0, &stageBar[wrp]);
                tma_load_2d(stg + IFR_STAGE_DEPTH, &p.tmap[1], bx0, by0, &stageBar[wrp]);
            }
            __syncwarp();
            mbar_wait(&stageBar[wrp], 0);
        }
    }
#else
    constexpr bool prevLoaded = false;
    const float* const stg = nullptr;
    constexpr int bi = 0;
#endif
    float d0 = 0.0f;
    if (inFrame) d0 = (p.doClear && scis) ? p.clearDepth : (prevLoaded ? stg[IFR_STAGE_DEPTH + bi] : p.depth[pidx]);
    const unsigned long long k0 = key_init(km, d0);
    // Per-sample visibility keys.  S == 1: the pixel's key lives in a register (`best`); shared memory is only the
    // mailbox through which small triangles (processed by other threads) reach this pixel, filled from `best`
    // once the big triangles are done.
    // S > 1: while every sample of the pixel holds the same key it lives in `best` too (`uni`); the per-sample keys in
    // shared memory are only written once a triangle covers the pixel partially (or a small-triangle pass follows).
    unsigned long long best = k0;
    bool uni = true;
    const bool deep = km.ordered && nBig >= IFR_HIZ_MIN_QUEUE;

    // ---- big triangles first: their result gives the small-triangle pass a far bound per 8x4 block ----
    if (nBig) {
        if (tid == 0) {
            mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (tid < 8) hz.warpFar[tid] = 0xffffffffu;
        __syncthreads();      // mbarriers initialised
        const TriS* bigSrc = p.bigQueue + bStart;
        int* covScratch = covOff[wrp];
        if (tid == 0) bulk_load(bigBuf, bigSrc, min(nBig, (uint32_t)p.bigChunk) * (uint32_t)sizeof(TriS), &mbar[0]);
#ifndef IFR_NO_HIZ
        if (deep) big_loop<S, true>(p, km, keys, covScratch, bigBuf, mbar, hz, bigSrc, nBig, tid, wrp, lx, ly, scis, d0, best, uni);
        else
#endif
            big_loop<S, false>(p, km, keys, covScratch, bigBuf, mbar, hz, bigSrc, nBig, tid, wrp, lx, ly, scis, d0, best, uni);
    }

    // ---- small triangles: one thread per (triangle, tile) pair, shared-memory atomicMin --------------
    if (nSmall) {
        if (S == 1) keys[tid] = best;
        else if (uni) {
#pragma unroll
            for (int s = 0; s < S; s++) keys[s * IFR_TILE_PIX + tid] = best;
            uni = false;
        }
        // per-pixel early-Z against the block's far bound (deep tiles: most small triangles are hidden behind the big
        // ones): depth first, coverage only for pixels that can still win a sample
        const bool useFar = deep;
        if (useFar) {
            const uint32_t f = block_far<S>(keys, tid, best, uni);
            if ((tid & 31) == 0) hz.warpFar[wrp] = f;
        }
        __syncthreads();      // keys complete, far bounds published
        for (uint32_t qi = tid; qi < nSmall; qi += IFR_CHUNK) {
            TriS s;
            int bx0, bx1, by0, by1;
            const uint32_t slot = __ldg(p.queue + sStart + qi);
            if (!setup_tri(p, slot, tileX, tileY, s, bx0, bx1, by0, by1)) continue;
            const int wide = s.flags & IFR_FLAG_WIDE;
            for (int y = by0; y <= by1; y++)
                for (int x = bx0; x <= bx1; x++) {
                    unsigned long long k = 0;
                    if (useFar) {
                        const float d = depth_at(s.dX, s.dY, s.E0, s.flags, s.ar, s.z0, s.z1, s.z2, x, y);
                        key_make(km, d, 0.0f, slot, k);      // ordered compare op: always a key
                        if ((uint32_t)(k >> 32) > hz.warpFar[(y >> 2) * 2 + (x >> 3)]) continue;
                    }
                    const uint32_t m = coverage<S>(p, s.dX, s.dY, s.T, s.E0, wide, 16 * x, 16 * y);
                    if (!m) continue;
                    const int gx = tileX * IFR_TILE + x, gy = tileY * IFR_TILE + y;
                    if (p.scissorEnable && !scissor_pass(p, gx, gy)) continue;
                    if (!useFar) {
                        const float d = depth_at(s.dX, s.dY, s.E0, s.flags, s.ar, s.z0, s.z1, s.z2, x, y);
                        float dd0 = 0.0f;
                        if (km.equal)
                            dd0 = p.doClear ? p.clearDepth : p.depth[(size_t)gy * p.W + gx];
                        if (!key_make(km, d, dd0, slot, k)) continue;
                    }
                    const int pi = pix_index(x, y);
#pragma unroll
                    for (int sm = 0; sm < S; sm++)
                        if ((m >> sm) & 1) key_min_shared(&keys[sm * IFR_TILE_PIX + pi], k);
                }
        }
        __syncthreads();
    }

    // ---- quad shading + resolve -----------------------------------------------------------------
    // A pixel whose samples all hold the same key (S == 1, `uni`, or found equal below) needs no search: one FS
    // invocation with count S.  Otherwise the per-sample slots are re-read from the shared-memory keys by every search
    // (edge pixels only), so that nothing but the accumulators stays live across the fragment shader.
    constexpr bool PLANES = FsPlanes<FS>::value;    // FAST textured shaders: varyings from attribute planes, depth from the key
    bool same = true;
    uint32_t slot0;
    uint32_t dword0 = 0;      // PLANES: depth word of slot0's key
    int todo = 0;             // samples that still wait for their fragment
    if (S == 1 || uni) {
        const unsigned long long k = (S == 1 && nSmall != 0) ? keys[tid] : best;
        slot0 = key_slot(km, k);
        if (PLANES) dword0 = km.ordered ? (uint32_t)(k >> 32) : (uint32_t)k;
        todo = slot0 != IFR_SLOT_NONE ? S : 0;
    } else {
        // the pixel's keys are final: replace each by its slot (low word, in place) so that the searches below are
        // 32-bit loads and compares; PLANES: the other word keeps (ordered compare ops) or receives the key's depth word
        uint32_t* kw = reinterpret_cast<uint32_t*>(keys);
        {
            const unsigned long long k = keys[tid];
            slot0 = key_slot(km, k);
            if (PLANES) { dword0 = km.ordered ? (uint32_t)(k >> 32) : (uint32_t)k; if (!km.ordered) kw[2 * tid + 1] = dword0; }
        }
        kw[2 * tid] = slot0;
        todo = slot0 != IFR_SLOT_NONE;
#pragma unroll
        for (int s = 1; s < S; s++) {
            const unsigned long long k = keys[s * IFR_TILE_PIX + tid];
            const uint32_t v = key_slot(km, k);
            if (PLANES && !km.ordered) kw[2 * (s * IFR_TILE_PIX + tid) + 1] = (uint32_t)k;
            kw[2 * (s * IFR_TILE_PIX + tid)] = v;
            same = same && v == slot0;
            todo += v != IFR_SLOT_NONE;
        }
    }
    if (!inFrame) todo = 0;
    if (p.primIds && inFrame) {
#pragma unroll
        for (int s = 0; s < S; s++) p.primIds[pidx * S + s] = -1;     // overwritten below for the covered samples
    }

    float4 acc = make_float4(0, 0, 0, 0);
    float dacc = 0.0f;
    int bg = S;
    unsigned int myFrags = 0;
    // One FS invocation per (pixel, slot) that owns >= 1 sample, slots in ascending order (CU:2264-2310).  A lane
    // shades alone (see interp), so the loop runs max-over-lanes(distinct slots) times and nothing is quad-uniform.
    // (Tried and measured slower.  Two tiles per CTA with both queue headers and both first chunks fetched up front:
    // c2 tile 0.280 -> 0.311 ms, c3 0.441 -> 0.506 — the prologue latency of a fresh CTA is already covered by the other
    // resident CTAs.  c3 tile 0.44 -> 0.57 ms: compacting the second / third slots of edge pixels into a
    // shared work list shaded one per thread — the kernel is latency bound, and the two CTA barriers plus one warp
    // shading the list while seven wait cost more than the divergent second pass they replace.)
    uint32_t lastp1 = 0;      // last shaded slot + 1
    while (todo > 0) {
        uint32_t cur = slot0;
        uint32_t dword = dword0;
        int cnt = S;
        if (S > 1 && !same) {
            // smallest slot above `last` and its sample count: slots are compared as (slot - (last + 1)) mod 2^32, which
            // sends the slots already shaded to the top of the range, just below the empty marker (0xffffffff - ...)
            const uint32_t* kw = reinterpret_cast<const uint32_t*>(keys);
            uint32_t c[S], m = 0xffffffffu;
#pragma unroll
            for (int s = 0; s < S; s++) { c[s] = kw[2 * (s * IFR_TILE_PIX + tid)] - lastp1; m = min(m, c[s]); }
            cnt = 0;
            int first = 0;
#pragma unroll
            for (int s = S - 1; s >= 0; s--) { cnt += c[s] == m; if (PLANES) first = c[s] == m ? s : first; }
            cur = m + lastp1;
            if (PLANES) dword = kw[2 * (first * IFR_TILE_PIX + tid) + 1];     // all samples of one slot carry the pixel's depth
        }
        float4 v[NV > 0 ? NV : 1], ddx[NV > 0 ? NV : 1], ddy[NV > 0 ? NV : 1];
        float depth = 0.0f;
        int prim = 0, fan = 0;
        if (PLANES) {
            interp_planes<FsTraits<FS>::needsDerivs>(p, cur, px, py, v, ddx, ddy);
            depth = depth_from_word(km, dword);
            fan = cur >= (uint32_t)p.ntri;
            prim = fan ? (int)(cur - (uint32_t)p.ntri) : (int)cur;
        } else
            interp<NV, FsTraits<FS>::needsDerivs, FS == IFRIT_FS_TEXTURED_FAST || FS == IFRIT_FS_TEXTURED_LOD0_FAST || FS == IFRIT_FS_TEXTURED_FAST_RA>(p, cur, px, py, v, ddx, ddy, depth, prim, fan);
        const float4 c = run_fs<FS>(p, v, ddx, ddy, prim, fan);
        const float fc = (float)cnt;
        acc.x = fmaf(fc, c.x, acc.x); acc.y = fmaf(fc, c.y, acc.y);
        acc.z = fmaf(fc, c.z, acc.z); acc.w = fmaf(fc, c.w, acc.w);
        dacc = fmaf(fc, depth, dacc);
        bg -= cnt;
        todo -= cnt;
        myFrags++;
        if (p.primIds) {
            if (S == 1 || same) {
#pragma unroll
                for (int s = 0; s < S; s++) p.primIds[pidx * S + s] = prim;
            } else {
#pragma unroll
                for (int s = 0; s < S; s++)
                    if (reinterpret_cast<const uint32_t*>(keys)[2 * (s * IFR_TILE_PIX + tid)] == cur) p.primIds[pidx * S + s] = prim;
            }
        }
        lastp1 = cur + 1u;
    }
    {
        const bool cleared = p.doClear && scis;
        const bool written = inFrame && (bg != S || cleared);
        float4 out = acc;
        float dout = dacc;
        if (written) {
            if (bg > 0) {
                const float fb = (float)bg;
                const float4 prev = cleared ? make_float4(p.clearColor[0], p.clearColor[1], p.clearColor[2], p.clearColor[3])
                                            : (prevLoaded ? reinterpret_cast<const float4*>(stg)[bi] : p.color[pidx]);
                out.x = fmaf(fb, prev.x, out.x); out.y = fmaf(fb, prev.y, out.y);
                out.z = fmaf(fb, prev.z, out.z); out.w = fmaf(fb, prev.w, out.w);
                dout = fmaf(fb, d0, dout);
            }
            if (bg == S) {   // nothing drawn here: plain clear
                out = make_float4(p.clearColor[0], p.clearColor[1], p.clearColor[2], p.clearColor[3]);
                dout = p.clearDepth;
            } else if (S > 1) {
                const float fs = (float)S;
                out.x = out.x / fs; out.y = out.y / fs; out.z = out.z / fs; out.w = out.w / fs;
                dout = dout / fs;
            }
        }
#if IFR_TMA_EPILOGUE
        if (p.useTma) {
            // every pixel of the block is either written or holds its previous value (prevLoaded: a block with a pixel
            // that is neither cleared nor off the frame has been fetched); pixels off the frame are clipped by the unit
            if (!written) {
                if (prevLoaded) { out = reinterpret_cast<const float4*>(stg)[bi]; dout = stg[IFR_STAGE_DEPTH + bi]; }
                else { out = make_float4(0, 0, 0, 0); dout = 0.0f; }
            }
            __syncwarp();        // previous contents read by every lane before the row is overwritten
            reinterpret_cast<float4*>(stg)[bi] = out;
            stg[IFR_STAGE_DEPTH + bi] = dout;
            if (p.present) reinterpret_cast<uint32_t*>(stg)[IFR_STAGE_PRESENT + bi] = ifr_pack_rgba8(out);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if ((tid & 31) == 0) {
                tma_store_2d(&p.tmap[0], bx0 * 4, by0, stg);
                tma_store_2d(&p.tmap[1], bx0, by0, stg + IFR_STAGE_DEPTH);
                if (p.present) tma_store_2d(&p.tmap[2], bx0, by0, stg + IFR_STAGE_PRESENT);
                tma_store_commit_wait();
            }
        } else
#endif
        if (written) {
            // write-once data: streaming (evict-first) stores keep the framebuffer from pushing the texture mip chain and the
            // triangle records out of L2 (c3: 166 MB of framebuffer per frame against a 126 MB L2)
            IFR_FB_STORE(&p.color[pidx], out);
            IFR_FB_STORE(&p.depth[pidx], dout);
            if (p.present) IFR_FB_STORE(&p.present[pidx], ifr_pack_rgba8(out));
        } else if (inFrame && p.present) {
            p.present[pidx] = ifr_pack_rgba8(p.color[pidx]);   // untouched pixel (no clear, nothing drawn): present what is there
        }
    }
    myFrags = __reduce_add_sync(0xffffffffu, myFrags);
    if ((tid & 31) == 0 && myFrags)   // no CTA barrier: warps retire independently; spread over 256 sectors (common.cuh)
        ifr_count_spread(p.counters, 1, (unsigned)(tile * 8 + wrp) * 2654435761u >> 16, (unsigned long long)myFrags);
}


#ifndef __CUDACC_RTC__
// dynamic shared memory of the tile kernel: keys [msaa][256] + staging [2][bigChunk]
static inline size_t ifr_tile_smem_bytes(const DrawParams& p) {
    return (size_t)p.msaa * IFR_TILE_PIX * sizeof(unsigned long long) + 2 * (size_t)p.bigChunk * sizeof(TriS);
}
template <int FS>
static inline cudaError_t launch_fs(const DrawParams& p, cudaStream_t s) {
    // one CTA per tile is the upper bound of the work list (surplus CTAs exit at once); the exact count when the host knows it
    const dim3 grid(p.gridHint ? p.gridHint : (uint32_t)(p.tilesX * p.tilesY));
    const size_t smem = ifr_tile_smem_bytes(p);
#define IFR_LAUNCH(SS)                                                                                   \
    case SS: {                                                                                           \
        static bool attrSet = false;                                                                     \
        if (!attrSet) {                                                                                  \
            cudaFuncSetAttribute(tile_kernel<FS, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); \
            attrSet = true;                                                                              \
        }                                                                                                \
        tile_kernel<FS, SS><<<grid, 256, smem, s>>>(p);                                                  \
        break;                                                                                           \
    }
    switch (p.msaa) {
        IFR_LAUNCH(1)
        IFR_LAUNCH(2)
        IFR_LAUNCH(4)
        IFR_LAUNCH(8)
        IFR_LAUNCH(16)
        default: return cudaErrorInvalidValue;
    }
#undef IFR_LAUNCH
    return cudaGetLastError();
}
#endif   // !__CUDACC_RTC__
