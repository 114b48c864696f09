// sampler.cuh — device texture sampling library + fragment-shader registry.
//
// Behaviour restated from include/ifrit/softgraphics/engine/math/ShaderBuiltinCuda.cuh ("SB"):
//   textureImplLegacy SB:106-182 (address modes, nearest / bilinear), textureImplLod SB:210-218
//   (trilinear), textureImplAdaptiveLod SB:233-243 (incl. the dx.y-for-dy.y slip at SB:237),
//   textureImplAnisotropicFilter SB:244-272, textureCubeCoordConversion SB:184-208,
//   textureCubeImplLod SB:220-231.  Texture memory layout SB:97-105: RGBA32F, level-major, layers
//   inside each level, row-major.  Deviations (documented in DESIGN.md §3.6): the sampled level is
//   clamped to the allocated chain and linear-filter texel indices are clamped (the reference reads
//   out of bounds there).  Op order matches oracle/ifrit_oracle.c; log2f is the only libm call whose
//   device / host results may differ in the last ulp.
#pragma once
#include "common.cuh"
#include "../../include/ifrit_b200.h"

__device__ __forceinline__ float wrapf(float c, float s, float invS) {
    const float q = floorf(c * invS);
    float r = fmaf(-q, s, c);
    if (r < 0.0f) r += s;
    if (r >= s) r -= s;
    return r;
}

__device__ __forceinline__ float addr_mode(int mode, float c, float s, float invS, float inv2S, bool& border) {
    switch (mode) {
        case IF_SAMPLER_ADDRESS_MODE_REPEAT:
            c = wrapf(c, s, invS);
            break;
        case IF_SAMPLER_ADDRESS_MODE_MIRRORED_REPEAT:
            c = wrapf(c, 2 * s, inv2S);
            c = fminf(c, 2 * s - c);
            break;
        case IF_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE:
            c = (c >= s - 1.0f) ? s - 1.0f : ((c <= 0.0f) ? 0.0f : c);
            break;
        case IF_SAMPLER_ADDRESS_MODE_CLAMP_TO_BORDER:
            if (c >= s || c < 0.0f) border = true;
            break;
        case IF_SAMPLER_ADDRESS_MODE_MIRROR_CLAMP_TO_EDGE:
            c = (c >= s - 1.0f) ? 0.0f : ((c <= 0.0f) ? s - 1.0f : c);
            break;
    }
    return c;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// single-instruction SFU forms, used by the FAST sampler variant only (colour bar: 1/255 on >= 99.9 % of pixels)
__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_fast(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ float4 lerp4(const float4& a, const float4& b, float t) {
    return make_float4(fmaf(b.x - a.x, t, a.x), fmaf(b.y - a.y, t, a.y), fmaf(b.z - a.z, t, a.z),
                       fmaf(b.w - a.w, t, a.w));
}

// Per-level sampling context: everything that depends on the level only, prepared once per
// texture() call so that the taps of an anisotropic footprint share it (SB:97-105 walks the level
// table on every fetch).
struct LevelCtx { const float4* base; unsigned long long hw; int w, h; float fw, fh, fw1, fh1, invW, invH, inv2W, inv2H; };

__device__ __forceinline__ LevelCtx level_ctx(const TexDesc& t, int layer, int lod, int hwFilter = 0) {
    lod = clampi(lod, 0, t.mips);
    const TexLevel& L = t.lvl[lod];
    LevelCtx c;
    c.w = L.w; c.h = L.h;
    c.fw = (float)c.w; c.fh = (float)c.h; c.fw1 = (float)(c.w - 1); c.fh1 = (float)(c.h - 1);
    c.invW = L.invW; c.invH = L.invH; c.inv2W = L.inv2W; c.inv2H = L.inv2H;
    c.base = t.data + L.off + (size_t)L.layerStride * layer;
    c.hw = (hwFilter && layer == 0) ? t.hw[lod] : 0ull;
    return c;
}

__device__ __forceinline__ float4 sample_level(const LevelCtx& L, const SampDesc& s, float u, float v, int offx, int offy) {
    const int w = L.w, h = L.h;
    float cx = fmaf(u, L.fw1, (float)offx);
    float cy = fmaf(v, L.fh1, (float)offy);
    bool border = false;
    cx = addr_mode(s.addrU, cx, L.fw, L.invW, L.inv2W, border);
    if (!border) cy = addr_mode(s.addrV, cy, L.fh, L.invH, L.inv2H, border);
    if (border) {
        const float b = s.border == 0 ? 0.0f : 1.0f;
        return make_float4(b, b, b, 1.0f);
    }
    const float4* base = L.base;
    if (s.filter == IF_FILTER_NEAREST) {
        int px = (int)roundf(cx), py = (int)roundf(cy);
        px = clampi(px % w, 0, w - 1);
        py = clampi(py % h, 0, h - 1);
        return __ldg(base + py * w + px);
    }
    if (L.hw && cx >= 0.0f && cy >= 0.0f) {
        // the same 2x2 footprint and weights (texel = floor(c), neighbour clamped to the edge), fetched and
        // blended by the texture unit: one TEX instead of 4 loads + 12 lerps; weights are quantised to 8 bits
        return tex2D<float4>((cudaTextureObject_t)L.hw, cx + 0.5f, cy + 0.5f);
    }
    const int px = clampi((int)cx, 0, w - 1), py = clampi((int)cy, 0, h - 1);
    const float fx = cx - (float)px, fy = cy - (float)py;
    const int px1 = min(px + 1, w - 1), py1 = min(py + 1, h - 1);
    const float4* r0 = base + py * w;
    const float4* r1 = base + py1 * w;
    const float4 c00 = __ldg(r0 + px), c01 = __ldg(r0 + px1);
    const float4 c10 = __ldg(r1 + px), c11 = __ldg(r1 + px1);
    const float4 a = lerp4(c00, c01, fx), b = lerp4(c10, c11, fx);
    return lerp4(a, b, fy);
}

// trilinear footprint shared by n taps: levels floor(lod), floor(lod)+1 and the blend factor
struct TriCtx { LevelCtx lo, hi; float d; };

__device__ __forceinline__ TriCtx tri_ctx(const TexDesc& t, int layer, float lod, int hwFilter = 0) {
    TriCtx c;
    const int lo = (int)floorf(lod);
    c.d = lod - (float)lo;
    c.lo = level_ctx(t, layer, lo, hwFilter);
    c.hi = level_ctx(t, layer, lo + 1, hwFilter);
    return c;
}

__device__ __forceinline__ float4 sample_trilinear(const TriCtx& c, const SampDesc& s, float u, float v, int offx, int offy) {
    const float4 a = sample_level(c.lo, s, u, v, offx, offy);
    if (c.d == 0.0f) return a;   // lerp(a, b, 0) == a: skip the second level (bit-identical for finite texels)
    const float4 b = sample_level(c.hi, s, u, v, offx, offy);
    return lerp4(a, b, c.d);
}

__device__ __forceinline__ float4 sample_trilinear(const TexDesc& t, const SampDesc& s, int layer, float u, float v,
                                                   int offx, int offy, float lod, int hwFilter = 0) {
    return sample_trilinear(tri_ctx(t, layer, lod, hwFilter), s, u, v, offx, offy);
}

template <bool FAST>
__device__ __forceinline__ float4 sample_adaptive(const TexDesc& t, const SampDesc& s, float u, float v, float2 dx,
                                                  float2 dy, float bias, int hwFilter = 0) {
    const float W = (float)t.w, H = (float)t.h;
    const float ax = dx.x * W, ay = dx.y * H;
    const float bx = dy.x * W, by = dx.y * H;   // (sic) SB:237
    const float a2 = fmaf(ay, ay, ax * ax), b2 = fmaf(by, by, bx * bx);
    const float d = fmaxf(a2, b2);
    const float lod = fmaxf(0.0f, fmaf(0.5f, FAST ? lg2_fast(d) : log2f(d), bias));
    return sample_trilinear(t, s, 0, u, v, 0, 0, lod, hwFilter);
}

// 1/k for k = 0..33 as correctly rounded floats (identical to the oracle's 1.0f / (float)k)
__constant__ float kRcpTab[34] = {
    0.0f, 1.0f / 1.0f, 1.0f / 2.0f, 1.0f / 3.0f, 1.0f / 4.0f, 1.0f / 5.0f, 1.0f / 6.0f, 1.0f / 7.0f, 1.0f / 8.0f,
    1.0f / 9.0f, 1.0f / 10.0f, 1.0f / 11.0f, 1.0f / 12.0f, 1.0f / 13.0f, 1.0f / 14.0f, 1.0f / 15.0f, 1.0f / 16.0f,
    1.0f / 17.0f, 1.0f / 18.0f, 1.0f / 19.0f, 1.0f / 20.0f, 1.0f / 21.0f, 1.0f / 22.0f, 1.0f / 23.0f, 1.0f / 24.0f,
    1.0f / 25.0f, 1.0f / 26.0f, 1.0f / 27.0f, 1.0f / 28.0f, 1.0f / 29.0f, 1.0f / 30.0f, 1.0f / 31.0f, 1.0f / 32.0f,
    1.0f / 33.0f};
__device__ __forceinline__ float rcp_int(int k) { return k <= 33 ? kRcpTab[k] : __fdiv_rn(1.0f, (float)k); }

// texel-space coordinates of one tap on one level (address modes without a border colour)
__device__ __forceinline__ float2 tap_coords(const LevelCtx& L, const SampDesc& s, float u, float v) {
    bool border = false;
    const float cx = addr_mode(s.addrU, fmaf(u, L.fw1, 0.0f), L.fw, L.invW, L.inv2W, border);
    const float cy = addr_mode(s.addrV, fmaf(v, L.fh1, 0.0f), L.fh, L.invH, L.inv2H, border);
    return make_float2(cx + 0.5f, cy + 0.5f);
}

template <bool FAST>
__device__ __forceinline__ float4 sample_aniso(const TexDesc& t, const SampDesc& s, float u, float v, float2 du, float2 dv,
                                               float bias, int hwFilter = 0) {
    const float dudx = du.x, dudy = du.y, dvdx = dv.x, dvdy = dv.y;
    const float lx2 = fmaf(dvdx, dvdx, dudx * dudx), ly2 = fmaf(dvdy, dvdy, dudy * dudy);
    const float px = (FAST ? sqrt_fast(lx2) : sqrtf(lx2)) * (float)t.w;
    const float py = (FAST ? sqrt_fast(ly2) : sqrtf(ly2)) * (float)t.h;
    const float pmax = fmaxf(px, py), pmin = fminf(px, py);
    const int maxAniso = (int)s.maxAniso;
    // n = (int)fmax(1.0, fmin((double)ceilf(pmax / pmin) + 1e-3, (double)maxAniso + 1e-3)) of SB:252,
    // without the double-precision detour: ceilf() is integral, NaN / inf fall through to maxAniso
    const float cr = ceilf(FAST ? pmax * rcp_fast(pmin) : pmax / pmin);
    const int n = max(1, (cr < (float)maxAniso) ? (int)cr : maxAniso);
    const float step = rcp_int(n + 1), invN = rcp_int(n);
    const float lambda = fmaxf(0.0f, FAST ? lg2_fast(pmax * invN) : log2f(pmax / (float)n));
    const float ddu = (px > py) ? dudx : dudy, ddv = (px > py) ? dvdx : dvdy;
    float4 sum = make_float4(0, 0, 0, 0);
    const TriCtx ctx = tri_ctx(t, 0, lambda + bias, hwFilter);
    if (FAST && ctx.lo.hw && (ctx.d == 0.0f || ctx.hi.hw) && s.filter == IF_FILTER_LINEAR &&
        s.addrU != IF_SAMPLER_ADDRESS_MODE_CLAMP_TO_BORDER && s.addrV != IF_SAMPLER_ADDRESS_MODE_CLAMP_TO_BORDER) {
        // every footprint goes through the texture unit: taps are issued two at a time (up to four fetches in
        // flight) and summed in tap order; REPEAT / REPEAT (the common sampler) runs a copy of the loop without the
        // per-tap switch on the address mode
        const bool two = ctx.d != 0.0f;
        auto taps = [&](auto coords) {
            for (int i = 0; i < n; i += 2) {
                const float t0 = (float)i * step - 0.5f, t1 = (float)(i + 1) * step - 0.5f;
                const float u0 = fmaf(ddu, t0, u), v0 = fmaf(ddv, t0, v), u1 = fmaf(ddu, t1, u), v1 = fmaf(ddv, t1, v);
                const bool has1 = i + 1 < n;
                const float2 a0 = coords(ctx.lo, u0, v0), a1 = coords(ctx.lo, u1, v1);
                float4 r0 = tex2D<float4>((cudaTextureObject_t)ctx.lo.hw, a0.x, a0.y);
                float4 r1 = make_float4(0, 0, 0, 0);
                if (has1) r1 = tex2D<float4>((cudaTextureObject_t)ctx.lo.hw, a1.x, a1.y);
                if (two) {
                    const float2 b0 = coords(ctx.hi, u0, v0), b1 = coords(ctx.hi, u1, v1);
                    const float4 h0 = tex2D<float4>((cudaTextureObject_t)ctx.hi.hw, b0.x, b0.y);
                    float4 h1 = make_float4(0, 0, 0, 0);
                    if (has1) h1 = tex2D<float4>((cudaTextureObject_t)ctx.hi.hw, b1.x, b1.y);
                    r0 = lerp4(r0, h0, ctx.d);
                    r1 = lerp4(r1, h1, ctx.d);
                }
                sum.x += r0.x; sum.y += r0.y; sum.z += r0.z; sum.w += r0.w;
                if (has1) { sum.x += r1.x; sum.y += r1.y; sum.z += r1.z; sum.w += r1.w; }
            }
        };
        if (s.addrU == IF_SAMPLER_ADDRESS_MODE_REPEAT && s.addrV == IF_SAMPLER_ADDRESS_MODE_REPEAT)
            taps([&](const LevelCtx& L, float uu, float vv) {
                return make_float2(wrapf(fmaf(uu, L.fw1, 0.0f), L.fw, L.invW) + 0.5f, wrapf(fmaf(vv, L.fh1, 0.0f), L.fh, L.invH) + 0.5f);
            });
        else
            taps([&](const LevelCtx& L, float uu, float vv) { return tap_coords(L, s, uu, vv); });
    } else {
        for (int i = 0; i < n; i++) {
            const float tt = (float)i * step - 0.5f;
            const float4 r = sample_trilinear(ctx, s, fmaf(ddu, tt, u), fmaf(ddv, tt, v), 0, 0);
            sum.x += r.x; sum.y += r.y; sum.z += r.z; sum.w += r.w;
        }
    }
    return make_float4(sum.x * invN, sum.y * invN, sum.z * invN, sum.w * invN);
}

// ---- the common sampler, specialised: LINEAR filter, REPEAT / REPEAT addressing, anisotropic (C2 / C3 of BASELINE.json).
// Same arithmetic as sample_aniso -> sample_trilinear -> sample_level with those modes (FAST = false: op for op; FAST = true:
// footprints through the texture unit), without the switches on the sampler state, with the per-level constants read
// from the constant bank where they are used instead of being carried in registers across the tap loop.  Selected by
// the launcher when the bound sampler matches (kernels_tile.cu: ifr_launch_tiles).
#ifndef IFR_TAP_UNROLL
#define IFR_TAP_UNROLL 2     // taps of the anisotropic loops in flight together (texture latency vs registers)
#endif
#define IFR_PRAGMA_(x) _Pragma(#x)
#define IFR_PRAGMA(x) IFR_PRAGMA_(x)
template <bool FAST>
__device__ __forceinline__ float4 bilinear_repeat(const TexDesc& t, int lod, float u, float v) {
    const TexLevel& L = t.lvl[lod];
    if (FAST) {
        // REPEAT wrap without the two last-ulp fix-ups of wrapf: a coordinate that lands a hair outside [0, size) is
        // clamped by the texture unit
        const float cx = u * L.fw1, cy = v * L.fh1;
        const float x = fmaf(-floorf(cx * L.invW), L.fw, cx) + 0.5f, y = fmaf(-floorf(cy * L.invH), L.fh, cy) + 0.5f;
        const unsigned long long hw = t.hw[lod];
        if (hw) return tex2D<float4>((cudaTextureObject_t)hw, x, y);
        const int w = L.w, h = L.h;      // levels the texture unit cannot bind (1-texel rows, unaligned pitch)
        const int ix = clampi((int)(x - 0.5f), 0, w - 1), iy = clampi((int)(y - 0.5f), 0, h - 1);
        const float fx = (x - 0.5f) - (float)ix, fy = (y - 0.5f) - (float)iy;
        const float4* b = t.data + L.off;
        const int ix1 = min(ix + 1, w - 1), iy1 = min(iy + 1, h - 1);
        return lerp4(lerp4(__ldg(b + iy * w + ix), __ldg(b + iy * w + ix1), fx), lerp4(__ldg(b + iy1 * w + ix), __ldg(b + iy1 * w + ix1), fx), fy);
    } else {
        const int w = L.w, h = L.h;
        const float cx = wrapf(fmaf(u, L.fw1, 0.0f), L.fw, L.invW), cy = wrapf(fmaf(v, L.fh1, 0.0f), L.fh, L.invH);
        const int px = clampi((int)cx, 0, w - 1), py = clampi((int)cy, 0, h - 1);
        const float fx = cx - (float)px, fy = cy - (float)py;
        const int px1 = min(px + 1, w - 1), py1 = min(py + 1, h - 1);
        const float4* r0 = t.data + L.off + py * w;
        const float4* r1 = t.data + L.off + py1 * w;
        const float4 c00 = __ldg(r0 + px), c01 = __ldg(r0 + px1), c10 = __ldg(r1 + px), c11 = __ldg(r1 + px1);
        return lerp4(lerp4(c00, c01, fx), lerp4(c10, c11, fx), fy);
    }
}

template <bool FAST>
__device__ __forceinline__ float4 sample_aniso_repeat(const TexDesc& t, float maxAnisoF, float u, float v, float2 du, float2 dv, float bias) {
    const float dudx = du.x, dudy = du.y, dvdx = dv.x, dvdy = dv.y;
    const float lx2 = fmaf(dvdx, dvdx, dudx * dudx), ly2 = fmaf(dvdy, dvdy, dudy * dudy);
    const float px = (FAST ? sqrt_fast(lx2) : sqrtf(lx2)) * (float)t.w;
    const float py = (FAST ? sqrt_fast(ly2) : sqrtf(ly2)) * (float)t.h;
    const float pmax = fmaxf(px, py), pmin = fminf(px, py);
    const int maxAniso = (int)maxAnisoF;
    const float cr = ceilf(FAST ? pmax * rcp_fast(pmin) : pmax / pmin);
    const int n = max(1, (cr < (float)maxAniso) ? (int)cr : maxAniso);
    const float step = rcp_int(n + 1), invN = rcp_int(n);
    const float lambda = fmaxf(0.0f, FAST ? lg2_fast(pmax * invN) : log2f(pmax / (float)n));
    const float ddu = (px > py) ? dudx : dudy, ddv = (px > py) ? dvdx : dvdy;
    const float lod = lambda + bias;
    const int lo0 = (int)floorf(lod);
    const float d = lod - (float)lo0;
    const int lo = clampi(lo0, 0, t.mips), hi = clampi(lo0 + 1, 0, t.mips);
    float4 sum = make_float4(0, 0, 0, 0);
    if (d == 0.0f) {      // lerp(a, b, 0) == a: one level (the magnified case)
IFR_PRAGMA(unroll IFR_TAP_UNROLL)
        for (int i = 0; i < n; i++) {
            const float tt = (float)i * step - 0.5f;
            const float4 r = bilinear_repeat<FAST>(t, lo, fmaf(ddu, tt, u), fmaf(ddv, tt, v));
            sum.x += r.x; sum.y += r.y; sum.z += r.z; sum.w += r.w;
        }
    } else {
IFR_PRAGMA(unroll IFR_TAP_UNROLL)
        for (int i = 0; i < n; i++) {
            const float tt = (float)i * step - 0.5f;
            const float uu = fmaf(ddu, tt, u), vv = fmaf(ddv, tt, v);
            const float4 a = bilinear_repeat<FAST>(t, lo, uu, vv), b = bilinear_repeat<FAST>(t, hi, uu, vv);
            const float4 r = lerp4(a, b, d);
            sum.x += r.x; sum.y += r.y; sum.z += r.z; sum.w += r.w;
        }
    }
    return make_float4(sum.x * invN, sum.y * invN, sum.z * invN, sum.w * invN);
}

__device__ __forceinline__ float4 sample_cube(const TexDesc& t, const SampDesc& s, float x, float y, float z, float lod) {
    const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    float u, v, norm;
    int face;
    if (az >= ax && az >= ay) {
        face = z < 0.0f ? 5 : 4; norm = 0.5f / az; u = z < 0.0f ? -x : x; v = -y;
    } else if (ay >= ax) {
        face = y < 0.0f ? 3 : 2; norm = 0.5f / ay; u = x; v = y < 0.0f ? -z : z;
    } else {
        face = x < 0.0f ? 1 : 0; norm = 0.5f / ax; u = x < 0.0f ? z : -z; v = -y;
    }
    u = fmaf(u, norm, 0.5f);
    v = fmaf(v, norm, 0.5f);
    return sample_trilinear(t, s, face, u, v, 0, 0, lod);
}

// texture(tex, samp, varyings, uvVaryId, bias) — SB:357-372
template <bool FAST>
__device__ __forceinline__ float4 tex_grad(const TexDesc& t, const SampDesc& s, float u, float v, float2 ddx,
                                           float2 ddy, float bias, int hwFilter = 0) {
    return s.aniso ? sample_aniso<FAST>(t, s, u, v, ddx, ddy, bias, hwFilter)
                   : sample_adaptive<FAST>(t, s, u, v, ddx, ddy, bias, hwFilter);
}

// ---- fragment shader registry (static functors inlined into the tile kernel) -----------------
// nv = number of varyings the shader reads (interpolation of the others is skipped)
// Internal ids (not part of the ABI): the textured shaders with the FAST sampler — bilinear footprints blended by
// the texture unit (8-bit weights) and single-instruction SFU forms of sqrt / divide / log2 in the LOD and
// lighting math.  Selected by the launcher when iftrcuSetSamplerHardwareFilter is on (the default); the
// plain ids keep the oracle's arithmetic op for op.
// a fragment shader compiled at run time from CUDA source (iftrcuCreateFragmentShaderFromSource): the user's function is
// declared here and defined by the generated translation unit (capi.cu: rtc_compile)
#ifdef __CUDACC_RTC__
struct IfritFsInput;
__device__ float4 ifrit_user_fs_entry(const DrawParams& p, const float4* v, const float4* ddx, const float4* ddy, int prim, int fan);
#endif
#define IFRIT_FS_TEXTURED_FAST 100
#define IFRIT_FS_TEXTURED_LOD0_FAST 101
// `textured` with the sampler state LINEAR / REPEAT / REPEAT / anisotropic compiled in (sample_aniso_repeat)
#define IFRIT_FS_TEXTURED_RA 102
#define IFRIT_FS_TEXTURED_FAST_RA 103
template <int FS> struct FsTraits { static constexpr bool needsDerivs = false; static constexpr int nv = 1; };
template <> struct FsTraits<IFRIT_FS_TEXTURED_FAST> { static constexpr bool needsDerivs = true; static constexpr int nv = 2; };
template <> struct FsTraits<IFRIT_FS_TEXTURED_LOD0_FAST> { static constexpr bool needsDerivs = false; static constexpr int nv = 2; };
template <> struct FsTraits<IFRIT_FS_TEXTURED_RA> { static constexpr bool needsDerivs = true; static constexpr int nv = 2; };
template <> struct FsTraits<IFRIT_FS_TEXTURED_FAST_RA> { static constexpr bool needsDerivs = true; static constexpr int nv = 2; };
template <> struct FsTraits<IFRIT_FS_PRIM_ID> { static constexpr bool needsDerivs = false; static constexpr int nv = 0; };
template <> struct FsTraits<IFRIT_FS_TEXTURED> { static constexpr bool needsDerivs = true; static constexpr int nv = 2; };
template <> struct FsTraits<IFRIT_FS_TEXTURED_LOD0> { static constexpr bool needsDerivs = false; static constexpr int nv = 2; };
template <> struct FsTraits<IFRIT_FS_DERIV_VIS> { static constexpr bool needsDerivs = true; static constexpr int nv = 2; };

// A shader that writes depth (FragmentShader::allowDepthModification, Shaders.h:61; `float* fragmentDepth`, CU:1803): its
// draws take the ordered path with the depth test before and after the shader (kernels_tile.cu: tile_blend_kernel).
__host__ __device__ constexpr bool fs_modifies_depth(int fs) { return fs == IFRIT_FS_DEPTH_MOD; }
template <int FS>
__device__ __forceinline__ float fs_depth(const DrawParams& p, const float4* v, float depth) {
    if (FS == IFRIT_FS_DEPTH_MOD) return fmaf(p.u0[22], v[0].x, depth);
    return depth;
}

template <int FS>
__device__ __forceinline__ float4 run_fs(const DrawParams& p, const float4* v, const float4* ddx,
                                         const float4* ddy, int prim, int fan) {
    const float* u = p.u0;   // uniform block, constant bank
#ifdef __CUDACC_RTC__
    if (FS == IFRIT_FS_USER) return ifrit_user_fs_entry(p, v, ddx, ddy, prim, fan);
#endif
    if (FS == IFRIT_FS_VARYING0 || FS == IFRIT_FS_DEPTH_MOD) return v[0];
    if (FS == IFRIT_FS_NORMAL_COLOR)
        return make_float4(fmaf(v[0].x, 0.5f, 0.5f), fmaf(v[0].y, 0.5f, 0.5f), fmaf(v[0].z, 0.5f, 0.5f), 0.5f);
    if (FS == IFRIT_FS_PRIM_ID) return make_float4((float)prim, (float)fan, 0.0f, 1.0f);
    if (FS == IFRIT_FS_TEXTURED || FS == IFRIT_FS_TEXTURED_LOD0 || FS == IFRIT_FS_TEXTURED_FAST || FS == IFRIT_FS_TEXTURED_LOD0_FAST ||
        FS == IFRIT_FS_TEXTURED_RA || FS == IFRIT_FS_TEXTURED_FAST_RA) {
        constexpr bool FAST = FS == IFRIT_FS_TEXTURED_FAST || FS == IFRIT_FS_TEXTURED_LOD0_FAST || FS == IFRIT_FS_TEXTURED_FAST_RA;
        float4 c;
        if (FS == IFRIT_FS_TEXTURED_RA || FS == IFRIT_FS_TEXTURED_FAST_RA)
            c = sample_aniso_repeat<FAST>(p.tex[0], p.samp[0].maxAniso, v[1].x, v[1].y, make_float2(ddx[1].x, ddx[1].y),
                                          make_float2(ddy[1].x, ddy[1].y), u[20]);
        else if (FS == IFRIT_FS_TEXTURED || FS == IFRIT_FS_TEXTURED_FAST)
            c = tex_grad<FAST>(p.tex[0], p.samp[0], v[1].x, v[1].y, make_float2(ddx[1].x, ddx[1].y),
                               make_float2(ddy[1].x, ddy[1].y), u[20], FAST ? 1 : 0);
        else
            c = sample_trilinear(p.tex[0], p.samp[0], 0, v[1].x, v[1].y, 0, 0, 0.0f, FAST ? 1 : 0);
        const float nx = v[0].x, ny = v[0].y, nz = v[0].z;
        const float len2 = fmaf(nz, nz, fmaf(ny, ny, nx * nx));
        const float dot = fmaf(nz, u[18], fmaf(ny, u[17], nx * u[16]));
        float ndl = FAST ? dot * rsqrt_fast(len2) : dot / sqrtf(len2);
        ndl = fmaxf(ndl, 0.2f);
        return make_float4(c.x * ndl, c.y * ndl, c.z * ndl, 1.0f);
    }
    if (FS == IFRIT_FS_SKYBOX) return sample_cube(p.tex[0], p.samp[0], v[0].x, v[0].y, v[0].z, u[20]);
    if (FS == IFRIT_FS_DERIV_VIS) {
        const float sc = u[21];
        return make_float4(ddx[1].x * sc, ddx[1].y * sc, ddy[1].x * sc, ddy[1].y * sc);
    }
    return make_float4(0, 0, 0, 0);
}
