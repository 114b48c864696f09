// tri_setup.cuh — per (triangle, tile) raster setup shared by the fill pass (which writes it into the
// big-triangle queue) and the tile kernel (which computes it in registers for small triangles).
// Replaces geometryParamPostprocKernel CU:1549-1644: edge equations are exact integers in 24.8 fixed
// point instead of FP32 planes.
#pragma once
#include "common.cuh"
#include "../../include/ifrit_b200.h"

// order-preserving float -> uint map of the visibility keys (oracle: key_* in ifrit_oracle.c)
__device__ __forceinline__ uint32_t mono(float d) {
    d = d + 0.0f;                                   // -0 -> +0
    const uint32_t b = __float_as_uint(d);
    // negative: all bits flipped, otherwise the sign bit set — one arithmetic shift + two logic ops
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}

// the vertex shader functors of the registry (position only: both pass the other attributes through)
template <int VS>
__device__ __forceinline__ float4 run_vs(const DrawParams& p, const float4 a0) {
    if (VS == IFRIT_VS_MVP) {
        const float4* m = reinterpret_cast<const float4*>(p.u0);   // constant bank
        const float4 r0 = m[0], r1 = m[1], r2 = m[2], r3 = m[3];
        float4 o;
        o.x = fmaf(r0.w, a0.w, fmaf(r0.z, a0.z, fmaf(r0.y, a0.y, __fmul_rn(r0.x, a0.x))));
        o.y = fmaf(r1.w, a0.w, fmaf(r1.z, a0.z, fmaf(r1.y, a0.y, __fmul_rn(r1.x, a0.x))));
        o.z = fmaf(r2.w, a0.w, fmaf(r2.z, a0.z, fmaf(r2.y, a0.y, __fmul_rn(r2.x, a0.x))));
        o.w = fmaf(r3.w, a0.w, fmaf(r3.z, a0.z, fmaf(r3.y, a0.y, __fmul_rn(r3.x, a0.x))));
        return o;
    }
    return a0;
}


struct RecRegs { int X[3], Y[3]; float Z[3]; float ar; uint32_t flags; };

__device__ __forceinline__ void load_rec(const DrawParams& p, uint32_t slot, RecRegs& rr) {
    const int4* r = reinterpret_cast<const int4*>(p.rec + slot);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int4 v = __ldg(r + j);
        rr.X[j] = v.x; rr.Y[j] = v.y; rr.Z[j] = __int_as_float(v.z);
    }
    const int4 w3 = __ldg(r + 3);
    rr.ar = __int_as_float(w3.x);
    rr.flags = (uint32_t)w3.y;
}

// Hi-Z extras of one (triangle, tile) pair: trivial-accept blocks (mask bits 8-15) and the most optimistic depth
// inside the tile-local box.  Computed by the tile kernel itself, on the staged queue entries (shared memory) of
// tiles whose big queue is deep enough for the reject to pay off — the fill pass and shallow tiles never see it.
//
// The depth the tile kernel computes, F(x, y) = fma chain of z_i * (float(E_i) * ar), follows the exact plane
// P = sum z_i E_i / d within eps * M, M = sum |z_i| |E_i / d| (convex), eps < 1.75 * 2^-22.  P is affine, so over
// the box F >= min_corners F - 2 eps max_corners M; the margin below is twice that.
__device__ __forceinline__ uint32_t setup_hiz(TriS* sp, int sampMinX, int sampMaxX, int sampMinY, int sampMaxY, uint32_t xhi) {
    const bool greater = xhi != 0;
    TriS& s = *sp;
    const int bx0 = s.box & 0xff, bx1 = (s.box >> 8) & 0xff, by0 = (s.box >> 16) & 0xff, by1 = s.box >> 24;
    uint32_t mask = (uint32_t)s.pad[0];
    if (!(s.flags & IFR_FLAG_WIDE)) {
        const int sLoX = sampMinX >> 4, sHiX = sampMaxX >> 4, sLoY = sampMinY >> 4, sHiY = sampMaxY >> 4;
#pragma unroll
        for (int blk = 0; blk < 8; blk++) {
            if (!((mask >> blk) & 1)) continue;
            const int x0 = (blk & 1) * 8, y0 = (blk >> 1) * 4;
            const int iLo = 16 * x0 + sLoX, iHi = 16 * (x0 + 7) + sHiX, jLo = 16 * y0 + sLoY, jHi = 16 * (y0 + 3) + sHiY;
            bool full = true;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const int worst = s.dX[i] * (s.dX[i] > 0 ? jLo : jHi) - s.dY[i] * (s.dY[i] > 0 ? iHi : iLo);
                full = full && worst >= s.T[i];
            }
            mask |= (uint32_t)full << (8 + blk);
        }
    }
    float lo = 3.0e38f, hi = -3.0e38f, mag = 0.0f;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int lx = (c & 1) ? bx1 : bx0, ly = (c & 2) ? by1 : by0;
        float beta[3];
#pragma unroll
        for (int e = 0; e < 3; e++) {
            const long long Ec = s.E0[e] + 256 * ((long long)s.dX[e] * ly - (long long)s.dY[e] * lx);
            beta[e] = __fmul_rn((float)Ec, s.ar);
        }
        const float d = fmaf(s.z2, beta[2], fmaf(s.z1, beta[1], __fmul_rn(s.z0, beta[0])));
        const float m = fabsf(s.z0 * beta[0]) + fabsf(s.z1 * beta[1]) + fabsf(s.z2 * beta[2]);
        lo = fminf(lo, d); hi = fmaxf(hi, d); mag = fmaxf(mag, m);
    }
    const float margin = mag * 1.9073486e-6f + 1e-30f;   // 2^-19 M
    float zbound = greater ? hi + margin : lo - margin;
    if (!(zbound == zbound)) zbound = greater ? 3.0e38f : -3.0e38f;   // NaN depth: never reject
    // stored as the key depth word (mono(z) ^ xhi, the form the tile kernel compares): smaller = more optimistic
    const uint32_t tb = mono(zbound) ^ xhi;
    s.pad[0] = (int)mask;
    s.pad[1] = (int)tb;
    return tb;
}

#define IFR_HIZ_MIN_QUEUE 12   // tiles with fewer big-queue entries run the tile kernel's big loop without the Hi-Z reject

template <bool COARSE>
__device__ __forceinline__ bool setup_from_rec(const DrawParams& p, const RecRegs& rr, uint32_t slot, int tileX, int tileY,
                                               TriS& s, int& bx0, int& bx1, int& by0, int& by1) {
    const int* X = rr.X; const int* Y = rr.Y;
    s.slot = slot;
    s.ar = rr.ar;
    s.flags = rr.flags;
    s.z0 = rr.Z[0]; s.z1 = rr.Z[1]; s.z2 = rr.Z[2];
    s.pad[0] = 0; s.pad[1] = 0;
    // tile origin <= 2^21 and |X|,|Y| <= 2^28: the differences fit int32, products are IMAD.WIDE
    const int sx0 = tileX * (IFR_TILE * IFR_SUBPIX), sy0 = tileY * (IFR_TILE * IFR_SUBPIX);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int a = (i + 1) % 3, b = (i + 2) % 3;
        s.dX[i] = X[b] - X[a]; s.dY[i] = Y[b] - Y[a];
        s.E0[i] = (long long)s.dX[i] * (long long)(sy0 - Y[a]) - (long long)s.dY[i] * (long long)(sx0 - X[a]);
        // inside <=> dX*j - dY*i >= T.  Non-wide edges (|dX|,|dY| <= 2^20) keep |dX*j - dY*i| < 2^29 over the
        // tile, so clamping T to +-2^30 changes no outcome and lets the tile kernel form (dX*j - dY*i - T) in int32
        long long T = ((-s.E0[i]) >> 4) + 1;
        T = T > (1LL << 30) ? (1LL << 30) : (T < -(1LL << 30) ? -(1LL << 30) : T);
        s.T[i] = (int)T;
    }
    const int minx = min(X[0], min(X[1], X[2])), maxx = max(X[0], max(X[1], X[2]));
    const int miny = min(Y[0], min(Y[1], Y[2])), maxy = max(Y[0], max(Y[1], Y[2]));
    const int px0 = ((minx - p.sampMaxX) >> 8) + 1, px1 = (maxx - p.sampMinX - 1) >> 8;
    const int py0 = ((miny - p.sampMaxY) >> 8) + 1, py1 = (maxy - p.sampMinY - 1) >> 8;
    const int ox = tileX * IFR_TILE, oy = tileY * IFR_TILE;
    bx0 = max(px0 - ox, 0); by0 = max(py0 - oy, 0);
    bx1 = min(min(px1, p.W - 1) - ox, IFR_TILE - 1); by1 = min(min(py1, p.H - 1) - oy, IFR_TILE - 1);
    s.box = (uint32_t)(bx0 & 0xff) | ((uint32_t)(bx1 & 0xff) << 8) | ((uint32_t)(by0 & 0xff) << 16) | ((uint32_t)(by1 & 0xff) << 24);
    const bool ok = bx0 <= bx1 && by0 <= by1;
    // Coarse raster (secondBinnerWorklistRasterizerKernel CU:1254 keeps 4x4-px sub-tiles): which of the tile's eight
    // 8x4-pixel blocks (one per warp of the tile kernel; bit = (y >> 2) * 2 + (x >> 3)) can hold a covered sample
    // (bits 0-7: exact trivial reject — an edge whose function is below threshold at the block's most-inside sample
    // corner) and which are covered at every sample (bits 8-15: trivial accept — every edge passes at the block's
    // most-outside sample corner, CU:287-313), so that the tile kernel skips the per-sample edge tests there.
    uint32_t mask = 0;
    if (COARSE && ok) {
        const int sLoX = p.sampMinX >> 4, sHiX = p.sampMaxX >> 4, sLoY = p.sampMinY >> 4, sHiY = p.sampMaxY >> 4;   // 1/16 px
        // Per edge, the function at the block's most-inside sample corner minus the threshold,
        //     v(blk) = dX * (16 y0 + (dX > 0 ? 48 + sHiY : sLoY)) - dY * (16 x0 + (dY > 0 ? sLoX : 112 + sHiX)) - T,
        // is its value at block 0 plus 64 dX per block row and - 128 dY for the right-hand column: one add per edge and block
        // instead of two selects and two multiplies (the fill pass of c5 spent a third of its instructions here).  Exact: the
        // same integers re-associated; |v| < 2^29 + 2^30 for non-wide edges (T is clamped to +-2^30 above).
        int v0[3], stepRow[3], stepCol[3];
        const bool wideTri = (rr.flags & IFR_FLAG_WIDE) != 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            v0[i] = s.dX[i] * (s.dX[i] > 0 ? 48 + sHiY : sLoY) - s.dY[i] * (s.dY[i] > 0 ? sLoX : 112 + sHiX) - s.T[i];
            stepRow[i] = 64 * s.dX[i]; stepCol[i] = -128 * s.dY[i];
        }
#pragma unroll
        for (int blk = 0; blk < 8; blk++) {
            const int x0 = (blk & 1) * 8, y0 = (blk >> 1) * 4;
            if (x0 + 7 < bx0 || x0 > bx1 || y0 + 3 < by0 || y0 > by1) continue;
            const int a0 = v0[0] + (blk >> 1) * stepRow[0] + (blk & 1) * stepCol[0];
            const int a1 = v0[1] + (blk >> 1) * stepRow[1] + (blk & 1) * stepCol[1];
            const int a2 = v0[2] + (blk >> 1) * stepRow[2] + (blk & 1) * stepCol[2];
            const bool may = wideTri || (a0 | a1 | a2) >= 0;
            mask |= (uint32_t)may << blk;
        }
    }
    s.pad[0] = (int)mask;
    s.pad[1] = 0;   // Hi-Z key bound: filled in by the tile kernel (setup_hiz) for deep queues
    return ok;
}

__device__ __forceinline__ bool setup_tri(const DrawParams& p, uint32_t slot, int tileX, int tileY, TriS& s,
                                          int& bx0, int& bx1, int& by0, int& by1) {
    RecRegs rr;
    load_rec(p, slot, rr);
    return setup_from_rec<false>(p, rr, slot, tileX, tileY, s, bx0, bx1, by0, by1);
}
