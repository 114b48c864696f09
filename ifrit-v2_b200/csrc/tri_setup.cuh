// tri_setup.cuh — per (triangle, tile) raster setup shared by the fill pass (which writes it into the
// big-triangle queue) and the tile kernel (which computes it in registers for small triangles).
// Replaces geometryParamPostprocKernel CU:1549-1644: edge equations are exact integers in 24.8 fixed
// point instead of FP32 planes.
#pragma once
#include "common.cuh"

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
        long long T = ((-s.E0[i]) >> 4) + 1;
        T = T > 2147483647LL ? 2147483647LL : (T < -2147483648LL ? -2147483648LL : T);
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
    return bx0 <= bx1 && by0 <= by1;
}

__device__ __forceinline__ bool setup_tri(const DrawParams& p, uint32_t slot, int tileX, int tileY, TriS& s,
                                          int& bx0, int& bx1, int& by0, int& by1) {
    RecRegs rr;
    load_rec(p, slot, rr);
    return setup_from_rec(p, rr, slot, tileX, tileY, s, bx0, bx1, by0, by1);
}
