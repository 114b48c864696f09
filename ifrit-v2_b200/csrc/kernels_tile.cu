// kernels_tile.cu — the per-tile raster + early-Z + quad shading + MSAA resolve kernel.
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
#include "common.cuh"
#include "sampler.cuh"

// ---- visibility keys (identical to oracle/ifrit_oracle.c key_*) --------------------------------
__device__ __forceinline__ uint32_t mono(float d) {
    d = d + 0.0f;
    const uint32_t b = __float_as_uint(d);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unmono(uint32_t m) {
    return __uint_as_float((m & 0x80000000u) ? (m & 0x7fffffffu) : ~m);
}
__device__ __forceinline__ unsigned long long key_init(int func, float d0) {
    switch (func) {
        case IF_COMPARE_OP_LESS: return ((unsigned long long)mono(d0) << 32) | 0u;
        case IF_COMPARE_OP_LESS_OR_EQUAL: return ((unsigned long long)mono(d0) << 32) | 0xffffffffu;
        case IF_COMPARE_OP_GREATER: return ((unsigned long long)(~mono(d0)) << 32) | 0u;
        case IF_COMPARE_OP_GREATER_OR_EQUAL: return ((unsigned long long)(~mono(d0)) << 32) | 0xffffffffu;
        default: return ((unsigned long long)0xffffffffu << 32) | __float_as_uint(d0 + 0.0f);
    }
}
__device__ __forceinline__ bool key_make(int func, float d, float d0, uint32_t slot, unsigned long long& k) {
    switch (func) {
        case IF_COMPARE_OP_LESS: k = ((unsigned long long)mono(d) << 32) | (slot + 1u); return true;
        case IF_COMPARE_OP_LESS_OR_EQUAL: k = ((unsigned long long)mono(d) << 32) | ~(slot + 1u); return true;
        case IF_COMPARE_OP_GREATER: k = ((unsigned long long)(~mono(d)) << 32) | (slot + 1u); return true;
        case IF_COMPARE_OP_GREATER_OR_EQUAL: k = ((unsigned long long)(~mono(d)) << 32) | ~(slot + 1u); return true;
        case IF_COMPARE_OP_ALWAYS: k = ((unsigned long long)(slot + 1u) << 32) | __float_as_uint(d + 0.0f); return true;
        case IF_COMPARE_OP_EQUAL:
            if (!(d == d0)) return false;
            k = ((unsigned long long)(slot + 1u) << 32) | __float_as_uint(d + 0.0f); return true;
        default: return false;
    }
}
__device__ __forceinline__ uint32_t key_slot(int func, unsigned long long k) {   // IFR_SLOT_NONE if empty
    uint32_t s;
    switch (func) {
        case IF_COMPARE_OP_LESS: case IF_COMPARE_OP_GREATER: s = (uint32_t)k; return s == 0 ? IFR_SLOT_NONE : s - 1;
        case IF_COMPARE_OP_LESS_OR_EQUAL: case IF_COMPARE_OP_GREATER_OR_EQUAL:
            s = ~(uint32_t)k; return s == 0 ? IFR_SLOT_NONE : s - 1;
        default: s = (uint32_t)(k >> 32); return s == 0xffffffffu ? IFR_SLOT_NONE : s - 1;
    }
}

__device__ __forceinline__ bool scissor_pass(const DrawParams& p, int x, int y) {
    if (!p.scissorEnable || p.nScissor == 0) return true;
    for (int i = 0; i < p.nScissor; i++) {
        const float* s = p.scissor[i];
        if ((float)x >= s[0] && (float)x <= s[0] + s[2] && (float)y >= s[1] && (float)y <= s[1] + s[3]) return true;
    }
    return false;
}

// ---- per (triangle, tile) setup ---------------------------------------------------------------
struct TriSetup {
    int dX[3], dY[3];
    long long E0[3];      // edge functions at the tile origin
    int T[3];             // int32 thresholds: inside <=> dX*j - dY*i >= T  (i, j in 1/16 px, tile local)
    float Z[3], ar;
    int bx0, bx1, by0, by1;   // tile-local pixel bbox that can contain covered samples
    uint32_t slot;
    int wide;             // edge deltas too large for the int32 path
};

__device__ __forceinline__ bool setup_tri(const DrawParams& p, uint32_t slot, int tileX, int tileY, TriSetup& s) {
    const int4* r = reinterpret_cast<const int4*>(p.rec + (size_t)slot * 3);
    int X[3], Y[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int4 v = __ldg(r + j);
        X[j] = v.x >> 2; Y[j] = v.y; s.Z[j] = __int_as_float(v.z);
    }
    s.slot = slot;
    // tile origin <= 2^21 and |X|,|Y| <= 2^28: the differences fit int32, products are IMAD.WIDE
    const int sx0 = tileX * (IFR_TILE * IFR_SUBPIX), sy0 = tileY * (IFR_TILE * IFR_SUBPIX);
    int wide = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int a = (i + 1) % 3, b = (i + 2) % 3;
        s.dX[i] = X[b] - X[a]; s.dY[i] = Y[b] - Y[a];
        s.E0[i] = (long long)s.dX[i] * (long long)(sy0 - Y[a]) - (long long)s.dY[i] * (long long)(sx0 - X[a]);
        wide |= (abs(s.dX[i]) > (1 << 22)) | (abs(s.dY[i]) > (1 << 22));
        long long T = ((-s.E0[i]) >> 4) + 1;
        T = T > 2147483647LL ? 2147483647LL : (T < -2147483648LL ? -2147483648LL : T);
        s.T[i] = (int)T;
    }
    s.wide = wide;
    const long long d = (long long)s.dX[1] * s.dY[2] - (long long)s.dX[2] * s.dY[1];
    s.ar = __frcp_rn((float)d);
    const int minx = min(X[0], min(X[1], X[2])), maxx = max(X[0], max(X[1], X[2]));
    const int miny = min(Y[0], min(Y[1], Y[2])), maxy = max(Y[0], max(Y[1], Y[2]));
    const int px0 = ((minx - p.sampMaxX) >> 8) + 1, px1 = (maxx - p.sampMinX - 1) >> 8;
    const int py0 = ((miny - p.sampMaxY) >> 8) + 1, py1 = (maxy - p.sampMinY - 1) >> 8;
    const int ox = tileX * IFR_TILE, oy = tileY * IFR_TILE;
    s.bx0 = max(px0 - ox, 0); s.by0 = max(py0 - oy, 0);
    s.bx1 = min(min(px1, p.W - 1) - ox, IFR_TILE - 1); s.by1 = min(min(py1, p.H - 1) - oy, IFR_TILE - 1);
    return s.bx0 <= s.bx1 && s.by0 <= s.by1;
}

// coverage mask of pixel (lx, ly) of the tile over the S samples
template <int S>
__device__ __forceinline__ uint32_t coverage(const DrawParams& p, const TriSetup& s, int lx, int ly) {
    uint32_t mask = 0;
    if (!s.wide) {
#pragma unroll
        for (int k = 0; k < S; k++) {
            const int i = 16 * lx + p.sampX[k], j = 16 * ly + p.sampY[k];
            const bool in = (s.dX[0] * j - s.dY[0] * i >= s.T[0]) && (s.dX[1] * j - s.dY[1] * i >= s.T[1]) &&
                            (s.dX[2] * j - s.dY[2] * i >= s.T[2]);
            mask |= (uint32_t)in << k;
        }
    } else {
        for (int k = 0; k < S; k++) {
            const long long i = 16 * lx + p.sampX[k], j = 16 * ly + p.sampY[k];
            bool in = true;
#pragma unroll
            for (int e = 0; e < 3; e++) in = in && (s.E0[e] + 16 * (s.dX[e] * j - s.dY[e] * i) > 0);
            mask |= (uint32_t)in << k;
        }
    }
    return mask;
}

// depth at the integer pixel (lx, ly): beta_i = E_i(corner) / d, from exact integer edge values
__device__ __forceinline__ float depth_at(const TriSetup& s, int lx, int ly) {
    float beta[3];
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const long long Ec = s.E0[e] + 256 * ((long long)s.dX[e] * ly - (long long)s.dY[e] * lx);
        beta[e] = __fmul_rn((float)Ec, s.ar);
    }
    return fmaf(s.Z[2], beta[2], fmaf(s.Z[1], beta[1], __fmul_rn(s.Z[0], beta[0])));
}

// tile-local pixel -> key index = index of the owning thread (quad order: 4 consecutive lanes = 2x2)
__device__ __forceinline__ int pix_index(int lx, int ly) {
    return ((ly >> 1) * 8 + (lx >> 1)) * 4 + (ly & 1) * 2 + (lx & 1);
}

__device__ __forceinline__ void key_min_shared(unsigned long long* addr, unsigned long long k) {
    if (k < *reinterpret_cast<volatile unsigned long long*>(addr)) atomicMin(addr, k);
}

#define IFR_SMALL_AREA 16
#define IFR_CHUNK 256

// ---- attribute interpolation at an integer pixel (pixelShadingExecKernel CU:2168-2215) ---------
__device__ __forceinline__ void interp(const DrawParams& p, uint32_t slot, int px, int py, float4* out, float& depth,
                                       int& prim, int& fan) {
    const int4* r = reinterpret_cast<const int4*>(p.rec + (size_t)slot * 3);
    int X[3], Y[3];
    float Z[3], IW[3];
    int flags = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int4 v = __ldg(r + j);
        if (j == 0) flags = v.x & 3;
        X[j] = v.x >> 2; Y[j] = v.y; Z[j] = __int_as_float(v.z); IW[j] = __int_as_float(v.w);
    }
    const int sx = px * IFR_SUBPIX, sy = py * IFR_SUBPIX;
    const long long d = (long long)(X[1] - X[0]) * (long long)(Y[2] - Y[0]) - (long long)(X[2] - X[0]) * (long long)(Y[1] - Y[0]);
    const float ar = __frcp_rn((float)d);
    float beta[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int a = (i + 1) % 3, b = (i + 2) % 3;
        const long long E = (long long)(X[b] - X[a]) * (long long)(sy - Y[a]) - (long long)(Y[b] - Y[a]) * (long long)(sx - X[a]);
        beta[i] = __fmul_rn((float)E, ar);
    }
    depth = fmaf(Z[2], beta[2], fmaf(Z[1], beta[1], __fmul_rn(Z[0], beta[0])));
    float b0 = __fmul_rn(beta[0], IW[0]), b1 = __fmul_rn(beta[1], IW[1]), b2 = __fmul_rn(beta[2], IW[2]);
    const float inv = __frcp_rn(__fadd_rn(__fadd_rn(b0, b1), b2));
    b0 = __fmul_rn(b0, inv); b1 = __fmul_rn(b1, inv); b2 = __fmul_rn(b2, inv);
    float B[3][2];
    if (flags & IFR_FLAG_CLIPPED) {
        const float* cb = p.clipBary + (size_t)slot * 6;
#pragma unroll
        for (int j = 0; j < 3; j++) { B[j][0] = __ldg(cb + 2 * j); B[j][1] = __ldg(cb + 2 * j + 1); }
    } else {   // unclipped: (v2, v3, v1) of the input triangle, first and last exchanged by a cull swap
        const bool sw = flags & IFR_FLAG_SWAPPED;
        B[0][0] = sw ? 1.0f : 0.0f; B[0][1] = sw ? 0.0f : 1.0f;
        B[1][0] = 0.0f; B[1][1] = 0.0f;
        B[2][0] = sw ? 0.0f : 1.0f; B[2][1] = sw ? 1.0f : 0.0f;
    }
    const float u = fmaf(b2, B[2][0], fmaf(b1, B[1][0], __fmul_rn(b0, B[0][0])));
    const float v = fmaf(b2, B[2][1], fmaf(b1, B[1][1], __fmul_rn(b0, B[0][1])));
    const float w2 = __fsub_rn(1.0f, __fadd_rn(u, v));
    fan = slot >= (uint32_t)p.ntri;
    prim = fan ? (int)(slot - (uint32_t)p.ntri) : (int)slot;
    const int i0 = __ldg(p.ib + 3 * (size_t)prim), i1 = __ldg(p.ib + 3 * (size_t)prim + 1), i2 = __ldg(p.ib + 3 * (size_t)prim + 2);
    for (int k = 0; k < p.nvary; k++) {
        const float4 v0 = __ldg(p.vary + (size_t)i0 * p.nvary + k), v1 = __ldg(p.vary + (size_t)i1 * p.nvary + k),
                     v2 = __ldg(p.vary + (size_t)i2 * p.nvary + k);
        out[k].x = fmaf(v2.x, w2, fmaf(v1.x, v, __fmul_rn(v0.x, u)));
        out[k].y = fmaf(v2.y, w2, fmaf(v1.y, v, __fmul_rn(v0.y, u)));
        out[k].z = fmaf(v2.z, w2, fmaf(v1.z, v, __fmul_rn(v0.z, u)));
        out[k].w = fmaf(v2.w, w2, fmaf(v1.w, v, __fmul_rn(v0.w, u)));
    }
}

// ------------------------------------------------------------------------------------------------
template <int FS, int S>
__global__ void __launch_bounds__(256) tile_kernel(const __grid_constant__ DrawParams p) {
    extern __shared__ unsigned long long keys[];        // [S][256], index = owning thread
    __shared__ TriSetup big[IFR_CHUNK];
    __shared__ int nbig;

    const int tileX = blockIdx.x, tileY = blockIdx.y;
    if (p.world > 1 && ifr_tile_owner(tileX, tileY, p.world) != p.rank) return;
    const int tile = tileY * p.tilesX + tileX;
    const int tid = threadIdx.x;
    const int q = tid >> 2, l = tid & 3;
    const int lx = (q & 7) * 2 + (l & 1), ly = (q >> 3) * 2 + (l >> 1);
    const int px = tileX * IFR_TILE + lx, py = tileY * IFR_TILE + ly;
    const bool inFrame = px < p.W && py < p.H;
    const bool scis = inFrame && scissor_pass(p, px, py);
    const uint32_t qStart = p.tileStart[tile], qEnd = p.tileStart[tile + 1];
    const bool overflow = p.counters[3] != 0;
    const uint32_t nq = overflow ? 0u : qEnd - qStart;
    if (nq == 0 && !p.doClear) return;

    const int func = p.depthFunc;
    const size_t pidx = (size_t)py * p.W + px;
    float d0 = 0.0f;
    if (inFrame) d0 = (p.doClear && scis) ? p.clearDepth : p.depth[pidx];
    {
        const unsigned long long k0 = key_init(func, d0);
#pragma unroll
        for (int s = 0; s < S; s++) keys[s * IFR_TILE_PIX + tid] = k0;
    }
    if (tid == 0) nbig = 0;
    __syncthreads();

    // ---- coverage + early-Z ---------------------------------------------------------------------
    for (uint32_t base = 0; base < nq; base += IFR_CHUNK) {
        const uint32_t qi = base + tid;
        if (qi < nq) {
            TriSetup s;
            const uint32_t slot = __ldg(p.queue + qStart + qi);
            if (setup_tri(p, slot, tileX, tileY, s)) {
                const int area = (s.bx1 - s.bx0 + 1) * (s.by1 - s.by0 + 1);
                if (area <= IFR_SMALL_AREA) {
                    // small triangle: this thread walks its few pixels, shared-memory atomicMin
                    for (int y = s.by0; y <= s.by1; y++)
                        for (int x = s.bx0; x <= s.bx1; x++) {
                            const uint32_t m = coverage<S>(p, s, x, y);
                            if (!m) continue;
                            const int gx = tileX * IFR_TILE + x, gy = tileY * IFR_TILE + y;
                            if (p.scissorEnable && !scissor_pass(p, gx, gy)) continue;
                            const float d = depth_at(s, x, y);
                            unsigned long long k;
                            float dd0 = 0.0f;
                            if (func == IF_COMPARE_OP_EQUAL)
                                dd0 = p.doClear ? p.clearDepth : p.depth[(size_t)gy * p.W + gx];
                            if (!key_make(func, d, dd0, slot, k)) continue;
                            const int pi = pix_index(x, y);
#pragma unroll
                            for (int sm = 0; sm < S; sm++)
                                if ((m >> sm) & 1) key_min_shared(&keys[sm * IFR_TILE_PIX + pi], k);
                        }
                } else {
                    big[atomicAdd(&nbig, 1)] = s;
                }
            }
        }
        __syncthreads();
        // large triangles: every thread tests its own pixel, plain stores to its own keys
        const int nb = nbig;
        if (scis) {
            for (int b = 0; b < nb; b++) {
                const TriSetup& s = big[b];
                if (lx < s.bx0 || lx > s.bx1 || ly < s.by0 || ly > s.by1) continue;
                const uint32_t m = coverage<S>(p, s, lx, ly);
                if (!m) continue;
                const float d = depth_at(s, lx, ly);
                unsigned long long k;
                if (!key_make(func, d, d0, s.slot, k)) continue;
#pragma unroll
                for (int sm = 0; sm < S; sm++)
                    if (((m >> sm) & 1) && k < keys[sm * IFR_TILE_PIX + tid]) keys[sm * IFR_TILE_PIX + tid] = k;
            }
        }
        __syncthreads();
        if (tid == 0) nbig = 0;
        __syncthreads();
    }

    // ---- quad shading + resolve -----------------------------------------------------------------
    uint32_t mySlots[S];
    float sampDepth = 0.0f;   // unused for S > 1
    (void)sampDepth;
#pragma unroll
    for (int s = 0; s < S; s++) mySlots[s] = key_slot(func, keys[s * IFR_TILE_PIX + tid]);

    float4 acc = make_float4(0, 0, 0, 0);
    float dacc = 0.0f;
    int bg = S;
    uint32_t last = IFR_SLOT_NONE;   // quad-uniform: last slot shaded (ascending order)
    bool first = true;
    unsigned int myFrags = 0;
    while (true) {
        uint32_t mine = IFR_SLOT_NONE;
#pragma unroll
        for (int s = 0; s < S; s++) {
            const uint32_t v = mySlots[s];
            if (v != IFR_SLOT_NONE && (first || v > last) && v < mine) mine = v;
        }
        uint32_t cur = min(mine, __shfl_xor_sync(0xffffffffu, mine, 1));
        cur = min(cur, __shfl_xor_sync(0xffffffffu, cur, 2));
        if (__ballot_sync(0xffffffffu, cur != IFR_SLOT_NONE) == 0) break;
        const bool active = cur != IFR_SLOT_NONE;
        float4 v[IFR_MAX_VARY];
        float depth = 0.0f;
        int prim = 0, fan = 0;
        if (active) interp(p, cur, px, py, v, depth, prim, fan);
        float4 ddx[IFR_MAX_VARY], ddy[IFR_MAX_VARY];
        if (FsTraits<FS>::needsDerivs) {
            for (int k = 0; k < p.nvary; k++) {
                // coarse derivatives: right - left of the row, bottom - top of the column (SB:36-76)
                const float4 a = v[k];
                float4 o1, o2;
                o1.x = __shfl_xor_sync(0xffffffffu, a.x, 1); o1.y = __shfl_xor_sync(0xffffffffu, a.y, 1);
                o1.z = __shfl_xor_sync(0xffffffffu, a.z, 1); o1.w = __shfl_xor_sync(0xffffffffu, a.w, 1);
                o2.x = __shfl_xor_sync(0xffffffffu, a.x, 2); o2.y = __shfl_xor_sync(0xffffffffu, a.y, 2);
                o2.z = __shfl_xor_sync(0xffffffffu, a.z, 2); o2.w = __shfl_xor_sync(0xffffffffu, a.w, 2);
                const bool right = l & 1, bottom = l & 2;
                ddx[k] = right ? make_float4(a.x - o1.x, a.y - o1.y, a.z - o1.z, a.w - o1.w)
                               : make_float4(o1.x - a.x, o1.y - a.y, o1.z - a.z, o1.w - a.w);
                ddy[k] = bottom ? make_float4(a.x - o2.x, a.y - o2.y, a.z - o2.z, a.w - o2.w)
                                : make_float4(o2.x - a.x, o2.y - a.y, o2.z - a.z, o2.w - a.w);
            }
        }
        if (active) {
            int cnt = 0;
#pragma unroll
            for (int s = 0; s < S; s++) cnt += mySlots[s] == cur;
            if (cnt && inFrame) {
                const float4 c = run_fs<FS>(p, v, ddx, ddy, prim, fan);
                const float fc = (float)cnt;
                acc.x = fmaf(fc, c.x, acc.x); acc.y = fmaf(fc, c.y, acc.y);
                acc.z = fmaf(fc, c.z, acc.z); acc.w = fmaf(fc, c.w, acc.w);
                dacc = fmaf(fc, depth, dacc);
                bg -= cnt;
                myFrags++;
                if (p.primIds) {
#pragma unroll
                    for (int s = 0; s < S; s++)
                        if (mySlots[s] == cur) p.primIds[pidx * S + s] = prim;
                }
            }
            last = cur;
            first = false;
        }
    }
    if (inFrame) {
        if (p.primIds) {
#pragma unroll
            for (int s = 0; s < S; s++)
                if (mySlots[s] == IFR_SLOT_NONE) p.primIds[pidx * S + s] = -1;
        }
        const bool cleared = p.doClear && scis;
        if (bg != S || cleared) {
            float4 out = acc;
            float dout = dacc;
            if (bg > 0) {
                const float fb = (float)bg;
                const float4 prev = cleared ? make_float4(p.clearColor[0], p.clearColor[1], p.clearColor[2], p.clearColor[3])
                                            : p.color[pidx];
                out.x = fmaf(fb, prev.x, out.x); out.y = fmaf(fb, prev.y, out.y);
                out.z = fmaf(fb, prev.z, out.z); out.w = fmaf(fb, prev.w, out.w);
                dout = fmaf(fb, d0, dout);
            }
            if (bg == S) {   // nothing drawn here: plain clear
                out = make_float4(p.clearColor[0], p.clearColor[1], p.clearColor[2], p.clearColor[3]);
                dout = p.clearDepth;
            } else {
                const float fs = (float)S;
                out.x = out.x / fs; out.y = out.y / fs; out.z = out.z / fs; out.w = out.w / fs;
                dout = dout / fs;
            }
            p.color[pidx] = out;
            p.depth[pidx] = dout;
        }
    }
    myFrags = __reduce_add_sync(0xffffffffu, myFrags);
    if ((tid & 31) == 0 && myFrags) atomicAdd(&p.counters[2], (unsigned long long)myFrags);   // no CTA barrier: warps retire independently
}

template <int FS>
static cudaError_t launch_fs(const DrawParams& p, cudaStream_t s) {
    const dim3 grid(p.tilesX, p.tilesY);
    const size_t smem = (size_t)p.msaa * IFR_TILE_PIX * sizeof(unsigned long long);
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

cudaError_t ifr_launch_tiles(const DrawParams& p, int fsId, cudaStream_t s) {
    switch (fsId) {
        case IFRIT_FS_VARYING0: return launch_fs<IFRIT_FS_VARYING0>(p, s);
        case IFRIT_FS_NORMAL_COLOR: return launch_fs<IFRIT_FS_NORMAL_COLOR>(p, s);
        case IFRIT_FS_PRIM_ID: return launch_fs<IFRIT_FS_PRIM_ID>(p, s);
        case IFRIT_FS_TEXTURED: return launch_fs<IFRIT_FS_TEXTURED>(p, s);
        case IFRIT_FS_TEXTURED_LOD0: return launch_fs<IFRIT_FS_TEXTURED_LOD0>(p, s);
        case IFRIT_FS_SKYBOX: return launch_fs<IFRIT_FS_SKYBOX>(p, s);
        case IFRIT_FS_DERIV_VIS: return launch_fs<IFRIT_FS_DERIV_VIS>(p, s);
        default: return cudaErrorInvalidValue;
    }
}
