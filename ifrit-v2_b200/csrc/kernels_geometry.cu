// kernels_geometry.cu — vertex shading, clip/cull/snap triangle setup, sort-middle binning.
//
// Replaces, B200-first (host-launched, count -> scan -> fill, exactly-sized queues):
//   vertexProcessingKernel            CU:330        -> vs_kernel
//   geometryClippingKernel            CU:1493/1343  -> setup_bin_kernel<false> (clip, viewport, snap, cull)
//   geometryParamPostprocKernel       CU:1549       -> folded into the tile kernel (48-byte snapped records
//                                                      are cheaper to stream than 112-byte setup records)
//   firstBinner* / *Alloc / *Gather   CU:721-1222   -> setup_bin_kernel<false> (count), scan_kernel,
//                                                      setup_bin_kernel<true> (fill)
// All arithmetic that decides coverage / visibility is integer (24.8 fixed point, 64-bit products) and
// follows oracle/ifrit_oracle.c op for op; float ops use explicit fmaf / rn intrinsics (-fmad=false).
#include <algorithm>
#include "common.cuh"
#include "tri_setup.cuh"
#include "../../include/ifrit_b200.h"

#define SNAP_LIMIT 268435456.0f

// ------------------------------------------------------------------------------------------------
// Vertex stage: one thread per vertex, 128-bit loads of the interleaved attributes, 128-bit stores.
// the vertex shader functors of the registry: run_vs, tri_setup.cuh
// Stand-alone vertex stage (used when the fused path is switched off, iftrcuSetVertexStageFused):
// one thread per vertex, 128-bit loads of the interleaved attributes, 128-bit stores.
template <int VS>
__global__ void __launch_bounds__(256) vs_kernel(const __grid_constant__ DrawParams p) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= p.nverts) return;
    const uint8_t* base = p.vb + (size_t)v * p.stride;
    p.pos[v] = run_vs<VS>(p, __ldg(reinterpret_cast<const float4*>(base + p.attrOff[0])));
    for (int k = 0; k < p.nvary; k++)
        p.vary[(size_t)v * p.nvary + k] = __ldg(reinterpret_cast<const float4*>(base + p.attrOff[k + 1]));
}

cudaError_t ifr_launch_vertex(const DrawParams& p, int vsId, cudaStream_t s) {
    if (p.nverts <= 0) return cudaSuccess;
    const int blocks = (p.nverts + 255) / 256;
    if (vsId == IFRIT_VS_MVP) vs_kernel<IFRIT_VS_MVP><<<blocks, 256, 0, s>>>(p);
    else vs_kernel<IFRIT_VS_PASSTHROUGH><<<blocks, 256, 0, s>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
struct ClipV { float4 p; float bx, by; };

__device__ __forceinline__ ClipV clip_lerp(const ClipV& a, const ClipV& b, float t) {
    ClipV r;
    r.p.x = fmaf(__fsub_rn(b.p.x, a.p.x), t, a.p.x);
    r.p.y = fmaf(__fsub_rn(b.p.y, a.p.y), t, a.p.y);
    r.p.z = fmaf(__fsub_rn(b.p.z, a.p.z), t, a.p.z);
    r.p.w = fmaf(__fsub_rn(b.p.w, a.p.w), t, a.p.w);
    r.bx = fmaf(__fsub_rn(b.bx, a.bx), t, a.bx);
    r.by = fmaf(__fsub_rn(b.by, a.by), t, a.by);
    return r;
}

__device__ __forceinline__ int snap(float ndc, float halfExtentSubpix) {
    float f = fmaf(ndc, halfExtentSubpix, halfExtentSubpix);
    f = fminf(fmaxf(f, -SNAP_LIMIT), SNAP_LIMIT);
    return __float2int_rn(f);
}

struct Assembled {
    int X[3], Y[3];
    float Z[3], IW[3];
    float bary[3][2];
    int flags;
    float ar;                 // 1 / float(d) of the stored (post-swap) winding
    int tx0, tx1, ty0, ty1;   // tile bbox (inclusive), valid when ok
    bool ok;
};

// viewport + snap + rejects + cull for one fan triangle (devGeometryCullClip CU:1399-1483)
template <bool ALLRANKS = false>
__device__ __forceinline__ void finish_triangle(const DrawParams& p, const ClipV& v0, const ClipV& v1,
                                                const ClipV& v2, bool clipped, Assembled& a) {
    a.ok = false;
    const bool cpu = p.convention == IFRIT_CONVENTION_CPU;
    const float hw = 128.0f * (float)p.W, hh = 128.0f * (float)p.H;
    const ClipV* v[3] = {&v0, &v1, &v2};
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float iw = __frcp_rn(v[j]->p.w);
        a.IW[j] = iw;
        a.Z[j] = __fmul_rn(v[j]->p.z, iw);
        a.X[j] = snap(__fmul_rn(v[j]->p.x, iw), hw);
        a.Y[j] = snap(__fmul_rn(v[j]->p.y, iw), hh);
        a.bary[j][0] = v[j]->bx;
        a.bary[j][1] = v[j]->by;
    }
    const int minx = min(a.X[0], min(a.X[1], a.X[2])), maxx = max(a.X[0], max(a.X[1], a.X[2]));
    const int miny = min(a.Y[0], min(a.Y[1], a.Y[2])), maxy = max(a.Y[0], max(a.Y[1], a.Y[2]));
    if (maxx < 0 || maxy < 0 || minx > p.W * IFR_SUBPIX || miny > p.H * IFR_SUBPIX) return;
    if (cpu) {
        const float zmin = fminf(a.Z[0], fminf(a.Z[1], a.Z[2])), zmax = fmaxf(a.Z[0], fmaxf(a.Z[1], a.Z[2]));
        if (zmax < 0.0f || zmin > 1.0f) return;
    }
    const long long d = (long long)(a.X[1] - a.X[0]) * (a.Y[2] - a.Y[0]) -
                        (long long)(a.X[2] - a.X[0]) * (a.Y[1] - a.Y[0]);
    if (d == 0) return;
    bool swp = false;
    switch (p.cullMode) {
        case IF_CULL_MODE_BACK: if (d < 0) return; break;
        case IF_CULL_MODE_FRONT: if (d > 0) return; swp = true; break;
        case IF_CULL_MODE_NONE: swp = d < 0; break;
        default: return;
    }
    a.flags = (clipped ? IFR_FLAG_CLIPPED : 0) | (swp ? IFR_FLAG_SWAPPED : 0);
    {
        // triangle-level setup shared by every tile: 1/d (sign follows the stored winding) and the
        // arithmetic class of the edge functions
        const long long dpost = swp ? -d : d;
        a.ar = __frcp_rn((float)dpost);
        long long span = 0;
        int wide = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int j = (i + 1) % 3;
            const long long adx = llabs((long long)a.X[j] - a.X[i]), ady = llabs((long long)a.Y[j] - a.Y[i]);
            span = max(span, adx + ady);
            wide |= (adx > (1 << 20)) | (ady > (1 << 20));
        }
        // |E_i| at any pixel corner within two pixels of a covered sample is below |d| + 1024 (|dX| + |dY|)
        if (llabs(d) + 1024 * span < 2147483647LL) a.flags |= IFR_FLAG_FITS32;
        if (wide) a.flags |= IFR_FLAG_WIDE;
    }
    if (swp) {
        int t;
        float f;
        t = a.X[0]; a.X[0] = a.X[2]; a.X[2] = t;
        t = a.Y[0]; a.Y[0] = a.Y[2]; a.Y[2] = t;
        f = a.Z[0]; a.Z[0] = a.Z[2]; a.Z[2] = f;
        f = a.IW[0]; a.IW[0] = a.IW[2]; a.IW[2] = f;
        f = a.bary[0][0]; a.bary[0][0] = a.bary[2][0]; a.bary[2][0] = f;
        f = a.bary[0][1]; a.bary[0][1] = a.bary[2][1]; a.bary[2][1] = f;
    }
    // exact contribution cull: pixel range that can contain a strictly-inside sample
    int px0 = ((minx - p.sampMaxX) >> 8) + 1, px1 = (maxx - p.sampMinX - 1) >> 8;
    int py0 = ((miny - p.sampMaxY) >> 8) + 1, py1 = (maxy - p.sampMinY - 1) >> 8;
    px0 = max(px0, 0); py0 = max(py0, 0);
    px1 = min(px1, p.W - 1); py1 = min(py1, p.H - 1);
    if (px0 > px1 || py0 > py1) return;
    const long long boxArea = (long long)(px1 - px0 + 1) * (py1 - py0 + 1);
    if (boxArea <= IFR_SMALL_AREA) a.flags |= IFR_FLAG_SMALL;
    a.tx0 = px0 >> 4; a.tx1 = px1 >> 4; a.ty0 = py0 >> 4; a.ty1 = py1 >> 4;
    if (!ALLRANKS && p.world > 1) {      // ALLRANKS: the classification pass of the distributed setup decides for every rank
        // sort-first: a triangle whose box touches no super-tile of this rank contributes nothing here — drop it
        // before the 64-byte record store and the fill pass (geometry is replicated; this keeps the replicated
        // part to reading it).  Boxes spanning >= world super-tiles in a row always contain an owned one.
        const int per = IFR_SUPER / IFR_TILE;
        const int sx0 = a.tx0 / per, sx1 = a.tx1 / per, sy0 = a.ty0 / per, sy1 = a.ty1 / per;
        bool owned = sx1 - sx0 + 1 >= p.world;
        for (int sy = sy0; sy <= sy1 && !owned; sy++)
            for (int sx = sx0; sx <= sx1 && !owned; sx++) owned = (sx + 3 * sy) % p.world == p.rank;
        if (!owned) return;
    }
    a.ok = true;
}

// clip one input triangle against the w plane; returns the number of fan triangles (0..2+)
// VSMODE: IFRIT_VS_* = vertex shader fused here (positions computed per triangle corner straight from the
// vertex buffer, no pos / vary round trip through HBM); -1 = positions precomputed by vs_kernel
template <int VSMODE>
__device__ __forceinline__ float4 fetch_pos(const DrawParams& p, int i) {
    if (VSMODE < 0) return __ldg(p.pos + i);
    return run_vs<VSMODE>(p, __ldg(reinterpret_cast<const float4*>(p.vb + (size_t)i * p.stride + p.attrOff[0])));
}

template <int VSMODE, bool ALLRANKS = false>
__device__ int assemble(const DrawParams& p, int t, Assembled out[2]) {
    const bool cpu = p.convention == IFRIT_CONVENTION_CPU;
    const int i0 = __ldg(p.ib + 3 * (size_t)t), i1 = __ldg(p.ib + 3 * (size_t)t + 1), i2 = __ldg(p.ib + 3 * (size_t)t + 2);
    ClipV src[3];
    src[0].p = fetch_pos<VSMODE>(p, i0); src[0].bx = 1.0f; src[0].by = 0.0f;
    src[1].p = fetch_pos<VSMODE>(p, i1); src[1].bx = 0.0f; src[1].by = 1.0f;
    src[2].p = fetch_pos<VSMODE>(p, i2); src[2].bx = 0.0f; src[2].by = 0.0f;
    const float w0 = src[0].p.w, w1 = src[1].p.w, w2 = src[2].p.w;
    out[0].ok = false; out[1].ok = false;
    if (!cpu) {
        if (w0 < 0 && w1 < 0 && w2 < 0) return 0;
    } else {
        const float mx = fmaxf(w0, fmaxf(w1, w2)), mn = fminf(w0, fminf(w1, w2));
        if (!(__fmul_rn(mx, mn) < 0) && !(mn > 0)) return 0;
    }
    if (w0 > 0 && w1 > 0 && w2 > 0) {
        // no sign change: Sutherland–Hodgman emits (v2, v3, v1) — CU:1361-1396
        finish_triangle<ALLRANKS>(p, src[1], src[2], src[0], false, out[0]);
        return 1;
    }
    ClipV poly[6];
    int np = 0;
    bool clipped = false;
    ClipV pc = src[0];
    for (int j = 0; j < 3; j++) {
        const ClipV& pn = src[(j + 1) % 3];
        const float npc = -pc.p.w, npn = -pn.p.w;
        if (__fmul_rn(npc, npn) < 0) {
            const float plane = cpu ? 1.0f : 1e-3f;
            const float tt = __fdiv_rn(__fsub_rn(pc.p.w, plane), __fsub_rn(pc.p.w, pn.p.w));
            poly[np++] = clip_lerp(pc, pn, tt);
            clipped = true;
        }
        if (cpu ? (npn < 1e-8f) : (npn < 0)) poly[np++] = pn;
        pc = pn;
    }
    if (np < 3) return 0;
    int nf = 0;
    for (int f = 0; f < np - 2 && f < 2; f++) {
        finish_triangle<ALLRANKS>(p, poly[0], poly[f + 1], poly[f + 2], clipped, out[f]);
        nf = f + 1;
    }
    return nf;
}

// ------------------------------------------------------------------------------------------------
// Binning.  A tile is kept for a triangle unless one edge function is <= 0 at every sample position
// of the tile (exact, 64-bit) — same predicate in the count and the fill pass.
struct EdgeSet { int dX[3], dY[3], Xa[3], Ya[3]; };

__device__ __forceinline__ void make_edges(const int X[3], const int Y[3], EdgeSet& e) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int a = (i + 1) % 3, b = (i + 2) % 3;
        e.dX[i] = X[b] - X[a]; e.dY[i] = Y[b] - Y[a]; e.Xa[i] = X[a]; e.Ya[i] = Y[a];
    }
}

__device__ __forceinline__ bool tile_may_cover(const EdgeSet& e, int tx, int ty) {
    // tile origin <= 2^21, |Xa|,|Ya| <= 2^28: the differences fit int32 and the products are single IMAD.WIDE
    const int sx0 = tx * (IFR_TILE * IFR_SUBPIX), sy0 = ty * (IFR_TILE * IFR_SUBPIX);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const long long E0 = (long long)e.dX[i] * (long long)(sy0 - e.Ya[i]) - (long long)e.dY[i] * (long long)(sx0 - e.Xa[i]);
        const long long add = 4080LL * ((long long)max(e.dX[i], 0) + (long long)max(-e.dY[i], 0));
        if (E0 + add <= 0) return false;
    }
    return true;
}

// One (triangle, tile) pair.  Count pass: bump the tile's counter of the triangle's class.  Fill
// pass: small class -> append the slot; big class -> compute the per-tile raster setup once, here,
// and append it, so that the tile kernel has no setup phase (and no barrier) for big triangles.
__device__ __forceinline__ bool tile_owned(const DrawParams& p, int tx, int ty) {
    if (p.world <= 1) return true;
    const int v = tx / (IFR_SUPER / IFR_TILE) + 3 * (ty / (IFR_SUPER / IFR_TILE));      // ifr_tile_owner without the runtime division
    return ((p.world & (p.world - 1)) == 0 ? (v & (p.world - 1)) : v % p.world) == p.rank;
}
// bit r set <=> rank r owns a tile of the box (ownership interleaves 64-px super-tiles, common.cuh: ifr_tile_owner; boxes of more
// than 16 super-tiles go to everyone) — a superset of the ranks finish_triangle keeps the triangle on
__device__ __forceinline__ uint32_t box_owner_mask(int world, int tx0, int tx1, int ty0, int ty1) {
    constexpr int ST = IFR_SUPER / IFR_TILE;
    const uint32_t all = world >= 32 ? 0xffffffffu : ((1u << world) - 1u);
    const int sx0 = tx0 / ST, sx1 = tx1 / ST, sy0 = ty0 / ST, sy1 = ty1 / ST;
    if ((sx1 - sx0 + 1) * (sy1 - sy0 + 1) > 16) return all;
    uint32_t m = 0;
    for (int sy = sy0; sy <= sy1; sy++)
        for (int sx = sx0; sx <= sx1; sx++) m |= 1u << ((sx + 3 * sy) % world);
    return m;
}
__device__ __forceinline__ void count_one(const DrawParams& p, int big, int tx, int ty) {
    if (!tile_owned(p, tx, ty)) return;
    atomicAdd(&p.tileCount[big * p.tilesX * p.tilesY + ty * p.tilesX + tx], 1u);
}
// reserve the queue position of a pair (the expensive part: an atomic round trip to L2)
__device__ __forceinline__ uint32_t reserve_one(const DrawParams& p, int big, int tx, int ty) {
    const int ntiles = p.tilesX * p.tilesY;
    const int tile = ty * p.tilesX + tx;
    return p.tileStart[big * (ntiles + 1) + tile] + atomicAdd(&p.tileCursor[big * ntiles + tile], 1u);
}
__device__ __forceinline__ void emit_one(const DrawParams& p, const RecRegs& rr, uint32_t slot, int big, int tx, int ty,
                                         uint32_t pos) {
    if (!big) {
        p.queue[pos] = slot;
    } else {
        TriS s;
        int bx0, bx1, by0, by1;
        const bool ok = setup_from_rec<true>(p, rr, slot, tx, ty, s, bx0, bx1, by0, by1);
        if (!ok) s.box = 0x00ff00ffu;   // empty: rows 255..0
        int4* dst = reinterpret_cast<int4*>(p.bigQueue + pos);
        const int4* src = reinterpret_cast<const int4*>(&s);
#pragma unroll
        for (int k = 0; k < 6; k++) dst[k] = src[k];
    }
}

#define IFR_BIG_TILES 16
// Warp-aggregated bin-queue atomics (north-star N2; the reference issues one atomic per (triangle, tile) pair,
// CU:863-871).  Lanes of a warp whose pairs fall into the same tile queue are found with __match_any_sync; the lowest
// lane of each group issues ONE atomic for the group and, in the fill pass, hands out the reserved positions by rank.
// MEASURED SLOWER on both kinds of scene (profiles/r02_atomics.txt: coherent 50k mesh, setup_count 21.0 -> 25.2 us; 10M
// random soup, setup_count 331 -> 390 us, light fill 149 -> 188 us): on the soup almost no two lanes share a tile, on
// the mesh the L2 already merges same-address REDs of a warp, and either way every pair pays a MATCH + vote + shuffle on
// top of its atomic.  Default OFF (per-pair RED / ATOM, which the L2 atomic units absorb); IFR_WARP_AGG = 1 builds the
// aggregated variant for the A/B.
#ifndef IFR_WARP_AGG
#define IFR_WARP_AGG 0
#endif
// all 32 lanes call; `act` lanes hold a queue index `q` (< 2^31).  Returns the group's size and the lane's rank in it.
__device__ __forceinline__ unsigned agg_group(bool act, uint32_t q, int lane, int& rank, int& leader) {
    const unsigned peers = __match_any_sync(0xffffffffu, act ? q : (0x80000000u | (uint32_t)lane));
    rank = __popc(peers & ((1u << lane) - 1u));
    leader = __ffs(peers) - 1;
    return peers;
}

// binning of boxes of <= 16 tiles by a group of G lanes that share the slot (lane `sub` of the group
// takes tiles sub, sub + G, ...; G = 1 in the count pass) + warp-cooperative binning of larger boxes
template <bool FILL, int G>
__device__ __forceinline__ void bin_triangle(const DrawParams& p, bool has, uint32_t slot, int big, int tx0, int tx1,
                                             int ty0, int ty1, const EdgeSet& e, int sub) {
    const int nx = tx1 - tx0 + 1, ny = ty1 - ty0 + 1;
    const int nt = has ? nx * ny : 0;
    if (IFR_WARP_AGG && !FILL) {
        // count pass, aggregated: the warp walks its boxes tile index by tile index in lockstep
        const int lane = threadIdx.x & 31;
        const int mynt = (nt >= 1 && nt <= IFR_BIG_TILES) ? nt : 0;
        const int maxnt = __reduce_max_sync(0xffffffffu, mynt);
        const bool test = nx > 1 && ny > 1;
        for (int i = 0; i < maxnt; i++) {
            bool act = i < mynt;
            int tx = 0, ty = 0;
            if (act) {
                tx = tx0 + i % nx; ty = ty0 + i / nx;
                act = !(test && !tile_may_cover(e, tx, ty)) && tile_owned(p, tx, ty);
            }
            const uint32_t q = (uint32_t)(big * p.tilesX * p.tilesY + ty * p.tilesX + tx);
            int rank, leader;
            const unsigned peers = agg_group(act, q, lane, rank, leader);
            if (act && lane == leader) atomicAdd(&p.tileCount[q], (uint32_t)__popc(peers));
        }
    } else if (nt >= 1 && nt <= IFR_BIG_TILES && sub < nt && (FILL || !(nx > 1 && ny > 1))) {
        RecRegs rr;
        if (FILL && big) load_rec(p, slot, rr);
        const bool test = nx > 1 && ny > 1;     // a one-tile-wide box is covered in every tile of it
        int tx = tx0 + sub % nx, ty = ty0 + sub / nx;
        for (int i = sub; i < nt; i += G) {
            // ownership first: under sort-first most tiles of a box belong to other ranks, and the exact test is the expensive part
            if (tile_owned(p, tx, ty) && !(test && !tile_may_cover(e, tx, ty))) {
                if (!FILL) atomicAdd(&p.tileCount[big * p.tilesX * p.tilesY + ty * p.tilesX + tx], 1u);
                else emit_one(p, rr, slot, big, tx, ty, reserve_one(p, big, tx, ty));
            }
            tx += G;
            while (tx > tx1) { tx -= nx; ty++; }
        }
    }
    // Warp-cooperative binning: one box at a time, a tile per lane.  Boxes of more than IFR_BIG_TILES tiles in both passes; in
    // the COUNT pass also every box that needs the exact test (two or more tiles in both directions) — there each thread holds
    // its own triangle, and a per-thread loop makes the whole warp wait for its largest box: on c5 (10 % of the triangles
    // span 4 - 16 tiles, three of them in an average warp) that loop was most of the 3 ms of the setup pass.
    unsigned wideBox = __ballot_sync(0xffffffffu, sub == 0 && (nt > IFR_BIG_TILES || (!FILL && nt >= 1 && nx > 1 && ny > 1)));
    const int lane = threadIdx.x & 31;
    while (wideBox) {
        const int src = __ffs(wideBox) - 1;
        wideBox &= wideBox - 1;
        const uint32_t s = __shfl_sync(0xffffffffu, slot, src);
        const int sbig = __shfl_sync(0xffffffffu, big, src);
        const int bx0 = __shfl_sync(0xffffffffu, tx0, src), bx1 = __shfl_sync(0xffffffffu, tx1, src);
        const int by0 = __shfl_sync(0xffffffffu, ty0, src), by1 = __shfl_sync(0xffffffffu, ty1, src);
        EdgeSet be;      // Xa / Ya are the vertices in rotated order: the deltas follow from them
#pragma unroll
        for (int i = 0; i < 3; i++) { be.Xa[i] = __shfl_sync(0xffffffffu, e.Xa[i], src); be.Ya[i] = __shfl_sync(0xffffffffu, e.Ya[i], src); }
#pragma unroll
        for (int i = 0; i < 3; i++) { be.dX[i] = be.Xa[(i + 1) % 3] - be.Xa[i]; be.dY[i] = be.Ya[(i + 1) % 3] - be.Ya[i]; }
        RecRegs wr;
        if (FILL && sbig) load_rec(p, s, wr);
        const int bnx = bx1 - bx0 + 1, total = bnx * (by1 - by0 + 1);
        int tx = bx0 + lane % bnx, ty = by0 + lane / bnx;
        for (int i = lane; i < total; i += 32) {
            if (tile_owned(p, tx, ty) && tile_may_cover(be, tx, ty)) {
                if (!FILL) atomicAdd(&p.tileCount[sbig * p.tilesX * p.tilesY + ty * p.tilesX + tx], 1u);
                else emit_one(p, wr, s, sbig, tx, ty, reserve_one(p, sbig, tx, ty));
            }
            tx += 32;
            while (tx > bx1) { tx -= bnx; ty++; }
        }
    }
}

__device__ __forceinline__ void store_record(const DrawParams& p, uint32_t slot, const Assembled& a) {
    int4* r = reinterpret_cast<int4*>(p.rec + slot);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        int4 v;
        v.x = a.X[j];
        v.y = a.Y[j];
        v.z = __float_as_int(a.Z[j]);
        v.w = __float_as_int(a.IW[j]);
        r[j] = v;
    }
    r[3] = make_int4(__float_as_int(a.ar), a.flags, 0, 0);
    if (a.flags & IFR_FLAG_CLIPPED) {
        float* b = p.clipBary + (size_t)slot * 6;
#pragma unroll
        for (int j = 0; j < 3; j++) { b[2 * j] = a.bary[j][0]; b[2 * j + 1] = a.bary[j][1]; }
    }
}

// pass 1: one thread per input triangle — clip, snap, cull, store the 64-byte record, count tiles
#ifndef IFR_SETUP_MINB
#define IFR_SETUP_MINB 3
#endif
// Sort-first with a DISTRIBUTED triangle setup (iftrcuClassifyTriangles): rank r assembles triangles [first, first + count)
// — 1 / world of the draw — only far enough to know which ranks own a tile of each triangle's box, and writes one owner-mask byte per
// triangle.  The ranks all-gather the masks (NCCL, world x ntri / world bytes), after which setup_count_kernel runs the assembly
// only for the triangles whose mask names this rank: the per-triangle arithmetic that every rank used to repeat for the whole
// draw (3 ms of c5 whatever the rank count) is divided by world, plus the overlap of triangles that straddle super-tiles.
template <int VSMODE>
__global__ void __launch_bounds__(256, IFR_SETUP_MINB) classify_kernel(const __grid_constant__ DrawParams p, int first, int count, uint8_t* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int t = first + i;
    Assembled a[2];
    a[0].ok = a[1].ok = false;
    assemble<VSMODE, true>(p, t, a);
    uint32_t m = 0;
#pragma unroll
    for (int f = 0; f < 2; f++)
        if (a[f].ok) m |= box_owner_mask(p.world, a[f].tx0, a[f].tx1, a[f].ty0, a[f].ty1);
    mask[t] = (uint8_t)m;
}
cudaError_t ifr_launch_classify(const DrawParams& p, int vsMode, int first, int count, uint8_t* mask, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const int blocks = (count + 255) / 256;
    if (vsMode == IFRIT_VS_MVP) classify_kernel<IFRIT_VS_MVP><<<blocks, 256, 0, s>>>(p, first, count, mask);
    else if (vsMode == IFRIT_VS_PASSTHROUGH) classify_kernel<IFRIT_VS_PASSTHROUGH><<<blocks, 256, 0, s>>>(p, first, count, mask);
    else classify_kernel<-1><<<blocks, 256, 0, s>>>(p, first, count, mask);
    return cudaGetLastError();
}

// the setup of one input triangle (all 32 lanes of a warp call together; `live` = this lane holds a triangle to assemble, t < ntri
// whenever `inRange`): records, boxes, fan list, tile counts.  Returns whether the first-fan slot goes to the heavy fill pass.
template <int VSMODE>
__device__ __forceinline__ bool setup_one(const DrawParams& p, int t, bool inRange, bool live, unsigned long long& nasm) {
    Assembled a[2];
    a[0].ok = a[1].ok = false;
    if (live) assemble<VSMODE>(p, t, a);
    bool heavy = false;
#pragma unroll
    for (int f = 0; f < 2; f++) {
        bool has = a[f].ok;
        const uint32_t slot = f == 0 ? (uint32_t)t : (uint32_t)(p.ntri + t);
        if (has) nasm++;
        const int big = has && !(a[f].flags & IFR_FLAG_SMALL);
        EdgeSet e;
        if (has) {
            store_record(p, slot, a[f]);
            make_edges(a[f].X, a[f].Y, e);
            if (f == 1) {
                const unsigned long long k = atomicAdd(&p.counters[4], 1ull);
                p.fan1List[k] = slot;
            }
        }
        if (inRange && (f == 0 || has)) {
            uint2 box;
            if (has) {
                box.x = (uint32_t)a[f].tx0 | ((uint32_t)a[f].tx1 << 16);
                box.y = (uint32_t)a[f].ty0 | ((uint32_t)a[f].ty1 << 16) | (big ? 0u : 0x80000000u);
            } else { box.x = 1u; box.y = 0u; }   // x0 = 1 > x1 = 0: empty
            p.tileBox[slot] = box;
        }
        if (f == 1 && __ballot_sync(0xffffffffu, has) == 0) continue;   // warp-uniform skip
        bin_triangle<false, 1>(p, has, slot, big, a[f].tx0, a[f].tx1, a[f].ty0, a[f].ty1, e, 0);
        // everything but single-tile small triangles of the first fan is left to the heavy fill pass (one bit per triangle;
        // second-fan slots are already listed in fan1List)
        if (f == 0) heavy = has && (big || a[f].tx1 > a[f].tx0 || a[f].ty1 > a[f].ty0);
    }
    return heavy;
}

template <int VSMODE>
__global__ void __launch_bounds__(256, IFR_SETUP_MINB) setup_count_kernel(const __grid_constant__ DrawParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nasm = 0;
    const bool heavy = setup_one<VSMODE>(p, t, t < p.ntri, t < p.ntri, nasm);
    // one bit per triangle, one plain 4-byte store per warp (no atomics)
    const unsigned hm = __ballot_sync(0xffffffffu, heavy);
    if ((threadIdx.x & 31) == 0) p.heavyBits[t >> 5] = hm;
    // block-aggregated statistics
    nasm = __reduce_add_sync(0xffffffffu, (unsigned)nasm);
    if ((threadIdx.x & 31) == 0 && nasm) ifr_count_spread(p.counters, 0, blockIdx.x * 8u + (threadIdx.x >> 5), nasm);
}

// The same pass under the distributed setup (DrawParams::triMask: some rank has established which ranks own a tile of each
// triangle): only 1 / world .. a few / world of the triangles are this rank's, scattered at random through the draw (c5 is a
// soup), so one thread per triangle leaves most lanes of every warp idle through the whole assembly (ncu, rank 0 of 8 on c5:
// 9.6 of 32 threads active per instruction, 1.04 G warp instructions, 2.23 ms).  Here a CTA takes IFR_MASKED_PER_CTA consecutive
// triangles, compacts the ones that name this rank into shared memory (empty boxes for the others on the way), and walks the
// compacted list with full warps.
#define IFR_MASKED_PER_CTA 2048
#ifndef IFR_MASKED_MINB
#define IFR_MASKED_MINB 4      // 64 registers (spills on the clip path); measured on c5 as rank 0 of 8: 0.96 ms, against 1.20 ms with 3 (80 registers) and 1.51 ms with 2 (112, no spills): the pass is bound by the latency of its gathers, occupancy wins
#endif
template <int VSMODE>
__global__ void __launch_bounds__(256, IFR_MASKED_MINB) setup_count_masked_kernel(const __grid_constant__ DrawParams p) {
    __shared__ int live[IFR_MASKED_PER_CTA];
    __shared__ unsigned heavyWords[IFR_MASKED_PER_CTA / 32];
    __shared__ int nLive;
    const int base = blockIdx.x * IFR_MASKED_PER_CTA;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) nLive = 0;
    if (tid < IFR_MASKED_PER_CTA / 32) heavyWords[tid] = 0u;
    __syncthreads();
    for (int i = tid; i < IFR_MASKED_PER_CTA; i += 256) {       // a warp reads 32 consecutive mask bytes per turn
        const int t = base + i;
        const bool in = t < p.ntri;
        const bool mine = in && ((p.triMask[t] >> p.rank) & 1);
        if (in && !mine) p.tileBox[t] = make_uint2(1u, 0u);     // not ours: empty box, no record, no bit
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        int pos = 0;
        if (lane == 0 && m) pos = atomicAdd(&nLive, __popc(m));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (mine) live[pos + __popc(m & ((1u << lane) - 1u))] = t;
    }
    __syncthreads();
    const int n = nLive;
    unsigned long long nasm = 0;
    for (int i0 = (tid >> 5) * 32; i0 < n; i0 += 256) {           // warp-uniform bound: setup_one votes and shuffles
        const bool has = i0 + lane < n;
        const int t = has ? live[i0 + lane] : 0;
        if (setup_one<VSMODE>(p, t, has, has, nasm)) atomicOr(&heavyWords[(t - base) >> 5], 1u << (t & 31));
    }
    __syncthreads();
    if (tid < IFR_MASKED_PER_CTA / 32 && base + tid * 32 < ((p.ntri + 31) & ~31)) p.heavyBits[(base >> 5) + tid] = heavyWords[tid];
    nasm = __reduce_add_sync(0xffffffffu, (unsigned)nasm);
    if (lane == 0 && nasm) ifr_count_spread(p.counters, 0, blockIdx.x * 8u + (tid >> 5), nasm);
}

// pass 3a: single-tile small triangles of the first fan — one thread per input triangle, nothing but
// tileBox -> {tileStart, cursor atomic} -> 4-byte store, so that it runs at full occupancy
__global__ void __launch_bounds__(256) fill_light_kernel(const __grid_constant__ DrawParams p) {
    if (p.counters[3]) return;   // a queue is too small: host re-issues the draw with larger ones
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#if IFR_WARP_AGG
    bool act = t < p.ntri;
    int tx0 = 0, ty0 = 0;
    if (act) {
        const uint2 box = p.tileBox[t];
        tx0 = box.x & 0xffff; ty0 = box.y & 0xffff;
        const int tx1 = box.x >> 16, ty1 = (box.y >> 16) & 0x7fff;
        act = tx0 == tx1 && ty0 == ty1 && (box.y >> 31) && tile_owned(p, tx0, ty0);   // else: empty, multi-tile or big (heavy list)
    }
    if (!__any_sync(0xffffffffu, act)) return;
    const int lane = threadIdx.x & 31;
    const uint32_t tile = (uint32_t)(ty0 * p.tilesX + tx0);
    int rank, leader;
    const unsigned peers = agg_group(act, tile, lane, rank, leader);
    uint32_t base = 0;
    if (act && lane == leader) base = atomicAdd(&p.tileCursor[tile], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (act) p.queue[p.tileStart[tile] + base + (uint32_t)rank] = (uint32_t)t;
#else
    if (t >= p.ntri) return;
    const uint2 box = p.tileBox[t];
    const int tx0 = box.x & 0xffff, tx1 = box.x >> 16, ty0 = box.y & 0xffff, ty1 = (box.y >> 16) & 0x7fff;
    if (tx0 != tx1 || ty0 != ty1 || !(box.y >> 31)) return;     // empty, multi-tile or big: heavy list
    if (!tile_owned(p, tx0, ty0)) return;
    p.queue[reserve_one(p, 0, tx0, ty0)] = (uint32_t)t;
#endif
}

// pass 3b: the heavy slots (multi-tile or big triangles via the per-warp bit words, second-fan slots via
// fan1List).  A warp takes 32 bit words (1024 triangles) at a time and compacts their set bits so
// that every lane gets a slot: no atomics, no list.
#define IFR_FILL_GROUP 4      // lanes that share one heavy slot (each takes every 4th tile of its box)
__device__ __forceinline__ void fill_heavy_slot(const DrawParams& p, bool has, uint32_t slot, int sub) {
    int tx0 = 0, tx1 = 0, ty0 = 0, ty1 = 0, big = 0;
    EdgeSet e;
    if (has) {
        const uint2 box = p.tileBox[slot];
        tx0 = box.x & 0xffff; tx1 = box.x >> 16; ty0 = box.y & 0xffff; ty1 = (box.y >> 16) & 0x7fff;
        big = !(box.y >> 31);
        has = tx0 <= tx1;
        if (has && (tx1 > tx0 || ty1 > ty0)) {
            const int4* r = reinterpret_cast<const int4*>(p.rec + slot);
            int X[3], Y[3];
#pragma unroll
            for (int j = 0; j < 3; j++) { const int4 v = __ldg(r + j); X[j] = v.x; Y[j] = v.y; }
            make_edges(X, Y, e);
        }
    }
    bin_triangle<true, IFR_FILL_GROUP>(p, has, slot, big, tx0, tx1, ty0, ty1, e, sub);
}

#ifndef IFR_FILL_MINB
#define IFR_FILL_MINB 4
#endif
__global__ void __launch_bounds__(128, IFR_FILL_MINB) fill_heavy_kernel(const __grid_constant__ DrawParams p, const int wordsPerWarp, const int split) {
    if (p.counters[3]) return;
    const unsigned long long nfan1 = p.counters[4];
    const long long nwords = (p.ntri + 31) / 32;
    const int lane = threadIdx.x & 31, grp = lane / IFR_FILL_GROUP, sub = lane % IFR_FILL_GROUP;
    constexpr int PER = 32 / IFR_FILL_GROUP;    // slots per warp pass
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // wordsPerWarp (1..32): small draws spread over many warps, large ones compact 1024 triangles per warp;
    // split (1 or 4): for small draws, `split` warps share one batch and take every split-th pass of it
    const long long warpId = gw / split, nwarps = (((long long)gridDim.x * blockDim.x) >> 5) / split;
    const int part = (int)(gw % split);
    for (long long wb = warpId * wordsPerWarp; wb < nwords; wb += nwarps * wordsPerWarp) {
        const uint32_t bits = (lane < wordsPerWarp && wb + lane < nwords) ? p.heavyBits[wb + lane] : 0u;
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const int excl = incl - cnt, total = __shfl_sync(0xffffffffu, incl, 31);
        for (int j = part * PER; j < total; j += PER * split) {
            const int item = j + grp;
            const bool has = item < total;
            int k = 0;   // number of words whose inclusive count is <= item = index of the word holding the item
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, incl, k + step - 1);
                if (has && v <= item) k += step;
            }
            const uint32_t wbits = __shfl_sync(0xffffffffu, bits, k);
            const int wexcl = __shfl_sync(0xffffffffu, excl, k);
            const uint32_t slot = has ? (uint32_t)((wb + k) * 32 + __fns(wbits, 0, item - wexcl + 1)) : 0u;
            fill_heavy_slot(p, has, slot, sub);
        }
    }
    for (unsigned long long i0 = (unsigned long long)gw * PER; i0 < nfan1; i0 += (unsigned long long)nwarps * split * PER) {
        const bool has = i0 + grp < nfan1;
        fill_heavy_slot(p, has, has ? p.fan1List[i0 + grp] : 0u, sub);
    }
}

// pass 2: exclusive scans over the tiles, four classes in one launch (blockIdx.y):
//   0 / 1  per-tile counts of the small / big queue  -> tileStart, cursor reset, totals -> counters[1] / counters[5],
//          overflow flag -> counters[3]
//   2      owned tiles with a non-empty queue         -> tileOrder[rank] = tileX | tileY << 16 (front of the work list), total -> counters[2]
//   3      owned tiles with nothing binned            -> tileOrder[ntiles - 1 - rank] = tile (back), total -> counters[7]
// Chained multi-CTA scan: CTA b scans 4096 items locally, publishes (epoch, total) and sums the totals of its
// predecessors.  The logical block index is a ticket taken on arrival (counters[8 + class], zeroed with the counters
// every draw), so every predecessor a CTA spins on has started before it: no assumption about the dispatch order of
// blockIdx or about the whole grid being resident (at most 32 CTAs per class at 8K).
#define IFR_SCAN_ITEMS 4096
__global__ void __launch_bounds__(1024) scan_kernel(const __grid_constant__ DrawParams p) {
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t prefix;
    __shared__ int ticket;
    const int n = p.tilesX * p.tilesY;
    const int cls = blockIdx.y, nblk = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) ticket = (int)atomicAdd(&p.counters[8 + cls], 1ull);
    __syncthreads();
    const int blk = ticket;
    volatile unsigned long long* partial = p.scanPartial + (size_t)cls * nblk;
    const int i0 = blk * IFR_SCAN_ITEMS + tid * 4;
    uint32_t v[4];
    if (cls < 2) {
        const uint32_t* count = p.tileCount + (size_t)cls * n;
        if (i0 + 3 < n && ((n & 3) == 0)) {
            const uint4 q = *reinterpret_cast<const uint4*>(count + i0);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = i0 + k < n ? count[i0 + k] : 0u;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = i0 + k;
            v[k] = 0u;
            if (i < n) {
                const bool work = (p.tileCount[i] | p.tileCount[n + i]) != 0u;
                const bool owned = p.world <= 1 || ifr_tile_owner(i % p.tilesX, i / p.tilesX, p.world) == p.rank;
                v[k] = (owned && (work == (cls == 2))) ? 1u : 0u;
            }
        }
    }
    const uint32_t sum = v[0] + v[1] + v[2] + v[3];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warpSums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = warpSums[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        warpSums[lane] = wi - w;
        const uint32_t total = __shfl_sync(0xffffffffu, wi, 31);
        if (lane == 0) {
            __threadfence();
            partial[blk] = ((unsigned long long)p.epoch << 32) | total;
        }
        // totals of the predecessors (spin until each has published this draw's epoch)
        uint32_t pre = 0;
        for (int b0 = 0; b0 < blk; b0 += 32) {
            uint32_t t = 0;
            if (b0 + lane < blk) {
                unsigned long long x;
                do { x = partial[b0 + lane]; } while ((uint32_t)(x >> 32) != p.epoch);
                t = (uint32_t)x;
            }
            pre += __reduce_add_sync(0xffffffffu, t);
        }
        if (lane == 0) prefix = pre;
        if (blk == nblk - 1 && lane == 0) {
            const uint32_t all = pre + total;
            if (cls < 2) {
                p.tileStart[(size_t)cls * (n + 1) + n] = all;
                p.counters[cls ? 5 : 1] = all;
                if (all > (cls ? p.bigCap : p.queueCap)) p.counters[3] = 1ull;
            } else {
                p.counters[cls == 2 ? 2 : 7] = all;
            }
        }
    }
    __syncthreads();
    uint32_t run = prefix + warpSums[wid] + (incl - sum);
    if (cls < 2) {
        uint32_t* start = p.tileStart + (size_t)cls * (n + 1);
        uint32_t* cursor = p.tileCursor + (size_t)cls * n;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n) { start[i0 + k] = run; cursor[i0 + k] = 0u; }
            run += v[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (v[k]) p.tileOrder[cls == 2 ? run : (uint32_t)n - 1u - run] = (uint32_t)((i0 + k) % p.tilesX) | ((uint32_t)((i0 + k) / p.tilesX) << 16);
            run += v[k];
        }
    }
}

cudaError_t ifr_launch_setup_count(const DrawParams& p, int vsMode, cudaStream_t s) {
    if (p.ntri <= 0) return cudaSuccess;
    if (p.triMask) {       // distributed setup: compacted walk over this rank's triangles
        const int mb = (p.ntri + IFR_MASKED_PER_CTA - 1) / IFR_MASKED_PER_CTA;
        if (vsMode == IFRIT_VS_MVP) setup_count_masked_kernel<IFRIT_VS_MVP><<<mb, 256, 0, s>>>(p);
        else if (vsMode == IFRIT_VS_PASSTHROUGH) setup_count_masked_kernel<IFRIT_VS_PASSTHROUGH><<<mb, 256, 0, s>>>(p);
        else setup_count_masked_kernel<-1><<<mb, 256, 0, s>>>(p);
        return cudaGetLastError();
    }
    const int blocks = (p.ntri + 255) / 256;
    if (vsMode == IFRIT_VS_MVP) setup_count_kernel<IFRIT_VS_MVP><<<blocks, 256, 0, s>>>(p);
    else if (vsMode == IFRIT_VS_PASSTHROUGH) setup_count_kernel<IFRIT_VS_PASSTHROUGH><<<blocks, 256, 0, s>>>(p);
    else setup_count_kernel<-1><<<blocks, 256, 0, s>>>(p);
    return cudaGetLastError();
}
cudaError_t ifr_launch_scan(const DrawParams& p, cudaStream_t s) {
    const int n = p.tilesX * p.tilesY;
    scan_kernel<<<dim3((n + IFR_SCAN_ITEMS - 1) / IFR_SCAN_ITEMS, 4), 1024, 0, s>>>(p);
    return cudaGetLastError();
}
cudaError_t ifr_launch_fill(const DrawParams& p, cudaStream_t s) {
    if (p.ntri <= 0) return cudaSuccess;
    fill_light_kernel<<<(p.ntri + 255) / 256, 256, 0, s>>>(p);
    const long long nwords = ((long long)p.ntri + 31) / 32;
    int wordsPerWarp = 1;
    while (wordsPerWarp < 32 && nwords / (wordsPerWarp * 2) >= 148 * 16 * IFR_FILL_GROUP) wordsPerWarp *= 2;
    const int split = wordsPerWarp == 1 ? 4 : 1;
    const int heavyBlocks = (int)std::min<long long>((nwords / wordsPerWarp + 1) * split / 4 + 1, 148 * 16);
    fill_heavy_kernel<<<heavyBlocks, 128, 0, s>>>(p, wordsPerWarp, split);
    return cudaGetLastError();
}

// ---- attribute planes of the FAST textured shaders ------------------------------------------------------------------------------
// The exact shaders interpolate every fragment from the triangle record, the three indices and the three vertices' varyings
// (tile_kernel.cuh: interp — 13 dependent loads and ~170 instructions per fragment with derivatives).  The perspective numerators
// N_i = beta_i * IW_i are AFFINE in the pixel position, and so are their sums with any per-vertex constant, so for the FAST
// shaders (colour within 1/255, not bit-exact) one thread per assembled slot writes, once per draw,
//     D = sum N_i,  U = sum N_i u_i,  V = sum N_i v_i,  Nrm = sum N_i n_i            (each as {value at the origin pixel, d/dx, d/dy})
// relative to an origin pixel in the middle of the triangle's on-screen box; a fragment then needs five 16-byte loads and
// (u, v) = (U, V) / D — the neighbours' values of ddx / ddy are one add away.  The normal is left undivided: `textured` only uses
// its direction (SB:394-417).  Computed in double from the exact integer edge functions; the clip-space barycentric mapping of
// clipped / swapped triangles (persp_weights) is folded into the per-vertex constants.
__global__ void __launch_bounds__(256) planes_kernel(const __grid_constant__ DrawParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t slot;
    if (i < p.ntri) {
        const uint2 box = p.tileBox[i];
        if ((box.x & 0xffffu) > (box.x >> 16)) return;          // nothing assembled in this slot
        slot = (uint32_t)i;
    } else {
        const int k = i - p.ntri;
        if (k >= p.ntri || (unsigned long long)k >= p.counters[4]) return;
        slot = p.fan1List[k];
    }
    const int4* r = reinterpret_cast<const int4*>(p.rec + slot);
    int X[3], Y[3];
    float IW[3];
#pragma unroll
    for (int j = 0; j < 3; j++) { const int4 v = r[j]; X[j] = v.x; Y[j] = v.y; IW[j] = __int_as_float(v.w); }
    const int4 w3 = r[3];
    const double ar = (double)__int_as_float(w3.x);
    const uint32_t flags = (uint32_t)w3.y;
    const uint32_t prim = slot >= (uint32_t)p.ntri ? slot - (uint32_t)p.ntri : slot;
    // clip-space barycentrics (u, v) of the record's three vertices: out = v0 u + v1 v + v2 (1 - u - v)
    float bu[3], bv[3];
    if (flags & IFR_FLAG_CLIPPED) {
        const float* g = p.clipBary + (size_t)slot * 6;
#pragma unroll
        for (int j = 0; j < 3; j++) { bu[j] = g[2 * j]; bv[j] = g[2 * j + 1]; }
    } else {
        const bool sw = flags & IFR_FLAG_SWAPPED;
        bu[0] = sw ? 1.0f : 0.0f; bv[0] = sw ? 0.0f : 1.0f;
        bu[1] = 0.0f; bv[1] = 0.0f;
        bu[2] = sw ? 0.0f : 1.0f; bv[2] = sw ? 1.0f : 0.0f;
    }
    // origin pixel: middle of the box clamped to the frame (keeps |fx|, |fy| of every shaded pixel small)
    int x0 = min(X[0], min(X[1], X[2])) >> 8, x1 = max(X[0], max(X[1], X[2])) >> 8;
    int y0 = min(Y[0], min(Y[1], Y[2])) >> 8, y1 = max(Y[0], max(Y[1], Y[2])) >> 8;
    x0 = max(0, min(x0, p.W - 1)); x1 = max(0, min(x1, p.W - 1)); y0 = max(0, min(y0, p.H - 1)); y1 = max(0, min(y1, p.H - 1));
    const int ox = (x0 + x1) >> 1, oy = (y0 + y1) >> 1;
    const long long sx = (long long)ox * IFR_SUBPIX, sy = (long long)oy * IFR_SUBPIX;
    double n0[3], nx[3], ny[3];
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const int a = (e + 1) % 3, b = (e + 2) % 3;
        const long long dx = (long long)X[b] - X[a], dy = (long long)Y[b] - Y[a];
        const long long E = dx * (sy - Y[a]) - dy * (sx - X[a]);
        const double k = ar * (double)IW[e];
        n0[e] = (double)E * k; nx[e] = -(double)(IFR_SUBPIX * dy) * k; ny[e] = (double)(IFR_SUBPIX * dx) * k;
    }
    // per-vertex constants of the record's vertices: normal (varying 0) and uv (varying 1)
    int idx[3];
#pragma unroll
    for (int j = 0; j < 3; j++) idx[j] = p.ib[3 * (size_t)prim + j];
    double c[3][5];
#pragma unroll
    for (int e = 0; e < 3; e++)
#pragma unroll
        for (int q = 0; q < 5; q++) c[e][q] = 0.0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (k >= p.nvary) continue;
        const uint8_t* vb = p.varyBase + p.varyOff[k];
        const float4 v0 = *reinterpret_cast<const float4*>(vb + (size_t)idx[0] * p.varyStride),
                     v1 = *reinterpret_cast<const float4*>(vb + (size_t)idx[1] * p.varyStride),
                     v2 = *reinterpret_cast<const float4*>(vb + (size_t)idx[2] * p.varyStride);
#pragma unroll
        for (int e = 0; e < 3; e++) {
            const double u = bu[e], v = bv[e], w = 1.0 - (u + v);
            if (k == 0) {
                c[e][2] = v0.x * u + v1.x * v + v2.x * w; c[e][3] = v0.y * u + v1.y * v + v2.y * w; c[e][4] = v0.z * u + v1.z * v + v2.z * w;
            } else {
                c[e][0] = v0.x * u + v1.x * v + v2.x * w; c[e][1] = v0.y * u + v1.y * v + v2.y * w;
            }
        }
    }
    float pl[18];        // D, U, V, Nx, Ny, Nz: {origin, d/dx, d/dy}
    pl[0] = (float)(n0[0] + n0[1] + n0[2]); pl[1] = (float)(nx[0] + nx[1] + nx[2]); pl[2] = (float)(ny[0] + ny[1] + ny[2]);
#pragma unroll
    for (int q = 0; q < 5; q++) {
        pl[3 + 3 * q] = (float)(n0[0] * c[0][q] + n0[1] * c[1][q] + n0[2] * c[2][q]);
        pl[4 + 3 * q] = (float)(nx[0] * c[0][q] + nx[1] * c[1][q] + nx[2] * c[2][q]);
        pl[5 + 3 * q] = (float)(ny[0] * c[0][q] + ny[1] * c[1][q] + ny[2] * c[2][q]);
    }
    float4* out = const_cast<float4*>(p.planes) + (size_t)slot * IFR_PLANE_QUADS;
    out[0] = make_float4(pl[0], pl[1], pl[2], pl[3]);
    out[1] = make_float4(pl[4], pl[5], pl[6], pl[7]);
    out[2] = make_float4(pl[8], pl[9], pl[10], pl[11]);
    out[3] = make_float4(pl[12], pl[13], pl[14], pl[15]);
    out[4] = make_float4(pl[16], pl[17], (float)ox, (float)oy);
}
cudaError_t ifr_launch_planes(const DrawParams& p, cudaStream_t s) {
    if (p.ntri <= 0 || !p.planes) return cudaSuccess;
    planes_kernel<<<(2 * (long long)p.ntri + 255) / 256, 256, 0, s>>>(p);
    return cudaGetLastError();
}
