// kernels_tile.cu — ordered alpha blending kernel and the fragment-shader dispatch of the tile kernel (tile_kernel.cuh).
#include "tile_kernel.cuh"

// ---- ordered alpha blending ----------------------------------------------------------------------
// pixelShadingAlphaBlendKernel CU:1649-1897: per pixel, primitives in submission order (the reference merges its three
// per-bin lists by original primitive id, CU:1853-1895; its bins are heap-sorted first, CU:563-719); a covered fragment
// is depth-tested against the depth the previous fragment left, shaded, blended into the colour and writes depth.
// Here: the tile's two queues are merged and sorted by (prim, fan) in shared memory (bitonic), then walked in batches
// of 64 — one thread per entry builds the raster setup, every thread tests its own pixel.  1x only, as the reference.
#define IFR_BLEND_CAP 4096
__device__ __forceinline__ bool depth_pass(int func, float d, float cur) {   // PS_DEPTHTEST, CU:1790-1800
    switch (func) {
        case IF_COMPARE_OP_NEVER: return false;
        case IF_COMPARE_OP_LESS: return d < cur;
        case IF_COMPARE_OP_EQUAL: return d == cur;
        case IF_COMPARE_OP_LESS_OR_EQUAL: return d <= cur;
        case IF_COMPARE_OP_GREATER: return d > cur;
        case IF_COMPARE_OP_NOT_EQUAL: return d != cur;
        case IF_COMPARE_OP_GREATER_OR_EQUAL: return d >= cur;
        default: return true;
    }
}
__device__ __forceinline__ float blend_factor(const float c[4], float oms, float omd) {   // CU:1817-1820, op order of the oracle
    return __fmul_rn(__fadd_rn(__fadd_rn(c[0], __fmul_rn(c[1], oms)), __fmul_rn(c[2], omd)), __fsub_rn(1.0f, c[3]));
}

template <int FS>
__global__ void __launch_bounds__(256, 2) tile_blend_kernel(const __grid_constant__ DrawParams p) {
    __shared__ uint32_t skey[IFR_BLEND_CAP];
    __shared__ __align__(16) TriS batch[64];
    __shared__ uint8_t batchOk[64];
    __shared__ uint32_t ctaSum;
    constexpr int NV = FsTraits<FS>::nv;
    const int tileX = blockIdx.x, tileY = blockIdx.y;
    if (p.world > 1 && ifr_tile_owner(tileX, tileY, p.world) != p.rank) return;
    const int ntiles = p.tilesX * p.tilesY;
    const int tile = tileY * p.tilesX + tileX;
    const int tid = threadIdx.x;
    const int wrp = tid >> 5, q = (tid >> 2) & 7, l = tid & 3;
    const int lx = (wrp & 1) * 8 + (q & 3) * 2 + (l & 1), ly = (wrp >> 1) * 4 + (q >> 2) * 2 + (l >> 1);
    const int px = tileX * IFR_TILE + lx, py = tileY * IFR_TILE + ly;
    const bool inFrame = px < p.W && py < p.H;
    const bool scis = inFrame && scissor_pass(p, px, py);
    const bool overflow = p.counters[3] != 0;
    const uint32_t sStart = p.tileStart[tile], bStart = p.tileStart[ntiles + 1 + tile];
    const uint32_t nSmall = overflow ? 0u : p.tileStart[tile + 1] - sStart;
    const uint32_t nBig = overflow ? 0u : p.tileStart[ntiles + 1 + tile + 1] - bStart;
    const size_t pidx = (size_t)py * p.W + px;
    const uint32_t nAll = nSmall + nBig;
    const bool cleared = p.doClear && scis;
    float4 dst = make_float4(p.clearColor[0], p.clearColor[1], p.clearColor[2], p.clearColor[3]);
    float cur = p.clearDepth;
    if (inFrame && !cleared) { dst = p.color[pidx]; cur = p.depth[pidx]; }
    int lastPrim = -1;
    bool touched = false;
    unsigned int myFrags = 0;
    // sort key = prim * 2 + fan: submission order (a slot appears at most once in a tile)
    auto key_of = [&](uint32_t i) -> uint32_t {
        const uint32_t slot = i < nSmall ? __ldg(p.queue + sStart + i) : p.bigQueue[bStart + (i - nSmall)].slot;
        return slot >= (uint32_t)p.ntri ? (slot - (uint32_t)p.ntri) * 2u + 1u : slot * 2u;
    };
    // entries with lo <= key < hi, summed over the CTA
    auto count_range = [&](uint32_t lo, uint32_t hi) -> uint32_t {
        uint32_t c = 0;
        for (uint32_t i = tid; i < nAll; i += 256) { const uint32_t k = key_of(i); c += (k >= lo && k < hi); }
        c = __reduce_add_sync(0xffffffffu, c);
        __syncthreads();
        if (tid == 0) ctaSum = 0;
        __syncthreads();
        if ((tid & 31) == 0 && c) atomicAdd(&ctaSum, c);
        __syncthreads();
        const uint32_t total = ctaSum;
        __syncthreads();          // everyone has read the sum before the word is reused (compute-sanitizer racecheck)
        return total;
    };
    // A tile with more than IFR_BLEND_CAP triangles is walked in rounds of consecutive key ranges [lo, hi), each holding at most
    // IFR_BLEND_CAP entries (hi found by bisection over the key space: ~30 counting passes over the tile's queues per round —
    // the spill path of very deep tiles; the usual tile takes the first branch and never counts).
    const uint32_t keyEnd = 4u * (uint32_t)p.ntri + 2u;      // above every key
    uint32_t lo = 0;
    for (;;) {
        uint32_t hi = keyEnd, n = nAll;
        if (nAll > IFR_BLEND_CAP) {
            n = count_range(lo, keyEnd);
            if (n > IFR_BLEND_CAP) {
                uint32_t a = lo + 1u, b = keyEnd;            // count(lo, a) <= 1: always feasible
                while (a < b) {
                    const uint32_t mid = a + (b - a + 1u) / 2u;
                    if (count_range(lo, mid) <= IFR_BLEND_CAP) a = mid; else b = mid - 1u;
                }
                hi = a;
                n = count_range(lo, hi);
            }
        } else if (lo) n = 0;
        if (!n) break;
        uint32_t npow2 = 2;
        while (npow2 < n) npow2 <<= 1;
        if (nAll <= IFR_BLEND_CAP) {
            for (uint32_t i = tid; i < npow2; i += 256) skey[i] = i < n ? key_of(i) : 0xffffffffu;
        } else {
            for (uint32_t i = tid; i < npow2; i += 256) skey[i] = 0xffffffffu;
            if (tid == 0) ctaSum = 0;
            __syncthreads();
            for (uint32_t i = tid; i < nAll; i += 256) {
                const uint32_t k = key_of(i);
                if (k >= lo && k < hi) skey[atomicAdd(&ctaSum, 1u)] = k;
            }
        }
        __syncthreads();
        for (uint32_t k = 2; k <= npow2; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = tid; i < npow2; i += 256) {
                    const uint32_t ixj = i ^ j;
                    if (ixj > i) {
                        const uint32_t a = skey[i], b = skey[ixj];
                        if ((a > b) == ((i & k) == 0)) { skey[i] = b; skey[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        for (uint32_t base = 0; base < n; base += 64) {
            const uint32_t cnt = min(64u, n - base);
            if ((uint32_t)tid < cnt) {
                const uint32_t key = skey[base + tid];
                const uint32_t slot = (key & 1u) ? (uint32_t)p.ntri + (key >> 1) : (key >> 1);
                TriS s;
                int bx0, bx1, by0, by1;
                batchOk[tid] = setup_tri(p, slot, tileX, tileY, s, bx0, bx1, by0, by1) ? 1 : 0;
                batch[tid] = s;
            }
            __syncthreads();
            for (uint32_t b = 0; b < cnt; b++) {
                if (!batchOk[b]) continue;
                const TriS& s = batch[b];
                const int bx0 = s.box & 0xff, bx1 = (s.box >> 8) & 0xff, by0 = (s.box >> 16) & 0xff, by1 = s.box >> 24;
                if (!scis || lx < bx0 || lx > bx1 || ly < by0 || ly > by1) continue;
                const int dX[3] = {s.dX[0], s.dX[1], s.dX[2]}, dY[3] = {s.dY[0], s.dY[1], s.dY[2]}, T[3] = {s.T[0], s.T[1], s.T[2]};
                const long long E0[3] = {s.E0[0], s.E0[1], s.E0[2]};
                if (!coverage<1>(p, dX, dY, T, E0, s.flags & IFR_FLAG_WIDE, 16 * lx, 16 * ly)) continue;
                const float d = depth_at(dX, dY, E0, s.flags, s.ar, s.z0, s.z1, s.z2, lx, ly);
                if (!depth_pass(p.depthFunc, d, cur)) continue;
                float4 v[NV > 0 ? NV : 1], ddx[NV > 0 ? NV : 1], ddy[NV > 0 ? NV : 1];
                float depth2 = 0.0f;
                int prim = 0, fan = 0;
                interp<NV, FsTraits<FS>::needsDerivs>(p, s.slot, px, py, v, ddx, ddy, depth2, prim, fan);
                const float4 c = run_fs<FS>(p, v, ddx, ddy, prim, fan);
                myFrags++;
                const float d2 = fs_depth<FS>(p, v, d);
                if (fs_modifies_depth(FS) && !depth_pass(p.depthFunc, d2, cur)) continue;     // late Z: second PS_DEPTHTEST, CU:1805
                const float oms = __fsub_rn(1.0f, c.w), omd = __fsub_rn(1.0f, dst.w);
                const float ms = blend_factor(p.blendCoef[0], oms, omd), md = blend_factor(p.blendCoef[1], oms, omd);
                const float msa = blend_factor(p.blendCoef[2], oms, omd), mda = blend_factor(p.blendCoef[3], oms, omd);
                dst.x = __fadd_rn(__fmul_rn(dst.x, md), __fmul_rn(c.x, ms));
                dst.y = __fadd_rn(__fmul_rn(dst.y, md), __fmul_rn(c.y, ms));
                dst.z = __fadd_rn(__fmul_rn(dst.z, md), __fmul_rn(c.z, ms));
                dst.w = __fadd_rn(__fmul_rn(dst.w, mda), __fmul_rn(c.w, msa));
                cur = d2;                    // CU:1828-1829: depth written, the next fragment tests against it
                lastPrim = prim;
                touched = true;
            }
            __syncthreads();
        }
        if (hi == keyEnd) break;
        lo = hi;
    }
    if (inFrame) {
        if (touched || cleared) {
            p.color[pidx] = dst;
            p.depth[pidx] = cur;
            if (p.present) p.present[pidx] = ifr_pack_rgba8(dst);
        } else if (p.present) {
            p.present[pidx] = ifr_pack_rgba8(p.color[pidx]);
        }
        if (p.primIds) p.primIds[pidx] = lastPrim;
    }
    myFrags = __reduce_add_sync(0xffffffffu, myFrags);
    if ((tid & 31) == 0 && myFrags) ifr_count_spread(p.counters, 1, (unsigned)(tile * 8 + wrp) * 2654435761u >> 16, (unsigned long long)myFrags);
}

template <int FS>
static cudaError_t launch_blend(const DrawParams& p, cudaStream_t s) {
    tile_blend_kernel<FS><<<dim3(p.tilesX, p.tilesY), 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t ifr_launch_tiles_blend(const DrawParams& p, int fsId, cudaStream_t s) {
    if (p.msaa != 1) return cudaErrorInvalidValue;
    switch (fsId) {
        case IFRIT_FS_VARYING0: return launch_blend<IFRIT_FS_VARYING0>(p, s);
        case IFRIT_FS_NORMAL_COLOR: return launch_blend<IFRIT_FS_NORMAL_COLOR>(p, s);
        case IFRIT_FS_PRIM_ID: return launch_blend<IFRIT_FS_PRIM_ID>(p, s);
        case IFRIT_FS_TEXTURED: return p.hwFilter ? launch_blend<IFRIT_FS_TEXTURED_FAST>(p, s) : launch_blend<IFRIT_FS_TEXTURED>(p, s);
        case IFRIT_FS_TEXTURED_LOD0: return p.hwFilter ? launch_blend<IFRIT_FS_TEXTURED_LOD0_FAST>(p, s) : launch_blend<IFRIT_FS_TEXTURED_LOD0>(p, s);
        case IFRIT_FS_SKYBOX: return launch_blend<IFRIT_FS_SKYBOX>(p, s);
        case IFRIT_FS_DERIV_VIS: return launch_blend<IFRIT_FS_DERIV_VIS>(p, s);
        case IFRIT_FS_DEPTH_MOD: return launch_blend<IFRIT_FS_DEPTH_MOD>(p, s);
        default: return cudaErrorInvalidValue;
    }
}

#define IFR_DECL_FS(ID) cudaError_t ifr_launch_fs_##ID(const DrawParams& p, cudaStream_t s);
IFR_DECL_FS(0) IFR_DECL_FS(1) IFR_DECL_FS(2) IFR_DECL_FS(3) IFR_DECL_FS(4) IFR_DECL_FS(5) IFR_DECL_FS(6)
IFR_DECL_FS(100) IFR_DECL_FS(101) IFR_DECL_FS(102) IFR_DECL_FS(103)
#undef IFR_DECL_FS

cudaError_t ifr_launch_tiles(const DrawParams& p, int fsId, cudaStream_t s) {
    // the FAST textured instantiations read attribute planes: the host must have run planes_kernel for this draw
    if (IFR_PLANES && p.hwFilter && !p.planes && (fsId == IFRIT_FS_TEXTURED || fsId == IFRIT_FS_TEXTURED_LOD0)) return cudaErrorInvalidValue;
    switch (fsId) {
        case IFRIT_FS_VARYING0: return ifr_launch_fs_0(p, s);
        case IFRIT_FS_NORMAL_COLOR: return ifr_launch_fs_1(p, s);
        case IFRIT_FS_PRIM_ID: return ifr_launch_fs_2(p, s);
#ifdef IFR_NO_FAST
        case IFRIT_FS_TEXTURED: return ifr_launch_fs_3(p, s);
        case IFRIT_FS_TEXTURED_LOD0: return ifr_launch_fs_4(p, s);
#else
        case IFRIT_FS_TEXTURED: {
            // the common sampler state has its own instantiation (sampler.cuh: sample_aniso_repeat)
            const SampDesc& sd = p.samp[0];
            const bool ra = sd.filter == IF_FILTER_LINEAR && sd.addrU == IF_SAMPLER_ADDRESS_MODE_REPEAT && sd.addrV == IF_SAMPLER_ADDRESS_MODE_REPEAT && sd.aniso;
            if (ra) return p.hwFilter ? ifr_launch_fs_103(p, s) : ifr_launch_fs_102(p, s);
            return p.hwFilter ? ifr_launch_fs_100(p, s) : ifr_launch_fs_3(p, s);
        }
        case IFRIT_FS_TEXTURED_LOD0: return p.hwFilter ? ifr_launch_fs_101(p, s) : ifr_launch_fs_4(p, s);
#endif
        case IFRIT_FS_SKYBOX: return ifr_launch_fs_5(p, s);
        case IFRIT_FS_DERIV_VIS: return ifr_launch_fs_6(p, s);
        default: return cudaErrorInvalidValue;
    }
}
