/* ifrit_oracle.c — TEST INFRASTRUCTURE ONLY.  See ifrit_oracle.h for the rules.
 *
 * A scalar CPU restatement of the softgraphics tile-raster path, written for
 * clarity, not speed: every triangle is rasterised by brute force over its
 * bounding box with exact 64-bit integer edge functions.  The product (the
 * CUDA library) must reproduce its coverage, primitive-id and depth outputs
 * bit-for-bit and its colour within 1/255.
 *
 * Build: see oracle/Makefile (-ffp-contract=off: every fused multiply-add in
 * this file is an explicit fmaf(), so the float op sequence is the spec).
 */
#define _GNU_SOURCE
#include "ifrit_oracle.h"
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>

#define SUBPIX 256              /* 24.8 fixed point */
#define SNAP_LIMIT 268435456.0f /* 2^28 sub-pixel units: guard band of the snapping grid */

/* ------------------------------------------------------------------------- */
/* MSAA sample positions, in 1/16 pixel.  Restates cMsaaSampleSequenceAgg
 * (CU:223-234) and its indexing `S-1+k` (CU:932-933): note that for 8x the
 * table yields the FIRST EIGHT ENTRIES OF THE 16x PATTERN, and 16x reads eight
 * zero-initialised entries past the initialiser list; both are replicated. */
static const int8_t k_sample_agg[31][2] = {
    {8, 8},
    {4, 4}, {12, 12},
    {6, 2}, {14, 6}, {2, 10}, {10, 14},
    {9, 9}, {7, 5}, {5, 10}, {12, 7}, {3, 6}, {10, 13}, {13, 11}, {11, 3},
    {6, 14}, {8, 1}, {4, 2}, {2, 12}, {0, 8}, {15, 4}, {14, 15}, {1, 0},
    {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};

static void sample_offset16(const orc_state* st, int k, int* ox, int* oy) {
    if (st->convention == ORC_CONV_CPU) { /* WK:1043-1055: the pixel's integer corner */
        *ox = 0; *oy = 0; return;
    }
    *ox = k_sample_agg[st->msaa - 1 + k][0];
    *oy = k_sample_agg[st->msaa - 1 + k][1];
}

/* ------------------------------------------------------------------------- */
/* Vertex stage: vertexProcessingKernel CU:330 / vertexProcessing WK:432.      */
int orc_vertex_stage(int vs_id, const void* uniforms, const uint8_t* vb, int stride,
                     const int32_t* attr_offsets, int nattr, int nvary, int nverts,
                     orc_f4* pos, orc_f4* vary) {
    const float* m = (const float*)uniforms;
    if (nvary > nattr - 1) return -1;
    for (int v = 0; v < nverts; v++) {
        const uint8_t* base = vb + (size_t)v * stride;
        orc_f4 a0;
        memcpy(&a0, base + attr_offsets[0], sizeof a0);
        if (vs_id == ORC_VS_MVP) {
            orc_f4 p;
            p.x = fmaf(m[3], a0.w, fmaf(m[2], a0.z, fmaf(m[1], a0.y, m[0] * a0.x)));
            p.y = fmaf(m[7], a0.w, fmaf(m[6], a0.z, fmaf(m[5], a0.y, m[4] * a0.x)));
            p.z = fmaf(m[11], a0.w, fmaf(m[10], a0.z, fmaf(m[9], a0.y, m[8] * a0.x)));
            p.w = fmaf(m[15], a0.w, fmaf(m[14], a0.z, fmaf(m[13], a0.y, m[12] * a0.x)));
            pos[v] = p;
        } else if (vs_id == ORC_VS_PASSTHROUGH) {
            pos[v] = a0;
        } else
            return -2;
        for (int k = 0; k < nvary; k++)
            memcpy(&vary[(size_t)v * nvary + k], base + attr_offsets[k + 1], sizeof(orc_f4));
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Clip + viewport + snap + cull: devGeometryCullClip CU:1343-1491,
 * geometryClippingKernel CU:1493-1524; CPU convention: triangleHomogeneousClip
 * WK:136-276, geometryProcessing WK:738-770, directTriangleSetup WK:503-727.  */
typedef struct { orc_f4 p; float bx, by; } clipv;

static clipv clip_lerp(const clipv* a, const clipv* b, float t) {
    clipv r;
    r.p.x = fmaf(b->p.x - a->p.x, t, a->p.x);
    r.p.y = fmaf(b->p.y - a->p.y, t, a->p.y);
    r.p.z = fmaf(b->p.z - a->p.z, t, a->p.z);
    r.p.w = fmaf(b->p.w - a->p.w, t, a->p.w);
    r.bx = fmaf(b->bx - a->bx, t, a->bx);
    r.by = fmaf(b->by - a->by, t, a->by);
    return r;
}

static int32_t snap(float ndc, float half_extent_subpix) {
    float f = fmaf(ndc, half_extent_subpix, half_extent_subpix); /* (ndc*0.5+0.5)*extent*256 */
    f = fminf(fmaxf(f, -SNAP_LIMIT), SNAP_LIMIT);
    return (int32_t)rintf(f); /* round half to even */
}

int orc_assemble(const orc_state* st, const orc_f4* pos, const int32_t* idx, int ntri, orc_tri* out) {
    const int cpu = st->convention == ORC_CONV_CPU;
    const float hw = 128.0f * (float)st->width, hh = 128.0f * (float)st->height;
    const int64_t wmax = (int64_t)st->width * SUBPIX, hmax = (int64_t)st->height * SUBPIX;
    int n = 0;
    for (int t = 0; t < ntri; t++) {
        clipv src[3];
        for (int j = 0; j < 3; j++) {
            src[j].p = pos[idx[3 * t + j]];
            src[j].bx = (j == 0) ? 1.0f : 0.0f;
            src[j].by = (j == 1) ? 1.0f : 0.0f;
        }
        const float w0 = src[0].p.w, w1 = src[1].p.w, w2 = src[2].p.w;
        clipv poly[6];
        int np = 0, clipped = 0;
        if (!cpu) {
            if (w0 < 0 && w1 < 0 && w2 < 0) continue; /* CU:1521 */
        } else {
            const float mx = fmaxf(w0, fmaxf(w1, w2)), mn = fminf(w0, fminf(w1, w2));
            if (!(mx * mn < 0) && !(mn > 0)) continue; /* WK:752-767 */
        }
        /* Sutherland–Hodgman against the single w plane, walking v1->v2->v3->v1 and
         * emitting into the list in the order (…, next) — CU:1361-1396 / WK:153-176 */
        {
            clipv pc = src[0];
            for (int j = 0; j < 3; j++) {
                const clipv* pn = &src[(j + 1) % 3];
                const float npc = -pc.p.w, npn = -pn->p.w;
                if (npc * npn < 0) {
                    const float plane = cpu ? 1.0f : 1e-3f;
                    const float t_ = (pc.p.w - plane) / (pc.p.w - pn->p.w);
                    poly[np++] = clip_lerp(&pc, pn, t_);
                    clipped = 1;
                }
                if (cpu ? (npn < 1e-8f) : (npn < 0)) poly[np++] = *pn;
                pc = *pn;
            }
        }
        if (np < 3) continue;
        for (int f = 0; f < np - 2; f++) {
            const clipv* v[3] = {&poly[0], &poly[f + 1], &poly[f + 2]};
            orc_tri r;
            r.prim = t; r.fan = f; r.clipped = clipped;
            float zmin = INFINITY, zmax = -INFINITY;
            for (int j = 0; j < 3; j++) {
                const float iw = 1.0f / v[j]->p.w; /* CU:1403 */
                r.IW[j] = iw;
                r.Z[j] = v[j]->p.z * iw;
                r.X[j] = snap(v[j]->p.x * iw, hw);
                r.Y[j] = snap(v[j]->p.y * iw, hh);
                r.bary[j][0] = v[j]->bx;
                r.bary[j][1] = v[j]->by;
                zmin = fminf(zmin, r.Z[j]);
                zmax = fmaxf(zmax, r.Z[j]);
            }
            /* screen-bbox reject: devViewSpaceClip CU:278-285 / triangleFrustumClip WK:277-299 */
            int64_t minx = r.X[0], maxx = r.X[0], miny = r.Y[0], maxy = r.Y[0];
            for (int j = 1; j < 3; j++) {
                if (r.X[j] < minx) minx = r.X[j];
                if (r.X[j] > maxx) maxx = r.X[j];
                if (r.Y[j] < miny) miny = r.Y[j];
                if (r.Y[j] > maxy) maxy = r.Y[j];
            }
            if (maxx < 0 || maxy < 0 || minx > wmax || miny > hmax) continue;
            if (cpu && (zmax < 0.0f || zmin > 1.0f)) continue; /* WK:283-286 */
            /* facing: devTriangleCull CU:271-276; d = X1(Y2-Y3)+X2(Y3-Y1)+X3(Y1-Y2), exact */
            const int64_t d = (int64_t)(r.X[1] - r.X[0]) * (r.Y[2] - r.Y[0]) -
                              (int64_t)(r.X[2] - r.X[0]) * (r.Y[1] - r.Y[0]);
            if (d == 0) continue; /* zero area covers no sample under the strict rule */
            int swap = 0;
            switch (st->cull_mode) { /* CU:1433-1462 */
                case ORC_CULL_BACK: if (d < 0) continue; break;
                case ORC_CULL_FRONT: if (d > 0) continue; swap = 1; break;
                case ORC_CULL_NONE: swap = d < 0; break;
                default: continue;
            }
            if (swap) {
#define SWP(T, a, b) do { T _t = a; a = b; b = _t; } while (0)
                SWP(int32_t, r.X[0], r.X[2]); SWP(int32_t, r.Y[0], r.Y[2]);
                SWP(float, r.Z[0], r.Z[2]); SWP(float, r.IW[0], r.IW[2]);
                SWP(float, r.bary[0][0], r.bary[2][0]); SWP(float, r.bary[0][1], r.bary[2][1]);
#undef SWP
            }
            /* contribution cull (CU:1464-1475, WK:258-262, WK:623-636): restated exactly as
             * "no sample row or column lies strictly inside the bounding box", which is
             * what the reference's rounding rules approximate.  It never changes the
             * image (a sample on the bbox boundary cannot be strictly inside). */
            {
                int omin_x = 255, omax_x = 0, omin_y = 255, omax_y = 0;
                const int ns = cpu ? 1 : st->msaa;
                for (int k = 0; k < ns; k++) {
                    int ox, oy;
                    sample_offset16(st, k, &ox, &oy);
                    ox *= 16; oy *= 16;
                    if (ox < omin_x) omin_x = ox;
                    if (ox > omax_x) omax_x = ox;
                    if (oy < omin_y) omin_y = oy;
                    if (oy > omax_y) omax_y = oy;
                }
                int64_t px0 = ((minx - omax_x) >> 8) + 1, px1 = (maxx - omin_x - 1) >> 8;
                int64_t py0 = ((miny - omax_y) >> 8) + 1, py1 = (maxy - omin_y - 1) >> 8;
                if (px0 < 0) px0 = 0;
                if (py0 < 0) py0 = 0;
                if (px1 > st->width - 1) px1 = st->width - 1;
                if (py1 > st->height - 1) py1 = st->height - 1;
                if (px0 > px1 || py0 > py1) continue;
            }
            r.slot = (f == 0) ? t : ntri + t;
            out[n++] = r;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* Texture sampling: SB:97-295.                                               */
static void level_info(const orc_texture* t, int layer, int lod, int* w, int* h, int64_t* off) {
    int64_t o = 0;
    int lw = t->width, lh = t->height;
    for (int i = 0; i < lod; i++) { /* textureLodPtroffImpl SB:97-105 */
        o += (int64_t)lw * lh * t->layers;
        lw = (lw + 1) >> 1;
        lh = (lh + 1) >> 1;
    }
    *w = lw; *h = lh; *off = o + (int64_t)lw * lh * layer;
}

/* exact float remainder of c by the positive size s (== fmodf for c >= 0, and == fmodf(c + s, s)
 * for -s <= c < 0, which is what SB:116-124 computes), written with a reciprocal multiply + one
 * fused multiply-add so that the device does not need the slow exact fmodf loop */
static float wrapf(float c, float s) {
    const float q = floorf(c * (1.0f / s));
    float r = fmaf(-q, s, c);
    if (r < 0.0f) r += s;
    if (r >= s) r -= s;
    return r;
}

static float addr_mode(int mode, float c, int size, int* border) {
    const float s = (float)size;
    switch (mode) { /* SB:116-160 */
        case ORC_ADDR_REPEAT:
            c = wrapf(c, s);
            break;
        case ORC_ADDR_MIRRORED_REPEAT:
            c = wrapf(c, 2 * s);
            c = fminf(c, 2 * s - c);
            break;
        case ORC_ADDR_CLAMP_TO_EDGE:
            c = (c >= s - 1.0f) ? s - 1.0f : ((c <= 0.0f) ? 0.0f : c);
            break;
        case ORC_ADDR_CLAMP_TO_BORDER:
            if (c >= s || c < 0.0f) *border = 1;
            break;
        case ORC_ADDR_MIRROR_CLAMP_TO_EDGE: /* mirrorclamp (sic) ShaderOpsCuda.cuh:36-39 */
            c = (c >= s - 1.0f) ? 0.0f : ((c <= 0.0f) ? s - 1.0f : c);
            break;
    }
    return c;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void texel(const orc_texture* t, int64_t off, int w, int x, int y, float o[4]) {
    memcpy(o, t->data + 4 * (off + (int64_t)y * w + x), 16);
}

/* textureImplLegacy SB:106-182 (lod clamped to the allocated chain: the reference reads
 * past the allocation for lod > mipLevels) */
static void sample_level(const orc_texture* t, const orc_sampler* s, int layer, float u, float v,
                         int offx, int offy, int lod, float out[4]) {
    int w, h;
    int64_t off;
    lod = clampi(lod, 0, t->mip_levels);
    level_info(t, layer, lod, &w, &h, &off);
    float cx = fmaf(u, (float)(w - 1), (float)offx);
    float cy = fmaf(v, (float)(h - 1), (float)offy);
    int border = 0;
    cx = addr_mode(s->addr_u, cx, w, &border);
    if (!border) cy = addr_mode(s->addr_v, cy, h, &border);
    if (border) {
        const float b = s->border == 0 ? 0.0f : 1.0f;
        out[0] = b; out[1] = b; out[2] = b; out[3] = 1.0f;
        return;
    }
    if (s->filter == ORC_FILTER_NEAREST) {
        int px = (int)roundf(cx), py = (int)roundf(cy);
        px = clampi(px % w, 0, w - 1);
        py = clampi(py % h, 0, h - 1);
        texel(t, off, w, px, py, out);
    } else {
        int px = clampi((int)cx, 0, w - 1), py = clampi((int)cy, 0, h - 1);
        const float fx = cx - (float)px, fy = cy - (float)py;
        const int px1 = px + 1 < w - 1 ? px + 1 : w - 1, py1 = py + 1 < h - 1 ? py + 1 : h - 1;
        float c00[4], c01[4], c10[4], c11[4];
        texel(t, off, w, px, py, c00);
        texel(t, off, w, px1, py, c01);
        texel(t, off, w, px, py1, c10);
        texel(t, off, w, px1, py1, c11);
        for (int c = 0; c < 4; c++) {
            const float a = fmaf(c01[c] - c00[c], fx, c00[c]);
            const float b = fmaf(c11[c] - c10[c], fx, c10[c]);
            out[c] = fmaf(b - a, fy, a);
        }
    }
}

/* textureImplLod SB:210-218 */
static void sample_trilinear(const orc_texture* t, const orc_sampler* s, int layer, float u, float v,
                             int offx, int offy, float lod, float out[4]) {
    const int lo = (int)floorf(lod);
    const float d = lod - (float)lo;
    float a[4], b[4];
    sample_level(t, s, layer, u, v, offx, offy, lo, a);
    if (d == 0.0f) { /* lerp(a, b, 0) == a for finite texels: the second level is not fetched */
        for (int c = 0; c < 4; c++) out[c] = a[c];
        return;
    }
    sample_level(t, s, layer, u, v, offx, offy, lo + 1, b);
    for (int c = 0; c < 4; c++) out[c] = fmaf(b[c] - a[c], d, a[c]);
}

/* textureImplAdaptiveLod SB:233-243 — `dyI.y` uses dx.y (sic, SB:237), replicated */
static void sample_adaptive(const orc_texture* t, const orc_sampler* s, float u, float v,
                            const float dx[2], const float dy[2], float bias, float out[4]) {
    const float W = (float)t->width, H = (float)t->height;
    const float ax = dx[0] * W, ay = dx[1] * H;
    const float bx = dy[0] * W, by = dx[1] * H;
    const float a2 = fmaf(ay, ay, ax * ax), b2 = fmaf(by, by, bx * bx);
    const float d = fmaxf(a2, b2);
    const float lod = fmaxf(0.0f, fmaf(0.5f, log2f(d), bias));
    sample_trilinear(t, s, 0, u, v, 0, 0, lod, out);
}

/* textureImplAnisotropicFilter SB:244-272, called as (du = ddx(uv), dv = ddy(uv)) SB:285-293 */
static void sample_aniso(const orc_texture* t, const orc_sampler* s, float u, float v,
                         const float du[2], const float dv[2], float bias, float out[4]) {
    const float dudx = du[0], dudy = du[1], dvdx = dv[0], dvdy = dv[1];
    const float px = sqrtf(fmaf(dvdx, dvdx, dudx * dudx)) * (float)t->width;
    const float py = sqrtf(fmaf(dvdy, dvdy, dudy * dudy)) * (float)t->height;
    const float pmax = fmaxf(px, py), pmin = fminf(px, py);
    const int max_aniso = (int)s->max_aniso;
    const int n = (int)fmax(1.0, fmin((double)ceilf(pmax / pmin) + 1e-3, (double)max_aniso + 1e-3));
    const float lambda = fmaxf(0.0f, log2f(pmax / (float)n));
    const float ddu = (px > py) ? dudx : dudy, ddv = (px > py) ? dvdx : dvdy;
    float sum[4] = {0, 0, 0, 0};
    const float step = 1.0f / (float)(n + 1), inv_n = 1.0f / (float)n;
    for (int i = 0; i < n; i++) {
        const float tt = (float)i * step - 0.5f;
        float r[4];
        sample_trilinear(t, s, 0, fmaf(ddu, tt, u), fmaf(ddv, tt, v), 0, 0, lambda + bias, r);
        for (int c = 0; c < 4; c++) sum[c] += r[c];
    }
    for (int c = 0; c < 4; c++) out[c] = sum[c] * inv_n;
}

/* textureCubeCoordConversion SB:184-208, textureCubeImplLod SB:220-231 */
static void sample_cube(const orc_texture* t, const orc_sampler* s, float x, float y, float z,
                        float lod, float out[4]) {
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
    sample_trilinear(t, s, face, u, v, 0, 0, lod, out);
}

void orc_texture_lod(const orc_texture* t, const orc_sampler* s, float u, float v, float lod, float out[4]) {
    sample_trilinear(t, s, 0, u, v, 0, 0, lod, out);
}
void orc_texture_grad(const orc_texture* t, const orc_sampler* s, float u, float v,
                      const float ddx[2], const float ddy[2], float bias, float out[4]) {
    if (s->aniso_enable) sample_aniso(t, s, u, v, ddx, ddy, bias, out);
    else sample_adaptive(t, s, u, v, ddx, ddy, bias, out);
}
void orc_texture_cube_lod(const orc_texture* t, const orc_sampler* s, float x, float y, float z,
                          float lod, float out[4]) {
    sample_cube(t, s, x, y, z, lod, out);
}

/* Batch form of the sampler entry points, same `kind` / argument layout as oracle/ref_sampler_shim.cu
 * (the reference's own ShaderBuiltinCuda.cuh compiled for the host): 0 = one integer level (a[0] = lod, a[1] = layer),
 * 1 = trilinear (a[0] = lod), 2 = adaptive lod (dx = a[0..1], dy = a[2..3], bias = a[4]), 3 = anisotropic
 * (du = a[0..1], dv = a[2..3], bias = a[4]; max anisotropy from the sampler), 4 = cube (x, y = uv, z = a[1], lod = a[0]). */
int orc_sample_batch(int kind, const orc_texture* t, const orc_sampler* s, int n, const float* uv, const float* args, float* out) {
    for (int i = 0; i < n; i++) {
        const float* a = args + 6 * (size_t)i;
        const float u = uv[2 * (size_t)i], v = uv[2 * (size_t)i + 1];
        float* o = out + 4 * (size_t)i;
        switch (kind) {
            case 0: sample_level(t, s, (int)a[1], u, v, 0, 0, (int)a[0], o); break;
            case 1: sample_trilinear(t, s, 0, u, v, 0, 0, a[0], o); break;
            case 2: sample_adaptive(t, s, u, v, a, a + 2, a[4], o); break;
            case 3: sample_aniso(t, s, u, v, a, a + 2, a[4], o); break;
            case 4: sample_cube(t, s, u, v, a[1], a[0], o); break;
            default: return -1;
        }
    }
    return 0;
}

int64_t orc_texture_total_texels(int w, int h, int mip_levels, int layers) {
    int64_t n = 0;
    for (int i = 0; i <= mip_levels; i++) {
        n += (int64_t)w * h * layers;
        w = (w + 1) >> 1;
        h = (h + 1) >> 1;
    }
    return n;
}

/* invokeMipmapGeneration CUIMG:132-155 + blitImage{Bilinear,Nearest}Kernel CUIMG:36-86 */
void orc_generate_mipmap(float* data, int w, int h, int mip_levels, int layers, int filter) {
    orc_texture t = {data, w, h, mip_levels, layers};
    int sw = w, sh = h;
    for (int l = 0; l < mip_levels; l++) {
        const int dw = (sw + 1) >> 1, dh = (sh + 1) >> 1;
        for (int layer = 0; layer < layers; layer++) {
            int a, b;
            int64_t soff, doff;
            level_info(&t, layer, l, &a, &b, &soff);
            level_info(&t, layer, l + 1, &a, &b, &doff);
            for (int y = 0; y < dh; y++)
                for (int x = 0; x < dw; x++) {
                    const float px = (float)x / (float)dw, py = (float)y / (float)dh;
                    const float sx = px * (float)sw, sy = py * (float)sh;
                    float* dst = data + 4 * (doff + (int64_t)y * dw + x);
                    if (filter == ORC_FILTER_LINEAR) {
                        const int ix = (int)sx, iy = (int)sy;
                        const float fx = sx - (float)ix, fy = sy - (float)iy;
                        const int ix1 = ix + 1 < sw - 1 ? ix + 1 : sw - 1, iy1 = iy + 1 < sh - 1 ? iy + 1 : sh - 1;
                        const float* c00 = data + 4 * (soff + (int64_t)iy * sw + ix);
                        const float* c01 = data + 4 * (soff + (int64_t)iy * sw + ix1);
                        const float* c10 = data + 4 * (soff + (int64_t)iy1 * sw + ix);
                        const float* c11 = data + 4 * (soff + (int64_t)iy1 * sw + ix1);
                        for (int c = 0; c < 4; c++) {
                            const float p = fmaf(c01[c] - c00[c], fx, c00[c]);
                            const float q = fmaf(c11[c] - c10[c], fx, c10[c]);
                            dst[c] = fmaf(q - p, fy, p);
                        }
                    } else {
                        int ix = (int)roundf(sx), iy = (int)roundf(sy);
                        if (ix > sw - 1) ix = sw - 1;
                        if (iy > sh - 1) iy = sh - 1;
                        memcpy(dst, data + 4 * (soff + (int64_t)iy * sw + ix), 16);
                    }
                }
        }
        sw = dw; sh = dh;
    }
}

/* ------------------------------------------------------------------------- */
/* Fragment shaders of the registry (ids shared with include/ifrit_b200.h).   */
typedef struct {
    const float* u;            /* uniform block (buffer slot 0) */
    const orc_texture* tex;
    const orc_sampler* samp;
} fs_env;

static void run_fs(int fs_id, const fs_env* e, const orc_f4* v, const orc_f4* ddx, const orc_f4* ddy,
                   int prim, int fan, float out[4]) {
    switch (fs_id) {
        case ORC_FS_VARYING0:
            out[0] = v[0].x; out[1] = v[0].y; out[2] = v[0].z; out[3] = v[0].w; break;
        case ORC_FS_NORMAL_COLOR: /* DemoFragmentShaderCuda, SG/demo/shader/DefaultDemoShaders.cu */
            out[0] = fmaf(v[0].x, 0.5f, 0.5f); out[1] = fmaf(v[0].y, 0.5f, 0.5f);
            out[2] = fmaf(v[0].z, 0.5f, 0.5f); out[3] = 0.5f; break;
        case ORC_FS_PRIM_ID:
            out[0] = (float)prim; out[1] = (float)fan; out[2] = 0.0f; out[3] = 1.0f; break;
        case ORC_FS_TEXTURED:
        case ORC_FS_TEXTURED_LOD0: {
            float c[4];
            if (fs_id == ORC_FS_TEXTURED) {
                const float dx[2] = {ddx[1].x, ddx[1].y}, dy[2] = {ddy[1].x, ddy[1].y};
                orc_texture_grad(&e->tex[0], &e->samp[0], v[1].x, v[1].y, dx, dy, e->u[20], c);
            } else
                orc_texture_lod(&e->tex[0], &e->samp[0], v[1].x, v[1].y, 0.0f, c);
            const float nx = v[0].x, ny = v[0].y, nz = v[0].z;
            const float len = sqrtf(fmaf(nz, nz, fmaf(ny, ny, nx * nx)));
            float ndl = fmaf(nz, e->u[18], fmaf(ny, e->u[17], nx * e->u[16])) / len;
            ndl = fmaxf(ndl, 0.2f);
            out[0] = c[0] * ndl; out[1] = c[1] * ndl; out[2] = c[2] * ndl; out[3] = 1.0f;
            break;
        }
        case ORC_FS_SKYBOX:
            orc_texture_cube_lod(&e->tex[0], &e->samp[0], v[0].x, v[0].y, v[0].z, e->u[20], out); break;
        case ORC_FS_DERIV_VIS:
            out[0] = ddx[1].x * e->u[21]; out[1] = ddx[1].y * e->u[21];
            out[2] = ddy[1].x * e->u[21]; out[3] = ddy[1].y * e->u[21]; break;
        case ORC_FS_DEPTH_MOD:      /* colour = varying 0; writes depth (fs_depth) */
            out[0] = v[0].x; out[1] = v[0].y; out[2] = v[0].z; out[3] = v[0].w; break;
        default: out[0] = out[1] = out[2] = out[3] = 0.0f;
    }
}

/* Fragment shaders that write depth (reference: FragmentShader::allowDepthModification, Shaders.h:61; the shader receives
 * `float* fragmentDepth`, CU:1803).  Such a draw takes the ordered path (CU:2517-2527: forceSort, pixelShadingAlphaBlendKernel)
 * with the depth test run twice: before the shader on the interpolated depth, after it on the depth the shader left. */
static int fs_modifies_depth(int fs_id) { return fs_id == ORC_FS_DEPTH_MOD; }
static float fs_depth(int fs_id, const fs_env* e, const orc_f4* v, float depth) {
    if (fs_id == ORC_FS_DEPTH_MOD) return fmaf(e->u[22], v[0].x, depth);      /* depth + u[22] * varying0.x */
    return depth;
}

/* ------------------------------------------------------------------------- */
/* Visibility keys.  Depth-test outcome per sample is the minimum of a 64-bit
 * key over all covering triangles, which makes the result independent of the
 * order triangles are processed in (pixelTaggingExecKernel CU:1901-2004 is
 * order dependent; its `ALWAYS keeps the lowest primitive id` rule, CU:1985-1992,
 * is kept and extended to break every tie deterministically).                */
static uint32_t mono(float d) {
    d = d + 0.0f; /* -0 -> +0 */
    uint32_t b;
    memcpy(&b, &d, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static float unmono(uint32_t m) {
    uint32_t b = (m & 0x80000000u) ? (m & 0x7fffffffu) : ~m;
    float d;
    memcpy(&d, &b, 4);
    return d;
}
static uint32_t fbits(float d) { uint32_t b; memcpy(&b, &d, 4); return b; }

static uint64_t key_init(int func, float d0) {
    switch (func) {
        case ORC_CMP_LESS: return ((uint64_t)mono(d0) << 32) | 0u;
        case ORC_CMP_LESS_OR_EQUAL: return ((uint64_t)mono(d0) << 32) | 0xffffffffu;
        case ORC_CMP_GREATER: return ((uint64_t)(~mono(d0)) << 32) | 0u;
        case ORC_CMP_GREATER_OR_EQUAL: return ((uint64_t)(~mono(d0)) << 32) | 0xffffffffu;
        default: return ((uint64_t)0xffffffffu << 32) | fbits(d0 + 0.0f); /* ALWAYS / EQUAL */
    }
}
/* returns 0 when the fragment cannot pass at all */
static int key_make(int func, float d, float d0, uint32_t slot, uint64_t* k) {
    switch (func) {
        case ORC_CMP_LESS: *k = ((uint64_t)mono(d) << 32) | (slot + 1u); return 1;
        case ORC_CMP_LESS_OR_EQUAL: *k = ((uint64_t)mono(d) << 32) | ~(slot + 1u); return 1;
        case ORC_CMP_GREATER: *k = ((uint64_t)(~mono(d)) << 32) | (slot + 1u); return 1;
        case ORC_CMP_GREATER_OR_EQUAL: *k = ((uint64_t)(~mono(d)) << 32) | ~(slot + 1u); return 1;
        case ORC_CMP_ALWAYS: *k = ((uint64_t)(slot + 1u) << 32) | fbits(d + 0.0f); return 1;
        case ORC_CMP_EQUAL:
            if (!(d == d0)) return 0;
            *k = ((uint64_t)(slot + 1u) << 32) | fbits(d + 0.0f); return 1;
        default: return 0; /* NEVER; NOT_EQUAL is order dependent and unsupported */
    }
}
static int key_slot(int func, uint64_t k, uint32_t* slot) {
    uint32_t s;
    switch (func) {
        case ORC_CMP_LESS: case ORC_CMP_GREATER: s = (uint32_t)k; if (s == 0) return 0; *slot = s - 1; return 1;
        case ORC_CMP_LESS_OR_EQUAL: case ORC_CMP_GREATER_OR_EQUAL:
            s = ~(uint32_t)k; if (s == 0) return 0; *slot = s - 1; return 1;
        default: s = (uint32_t)(k >> 32); if (s == 0xffffffffu) return 0; *slot = s - 1; return 1;
    }
}
static float key_depth(int func, uint64_t k) {
    switch (func) {
        case ORC_CMP_LESS: case ORC_CMP_LESS_OR_EQUAL: return unmono((uint32_t)(k >> 32));
        case ORC_CMP_GREATER: case ORC_CMP_GREATER_OR_EQUAL: return unmono(~(uint32_t)(k >> 32));
        default: { uint32_t b = (uint32_t)k; float d; memcpy(&d, &b, 4); return d; }
    }
}

static int scissor_pass(const orc_state* st, int x, int y) { /* CU:1946-1957, CU:3345-3357 */
    if (!st->scissor_enable || st->nscissor == 0) return 1;
    for (int i = 0; i < st->nscissor; i++) {
        const float* s = st->scissor[i];
        if ((float)x >= s[0] && (float)x <= s[0] + s[2] && (float)y >= s[1] && (float)y <= s[1] + s[3]) return 1;
    }
    return 0;
}

void orc_clear(const orc_state* st, float* color, float* depth) { /* CU:3337-3391 */
    for (int y = 0; y < st->height; y++)
        for (int x = 0; x < st->width; x++) {
            if (!scissor_pass(st, x, y)) continue;
            const size_t i = (size_t)y * st->width + x;
            memcpy(color + 4 * i, st->clear_color, 16);
            depth[i] = st->clear_depth;
        }
}

/* exact edge functions at a point given in sub-pixel units; E_i > 0 for all i <=> inside
 * (strict `<` of CU:951-953, no top-left rule) */
static void edges_at(const orc_tri* r, int64_t sx, int64_t sy, int64_t e[3]) {
    /* e[0] is opposite vertex 0 (edge v1->v2), e[1] opposite vertex 1, e[2] opposite vertex 2 */
    e[0] = (int64_t)(r->X[2] - r->X[1]) * (sy - r->Y[1]) - (int64_t)(r->Y[2] - r->Y[1]) * (sx - r->X[1]);
    e[1] = (int64_t)(r->X[0] - r->X[2]) * (sy - r->Y[2]) - (int64_t)(r->Y[0] - r->Y[2]) * (sx - r->X[2]);
    e[2] = (int64_t)(r->X[1] - r->X[0]) * (sy - r->Y[0]) - (int64_t)(r->Y[1] - r->Y[0]) * (sx - r->X[0]);
}

/* screen-space barycentrics at the INTEGER pixel (px,py): the reference evaluates depth and
 * attributes at the pixel corner while coverage is taken at the sample (CU:1948 vs CU:934);
 * replicated.  beta_i = E_i(p) / d, from the exact integer edge functions. */
static void bary_at(const orc_tri* r, int px, int py, float beta[3]) {
    int64_t e[3];
    edges_at(r, (int64_t)px * SUBPIX, (int64_t)py * SUBPIX, e);
    const int64_t d = (int64_t)(r->X[1] - r->X[0]) * (r->Y[2] - r->Y[0]) -
                      (int64_t)(r->X[2] - r->X[0]) * (r->Y[1] - r->Y[0]);
    const float ar = 1.0f / (float)d;
    beta[0] = (float)e[0] * ar;
    beta[1] = (float)e[1] * ar;
    beta[2] = (float)e[2] * ar;
}
static float depth_at(const orc_tri* r, const float beta[3]) { /* CU:1974-1980 */
    return fmaf(r->Z[2], beta[2], fmaf(r->Z[1], beta[1], r->Z[0] * beta[0]));
}

/* perspective-correct attribute weights, pixelShadingExecKernel CU:2185-2196 */
static void weights_at(const orc_tri* r, const float beta[3], float wgt[3]) {
    float b0 = beta[0] * r->IW[0], b1 = beta[1] * r->IW[1], b2 = beta[2] * r->IW[2];
    const float inv = 1.0f / ((b0 + b1) + b2);
    b0 *= inv; b1 *= inv; b2 *= inv;
    const float u = fmaf(b2, r->bary[2][0], fmaf(b1, r->bary[1][0], b0 * r->bary[0][0]));
    const float v = fmaf(b2, r->bary[2][1], fmaf(b1, r->bary[1][1], b0 * r->bary[0][1]));
    wgt[0] = u; wgt[1] = v; wgt[2] = 1.0f - (u + v);
}

static void interp(const orc_tri* r, int px, int py, const orc_f4* vary, int nvary, const int32_t* idx,
                   orc_f4* out) {
    float beta[3], w[3];
    bary_at(r, px, py, beta);
    weights_at(r, beta, w);
    const int32_t* a = idx + 3 * (size_t)r->prim;
    for (int k = 0; k < nvary; k++) { /* CU:2197-2215 */
        const orc_f4 v0 = vary[(size_t)a[0] * nvary + k], v1 = vary[(size_t)a[1] * nvary + k],
                     v2 = vary[(size_t)a[2] * nvary + k];
        out[k].x = fmaf(v2.x, w[2], fmaf(v1.x, w[1], v0.x * w[0]));
        out[k].y = fmaf(v2.y, w[2], fmaf(v1.y, w[1], v0.y * w[0]));
        out[k].z = fmaf(v2.z, w[2], fmaf(v1.z, w[1], v0.z * w[0]));
        out[k].w = fmaf(v2.w, w[2], fmaf(v1.w, w[1], v0.w * w[0]));
    }
}

#define MAXV 4
#define MAXS 16

/* Direct depth compare of the blending path (PS_DEPTHTEST, CU:1790-1800) */
static int depth_pass(int func, float d, float cur) {
    switch (func) {
        case ORC_CMP_NEVER: return 0;
        case ORC_CMP_LESS: return d < cur;
        case ORC_CMP_EQUAL: return d == cur;
        case ORC_CMP_LESS_OR_EQUAL: return d <= cur;
        case ORC_CMP_GREATER: return d > cur;
        case ORC_CMP_NOT_EQUAL: return d != cur;
        case ORC_CMP_GREATER_OR_EQUAL: return d >= cur;
        default: return 1;
    }
}

/* (s.x + s.y (1 - src.a) + s.z (1 - dst.a)) (1 - s.w), CU:1817-1820 — op order fixed here and in the kernel */
static float blend_factor(const float c[4], float oms, float omd) {
    return ((c[0] + c[1] * oms) + c[2] * omd) * (1.0f - c[3]);
}

/* Ordered alpha blending (pixelShadingAlphaBlendKernel CU:1649-1897): primitives in submission order (the reference
 * merges its three per-bin lists by original primitive id, CU:1853-1895); per covered pixel: depth test against the
 * depth the previous fragment left, FS, blend into the colour, depth written.  1x only (the kernel has no sample loop). */
static int render_blend(const orc_state* st, const orc_f4* vary, int nvary, const int32_t* idx, const orc_tri* tris, int ntris,
                        int fs_id, const fs_env* env, float* color, float* depth, int32_t* prim_id, int64_t* stats) {
    const int W = st->width, H = st->height, func = st->depth_func;
    int64_t frags = 0, covered = 0;
    for (size_t p = 0; p < (size_t)W * H; p++) prim_id[p] = -1;
    for (int i = 0; i < ntris; i++) {
        const orc_tri* r = &tris[i];
        int64_t minx = r->X[0], maxx = r->X[0], miny = r->Y[0], maxy = r->Y[0];
        for (int j = 1; j < 3; j++) {
            if (r->X[j] < minx) minx = r->X[j];
            if (r->X[j] > maxx) maxx = r->X[j];
            if (r->Y[j] < miny) miny = r->Y[j];
            if (r->Y[j] > maxy) maxy = r->Y[j];
        }
        int x0 = (int)(minx >> 8) - 1, x1 = (int)(maxx >> 8) + 1;
        int y0 = (int)(miny >> 8) - 1, y1 = (int)(maxy >> 8) + 1;
        if (x0 < 0) x0 = 0;
        if (y0 < 0) y0 = 0;
        if (x1 > W - 1) x1 = W - 1;
        if (y1 > H - 1) y1 = H - 1;
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++) {
                int ox, oy;
                sample_offset16(st, 0, &ox, &oy);
                int64_t e[3];
                edges_at(r, (int64_t)x * SUBPIX + ox * 16, (int64_t)y * SUBPIX + oy * 16, e);
                if (!(e[0] > 0 && e[1] > 0 && e[2] > 0)) continue;
                if (!scissor_pass(st, x, y)) continue;
                covered++;
                const size_t p = (size_t)y * W + x;
                float beta[3];
                bary_at(r, x, y, beta);
                const float d = depth_at(r, beta);
                if (!depth_pass(func, d, depth[p])) continue;
                /* the pixel and its two quad neighbours (coarse derivatives, SB:36-76) */
                const int qx = x & ~1, qy = y & ~1, l = (x & 1) + 2 * (y & 1);
                orc_f4 v[4][MAXV];
                for (int k = 0; k < 4; k++) interp(r, qx + (k & 1), qy + (k >> 1), vary, nvary, idx, v[k]);
                orc_f4 ddx[MAXV], ddy[MAXV];
                const int row = l >> 1, col = l & 1;
                for (int k = 0; k < nvary; k++) {
                    const orc_f4 *xl = &v[row * 2][k], *xr = &v[row * 2 + 1][k];
                    const orc_f4 *yt = &v[col][k], *yb = &v[2 + col][k];
                    ddx[k].x = xr->x - xl->x; ddx[k].y = xr->y - xl->y; ddx[k].z = xr->z - xl->z; ddx[k].w = xr->w - xl->w;
                    ddy[k].x = yb->x - yt->x; ddy[k].y = yb->y - yt->y; ddy[k].z = yb->z - yt->z; ddy[k].w = yb->w - yt->w;
                }
                float c[4];
                run_fs(fs_id, env, v[l], ddx, ddy, r->prim, r->fan, c);
                frags++;
                const float d2 = fs_depth(fs_id, env, v[l], d);
                if (!depth_pass(func, d2, depth[p])) continue;      /* second PS_DEPTHTEST, CU:1805 (late Z) */
                float* dst = color + 4 * p;
                const float oms = 1.0f - c[3], omd = 1.0f - dst[3];
                const float ms = blend_factor(st->blend_coef[0], oms, omd), md = blend_factor(st->blend_coef[1], oms, omd);
                const float msa = blend_factor(st->blend_coef[2], oms, omd), mda = blend_factor(st->blend_coef[3], oms, omd);
                const float o0 = dst[0] * md + c[0] * ms, o1 = dst[1] * md + c[1] * ms, o2 = dst[2] * md + c[2] * ms;
                const float o3 = dst[3] * mda + c[3] * msa;
                dst[0] = o0; dst[1] = o1; dst[2] = o2; dst[3] = o3;
                depth[p] = d2;           /* CU:1828-1829: depth written, next fragment tests against it */
                prim_id[p] = r->prim;    /* last writer */
            }
    }
    if (stats) { stats[0] = frags; stats[1] = covered; stats[2] = ntris; }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Point and line polygon modes.  PARITY UNPINNED: the reference implements them only in its CUDA
 * renderer, which does not build here, so these restate CU:2782-2973 (points) and CU:2976-3334 (lines) from the source
 * alone.  Both work per PIXEL: every candidate that lands on the pixel competes with the smallest z (strict <; the
 * reference keeps whichever candidate its atomics listed first on a tie — here the lowest candidate id), the winner is
 * shaded if its z is strictly below the stored depth (the compare op is ignored, CU:2917 / CU:3280), colour replaces,
 * depth = z.  No MSAA, no blending, no derivatives (zeros).  Deliberate deviations, each a slip of the reference:
 *   - pixelShadingKernel derives the row from `globalInvo / csFrameHeight` (CU:2859 / CU:3175): correct only for
 *     square frames; restated with the frame width;
 *   - devLineCullClip tests v1 twice against the upper bound (CU:3012): a segment whose SECOND end lies beyond 1.0 is
 *     kept and rasterised past the frame edge; here its pixels outside the frame are skipped. */
static float to_unit(float v) { return (float)((double)v * 0.5 + 0.5); }       /* `v * 0.5 + 0.5` with double literals */

typedef struct { float x, y, z, w; } prim_v;
static int point_pixel(const orc_state* st, orc_f4 v, int* px, int* py, float* z) {
    if (v.w < 1e-8f) return 0;                                  /* CU_EPS, CU:2802 */
    v.x /= v.w; v.y /= v.w; v.z /= v.w;
    v.x = to_unit(v.x); v.y = to_unit(v.y);
    if ((double)v.x < 0.0 || (double)v.x > 1.0 || (double)v.y < 0.0 || (double)v.y > 1.0) return 0;
    *px = (int)(v.x * (float)(st->width - 1));
    *py = (int)(v.y * (float)(st->height - 1));
    *z = v.z;
    return 1;
}
static float lerpf_ref(float a, float b, float t) { return a + (b - a) * t; }    /* ShaderOpsCuda.cuh:107-109 */

/* clip + map of one segment (devLineCullClip CU:2976-3019); returns 0 if dropped */
static int line_setup(orc_f4 v1, orc_f4 v2, orc_f4* o1, orc_f4* o2) {
    if (v1.w < 0 && v2.w < 0) return 0;
    if (v1.w < 0) {
        const float deno = v2.w - v1.w, numo = (float)(-1.0 + (double)v2.w), t = numo / deno;
        v1.x = lerpf_ref(v2.x, v1.x, t); v1.y = lerpf_ref(v2.y, v1.y, t); v1.z = lerpf_ref(v2.z, v1.z, t); v1.w = lerpf_ref(v2.w, v1.w, t);
    }
    if (v2.w < 0) {
        const float deno = v1.w - v2.w, numo = (float)(-1.0 + (double)v1.w), t = numo / deno;
        v2.x = lerpf_ref(v1.x, v2.x, t); v2.y = lerpf_ref(v1.y, v2.y, t); v2.z = lerpf_ref(v1.z, v2.z, t); v2.w = lerpf_ref(v1.w, v2.w, t);
    }
    v1.x /= v1.w; v1.y /= v1.w; v1.z /= v1.w;
    v2.x /= v2.w; v2.y /= v2.w; v2.z /= v2.w;
    v1.x = to_unit(v1.x); v1.y = to_unit(v1.y); v2.x = to_unit(v2.x); v2.y = to_unit(v2.y);
    if (v1.x < 0.0f || v1.y < 0.0f || v1.x > 1.0f || v1.y > 1.0f) return 0;
    if (v2.x < 0.0f || v2.y < 0.0f) return 0;                   /* (sic) upper bound of v2 not tested, CU:3012 */
    *o1 = v1; *o2 = v2;
    return 1;
}
/* position of pixel (px, py) along the segment (getPercent CU:3181-3197) */
static float line_percent(const orc_state* st, orc_f4 v1, orc_f4 v2, int px, int py) {
    v1.x *= (float)st->width; v1.y *= (float)st->height; v2.x *= (float)st->width; v2.y *= (float)st->height;
    const float dx = fabsf(v2.x - v1.x), dy = fabsf(v2.y - v1.y);
    if (dx >= dy) return ((float)px - v1.x) / (v2.x - v1.x);
    return ((float)py - v1.y) / (v2.y - v1.y);
}

int orc_render_prims(const orc_state* st, int mode, const orc_f4* pos, const orc_f4* vary, int nvary, const int32_t* idx,
                     int nidx, int fs_id, const void* uniforms, const orc_texture* textures, const orc_sampler* samplers,
                     float* color, float* depth, int32_t* prim_id, int64_t* stats) {
    const int W = st->width, H = st->height;
    if (st->msaa != 1 || nvary > MAXV || (mode != 1 && mode != 2)) return -1;
    if (st->do_clear) orc_clear(st, color, depth);
    const size_t np = (size_t)W * H;
    float* bestz = (float*)malloc(sizeof(float) * np);
    int32_t* best = (int32_t*)malloc(sizeof(int32_t) * np);
    for (size_t p = 0; p < np; p++) { bestz[p] = 1e9f; best[p] = -1; prim_id[p] = -1; }
    fs_env env = {(const float*)uniforms, textures, samplers};
    int64_t frags = 0;
    if (mode == 2) {                                            /* IF_POLYGON_MODE_POINT: one point per index */
        for (int i = 0; i < nidx; i++) {
            int px, py; float z;
            if (!point_pixel(st, pos[idx[i]], &px, &py, &z)) continue;
            const size_t p = (size_t)py * W + px;
            if (z < bestz[p]) { bestz[p] = z + 0.0f; best[p] = i; }
        }
        for (size_t p = 0; p < np; p++) {
            if (best[p] < 0 || !(bestz[p] < depth[p])) continue;
            const int vi = idx[best[p]];
            orc_f4 v[MAXV], zero[MAXV];
            memset(zero, 0, sizeof zero);
            for (int k = 0; k < nvary; k++) v[k] = vary[(size_t)vi * nvary + k];
            float c[4];
            run_fs(fs_id, &env, v, zero, zero, vi, 0, c);
            frags++;
            memcpy(color + 4 * p, c, sizeof c);
            depth[p] = bestz[p];
            prim_id[p] = vi;                                   /* dAtriOriginalPrimId = the vertex index, CU:2815 */
        }
    } else {                                                    /* IF_POLYGON_MODE_LINE: the three edges of every triangle */
        const int nl = (nidx / 3) * 3;
        orc_f4* a = (orc_f4*)malloc(sizeof(orc_f4) * 2 * (size_t)(nl > 0 ? nl : 1));
        uint8_t* ok = (uint8_t*)calloc((size_t)(nl > 0 ? nl : 1), 1);
        for (int l = 0; l < nl; l++) {
            const int t = l / 3, e = l % 3;
            const orc_f4 v1 = pos[idx[3 * t + e]], v2 = pos[idx[3 * t + (e + 1) % 3]];
            ok[l] = (uint8_t)line_setup(v1, v2, &a[2 * l], &a[2 * l + 1]);
            if (!ok[l]) continue;
            /* bresenhamRasterizationKernel CU:3060-3127 */
            int x1 = (int)(a[2 * l].x * (float)(W - 1)), y1 = (int)(a[2 * l].y * (float)(H - 1));
            int x2 = (int)(a[2 * l + 1].x * (float)(W - 1)), y2 = (int)(a[2 * l + 1].y * (float)(H - 1));
            const int dxi = abs(x1 - x2), dyi = abs(y1 - y2);
            const int major_x = dxi > dyi;
            if (major_x ? (x2 < x1) : (y2 < y1)) { int tq = x1; x1 = x2; x2 = tq; tq = y1; y1 = y2; y2 = tq; }
            int cx = x1, cy = y1;
            const int n = major_x ? (x2 - x1) : (y2 - y1);
            const float dslope = major_x ? fabsf(1.0f * (float)(y2 - y1) / (float)(x2 - x1)) : fabsf(1.0f * (float)(x2 - x1) / (float)(y2 - y1));
            const int step = major_x ? ((y2 - y1) > 0 ? 1 : -1) : ((x2 - x1) > 0 ? 1 : -1);
            float d = 0.0f;
            for (int i = 0; i <= n; i++) {
                if (cx >= 0 && cx < W && cy >= 0 && cy < H) {
                    const size_t p = (size_t)cy * W + cx;
                    const float pc = line_percent(st, a[2 * l], a[2 * l + 1], cx, cy);
                    const float z = lerpf_ref(a[2 * l].z, a[2 * l + 1].z, pc);
                    if (z < bestz[p]) { bestz[p] = z + 0.0f; best[p] = l; }      /* ascending l: ties keep the lowest id */
                }
                d += dslope;
                if (d > 0.5f) { if (major_x) cy += step; else cx += step; d -= 1.0f; }
                if (major_x) cx++; else cy++;
            }
        }
        for (size_t p = 0; p < np; p++) {
            if (best[p] < 0 || !(bestz[p] < depth[p])) continue;
            const int l = best[p], t = l / 3, e = l % 3;
            const orc_f4 v1 = a[2 * l], v2 = a[2 * l + 1];
            const float pc = line_percent(st, v1, v2, (int)(p % W), (int)(p / W));
            const float z1 = v1.w, z2 = v2.w;
            const float zt = z1 * z2 / (z2 + pc * (z1 - z2));
            const int i1 = idx[3 * t + e], i2 = idx[3 * t + (e + 1) % 3];
            orc_f4 v[MAXV], zero[MAXV];
            memset(zero, 0, sizeof zero);
            for (int k = 0; k < nvary; k++) {
                orc_f4 a1 = vary[(size_t)i1 * nvary + k], a2 = vary[(size_t)i2 * nvary + k];
                a1.x /= z1; a1.y /= z1; a1.z /= z1; a1.w /= z1; a2.x /= z2; a2.y /= z2; a2.z /= z2; a2.w /= z2;
                v[k].x = (pc * (a2.x - a1.x) + a1.x) * zt; v[k].y = (pc * (a2.y - a1.y) + a1.y) * zt;
                v[k].z = (pc * (a2.z - a1.z) + a1.z) * zt; v[k].w = (pc * (a2.w - a1.w) + a1.w) * zt;
            }
            float c[4];
            run_fs(fs_id, &env, v, zero, zero, t, 0, c);
            frags++;
            memcpy(color + 4 * p, c, sizeof c);
            depth[p] = bestz[p];
            prim_id[p] = l;                                    /* dAtriOriginalPrimId = triangle * 3 + edge, CU:3048-3050 */
        }
        free(a); free(ok);
    }
    free(bestz); free(best);
    if (stats) { stats[0] = frags; stats[1] = frags; stats[2] = 0; }
    return 0;
}

/* one horizontal band of the frame (rows y_begin .. y_end - 1, whole quad rows): coverage + visibility keys, then
 * quad shading + resolve.  Bands touch disjoint pixels, so the threads share nothing they write. */
typedef struct {
    const orc_state* st; const orc_f4* vary; int nvary; const int32_t* idx; const orc_tri* tris; int ntris;
    const int32_t* by_slot; int fs_id; const void* uniforms; const orc_texture* textures; const orc_sampler* samplers;
    float* color; float* depth; int32_t* prim_id; uint64_t* keys; int y_begin, y_end; int64_t covered, frags;
} band_job;

static int orc_thread_count(void) {
    const char* e = getenv("ORC_THREADS");
    if (e && atoi(e) > 0) return atoi(e);
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 32 ? 32 : (int)n);
}

static void* band_worker(void* arg) {
    band_job* j = (band_job*)arg;
    const orc_state* st = j->st;
    const int W = st->width, H = st->height, S = st->msaa, func = st->depth_func;
    const orc_f4* vary = j->vary; const int nvary = j->nvary; const int32_t* idx = j->idx;
    const orc_tri* tris = j->tris; const int ntris = j->ntris; const int32_t* by_slot = j->by_slot;
    const int fs_id = j->fs_id; const void* uniforms = j->uniforms;
    const orc_texture* textures = j->textures; const orc_sampler* samplers = j->samplers;
    float* color = j->color; float* depth = j->depth; int32_t* prim_id = j->prim_id; uint64_t* keys = j->keys;
    int64_t covered = 0, frags = 0;
    for (int i = 0; i < ntris; i++) {
        const orc_tri* r = &tris[i];
        int64_t minx = r->X[0], maxx = r->X[0], miny = r->Y[0], maxy = r->Y[0];
        for (int j = 1; j < 3; j++) {
            if (r->X[j] < minx) minx = r->X[j];
            if (r->X[j] > maxx) maxx = r->X[j];
            if (r->Y[j] < miny) miny = r->Y[j];
            if (r->Y[j] > maxy) maxy = r->Y[j];
        }
        int x0 = (int)(minx >> 8) - 1, x1 = (int)(maxx >> 8) + 1;
        int y0 = (int)(miny >> 8) - 1, y1 = (int)(maxy >> 8) + 1;
        if (x0 < 0) x0 = 0;
        if (y0 < 0) y0 = 0;
        if (x1 > W - 1) x1 = W - 1;
        if (y1 > H - 1) y1 = H - 1;
        if (y0 < j->y_begin) y0 = j->y_begin;
        if (y1 > j->y_end - 1) y1 = j->y_end - 1;
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++) {
                int mask = 0;
                for (int s = 0; s < S; s++) {
                    int ox, oy;
                    sample_offset16(st, s, &ox, &oy);
                    int64_t e[3];
                    edges_at(r, (int64_t)x * SUBPIX + ox * 16, (int64_t)y * SUBPIX + oy * 16, e);
                    if (e[0] > 0 && e[1] > 0 && e[2] > 0) mask |= 1 << s;
                }
                if (!mask) continue;
                if (!scissor_pass(st, x, y)) continue;
                float beta[3];
                bary_at(r, x, y, beta);
                const float d = depth_at(r, beta);
                const size_t p = (size_t)y * W + x;
                uint64_t k;
                if (!key_make(func, d, depth[p], (uint32_t)r->slot, &k)) continue;
                for (int s = 0; s < S; s++)
                    if ((mask >> s) & 1) {
                        covered++;
                        if (k < keys[p * S + s]) keys[p * S + s] = k;
                    }
            }
    }

    /* shading: 2x2 quads, one FS invocation per (pixel, distinct visible triangle);
     * every lane of the quad interpolates so that ddx/ddy exist (CU:2264-2310, SB:36-76) */
    fs_env env = {(const float*)uniforms, textures, samplers};
    for (int qy = j->y_begin; qy < j->y_end; qy += 2)
        for (int qx = 0; qx < W; qx += 2) {
            uint32_t slots[4 * MAXS];
            int nslots = 0;
            for (int l = 0; l < 4; l++) {
                const int x = qx + (l & 1), y = qy + (l >> 1);
                if (x >= W || y >= H) continue;
                for (int s = 0; s < S; s++) {
                    uint32_t sl;
                    if (!key_slot(func, keys[((size_t)y * W + x) * S + s], &sl)) continue;
                    int j = 0;
                    while (j < nslots && slots[j] != sl) j++;
                    if (j == nslots) slots[nslots++] = sl;
                }
            }
            if (!nslots) {
                for (int l = 0; l < 4; l++) {
                    const int x = qx + (l & 1), y = qy + (l >> 1);
                    if (x >= W || y >= H) continue;
                    for (int s = 0; s < S; s++) prim_id[((size_t)y * W + x) * S + s] = -1;
                }
                continue;
            }
            /* ascending slot order */
            for (int a = 1; a < nslots; a++) {
                uint32_t v = slots[a];
                int b = a - 1;
                while (b >= 0 && slots[b] > v) { slots[b + 1] = slots[b]; b--; }
                slots[b + 1] = v;
            }
            float acc[4][4], dacc[4];
            int cnt_bg[4];
            memset(acc, 0, sizeof acc);
            memset(dacc, 0, sizeof dacc);
            for (int l = 0; l < 4; l++) cnt_bg[l] = S;
            for (int j = 0; j < nslots; j++) {
                const orc_tri* r = &tris[by_slot[slots[j]]];
                orc_f4 v[4][MAXV];
                for (int l = 0; l < 4; l++) interp(r, qx + (l & 1), qy + (l >> 1), vary, nvary, idx, v[l]);
                for (int l = 0; l < 4; l++) {
                    const int x = qx + (l & 1), y = qy + (l >> 1);
                    if (x >= W || y >= H) continue;
                    int cnt = 0;
                    for (int s = 0; s < S; s++) {
                        uint32_t sl;
                        const size_t si = ((size_t)y * W + x) * S + s;
                        if (key_slot(func, keys[si], &sl) && sl == slots[j]) { cnt++; prim_id[si] = r->prim; }
                    }
                    if (!cnt) continue; /* helper lane */
                    orc_f4 ddx[MAXV], ddy[MAXV];
                    const int row = l >> 1, col = l & 1;
                    for (int k = 0; k < nvary; k++) {
                        const orc_f4 *xl = &v[row * 2][k], *xr = &v[row * 2 + 1][k];
                        const orc_f4 *yt = &v[col][k], *yb = &v[2 + col][k];
                        ddx[k].x = xr->x - xl->x; ddx[k].y = xr->y - xl->y; ddx[k].z = xr->z - xl->z; ddx[k].w = xr->w - xl->w;
                        ddy[k].x = yb->x - yt->x; ddy[k].y = yb->y - yt->y; ddy[k].z = yb->z - yt->z; ddy[k].w = yb->w - yt->w;
                    }
                    float c[4];
                    run_fs(fs_id, &env, v[l], ddx, ddy, r->prim, r->fan, c);
                    frags++;
                    float beta[3];
                    bary_at(r, x, y, beta);
                    const float d = depth_at(r, beta);
                    for (int ch = 0; ch < 4; ch++) acc[l][ch] = fmaf((float)cnt, c[ch], acc[l][ch]);
                    dacc[l] = fmaf((float)cnt, d, dacc[l]);
                    cnt_bg[l] -= cnt;
                }
            }
            /* resolve: pixelMultisampleCoverageModulationKernel CU:2319-2361 (box average;
             * uncovered samples keep the previous colour / depth) */
            for (int l = 0; l < 4; l++) {
                const int x = qx + (l & 1), y = qy + (l >> 1);
                if (x >= W || y >= H) continue;
                const size_t p = (size_t)y * W + x;
                for (int s = 0; s < S; s++) {
                    uint32_t sl;
                    if (!key_slot(func, keys[p * S + s], &sl)) prim_id[p * S + s] = -1;
                }
                if (cnt_bg[l] == S) continue;
                for (int ch = 0; ch < 4; ch++) {
                    float a = acc[l][ch];
                    if (cnt_bg[l] > 0) a = fmaf((float)cnt_bg[l], color[4 * p + ch], a);
                    color[4 * p + ch] = a / (float)S;
                }
                float da = dacc[l];
                if (cnt_bg[l] > 0) da = fmaf((float)cnt_bg[l], depth[p], da);
                depth[p] = da / (float)S;
            }
        }
    j->covered = covered; j->frags = frags;
    return NULL;
}

int orc_render(const orc_state* st, const orc_f4* pos, const orc_f4* vary, int nvary,
               const int32_t* idx, int ntri, int fs_id, const void* uniforms,
               const orc_texture* textures, const orc_sampler* samplers,
               float* color, float* depth, int32_t* prim_id, int64_t* stats) {
    const int W = st->width, H = st->height, S = st->msaa, func = st->depth_func;
    if (nvary > MAXV || S > MAXS) return -1;
    if (func == ORC_CMP_NOT_EQUAL && !st->blend_enable && !fs_modifies_depth(fs_id)) return -2;
    orc_tri* tris = (orc_tri*)malloc(sizeof(orc_tri) * 2 * (size_t)(ntri > 0 ? ntri : 1));
    const int ntris = orc_assemble(st, pos, idx, ntri, tris);
    /* slot -> assembled triangle */
    int32_t* by_slot = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(ntri > 0 ? ntri : 1));
    for (int i = 0; i < 2 * ntri; i++) by_slot[i] = -1;
    for (int i = 0; i < ntris; i++) by_slot[tris[i].slot] = i;

    if (st->do_clear) orc_clear(st, color, depth);

    if (st->blend_enable || fs_modifies_depth(fs_id)) {
        if (S != 1) { free(by_slot); free(tris); return -3; }
        fs_env benv = {(const float*)uniforms, textures, samplers};
        const int rc = render_blend(st, vary, nvary, idx, tris, ntris, fs_id, &benv, color, depth, prim_id, stats);
        free(by_slot); free(tris);
        return rc;
    }

    uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)W * H * S);
    for (size_t p = 0; p < (size_t)W * H; p++)
        for (int s = 0; s < S; s++) keys[p * S + s] = key_init(func, depth[p]);

    /* The frame is cut into horizontal bands of whole quad rows, one thread per band: every thread walks all
     * triangles but only the rows of its band, so the per-sample minimum and the per-quad shading are computed by
     * exactly the loops below whatever the thread count (results do not depend on it). */
    int nthreads = orc_thread_count();
    if (nthreads > (H + 1) / 2) nthreads = (H + 1) / 2;
    if (nthreads < 1) nthreads = 1;
    band_job* jobs = (band_job*)calloc((size_t)nthreads, sizeof(band_job));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    const int qrows = (H + 1) / 2;
    for (int t = 0; t < nthreads; t++) {
        band_job* j = &jobs[t];
        j->st = st; j->vary = vary; j->nvary = nvary; j->idx = idx; j->tris = tris; j->ntris = ntris; j->by_slot = by_slot;
        j->fs_id = fs_id; j->uniforms = uniforms; j->textures = textures; j->samplers = samplers;
        j->color = color; j->depth = depth; j->prim_id = prim_id; j->keys = keys;
        j->y_begin = 2 * (int)((int64_t)qrows * t / nthreads);
        j->y_end = 2 * (int)((int64_t)qrows * (t + 1) / nthreads);
        if (j->y_end > H) j->y_end = H;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, band_worker, &jobs[t]);
    band_worker(&jobs[0]);
    int64_t covered = 0, frags = 0;
    for (int t = 0; t < nthreads; t++) {
        if (t) pthread_join(th[t], NULL);
        covered += jobs[t].covered; frags += jobs[t].frags;
    }
    free(jobs); free(th);
    if (stats) { stats[0] = frags; stats[1] = covered; stats[2] = ntris; }
    free(keys); free(by_slot); free(tris);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Shader callbacks for the UNMODIFIED reference CPU renderer (see header).    */
static float g_uniforms[64];
static orc_texture g_ref_tex;
static const orc_sampler g_ref_samp = {ORC_FILTER_LINEAR, ORC_ADDR_REPEAT, ORC_ADDR_REPEAT, ORC_ADDR_REPEAT, 0, 0, 1.0f};

void orc_ref_set_uniforms(const void* data, int nbytes) {
    memcpy(g_uniforms, data, nbytes < (int)sizeof g_uniforms ? (size_t)nbytes : sizeof g_uniforms);
}
void orc_ref_set_texture(const float* data, int w, int h) {
    g_ref_tex.data = data; g_ref_tex.width = w; g_ref_tex.height = h;
    g_ref_tex.mip_levels = 0; g_ref_tex.layers = 1;
}
static orc_f4 ref_mvp(const orc_f4 a0) {
    const float* m = g_uniforms;
    orc_f4 p;
    p.x = fmaf(m[3], a0.w, fmaf(m[2], a0.z, fmaf(m[1], a0.y, m[0] * a0.x)));
    p.y = fmaf(m[7], a0.w, fmaf(m[6], a0.z, fmaf(m[5], a0.y, m[4] * a0.x)));
    p.z = fmaf(m[11], a0.w, fmaf(m[10], a0.z, fmaf(m[9], a0.y, m[8] * a0.x)));
    p.w = fmaf(m[15], a0.w, fmaf(m[14], a0.z, fmaf(m[13], a0.y, m[12] * a0.x)));
    return p;
}
void orc_ref_vs_passthrough1(const void* const* in, orc_f4* out_pos, orc_f4* const* out_vary) {
    *out_pos = *(const orc_f4*)in[0];
    *out_vary[0] = *(const orc_f4*)in[1];
}
void orc_ref_vs_passthrough2(const void* const* in, orc_f4* out_pos, orc_f4* const* out_vary) {
    *out_pos = *(const orc_f4*)in[0];
    *out_vary[0] = *(const orc_f4*)in[1];
    *out_vary[1] = *(const orc_f4*)in[2];
}
void orc_ref_vs_mvp1(const void* const* in, orc_f4* out_pos, orc_f4* const* out_vary) {
    *out_pos = ref_mvp(*(const orc_f4*)in[0]);
    *out_vary[0] = *(const orc_f4*)in[1];
}
void orc_ref_vs_mvp2(const void* const* in, orc_f4* out_pos, orc_f4* const* out_vary) {
    *out_pos = ref_mvp(*(const orc_f4*)in[0]);
    *out_vary[0] = *(const orc_f4*)in[1];
    *out_vary[1] = *(const orc_f4*)in[2];
}
void orc_ref_fs_varying0(const void* vary, void* color, float* depth) {
    (void)depth;
    memcpy(color, vary, 16);
}
void orc_ref_fs_normal_color(const void* vary, void* color, float* depth) {
    (void)depth;
    fs_env e = {g_uniforms, &g_ref_tex, &g_ref_samp};
    run_fs(ORC_FS_NORMAL_COLOR, &e, (const orc_f4*)vary, NULL, NULL, 0, 0, (float*)color);
}
void orc_ref_fs_textured_lod0(const void* vary, void* color, float* depth) {
    (void)depth;
    fs_env e = {g_uniforms, &g_ref_tex, &g_ref_samp};
    run_fs(ORC_FS_TEXTURED_LOD0, &e, (const orc_f4*)vary, NULL, NULL, 0, 0, (float*)color);
}
