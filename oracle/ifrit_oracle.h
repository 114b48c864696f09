/* ifrit_oracle.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement, in plain C, of the softgraphics tile-raster path of
 * Aeroraven/Ifrit-v2 with the fixed-point snapping rules that DESIGN.md §3
 * defines.  Every function cites the reference file:line it restates.
 * Short names used in citations:
 *   CU    = projects/softgraphics/src/engine/tilerastercuda/TileRasterCoreInvocationCuda.cu
 *   CUIMG = projects/softgraphics/src/engine/tilerastercuda/TileRasterImageOpInvocationsCuda.cu
 *   WK    = projects/softgraphics/src/engine/tileraster/TileRasterWorker.cpp
 *   SB    = include/ifrit/softgraphics/engine/math/ShaderBuiltinCuda.cuh
 *
 * Parity pin: the reference has no tests/golden vectors, so
 * this restatement is pinned by running the reference's own CPU renderer
 * (oracle/_ref/libifrit_ref.so, built from /root/reference by oracle/Makefile)
 * on the same seeded scenes: tests/test_oracle_vs_reference.py and the
 * fixtures under tests/golden/ made by tests/golden/make_golden.py.
 */
#ifndef IFRIT_ORACLE_H
#define IFRIT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, z, w; } orc_f4;

/* convention: which of the reference's two renderers is being restated */
enum { ORC_CONV_CUDA = 0, /* TileRasterRendererCuda: centre sampling, w=1e-3 clip */
       ORC_CONV_CPU  = 1  /* TileRasterRenderer (AVX2): corner sampling, w=1 clip, z reject */ };

/* enums mirror include/ifrit/softgraphics/engine/base/Constants.h:23-96 */
enum { ORC_CULL_NONE = 0, ORC_CULL_FRONT = 1, ORC_CULL_BACK = 2, ORC_CULL_FRONT_AND_BACK = 3 };
enum { ORC_CMP_NEVER = 0, ORC_CMP_LESS = 1, ORC_CMP_EQUAL = 2, ORC_CMP_LESS_OR_EQUAL = 3,
       ORC_CMP_GREATER = 4, ORC_CMP_NOT_EQUAL = 5, ORC_CMP_GREATER_OR_EQUAL = 6, ORC_CMP_ALWAYS = 7 };
enum { ORC_ADDR_REPEAT = 0, ORC_ADDR_MIRRORED_REPEAT = 1, ORC_ADDR_CLAMP_TO_EDGE = 2,
       ORC_ADDR_CLAMP_TO_BORDER = 3, ORC_ADDR_MIRROR_CLAMP_TO_EDGE = 4 };
enum { ORC_FILTER_NEAREST = 0, ORC_FILTER_LINEAR = 1 };

/* vertex / fragment shader registry ids (same ids as include/ifrit_b200.h) */
enum { ORC_VS_PASSTHROUGH = 0, ORC_VS_MVP = 1 };
enum { ORC_FS_VARYING0 = 0, ORC_FS_NORMAL_COLOR = 1, ORC_FS_PRIM_ID = 2, ORC_FS_TEXTURED = 3,
       ORC_FS_TEXTURED_LOD0 = 4, ORC_FS_SKYBOX = 5, ORC_FS_DERIV_VIS = 6, ORC_FS_DEPTH_MOD = 7 };

typedef struct {
    int32_t width, height;
    int32_t msaa;          /* 1, 2, 4, 8, 16 */
    int32_t cull_mode;
    int32_t depth_func;    /* effective compare op (ALWAYS when the depth test is disabled) */
    int32_t convention;
    int32_t scissor_enable, nscissor;
    float   scissor[16][4]; /* x, y, w, h — inclusive both ends (CU:3351) */
    int32_t do_clear;
    float   clear_color[4];
    float   clear_depth;
    /* ordered alpha blending (pixelShadingAlphaBlendKernel CU:1649-1897, updateBlendStateKernel CU:3436-3537):
     * blend_coef[0..3] = colour src, colour dst, alpha src, alpha dst, each {x1, x(1-src.a), x(1-dst.a), zero flag} */
    int32_t blend_enable;
    float   blend_coef[4][4];
} orc_state;

typedef struct {
    const float* data;     /* RGBA32F, level-major, then layer, then row-major (SB:97-105) */
    int32_t width, height;
    int32_t mip_levels;    /* number of levels BEYOND level 0 (CU:3561-3577) */
    int32_t layers;
} orc_texture;

typedef struct {
    int32_t filter, addr_u, addr_v, addr_w, border, aniso_enable;
    float   max_aniso;
} orc_sampler;

/* one assembled (post clip / cull / snap) triangle */
typedef struct {
    int32_t prim, fan, clipped, slot;
    int32_t X[3], Y[3];    /* 24.8 fixed-point pixel coordinates */
    float   Z[3], IW[3];   /* z/w and 1/w */
    float   bary[3][2];    /* clip barycentrics (x,y) of the three fan vertices */
} orc_tri;

int orc_vertex_stage(int vs_id, const void* uniforms, const uint8_t* vb, int stride,
                     const int32_t* attr_offsets, int nattr, int nvary, int nverts,
                     orc_f4* pos, orc_f4* vary);

/* returns number of assembled triangles written to out (capacity 2*ntri) */
int orc_assemble(const orc_state* st, const orc_f4* pos, const int32_t* idx, int ntri, orc_tri* out);

/* full draw.  color: W*H*4, depth: W*H (in/out).  prim_id: W*H*msaa (out, -1 = none).
 * stats[0] = FS invocations, stats[1] = covered samples, stats[2] = assembled triangles. */
int orc_render(const orc_state* st, const orc_f4* pos, const orc_f4* vary, int nvary,
               const int32_t* idx, int ntri, int fs_id, const void* uniforms,
               const orc_texture* textures, const orc_sampler* samplers,
               float* color, float* depth, int32_t* prim_id, int64_t* stats);

/* point (mode 2) / line (mode 1) polygon modes: one point per index / the three edges of every triangle; 1x only */
int orc_render_prims(const orc_state* st, int mode, const orc_f4* pos, const orc_f4* vary, int nvary, const int32_t* idx,
                     int nidx, int fs_id, const void* uniforms, const orc_texture* textures, const orc_sampler* samplers,
                     float* color, float* depth, int32_t* prim_id, int64_t* stats);

void orc_clear(const orc_state* st, float* color, float* depth);

/* mip chain by repeated 2:1 blit (CUIMG:36-155); data holds all levels */
void orc_generate_mipmap(float* data, int w, int h, int mip_levels, int layers, int filter);
int64_t orc_texture_total_texels(int w, int h, int mip_levels, int layers);

/* sampler KAT entry points (SB:97-295) */
void orc_texture_lod(const orc_texture* t, const orc_sampler* s, float u, float v, float lod, float out[4]);
void orc_texture_grad(const orc_texture* t, const orc_sampler* s, float u, float v,
                      const float ddx[2], const float ddy[2], float bias, float out[4]);
void orc_texture_cube_lod(const orc_texture* t, const orc_sampler* s, float x, float y, float z,
                          float lod, float out[4]);

/* batch sampler entry point (see ifrit_oracle.c); mirrors oracle/ref_sampler_shim.cu */
int orc_sample_batch(int kind, const orc_texture* t, const orc_sampler* s, int n, const float* uv, const float* args, float* out);

/* --- shader functions handed to the UNMODIFIED reference CPU renderer through
 * iftrBind{Vertex,Fragment}ShaderFunc (TileRasterRendererExport.h:45-51).  They
 * read the globals set by orc_ref_set_uniforms / orc_ref_set_texture. */
void orc_ref_set_uniforms(const void* data, int nbytes);
void orc_ref_set_texture(const float* data, int w, int h);
void orc_ref_vs_passthrough1(const void* const* in, orc_f4* out_pos, orc_f4* const* out_vary);
void orc_ref_vs_passthrough2(const void* const* in, orc_f4* out_pos, orc_f4* const* out_vary);
void orc_ref_vs_mvp1(const void* const* in, orc_f4* out_pos, orc_f4* const* out_vary);
void orc_ref_vs_mvp2(const void* const* in, orc_f4* out_pos, orc_f4* const* out_vary);
void orc_ref_fs_varying0(const void* vary, void* color, float* depth);
void orc_ref_fs_normal_color(const void* vary, void* color, float* depth);
void orc_ref_fs_textured_lod0(const void* vary, void* color, float* depth);

#ifdef __cplusplus
}
#endif
#endif
