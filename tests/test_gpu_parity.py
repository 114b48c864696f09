"""GPU parity tests proper (-m gpu): the CUDA path, called through the C-ABI, against the oracle on the
same seeded inputs.  Bar: coverage, primitive id and depth bit-exact; colour within 1/255 on >= 99.9 % of
pixels (BASELINE.json north_star); where no libm call is involved colour is checked bit-exact too."""
import os
import numpy as np
import pytest
import ifrit_v2_b200  # noqa: F401
from ifrit_v2_b200 import scenes
from oracle import oracle as O
from util import oracle_draw, gpu_draw, compare

pytestmark = pytest.mark.gpu
IDENT = scenes.uniform_block(np.eye(4))


@pytest.fixture(scope="module")
def r(gpu_renderer_factory):
    return gpu_renderer_factory()


@pytest.mark.parametrize("conv", [0, 1])
def test_kat1_kat2(r, conv):
    for vb, idx in (scenes.kat1(), scenes.kat2()):
        kw = dict(cull=2, convention=conv)
        o = oracle_draw(64, 64, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
        g = gpu_draw(r, 64, 64, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
        compare(o, g, exact_color=True)
        assert g["fragments_shaded"] == o["frags"]
    # pinned counts of SURVEY.md §8c (corner sampling): KAT-1 oracle = reference minus the 31 exact-tie pixels
    o = oracle_draw(64, 64, *scenes.kat1(), 0, O.FS_VARYING0, 1, IDENT, cull=2, convention=1)
    assert (o["prim_id"] >= 0).sum() == 961 - 31


@pytest.mark.parametrize("msaa", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("cull", [0, 1, 2])
def test_random_triangles(r, msaa, cull):
    W, H = 200, 120   # ragged: not a multiple of the 16-px tile
    for seed, neg, size, n in [(11, 0.0, 0.25, 200), (12, 0.2, 0.4, 150), (13, 0.0, 0.02, 2500)]:
        vb, idx = scenes.random_tris(n, seed, neg_w_frac=neg, size=size)
        kw = dict(msaa=msaa, cull=cull, clear_color=(0.1, 0.2, 0.3, 1.0), clear_depth=1.0)
        o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
        g = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
        compare(o, g, exact_color=True)
        assert g["triangles_assembled"] == o["assembled"]
        assert g["fragments_shaded"] == o["frags"]


@pytest.mark.parametrize("func", [O.CMP_NEVER, O.CMP_LESS, O.CMP_EQUAL, O.CMP_LEQUAL, O.CMP_GREATER, O.CMP_GEQUAL,
                                  O.CMP_ALWAYS])
def test_depth_funcs(r, func):
    vb, idx = scenes.random_tris(300, 21, size=0.3)
    vb[:, 2] = np.round(vb[:, 2] / vb[:, 3] * 4) / 4 * vb[:, 3]   # quantised z: many exact depth ties
    kw = dict(msaa=4, depth_func=func, clear_depth=0.5)
    o = oracle_draw(160, 96, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    g = gpu_draw(r, 160, 96, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    compare(o, g, exact_color=True)


@pytest.mark.parametrize("msaa", [1, 8])
@pytest.mark.parametrize("func", [O.CMP_LESS, O.CMP_LEQUAL, O.CMP_GREATER, O.CMP_GEQUAL])
def test_hierarchical_z_heavy_overdraw(r, func, msaa):
    """Screen-filling triangles with steep depth planes, ~60x overdraw: the tile kernel's warp-level Hi-Z reject and
    its trivial-accept blocks must not change a single sample (ids / depth bit-exact), with and without a clear."""
    W, H = 200, 120
    vb, idx = scenes.random_tris(120, 77, size=1.6, extent=0.8)
    vb2, idx2 = scenes.random_tris(400, 78, size=0.25)
    vb = np.concatenate([vb, vb2]); idx = np.arange(len(vb), dtype=np.int32)
    clear_depth = 1.0 if func in (O.CMP_LESS, O.CMP_LEQUAL) else 0.0
    kw = dict(msaa=msaa, depth_func=func, clear_depth=clear_depth, cull=0)
    o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    g = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    compare(o, g, exact_color=True)
    assert g["fragments_shaded"] == o["frags"]
    # second draw over the first frame without a clear: the far bound starts from the stored depth
    vb3, idx3 = scenes.random_tris(60, 79, size=1.2, extent=0.9)
    kw2 = dict(kw, clear=False, color=o["color"], depth=o["depth"])
    o2 = oracle_draw(W, H, vb3, idx3, 0, O.FS_VARYING0, 1, IDENT, **kw2)
    g2 = gpu_draw(r, W, H, vb3, idx3, 0, O.FS_VARYING0, 1, IDENT, **kw2)
    compare(o2, g2, exact_color=True)


def test_double_buffered_readback_lags_one_frame(gpu_renderer_factory):
    """useDoubleBuffer (CUHOST:46-63, CU:3899-3909): drawElements() k copies frame k - 1 into the bound host image on
    a second stream while frame k renders; the read-back verbs return the frame just drawn."""
    from ifrit_v2_b200 import renderer as R
    W, H = 160, 96
    r2 = gpu_renderer_factory()
    host = np.full((H, W, 4), 7.0, np.float32)
    r2.bindFrameBuffer(host, None, W, H, useDoubleBuffer=True)
    frames = []
    for seed in (61, 62, 63):
        vb, idx = scenes.random_tris(80, seed, size=0.4)
        frames.append(oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, clear_color=(.1, .2, .3, 1))["color"])
        R.setup_scene(r2, vb, idx, 0, O.FS_VARYING0, 1, IDENT)
        r2.setClearValues([(.1, .2, .3, 1)], 1.0)
        r2.clear()
        r2.drawElements()
        r2.synchronize()
        assert np.array_equal(r2.readColor().view(np.uint32), frames[-1].view(np.uint32))
        if len(frames) == 1:
            assert np.all(host == 7.0)          # nothing has been copied back yet
        else:
            assert np.array_equal(host.view(np.uint32), frames[-2].view(np.uint32))


def _pack8(color):
    return np.rint(np.clip(color, 0, 1) * np.float32(255)).astype(np.uint8)


def test_rgba8_present_image(gpu_renderer_factory):
    """the packed presentation image is written by the same kernels as the float colour and is what
    the per-draw copy-back carries; it equals round(saturate(colour) * 255) of the float frame for every pixel —
    drawn, cleared, scissored out or left untouched by a draw without clear — with and without double buffering."""
    from ifrit_v2_b200 import renderer as R
    W, H = 200, 120
    r2 = gpu_renderer_factory()
    host8 = np.zeros((H, W, 4), np.uint8)
    start = np.random.default_rng(5).random((H, W, 4), dtype=np.float32)
    r2.bindFrameBuffer(start.copy(), None, W, H, useDoubleBuffer=False)
    r2.bindPresentBuffer(host8)
    vb, idx = scenes.random_tris(60, 91, size=0.4)
    vb[:, 4:8] = vb[:, 4:8] * 1.6 - 0.3          # colours outside [0, 1]: saturation
    sc = [(10, 8, 120, 70)]
    kw = dict(msaa=4, scissors=sc, clear=False, color=start, depth=np.zeros((H, W), np.float32), depth_func=O.CMP_ALWAYS)
    o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    R.setup_scene(r2, vb, idx, 0, O.FS_VARYING0, 1, IDENT)
    r2.setMsaaSamples(4); r2.SetDepthFunc(O.CMP_ALWAYS)
    r2.SetScissors(sc); r2.setScissorTestEnable(True)
    r2.drawElements()                                # no clear: untouched pixels present what the framebuffer holds
    r2.synchronize()
    col = r2.readColor()
    assert (np.abs(col - o["color"]).max(-1) <= 1e-6).mean() > 0.999
    assert np.array_equal(host8, _pack8(col))
    assert np.array_equal(r2.readPresent(), host8)
    # cleared frame, double buffered: the host image lags one draw
    r2.bindFrameBuffer(np.zeros((H, W, 4), np.float32), None, W, H, useDoubleBuffer=True)
    r2.setScissorTestEnable(False); r2.setMsaaSamples(1)
    r2.setClearValues([(.25, .5, .75, 1)], 1.0)
    frames = []
    for seed in (92, 93):
        vb, idx = scenes.random_tris(60, seed, size=0.4)
        R.setup_scene(r2, vb, idx, 0, O.FS_VARYING0, 1, IDENT)
        r2.clear(); r2.drawElements(); r2.synchronize()
        frames.append(_pack8(r2.readColor()))
        assert np.array_equal(r2.readPresent(), frames[-1])
    assert np.array_equal(host8, frames[0])
    r2.bindPresentBuffer(None, 0)


@pytest.mark.parametrize("pipelining", [True, False])
def test_back_to_back_draws_with_rebinding(gpu_renderer_factory, pipelining):
    """Frame pipelining (geometry stage of draw k + 1 on its own stream, overlapping draw k's tile kernel): four draws
    issued back to back without a host sync, each after re-binding the vertex / index buffers and rewriting the uniform
    block, must give the frame the oracle gets by drawing them in order (second pair without a clear)."""
    from ifrit_v2_b200 import renderer as R
    W, H = 320, 180
    r2 = gpu_renderer_factory()
    r2.setFramePipelining(pipelining)
    r2.bindFrameBuffer(None, None, W, H)
    vbm, idxm = scenes.mesh_sphere(n=32)
    col = dep = None
    r2.bindVertexShader(R.VS_MVP, 2); r2.bindFragmentShader(O.FS_NORMAL_COLOR)
    r2.CreateBuffer(0, scenes.uniform_block(np.eye(4)).nbytes)
    r2.SetCullMode(2); r2.setClearValues([(0, 0, 0, 0)], 1.0)
    for i in range(4):
        mvp = scenes.default_mvp(W, H) @ np.diag([1 + 0.15 * i, 1 - 0.1 * i, 1, 1]).astype(np.float32)
        u = scenes.uniform_block(mvp)
        vb = vbm.copy(); vb[:, 0] += 0.2 * i - 0.3
        idx = idxm if i % 2 == 0 else idxm[: 3 * (len(idxm) // 6)]
        clear = i in (0, 2)
        o = oracle_draw(W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, cull=2, clear=clear, color=col, depth=dep)
        col, dep = o["color"], o["depth"]
        r2.bindVertexBuffer(np.ascontiguousarray(vb)); r2.bindIndexBuffer(np.ascontiguousarray(idx))
        r2.copyHostBufferToBuffer(u, 0)
        if clear:
            r2.clear()
        r2.drawElements()                       # no synchronize() between the draws
    assert np.array_equal(r2.readDepth().view(np.uint32), dep.view(np.uint32))
    assert np.array_equal(r2.readColor().view(np.uint32), col.view(np.uint32))


@pytest.mark.parametrize("pipelining", [True, False])
def test_queue_overflow_between_unsynchronised_draws(gpu_renderer_factory, pipelining):
    """Advisor (round 1, high): the (triangle, tile) queues are sized by the first frame.  A camera that zooms in between
    two draws issued WITHOUT a synchronize makes the pair count grow ~50x past the capacity — the second draw must
    still come out complete (queues re-sized before its fill pass), and so must the third, drawn far away again."""
    from ifrit_v2_b200 import renderer as R
    W, H = 512, 384
    rng = np.random.default_rng(11)
    n = 96                                          # 96 layered triangles; zoomed in, each covers most of the 768 tiles
    vb = np.zeros((3 * n, 8), np.float32)
    for t in range(n):
        c = rng.uniform(-0.2, 0.2, 2)
        for k, (dx, dy) in enumerate(((-0.9, -0.8), (0.9, -0.7), (0.0, 0.95))):
            vb[3 * t + k, :4] = (c[0] + dx, c[1] + dy, 0.1 + 0.8 * rng.random(), 1.0)
            vb[3 * t + k, 4:] = rng.random(4)
    idx = np.arange(3 * n, dtype=np.int32)
    far = np.diag([0.02, 0.02, 1, 1]).astype(np.float32)
    near = np.eye(4, dtype=np.float32)
    r2 = gpu_renderer_factory()
    r2.setFramePipelining(pipelining)
    r2.bindFrameBuffer(None, None, W, H)
    r2.bindVertexBuffer(vb); r2.bindIndexBuffer(idx)
    r2.bindVertexShader(R.VS_MVP, 1); r2.bindFragmentShader(O.FS_VARYING0)
    r2.CreateBuffer(0, scenes.uniform_block(far).nbytes)
    r2.setClearValues([(0, 0, 0, 0)], 1.0)
    for m in (far, near, far):
        r2.copyHostBufferToBuffer(scenes.uniform_block(m), 0)
        r2.clear()
        r2.drawElements()                           # no synchronize() in between
    o_far = oracle_draw(W, H, vb, idx, 1, O.FS_VARYING0, 1, scenes.uniform_block(far))
    assert np.array_equal(r2.readColor().view(np.uint32), o_far["color"].view(np.uint32))
    # and the zoomed frame itself, read right after an unsynchronised far -> near sequence
    for m in (far, near):
        r2.copyHostBufferToBuffer(scenes.uniform_block(m), 0)
        r2.clear()
        r2.drawElements()
    o_near = oracle_draw(W, H, vb, idx, 1, O.FS_VARYING0, 1, scenes.uniform_block(near))
    assert (o_near["prim_id"] >= 0).mean() > 0.5
    assert np.array_equal(r2.readDepth().view(np.uint32), o_near["depth"].view(np.uint32))
    assert np.array_equal(r2.readColor().view(np.uint32), o_near["color"].view(np.uint32))
    assert r2.stats()["tile_pairs"] > 20 * (n // 4 + (W // 16) * (H // 16)) / 10


def test_ordered_path_beyond_4096_triangles_per_tile(r):
    """The ordered (blending) path sorts a tile's triangles in shared memory, 4096 keys at a time: a tile with more than
    that is walked in rounds of consecutive submission-order ranges (round 1 returned IFRIT_ERR_UNSUPPORTED here).  ~9000
    triangles, most of them crowded into two tiles, clipped fans included; colour / depth / last writer bit-exact."""
    W, H = 96, 64
    vb, idx = scenes.random_tris(9000, 411, extent=0.12, size=0.16, neg_w_frac=0.02)
    vb2, idx2 = scenes.random_tris(60, 412, size=0.5)
    vb = np.concatenate([vb, vb2]); idx = np.arange(len(vb), dtype=np.int32)
    vb[:, 7] = 0.05 + 0.3 * vb[:, 7]
    kw = dict(depth_func=O.CMP_ALWAYS, clear_color=(.2, .3, .4, .5), clear_depth=0.5, blend=(6, 7, 1, 0))
    o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    g = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    compare(o, g, exact_color=True)
    assert g["tile_pairs"] > 12000 and g["fragments_shaded"] == o["frags"]


@pytest.mark.parametrize("func", [O.CMP_LESS, O.CMP_GEQUAL, O.CMP_ALWAYS])
@pytest.mark.parametrize("blend", [None, (6, 7, 1, 0)])
def test_fragment_shader_depth_write_late_z(r, func, blend):
    """SURVEY §8f rank 1, late-Z arm: a fragment shader with allowDepthModification (Shaders.h:61) sends the draw down the
    ordered path (forceSort, CU:2517-2527): per pixel, primitives in submission order, depth test on the interpolated
    depth, shader, depth test again on the depth the SHADER left (CU:1803-1805), then colour (blended or replaced) and that
    depth are written.  `depth_mod`: depth += u[22] * varying0.x.  Bit-exact against the oracle, blending on and off."""
    W, H = 200, 120
    vb, idx = scenes.random_tris(140, 311, neg_w_frac=0.1, size=0.35)
    vb2, idx2 = scenes.random_tris(700, 312, size=0.03)
    vb = np.concatenate([vb, vb2]); idx = np.arange(len(vb), dtype=np.int32)
    vb[:, 7] = 0.2 + 0.6 * vb[:, 7]
    u = IDENT.copy() if IDENT.size >= 32 else np.concatenate([IDENT.ravel(), np.zeros(32 - IDENT.size, np.float32)])
    u[22] = -0.35                                   # large enough to reorder overlapping fragments
    kw = dict(depth_func=func, clear_color=(.2, .3, .4, .5), clear_depth=0.6, blend=blend)
    o = oracle_draw(W, H, vb, idx, 0, O.FS_DEPTH_MOD, 1, u, **kw)
    g = gpu_draw(r, W, H, vb, idx, 0, O.FS_DEPTH_MOD, 1, u, **kw)
    compare(o, g, exact_color=True)
    # the shader's depth is what lands in the depth buffer: it differs from the plain draw of the same scene
    plain = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, u, depth_func=func, clear_color=(.2, .3, .4, .5), clear_depth=0.6)
    assert (o["depth"] != plain["depth"]).mean() > 0.2
    from ifrit_v2_b200 import renderer as R
    r.setMsaaSamples(4)
    with pytest.raises(R.IfritError) as e:          # 1x only, as the reference's ordered kernel
        r.drawElements()
    assert e.value.code == R.IFRIT_ERR_UNSUPPORTED
    r.setMsaaSamples(1)


@pytest.mark.parametrize("func", [O.CMP_ALWAYS, O.CMP_LESS, O.CMP_GEQUAL])
@pytest.mark.parametrize("blend", [(6, 7, 1, 0), (1, 9, 9, 1), (8, 1, 0, 6)])
def test_ordered_alpha_blending(r, blend, func):
    """SURVEY §8f rank 1 (pixelShadingAlphaBlendKernel CU:1649-1897): fragments of a pixel are depth-tested against the
    depth the previous fragment left, shaded and blended in submission order.  Colour, depth and last-writer id
    bit-exact against the oracle; small + large triangles, clipped fans, a scissor, ragged frame."""
    W, H = 200, 120
    vb, idx = scenes.random_tris(160, 301, neg_w_frac=0.1, size=0.35)
    vb2, idx2 = scenes.random_tris(900, 302, size=0.03)
    vb = np.concatenate([vb, vb2]); idx = np.arange(len(vb), dtype=np.int32)
    vb[:, 7] = 0.15 + 0.7 * vb[:, 7]            # alpha
    kw = dict(depth_func=func, clear_color=(.2, .3, .4, .5), clear_depth=0.5, blend=blend, scissors=[(5, 3, 180, 110)])
    o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    g = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    compare(o, g, exact_color=True)
    assert g["fragments_shaded"] == o["frags"] > 5000
    # on top of the first frame without a clear
    vb3, idx3 = scenes.random_tris(60, 303, size=0.5)
    kw2 = dict(kw, clear=False, color=o["color"], depth=o["depth"], scissors=None)
    o2 = oracle_draw(W, H, vb3, idx3, 0, O.FS_VARYING0, 1, IDENT, **kw2)
    g2 = gpu_draw(r, W, H, vb3, idx3, 0, O.FS_VARYING0, 1, IDENT, **kw2)
    compare(o2, g2, exact_color=True)
    r.setBlendFunc(False)


def test_alpha_blending_textured_and_limits(r):
    from ifrit_v2_b200 import renderer as R
    W, H = 320, 180
    vb, idx, u = _mesh_case(W, H)
    tex = scenes.texture_checker(size=128, seed=7)
    chain = O.texture_chain(tex, 7)
    samp = dict(filterMode=1, anisotropyEnable=True, maxAnisotropy=16.0)
    kw = dict(cull=0, depth_func=O.CMP_ALWAYS, blend=(1, 1, 1, 1), mip_levels=7, sampler=samp)   # additive: both faces of the sphere
    o = oracle_draw(W, H, vb, idx, 1, O.FS_TEXTURED, 2, u, texture_chain=chain, tex_dims=(128, 128), **kw)
    g = gpu_draw(r, W, H, vb, idx, 1, O.FS_TEXTURED, 2, u, texture=tex, **kw)
    compare(o, g)
    r.setMsaaSamples(4)
    with pytest.raises(R.IfritError) as e:      # 1x only, as the reference's blend kernel
        r.drawElements()
    assert e.value.code == R.IFRIT_ERR_UNSUPPORTED
    r.setMsaaSamples(1)
    r.setBlendFunc(False)


def test_standalone_vertex_stage_matches_fused(r):
    """vertexProcessingKernel CU:330 as its own kernel (pos / vary buffers) vs fused into triangle setup."""
    W, H = 320, 200
    vb, idx = scenes.mesh_sphere(n=40)
    u = scenes.uniform_block(scenes.default_mvp(W, H))
    o = oracle_draw(W, H, vb, idx, O.VS_MVP, O.FS_NORMAL_COLOR, 2, u, cull=2, msaa=4)
    for fused in (False, True):
        g = gpu_draw(r, W, H, vb, idx, O.VS_MVP, O.FS_NORMAL_COLOR, 2, u, cull=2, msaa=4, fused_vs=fused)
        compare(o, g, exact_color=True)
    r.setVertexStageFused(True)


def test_depth_test_disabled_is_always(r):
    vb, idx = scenes.random_tris(100, 22, size=0.3)
    o = oracle_draw(96, 96, vb, idx, 0, O.FS_VARYING0, 1, IDENT, depth_func=O.CMP_ALWAYS)
    g = gpu_draw(r, 96, 96, vb, idx, 0, O.FS_VARYING0, 1, IDENT, depth_func=O.CMP_LESS, depth_test=False)
    compare(o, g, exact_color=True)


def test_assembled_records_bit_exact(r):
    """Stage-level parity: clip / viewport / snap / cull records (CU:1343-1491) field by field."""
    W, H = 320, 200
    vb, idx = scenes.random_tris(4000, 31, neg_w_frac=0.25, size=0.5, extent=1.6)
    for conv in (0, 1):
        for cull in (0, 1, 2):
            gpu_draw(r, W, H, vb, idx, 0, O.FS_PRIM_ID, 0, IDENT, cull=cull, convention=conv)
            got, n = r.readAssembled(2 * len(idx) // 3)
            st = O.make_state(W, H, cull=cull, convention=conv)
            pos, _ = O.vertex_stage(0, IDENT, vb, 0)
            exp = O.assemble(st, pos, idx)
            exp = exp[np.argsort(exp["slot"], kind="stable")]
            assert n == len(exp) and n > 0
            for f in ("prim", "fan", "clipped", "slot", "X", "Y"):
                assert np.array_equal(got[f], exp[f]), f
            for f in ("Z", "IW", "bary"):
                assert np.array_equal(got[f].view(np.uint32), exp[f].view(np.uint32)), f


def test_empty_and_degenerate_inputs(r):
    W, H = 64, 48
    vb = np.zeros((3, 8), np.float32)
    vb[:, 3] = 1.0
    idx = np.arange(3, dtype=np.int32)   # zero-area triangle
    o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, clear_color=(1, 0, 1, 0), clear_depth=7.0)
    g = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, clear_color=(1, 0, 1, 0), clear_depth=7.0)
    compare(o, g, exact_color=True)
    assert (g["prim_id"] == -1).all() and (g["depth"] == 7.0).all()
    # no indices at all: the draw degenerates to the (fused) clear
    g = gpu_draw(r, W, H, vb, idx[:0], 0, O.FS_VARYING0, 1, IDENT, clear_color=(0, 1, 0, 1), clear_depth=3.0)
    assert (g["color"] == np.array([0, 1, 0, 1], np.float32)).all() and (g["depth"] == 3.0).all()
    # vertices at w == 0 and behind the eye
    vb2, idx2 = scenes.random_tris(50, 5, neg_w_frac=0.5)
    vb2[::7, 3] = 0.0
    o = oracle_draw(W, H, vb2, idx2, 0, O.FS_VARYING0, 1, IDENT)
    g = gpu_draw(r, W, H, vb2, idx2, 0, O.FS_VARYING0, 1, IDENT)
    compare(o, g, exact_color=True)


def test_huge_triangles_guard_band(r):
    """Triangles far larger than the screen: warp-cooperative binning + the 64-bit edge path."""
    W, H = 256, 160
    vb, idx = scenes.random_tris(12, 41, size=900.0, extent=3.0)
    vb2, idx2 = scenes.random_tris(40, 42, size=2.5, extent=1.0)
    vb = np.concatenate([vb, vb2]); idx = np.arange(len(vb), dtype=np.int32)
    for msaa in (1, 8):
        o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, msaa=msaa)
        g = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, msaa=msaa)
        compare(o, g)


def test_scissor_and_no_clear_second_draw(r):
    W, H = 128, 96
    vb, idx = scenes.random_tris(120, 51, size=0.4)
    sc = [(10, 8, 50, 40), (70, 50, 30, 30)]
    o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, msaa=4, scissors=sc, clear_color=(.5, .5, .5, 1))
    g = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, msaa=4, scissors=sc, clear_color=(.5, .5, .5, 1))
    compare(o, g, exact_color=True)
    # second draw on top of the first, load-op LOAD (no clear): depth test against the stored depth
    vb2, idx2 = scenes.random_tris(120, 52, size=0.4)
    o2 = oracle_draw(W, H, vb2, idx2, 0, O.FS_VARYING0, 1, IDENT, clear=False, color=o["color"], depth=o["depth"])
    g2 = gpu_draw(r, W, H, vb2, idx2, 0, O.FS_VARYING0, 1, IDENT, clear=False, color=g["color"], depth=g["depth"])
    compare(o2, g2, exact_color=True)


def _mesh_case(W, H, n=40):
    vb, idx = scenes.mesh_sphere(n=n)
    u = scenes.uniform_block(scenes.default_mvp(W, H))
    return vb, idx, u


def test_mesh_normal_color_mvp(r):
    W, H = 320, 180
    vb, idx, u = _mesh_case(W, H)
    for cull, msaa in ((0, 1), (2, 8)):
        o = oracle_draw(W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, cull=cull, msaa=msaa)
        g = gpu_draw(r, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, cull=cull, msaa=msaa)
        compare(o, g, exact_color=True)
        assert (g["prim_id"] >= 0).mean() > 0.3


def test_derivatives_quad(r):
    """ddx / ddy of the uv varying (ShaderBuiltinCuda.cuh:36-76) incl. helper lanes at triangle edges."""
    W, H = 320, 180
    vb, idx, _ = _mesh_case(W, H)
    u = scenes.uniform_block(scenes.default_mvp(W, H), deriv_scale=64.0)
    for msaa in (1, 4):
        o = oracle_draw(W, H, vb, idx, 1, O.FS_DERIV_VIS, 2, u, msaa=msaa, cull=2)
        g = gpu_draw(r, W, H, vb, idx, 1, O.FS_DERIV_VIS, 2, u, msaa=msaa, cull=2)
        compare(o, g, exact_color=True)


@pytest.mark.parametrize("hw", [False, True])
@pytest.mark.parametrize("aniso", [False, True])
@pytest.mark.parametrize("addr", [0, 1, 2, 3, 4])
def test_textured_mesh(r, aniso, addr, hw):
    """hw = False (default): four loads + fp32 lerps (the oracle's arithmetic); hw = True (opt-in FAST): the same 2x2 footprint
    blended by the texture unit with 8-bit weights.  Both must meet the 1/255 on >= 99.9 % bar."""
    W, H = 384, 216
    vb, idx, u = _mesh_case(W, H)
    vb = vb.copy(); vb[:, 8:10] = vb[:, 8:10] * 1.5 - 0.25   # exercise the address modes outside [0,1]
    tex = scenes.texture_checker(size=256, seed=7)
    samp = dict(filterMode=1, addressModeU=addr, addressModeV=addr, borderColor=1, anisotropyEnable=aniso, maxAnisotropy=16.0)
    chain = O.texture_chain(tex, 8)
    o = oracle_draw(W, H, vb, idx, 1, O.FS_TEXTURED, 2, u, cull=2, texture_chain=chain, tex_dims=(256, 256), mip_levels=8, sampler=samp)
    g = gpu_draw(r, W, H, vb, idx, 1, O.FS_TEXTURED, 2, u, cull=2, texture=tex, mip_levels=8, sampler=samp, hw_filter=hw)
    mx = compare(o, g)   # log2f may differ in the last ulp between libm and the device: tolerance 1/255
    assert mx < 0.05
    if not hw:           # software path: beyond the rare log2f-ulp flips the colour is bit-identical
        assert (o["color"].view(np.uint32) == g["color"].view(np.uint32)).mean() > 0.99


@pytest.mark.parametrize("msaa", [1, 8])
@pytest.mark.parametrize("depth_func", ["LESS", "GEQUAL", "ALWAYS", "EQUAL"])
def test_fast_shader_attribute_planes(r, depth_func, msaa):
    """The FAST textured shaders take their varyings from per-triangle attribute planes (kernels_geometry.cu: planes_kernel)
    and the fragment depth from the winning visibility key: a camera INSIDE the sphere (near-plane clipped fans, second
    fan slots, no cull / front cull = swapped winding), ordered and unordered compare ops, 1x and 8x — ids and depth
    bit-exact, colour within 1/255 on >= 99.9 % like every FAST frame."""
    W, H = 400, 232
    vb, idx = scenes.mesh_sphere(n=40)
    import math
    proj = scenes.perspective(60.0 * math.pi / 180.0, W / H, 0.1, 100.0)
    view = scenes.look_at((0.0, 0.0, 1.0), (0.5, 0.3, 0.9), (0, 1, 0))              # eye ON the surface: 5 clipped triangles, 2 second fans
    u = scenes.uniform_block((proj.astype(np.float64) @ view.astype(np.float64)).astype(np.float32))
    tex = scenes.texture_checker(size=256, seed=5)
    chain = O.texture_chain(tex, 8)
    samp = dict(filterMode=1, anisotropyEnable=True, maxAnisotropy=16.0)
    func = getattr(O, "CMP_" + depth_func)
    for cull in (0, 1):
        kw = dict(cull=cull, msaa=msaa, depth_func=func, mip_levels=8, sampler=samp,
                  clear_depth=0.0 if depth_func == "GEQUAL" else 1.0)
        o = oracle_draw(W, H, vb, idx, 1, O.FS_TEXTURED, 2, u, texture_chain=chain, tex_dims=(256, 256), **kw)
        g = gpu_draw(r, W, H, vb, idx, 1, O.FS_TEXTURED, 2, u, texture=tex, hw_filter=True, **kw)
        compare(o, g)
        if depth_func != "EQUAL":
            assert (g["prim_id"] >= 0).mean() > 0.3
        assert g["fragments_shaded"] == o["frags"]


def test_textured_msaa8_nearest_and_lod0(r):
    W, H = 256, 144
    vb, idx, u = _mesh_case(W, H)
    tex = scenes.texture_checker(size=128, seed=9)
    chain = O.texture_chain(tex, 7)
    for fs, samp in ((O.FS_TEXTURED_LOD0, dict(filterMode=1)), (O.FS_TEXTURED, dict(filterMode=0))):
        o = oracle_draw(W, H, vb, idx, 1, fs, 2, u, cull=2, msaa=8, texture_chain=chain, tex_dims=(128, 128), mip_levels=7, sampler=samp)
        for hw in (False, True):
            g = gpu_draw(r, W, H, vb, idx, 1, fs, 2, u, cull=2, msaa=8, texture=tex, mip_levels=7, sampler=samp, hw_filter=hw)
            compare(o, g)


def test_mipmap_generation_bit_exact(r):
    """invokeMipmapGeneration (CUIMG:132-155): repeated bilinear 2:1 blits, odd sizes, 2 layers."""
    rng = np.random.default_rng(3)
    for (w, h, mips, layers, filt) in ((64, 64, 6, 1, 1), (100, 37, 5, 2, 1), (33, 17, 3, 1, 0)):
        lvl0 = rng.random((layers, h, w, 4), dtype=np.float32)
        exp = O.texture_chain(lvl0, mips, layers=layers, filter=filt)
        r.CreateTexture(1, w, h, mips, layers)
        r.copyHostBufferToImage(lvl0, 1, w, h, layers=tuple(range(layers)))
        r.generateMipmap(1, filt)
        got = r.copyImageToHost(1, exp.size // 4)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


def test_cube_map_skybox(r):
    W, H = 160, 120
    rng = np.random.default_rng(8)
    faces = rng.random((6, 32, 32, 4), dtype=np.float32)
    chain = O.texture_chain(faces, 3, layers=6)
    # a full-screen quad whose varying is the view direction
    vb = np.zeros((4, 8), np.float32)
    vb[:, :4] = [[-1, -1, .5, 1], [1, -1, .5, 1], [1, 1, .5, 1], [-1, 1, .5, 1]]
    vb[:, 4:7] = [[-1, -1, 1], [1.3, -1, -0.2], [1, 1, -0.6], [-1, 0.7, 0.4]]   # no zero vector on the diagonal
    idx = np.array([0, 1, 2, 0, 2, 3], np.int32)
    u = scenes.uniform_block(np.eye(4), lod_bias=1.25)
    samp = dict(filterMode=1, addressModeU=2, addressModeV=2)
    o = oracle_draw(W, H, vb, idx, 0, O.FS_SKYBOX, 1, u, texture_chain=chain, tex_dims=(32, 32), mip_levels=3, layers=6, sampler=samp)
    g = gpu_draw(r, W, H, vb, idx, 0, O.FS_SKYBOX, 1, u, texture=faces, mip_levels=3, layers=6, sampler=samp)
    compare(o, g, exact_color=True)


def test_small_triangle_soup_binning_stress(r):
    """S2-style soup at reduced size: queue sizing, shared-memory atomics, ties."""
    W, H = 480, 270
    vb, idx = scenes.soup(ntri=60000, width=W, height=H, seed=42)
    for msaa in (1, 8):
        o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, msaa=msaa, cull=2)
        g = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, msaa=msaa, cull=2)
        compare(o, g, exact_color=True)
        assert g["tile_pairs"] >= g["triangles_assembled"] > 10000


def test_unsupported_verbs_return_status(r):
    from ifrit_v2_b200 import renderer as R
    for fn, args in ((r.bindGeometryShader, ()), (r.SetDepthFunc, (R.IF_COMPARE_OP_NOT_EQUAL,))):
        with pytest.raises(R.IfritError) as e:
            fn(*args)
        assert e.value.code == R.IFRIT_ERR_UNSUPPORTED
    with pytest.raises(R.IfritError) as e:
        r.DrawMeshTasks(1, 0)                 # no mesh shader bound (tests/test_mesh_shaders.py covers the front end)
    assert e.value.code == R.IFRIT_ERR_STATE
    with pytest.raises(R.IfritError) as e:
        r.setBlendFunc(True, 2, 3, 1, 0)      # SRC_COLOR / ONE_MINUS_SRC_COLOR: the reference aborts on them (CU:3466)
    assert e.value.code == R.IFRIT_ERR_UNSUPPORTED
    r.setBlendFunc(blendEnable=False)


@pytest.mark.parametrize("W,H", [(400, 230), (333, 77), (64, 64)])
def test_tma_framebuffer_path_equals_per_thread_stores(r, W, H):
    """North-star N1: framebuffer blocks leave (and, for LOAD-op draws, arrive) through cp.async.bulk.tensor; the
    per-thread store path (iftrcuSetTmaFramebuffer(0); always taken when the row pitch is not a multiple of 16 bytes,
    W = 333) writes the same bits.  Covers: fused clear, empty tiles, LOAD-op draw with scissors on top, MSAA."""
    vb, idx, u = _mesh_case(W, H)
    rng = np.random.default_rng(5)
    start = rng.random((H, W, 4), dtype=np.float32)
    sc = [(W * 0.25, H * 0.2, W * 0.4, H * 0.5)]
    res = {}
    for tma in (True, False):
        a = gpu_draw(r, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, msaa=4, cull=2, clear_color=(.1, .2, .3, 1), tma=tma)
        b = gpu_draw(r, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, msaa=1, cull=1, clear=False, color=start,
                     depth=np.full((H, W), 0.9, np.float32), scissors=sc, tma=tma)
        res[tma] = (a, b)
    for k in (0, 1):
        for f in ("color", "depth", "prim_id"):
            x, y = res[True][k][f], res[False][k][f]
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (k, f)
    # the LOAD-op draw leaves everything outside the scissor exactly as it was
    b = res[True][1]
    outside = np.ones((H, W), bool)
    outside[int(H * 0.2) + 1:int(H * 0.7), int(W * 0.25) + 1:int(W * 0.65)] = False
    assert np.array_equal(b["color"][outside], start[outside])
    o = oracle_draw(W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, msaa=1, cull=1, clear=False, color=start,
                    depth=np.full((H, W), 0.9, np.float32), scissors=sc)
    compare(o, b, exact_color=True)


@pytest.mark.parametrize("mode", [2, 1])
def test_point_and_line_polygon_modes(r, mode):
    """SURVEY §8f rank 4 (unifiedPointRasterEngineKernel CU:2932 / unifiedLineRasterEngineKernel CU:3295): one point per
    index / the three Bresenham edges of every triangle; per pixel the candidate with the smallest z wins, is tested
    strictly against the stored depth, shaded from the vertex's / the perspective-interpolated varyings, and replaces the
    colour.  Bit-exact against the restatement (parity unpinned: the reference's CUDA renderer does not build)."""
    from ifrit_v2_b200 import renderer as R
    W, H = 320, 200
    vb, idx, u = _mesh_case(W, H, n=48)
    rng = np.random.default_rng(3)
    start_c = rng.random((H, W, 4), dtype=np.float32)
    start_d = np.full((H, W), 0.9, np.float32)           # part of the sphere lies behind it
    for kw in (dict(clear_color=(.1, .2, .3, 1), clear_depth=255.0), dict(clear=False, color=start_c, depth=start_d)):
        for fs in (O.FS_NORMAL_COLOR, O.FS_PRIM_ID):
            o = oracle_draw(W, H, vb, idx, 1, fs, 2, u, polygon_mode=mode, **kw)
            g = gpu_draw(r, W, H, vb, idx, 1, fs, 2, u, polygon_mode=mode, **kw)
            compare(o, g, exact_color=True)
            assert g["fragments_shaded"] == o["frags"] > (100 if mode == 2 else 1000)
    # segments that cross w = 0 are clipped at w = 1; vertices behind the eye are dropped in point mode
    vbr, idxr = scenes.random_tris(300, 77, neg_w_frac=0.3, size=0.5)
    o = oracle_draw(W, H, vbr, idxr, 0, O.FS_VARYING0, 1, IDENT, polygon_mode=mode)
    g = gpu_draw(r, W, H, vbr, idxr, 0, O.FS_VARYING0, 1, IDENT, polygon_mode=mode)
    compare(o, g, exact_color=True)
    r.setMsaaSamples(4)
    with pytest.raises(R.IfritError) as e:
        r.drawElements()
    assert e.value.code == R.IFRIT_ERR_UNSUPPORTED
    r.setMsaaSamples(1)
    r.setRasterizerPolygonMode(0)


def test_cpp_host_through_cabi():
    """C++ host code (tests/cpp/host_demo.cpp, written like the reference's GPU demo) over the C-ABI."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "host_demo")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "host_demo.cpp"),
                           "-L" + os.path.join(root, "ifrit-v2_b200", "lib"), "-lifrit_b200",
                           "-Wl,-rpath," + os.path.join(root, "ifrit-v2_b200", "lib"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "covered=" in out.stdout


def test_sort_first_partition_is_bit_identical(r, gpu_renderer_factory):
    """N tile-interleaved instances + pack / unpack == the single-instance frame."""
    import torch
    W, H = 400, 230
    vb, idx, u = _mesh_case(W, H)
    full = gpu_draw(r, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, msaa=4, cull=2, clear_color=(.2, .3, .4, 1))
    from ifrit_v2_b200 import renderer as R
    for world in (2, 3, 8):
        root = gpu_renderer_factory()
        root.bindFrameBuffer(np.zeros((H, W, 4), np.float32), np.zeros((H, W), np.float32))
        for rank in range(world):
            part = gpu_renderer_factory()
            gpu_draw(part, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, msaa=4, cull=2, clear_color=(.2, .3, .4, 1),
                     partition=(rank, world))
            nbytes = R.owned_tile_bytes(W, H, rank, world)
            buf = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
            part.packOwnedTiles(buf.data_ptr(), nbytes)
            part.synchronize()
            torch.cuda.synchronize()
            root.unpackOwnedTiles(buf.data_ptr(), rank, world)
            root.synchronize()
            part.close()
        assert np.array_equal(root.readColor().view(np.uint32), full["color"].view(np.uint32))
        root.close()


def test_distributed_setup_owner_masks_single_gpu(r, gpu_renderer_factory):
    """the distributed triangle setup on one GPU — every emulated rank classifies the draw (iftrcuClassifyTriangles), takes
    the owner masks (iftrcuSetTriangleOwnerMask: the setup pass walks a per-CTA compacted list of its own triangles) and renders its
    tiles; the union is the single-instance frame bit for bit.  Soup of small triangles + large clipped ones, ragged last CTA."""
    import torch
    from ifrit_v2_b200 import renderer as R
    W, H = 400, 230
    vb, idx = scenes.random_tris(5000, 41, size=0.05)
    vb2, idx2 = scenes.random_tris(150, 42, size=0.9, extent=0.9, neg_w_frac=0.2)
    vb = np.concatenate([vb, vb2]); idx = np.arange(len(vb), dtype=np.int32)
    ntri = len(idx) // 3
    kw = dict(msaa=4, cull=0, clear_color=(.2, .3, .4, 1))
    full = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    for world in (2, 8):
        root = gpu_renderer_factory()
        root.bindFrameBuffer(np.zeros((H, W, 4), np.float32), np.zeros((H, W), np.float32))
        live = 0
        for rank in range(world):
            part = gpu_renderer_factory()
            gpu_draw(part, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, partition=(rank, world), **kw)
            mask = torch.zeros((ntri + 15) // 16 * 16, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            part.classifyTriangles(0, ntri, mask.data_ptr())
            part.setTriangleOwnerMask(mask.data_ptr())
            part.clear()
            part.drawElements()
            part.synchronize()
            live += int(((mask[:ntri] >> rank) & 1).sum().item())
            nbytes = R.owned_tile_bytes(W, H, rank, world)
            buf = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
            part.packOwnedTiles(buf.data_ptr(), nbytes)
            part.synchronize()
            torch.cuda.synchronize()
            root.unpackOwnedTiles(buf.data_ptr(), rank, world)
            root.synchronize()
            part.setTriangleOwnerMask(0)
            part.close()
        assert np.array_equal(root.readColor().view(np.uint32), full["color"].view(np.uint32))
        assert ntri * 0.5 < live < ntri * world       # the masks do split the draw
        root.close()


def test_distributed_read_back_is_bit_identical(r, gpu_renderer_factory):
    """SURVEY §8e, read-back side: N tile-interleaved instances each store the tiles they own straight into ONE host image
    (iftrcuCopyOwnedTilesToHost, mapped pinned memory) == the single-instance frame; pixels of other ranks are left alone."""
    from ifrit_v2_b200 import multigpu as MG
    W, H = 400, 230
    vb, idx, u = _mesh_case(W, H)
    full = gpu_draw(r, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, msaa=4, cull=2, clear_color=(.2, .3, .4, 1))
    for world in (1, 2, 8):
        img = MG.SharedHostImage(W, H)
        img.array[...] = -5.0
        for rank in range(world):
            part = gpu_renderer_factory()
            gpu_draw(part, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, msaa=4, cull=2, clear_color=(.2, .3, .4, 1),
                     partition=(rank, world))
            part.copyOwnedTilesToHost(img.ptr)
            part.synchronize()
            if world > 1 and rank == 0:
                other = ~MG.owned_mask(W, H, 0, world)
                assert (img.array[other] == -5.0).all()
            part.close()
        assert np.array_equal(img.array.view(np.uint32), full["color"].view(np.uint32))


def test_full_size_properties_4k(r):
    """BASELINE.json full size (4K, mesh50k): size-independent properties — determinism, idempotence of a
    re-draw under LESS, and BACK-culled == unculled visibility on a closed mesh."""
    W, H = 3840, 2160
    vb, idx = scenes.mesh_sphere(n=160)
    u = scenes.uniform_block(scenes.default_mvp(W, H))
    a = gpu_draw(r, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, cull=2)
    b = gpu_draw(r, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, cull=2)
    assert np.array_equal(a["color"].view(np.uint32), b["color"].view(np.uint32))
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    cov = (a["prim_id"][..., 0] >= 0)
    assert 0.3 < cov.mean() < 0.9
    c = gpu_draw(r, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, cull=0)
    # the sphere is closed: with depth LESS the nearest surface is a front face, so NONE == BACK
    same = (a["prim_id"][..., 0] == c["prim_id"][..., 0]).mean()
    assert same > 0.995
    # re-drawing the same geometry without clearing changes nothing under LESS (idempotence)
    d = gpu_draw(r, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, cull=2, clear=False, color=a["color"], depth=a["depth"])
    assert np.array_equal(d["color"].view(np.uint32), a["color"].view(np.uint32))
    assert (d["prim_id"] == -1).all()


@pytest.mark.gpu
def test_multi_gpu_frame_is_bit_identical():
    """>= 2 GPUs only: peer-memory (NVLink stores into rank 0's framebuffer) and NCCL-gather sort-first frames
    equal the single-GPU frame bit for bit (tests/mgpu_check.py under torchrun)."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(here, "mgpu_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert "MGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---- BASELINE.json configurations at FULL size, against the oracle -------------------------------------------------
def _bench_scene(name):
    import bench
    cfg = bench.CONFIGS[name]
    return cfg, bench.build_scene(cfg)


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("name", ["c2", "c3"])
def test_full_size_c2_c3_vs_oracle(r, name, fast):
    """C2 / C3 exactly as bench.py builds them (mesh50k, 3840x2160, 2048^2 RGBA32F texture with 11 mips, aniso 16;
    C3: 8x MSAA + BACK cull): primitive ids per sample and depth bit-exact, colour within 1/255 on >= 99.9 % of the
    pixels, for the default sampler and for the opt-in FAST (texture-unit) sampler; FS invocation counts equal."""
    cfg, sc = _bench_scene(name)
    W, H = cfg["W"], cfg["H"]
    samp = dict(filterMode=1, anisotropyEnable=True, maxAnisotropy=16.0)
    chain = O.texture_chain(sc["tex"], sc["mips"])
    s = sc["tex"].shape[0]
    kw = dict(msaa=cfg["msaa"], cull=cfg["cull"], clear_depth=255.0, mip_levels=sc["mips"], sampler=samp)
    o = oracle_draw(W, H, sc["vb"], sc["idx"], sc["vs"], O.FS_TEXTURED, sc["nvary"], sc["u"], texture_chain=chain, tex_dims=(s, s), **kw)
    g = gpu_draw(r, W, H, sc["vb"], sc["idx"], sc["vs"], O.FS_TEXTURED, sc["nvary"], sc["u"], texture=sc["tex"], hw_filter=fast, **kw)
    compare(o, g)
    assert g["fragments_shaded"] == o["frags"] and g["triangles_assembled"] == o["assembled"]
    if not fast:   # default sampler: the oracle's arithmetic op for op; only log2f's last ulp may differ
        assert (o["color"].view(np.uint32) == g["color"].view(np.uint32)).mean() > 0.99


def test_full_size_c4_soup_vs_oracle(r):
    """C4 scene generator at 4K, 1 M of its triangles (small-triangle binning path): everything bit-exact."""
    cfg, _ = None, None
    W, H = 3840, 2160
    vb, idx = scenes.soup(ntri=1_000_000, width=W, height=H)
    kw = dict(cull=2, clear_depth=255.0)
    o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    g = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    compare(o, g, exact_color=True)
    assert g["fragments_shaded"] == o["frags"] and g["triangles_assembled"] == o["assembled"]


def test_full_size_c5_mix_vs_oracle(r):
    """C5 scene generator at 7680x4320, 8x MSAA, 1 M of its triangles (90 % tiny, 9.9 % mid, 0.1 % up to 512 px:
    both queue classes, Hi-Z on deep tiles): everything bit-exact."""
    W, H = 7680, 4320
    vb, idx = scenes.scene_mix(ntri=1_000_000, width=W, height=H)
    kw = dict(msaa=8, cull=2, clear_depth=255.0)
    o = oracle_draw(W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    g = gpu_draw(r, W, H, vb, idx, 0, O.FS_VARYING0, 1, IDENT, **kw)
    compare(o, g, exact_color=True)
    assert g["fragments_shaded"] == o["frags"] and g["triangles_assembled"] == o["assembled"]


def test_fast_sampler_precision_contract(r):
    """include/ifrit_b200.h, iftrcuSetSamplerHardwareFilter: the default sampler follows the oracle for texels of any
    range; the opt-in FAST sampler (texture-unit blend, 8-bit weights) is held to 1/255 only for texels in [0, 1] —
    on an HDR-range texture (x 64) its error scales with the range while the default stays exact."""
    W, H = 384, 216
    vb, idx, u = _mesh_case(W, H)
    tex = scenes.texture_checker(size=256, seed=7) * np.float32(64.0)
    samp = dict(filterMode=1, anisotropyEnable=True, maxAnisotropy=16.0)
    chain = O.texture_chain(tex, 8)
    o = oracle_draw(W, H, vb, idx, 1, O.FS_TEXTURED, 2, u, cull=2, texture_chain=chain, tex_dims=(256, 256), mip_levels=8, sampler=samp)
    g = gpu_draw(r, W, H, vb, idx, 1, O.FS_TEXTURED, 2, u, cull=2, texture=tex, mip_levels=8, sampler=samp, hw_filter=False)
    compare(o, g)
    assert (o["color"].view(np.uint32) == g["color"].view(np.uint32)).mean() > 0.99
    f = gpu_draw(r, W, H, vb, idx, 1, O.FS_TEXTURED, 2, u, cull=2, texture=tex, mip_levels=8, sampler=samp, hw_filter=True)
    assert np.array_equal(o["prim_id"], f["prim_id"]) and np.array_equal(o["depth"].view(np.uint32), f["depth"].view(np.uint32))
    d = np.abs(o["color"] - f["color"]).max(-1)
    assert (d <= 64.0 / 255).mean() >= 0.999      # the documented bound: 1/255 x texel range


def test_unsynchronised_draws_with_growing_pair_count(gpu_renderer_factory):
    """draws streamed without a synchronize while the camera zooms in, so that the (triangle, tile)
    pair count more than doubles from frame to frame — every frame must still be complete (the queues are re-sized
    whenever an input of the geometry stage changed), checked against the oracle frame by frame via the
    double-buffered host image."""
    from ifrit_v2_b200 import renderer as R
    W, H = 640, 360
    r2 = gpu_renderer_factory()
    vb, idx = scenes.mesh_sphere(n=96)
    host = np.zeros((H, W, 4), np.float32)
    r2.bindFrameBuffer(host, None, W, H, useDoubleBuffer=True)
    eyes = [(0.0, 0.0, 30.0), (0.0, 0.0, 8.0), (0.0, 0.0, 3.0), (0.0, 0.0, 1.7), (0.0, 0.0, 1.7)]
    R.setup_scene(r2, vb, idx, 1, O.FS_NORMAL_COLOR, 2, scenes.uniform_block(scenes.default_mvp(W, H, eye=eyes[0])))
    r2.setClearValues([(.1, .2, .3, 1)], 1.0)
    r2.clear()
    r2.drawElements()
    first_pairs = r2.stats()["tile_pairs"]         # (synchronises) the far view, before the streamed sequence
    exp = []
    for k, eye in enumerate(eyes):
        u = scenes.uniform_block(scenes.default_mvp(W, H, eye=eye))
        exp.append(oracle_draw(W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, clear_color=(.1, .2, .3, 1))["color"])
        r2.copyHostBufferToBuffer(u, 0)            # new MVP, no synchronize between the draws
        r2.clear()
        r2.drawElements()
    r2.synchronize()
    # double buffered: the host image holds the frame before the last one, the device the last one
    assert np.array_equal(host.view(np.uint32), exp[-2].view(np.uint32)), "frame before last incomplete"
    assert np.array_equal(r2.readColor().view(np.uint32), exp[-1].view(np.uint32)), "last frame incomplete"
    assert r2.stats()["tile_pairs"] > 2 * first_pairs
    # and every intermediate frame, one by one, without a synchronize between its draw and the next
    for k in range(1, len(eyes)):
        for j in (k - 1, k):
            r2.copyHostBufferToBuffer(scenes.uniform_block(scenes.default_mvp(W, H, eye=eyes[j])), 0)
            r2.clear()
            r2.drawElements()
        r2.synchronize()
        assert np.array_equal(host.view(np.uint32), exp[k - 1].view(np.uint32)), f"frame {k - 1} incomplete"
    r2.close()


def test_argument_validation(gpu_renderer_factory):
    """out-of-range blit rectangles / layers, a skybox texture without 6 layers and a too-small
    uniform block are rejected with a status code instead of reaching a kernel."""
    from ifrit_v2_b200 import renderer as R
    r2 = gpu_renderer_factory()
    r2.CreateTexture(1, 32, 32, 2, 1)

    def blit(srcLevel=0, dstLevel=1, srcLayer=0, dstLayer=0, srcRect=(0, 0, 32, 32), dstRect=(0, 0, 16, 16)):
        b = R.IfritImageBlit()
        b.srcSubresource.mipLevel, b.srcSubresource.baseArrayLayer = srcLevel, srcLayer
        b.dstSubresource.mipLevel, b.dstSubresource.baseArrayLayer = dstLevel, dstLayer
        b.srcExtentSt.width, b.srcExtentSt.height, b.srcExtentEd.width, b.srcExtentEd.height = srcRect
        b.dstExtentSt.width, b.dstExtentSt.height, b.dstExtentEd.width, b.dstExtentEd.height = dstRect
        r2.blitImage(1, 1, b, 1)
    blit()                                      # the valid one
    for args in (dict(srcRect=(0, 0, 33, 32)), dict(dstRect=(0, 0, 16, 17)), dict(srcLayer=1), dict(dstRect=(4, 4, 4, 8)), dict(dstLevel=3)):
        with pytest.raises(R.IfritError):
            blit(**args)
    W, H = 64, 64
    vb, idx = scenes.kat2()
    r2.bindFrameBuffer(np.zeros((H, W, 4), np.float32), np.zeros((H, W), np.float32))
    R.setup_scene(r2, vb, idx, 0, O.FS_SKYBOX, 1, IDENT, texture=np.zeros((1, 8, 8, 4), np.float32), mip_levels=0)
    with pytest.raises(R.IfritError):
        r2.drawElements()                       # one layer, the cube sampler reads six
    R.setup_scene(r2, vb, idx, 0, O.FS_DERIV_VIS, 1, IDENT)
    vb2 = np.zeros((3, 12), np.float32); vb2[:, :8] = vb
    r2.bindVertexBuffer(vb2)
    r2.bindVertexShader(0, 2)
    r2.CreateBuffer(0, 64)                      # MVP only: deriv_vis reads float 21
    with pytest.raises(R.IfritError):
        r2.drawElements()
    r2.close()
