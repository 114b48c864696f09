"""Fragment shaders compiled at run time from CUDA source (reference: the shader JIT behind
engine/base/ShaderRuntime.h:24-37).  CPU: the compile-only check (NVRTC needs no device).  GPU: the registry's shaders
re-written as source produce the built-in kernels' frames bit for bit."""
import numpy as np
import pytest
import ifrit_v2_b200  # noqa: F401
from ifrit_v2_b200 import renderer as R, scenes

NORMAL_COLOR = """
__device__ float4 fs_main(const IfritFsInput& in) {     // DemoFragmentShaderCuda
    const float4 v = in.varying[0];
    return make_float4(fmaf(v.x, 0.5f, 0.5f), fmaf(v.y, 0.5f, 0.5f), fmaf(v.z, 0.5f, 0.5f), 0.5f);
}
"""
TEXTURED = """
__device__ float4 fs_main(const IfritFsInput& in) {     // texture(0, 0, uv, bias) * max(N.L, 0.2)
    const float4 n = in.varying[0], t = in.varying[1];
    const float4 c = ifritTexture(in, 0, 0, make_float2(t.x, t.y), make_float2(in.ddx[1].x, in.ddx[1].y),
                                  make_float2(in.ddy[1].x, in.ddy[1].y), in.uniform[20]);
    const float len2 = fmaf(n.z, n.z, fmaf(n.y, n.y, n.x * n.x));
    const float dot = fmaf(n.z, in.uniform[18], fmaf(n.y, in.uniform[17], n.x * in.uniform[16]));
    const float ndl = fmaxf(dot / sqrtf(len2), 0.2f);
    return make_float4(c.x * ndl, c.y * ndl, c.z * ndl, 1.0f);
}
"""
STRIPES = """
__device__ float4 fs_main(const IfritFsInput& in) {     // not in the registry: primitive id stripes lit by varying 0
    const float s = (in.primitiveId & 1) ? 1.0f : 0.25f;
    return make_float4(s * fabsf(in.varying[0].x), s * fabsf(in.varying[0].y), (float)in.fanIndex, 1.0f);
}
"""


def _have_nvrtc():
    rc, _ = R.check_fragment_shader_source(NORMAL_COLOR, 1)
    return rc != R.IFRIT_ERR_UNSUPPORTED


def test_compile_only_check_accepts_and_rejects():
    if not _have_nvrtc():
        pytest.skip("libnvrtc.so.12 not loadable here")
    for src, nv, dv, msaa in ((NORMAL_COLOR, 1, False, 1), (TEXTURED, 2, True, 8)):
        rc, log = R.check_fragment_shader_source(src, nv, dv, msaa)
        assert rc == 0, log
    rc, log = R.check_fragment_shader_source("__device__ float4 fs_main(const IfritFsInput& in) { return nope; }", 1)
    assert rc == R.IFRIT_ERR_INVALID
    assert '"nope" is undefined' in log and "user_fs(1)" in log        # line numbers are the user's
    assert R.check_fragment_shader_source(NORMAL_COLOR, 9)[0] == R.IFRIT_ERR_INVALID     # more varyings than the pipeline carries


def _draw(r, W, H, fs, msaa, vb, idx, u, tex=None):
    color = np.zeros((H, W, 4), np.float32)
    depth = np.zeros((H, W), np.float32)
    r.bindFrameBuffer(color, depth)
    kw = dict(texture=tex, mip_levels=7, sampler=dict(filterMode=1, anisotropyEnable=True, maxAnisotropy=16.0)) if tex is not None else {}
    R.setup_scene(r, vb, idx, R.VS_MVP, fs, 2, u, **kw)
    r.setMsaaSamples(msaa)
    r.SetCullMode(R.IF_CULL_MODE_BACK)
    r.SetDepthFunc(1)
    r.setClearValues([(0.1, 0.2, 0.3, 1.0)], 255.0)
    r.setSamplerHardwareFilter(False)
    r.clear()
    r.drawElements()
    r.synchronize()
    return r.readColor().copy(), r.readDepth().copy()


@pytest.mark.gpu
def test_registry_shaders_from_source_match_builtins(gpu_renderer_factory):
    W, H = 360, 200
    vb, idx = scenes.mesh_sphere(n=40)
    u = scenes.uniform_block(scenes.default_mvp(W, H))
    tex = scenes.texture_checker(size=128, seed=7)
    r = gpu_renderer_factory()
    nc = r.createFragmentShaderFromSource(NORMAL_COLOR, 1, False, name="normal_color.cu")
    tx = r.createFragmentShaderFromSource(TEXTURED, 2, True, name="textured.cu")
    assert nc >= 1000 and tx == nc + 1
    for msaa in (1, 8):                      # 8x: compiled lazily at the first draw with that sample count
        want = _draw(r, W, H, R.FS_NORMAL_COLOR, msaa, vb, idx, u)
        got = _draw(r, W, H, nc, msaa, vb, idx, u)
        assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32)) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))
    want = _draw(r, W, H, R.FS_TEXTURED, 4, vb, idx, u, tex=tex)
    got = _draw(r, W, H, tx, 4, vb, idx, u, tex=tex)
    assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32))
    assert (got[0][..., 3] == 1.0).mean() > 0.2


@pytest.mark.gpu
def test_custom_shader_and_error_reporting(gpu_renderer_factory):
    W, H = 256, 160
    vb, idx = scenes.mesh_sphere(n=24)
    u = scenes.uniform_block(scenes.default_mvp(W, H))
    r = gpu_renderer_factory()
    sid = r.createFragmentShaderFromSource(STRIPES, 1)
    c, _ = _draw(r, W, H, sid, 1, vb, idx, u)
    r.setDebugPrimitiveIds(True)
    ids = None
    c2, _ = _draw(r, W, H, sid, 1, vb, idx, u)
    ids = r.readPrimitiveIds()[..., 0]
    cov = ids >= 0
    assert cov.mean() > 0.2 and np.array_equal(c, c2)
    # stripes follow the primitive id the visibility pass resolved
    n0 = np.abs(c2[..., 0][cov])
    odd = (ids[cov] & 1) == 1
    assert np.all(c2[..., 3][cov] == 1.0) and np.all(c2[..., 2][cov] == 0.0)
    assert n0[odd].mean() > 2.0 * n0[~odd].mean()
    with pytest.raises(R.IfritError) as e:
        r.createFragmentShaderFromSource("__device__ float4 fs_main(const IfritFsInput& in) { return 1; }", 1)
    assert e.value.code == R.IFRIT_ERR_INVALID and "error" in str(e.value)
    r.setBlendFunc(True, 6, 7, 1, 0)
    r.bindFragmentShader(sid)
    with pytest.raises(R.IfritError) as e:
        r.drawElements()
    assert e.value.code == R.IFRIT_ERR_UNSUPPORTED
    r.setBlendFunc(False)
