"""Mesh / task front end (meshShadingEntryPoint / meshShadingWithTaskShaderEntryPoint CU:375-500, RC:88-89, 115): mesh
and task shaders are CUDA source compiled at run time; their output is compacted in work-group order and drawn like drawElements.
CPU: compile-only checks.  GPU: a meshlet mesh shader that re-emits an indexed mesh from a buffer reproduces the drawElements frame bit
for bit (ids included), with and without a task shader in front."""
import numpy as np
import pytest
import ifrit_v2_b200  # noqa: F401
from ifrit_v2_b200 import renderer as R, scenes

# buffer slot 1: interleaved vertices (3 float4 each), slot 2: indices as floats (exact below 2^24); 64 triangles per work group
MESHLET_MS = """
__device__ void ms_main(const IfritMeshInput& in, float4* outPos, float4* outVaryings, int* outIndices, int& outNumVertices, int& outNumIndices) {
    const float4* vb = reinterpret_cast<const float4*>(ifritBuffer(in, 1));
    const float* ib = ifritBuffer(in, 2);
    const int ntri = (int)in.uniform[24];
    int wg = in.workGroupId;
    if (in.payload) wg = reinterpret_cast<const int*>(in.payload)[0] + in.workGroupId;     // task path: first meshlet of the parent
    const int t0 = wg * 64, n = min(64, ntri - t0);
    const int t = in.localInvocation.x;
    if (t == 0) { outNumVertices = n > 0 ? 3 * n : 0; outNumIndices = n > 0 ? 3 * n : 0; }
    if (t >= n) return;
    for (int j = 0; j < 3; j++) {
        const int vi = (int)ib[3 * (t0 + t) + j];
        outPos[3 * t + j] = ifritMvp(in, vb[3 * vi]);
        outVaryings[(3 * t + j) * 2 + 0] = vb[3 * vi + 1];
        outVaryings[(3 * t + j) * 2 + 1] = vb[3 * vi + 2];
        outIndices[3 * t + j] = 3 * t + j;
    }
}
"""
# every task work group launches 4 meshlets of 64 threads and tells them where its range starts
MESHLET_TS = """
__device__ void ts_main(int workGroupId, void* outPayload, int3* outMeshWorkGroups, int& outNumMeshWorkGroups) {
    reinterpret_cast<int*>(outPayload)[0] = 4 * workGroupId;
    for (int i = 0; i < 4; i++) outMeshWorkGroups[i] = make_int3(64, 1, 1);
    outNumMeshWorkGroups = 4;
}
"""


def test_compile_only_checks():
    rc, log = R.check_mesh_shader_source(MESHLET_MS)
    if rc == R.IFRIT_ERR_UNSUPPORTED:
        pytest.skip("libnvrtc.so.12 not loadable here")
    assert rc == 0, log
    rc, log = R.check_mesh_shader_source(MESHLET_TS, is_task=True)
    assert rc == 0, log
    rc, log = R.check_mesh_shader_source("__device__ void ms_main(int wrong) {}")
    assert rc == R.IFRIT_ERR_INVALID and "error" in log


def _frame(r, W, H):
    return r.readColor().copy(), r.readDepth().copy(), r.readPrimitiveIds().copy()


@pytest.mark.gpu
@pytest.mark.parametrize("with_task", [False, True])
def test_meshlet_shader_reproduces_indexed_draw(gpu_renderer_factory, with_task):
    W, H = 360, 200
    vb, idx = scenes.mesh_sphere(n=40)
    ntri = len(idx) // 3
    u = scenes.uniform_block(scenes.default_mvp(W, H))
    u[24] = ntri
    r = gpu_renderer_factory()
    r.bindFrameBuffer(None, None, W, H)
    R.setup_scene(r, vb, idx, R.VS_MVP, R.FS_NORMAL_COLOR, 2, u)
    r.setMsaaSamples(4)
    r.SetCullMode(R.IF_CULL_MODE_BACK)
    r.SetDepthFunc(1)
    r.setDebugPrimitiveIds(True)
    r.setClearValues([(0.1, 0.2, 0.3, 1.0)], 255.0)
    r.clear(); r.drawElements(); r.synchronize()
    want = _frame(r, W, H)
    # the same mesh through the mesh front end: vertices and indices live in buffer slots 1 and 2
    r.CreateBuffer(1, vb.nbytes); r.copyHostBufferToBuffer(np.ascontiguousarray(vb, np.float32), 1)
    fidx = np.ascontiguousarray(idx, np.float32)
    r.CreateBuffer(2, fidx.nbytes); r.copyHostBufferToBuffer(fidx, 2)
    ms = r.createMeshShaderFromSource(MESHLET_MS, name="meshlet.cu")
    r.bindMeshShader(ms, 2, (64, 1, 1))
    nmeshlets = (ntri + 63) // 64
    if with_task:
        ts = r.createTaskShaderFromSource(MESHLET_TS, name="meshlet_task.cu")
        r.bindTaskShader(ts)
        groups = (nmeshlets + 3) // 4
    else:
        groups = nmeshlets
    r.clear(); r.DrawMeshTasks(groups, 0); r.synchronize()
    got = _frame(r, W, H)
    assert np.array_equal(got[2], want[2])                                   # primitive ids: ordered compaction
    assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))
    assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32))
    assert (got[2] >= 0).mean() > 0.2
    # the user's vertex / index buffers are bound again afterwards
    r.clear(); r.drawElements(); r.synchronize()
    assert np.array_equal(r.readColor().view(np.uint32), want[0].view(np.uint32))
