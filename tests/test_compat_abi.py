"""The reference's shared-object C ABI for its data objects (include/ifrit_b200_compat.h) and the renderer verbs that
take those handles — SURVEY §8(b) row 1.  CPU tests: every declared symbol is exported, the objects behave like the
reference's (checked against the reference's own library where oracle/_ref holds it).  GPU tests: a scene built through
ifcr* / iftr* / ifbuf* handles renders bit-identically to the pointer path, and host code written against the
reference's real C++ classes (tests/cpp/ref_headers_bind.cpp) drives the renderer through the C++ mirror."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
import pytest
import ifrit_v2_b200  # noqa: F401
from ifrit_v2_b200 import renderer as R, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libifrit_ref.so")
REF_BIND = os.path.join(ROOT, "oracle", "_ref", "ref_bind")
vp = C.c_void_p
IFTP_FLOAT4 = 4


def _declare(L):
    for name in ("ifcrCreateImageFP32", "ifcrGetImageRawDataFP32", "iftrCreateFrameBuffer", "iftrCreateVaryingDescriptor",
                 "iftrCreateVertexBuffer", "ifbufCreateBufferManager", "ifbufCreateBuffer"):
        getattr(L, name).restype = vp
    L.ifcrCreateImageFP32.argtypes = [C.c_size_t] * 3
    L.ifcrGetImageRawDataFP32.argtypes = [vp]
    L.ifcrDestroyImageFP32.argtypes = [vp]
    L.iftrSetFrameBufferColorAttachmentFP32.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    L.iftrSetFrameBufferDepthAttachmentFP32.argtypes = [vp, vp]
    L.iftrDestroyFrameBuffer.argtypes = [vp]
    L.iftrWriteVaryingDescriptor.argtypes = [vp, vp, C.c_size_t]
    L.iftrDestroyVaryingDescriptor.argtypes = [vp]
    L.iftrSetVertexBufferLayout.argtypes = [vp, vp, C.c_size_t]
    L.iftrSetVertexBufferSize.argtypes = [vp, C.c_size_t]
    L.iftrAllocateVertexBuffer.argtypes = [vp]
    L.iftrSetVertexBufferValueFloat4.argtypes = [vp, C.c_int, C.c_int, vp]
    L.iftrDestroyVertexBuffer.argtypes = [vp]
    L.ifbufCreateBuffer.argtypes = [vp, C.c_size_t]
    L.ifbufBufferData.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    L.ifbufDestroyBuffer.argtypes = [vp]
    L.ifbufDestroyBufferManager.argtypes = [vp]
    return L


def _ours():
    L = _declare(R.load_library())
    L.iftrcuBindFrameBufferHandle.argtypes = [vp, vp, C.c_int]
    L.iftrcuBindVertexBufferHandle.argtypes = [vp, vp]
    L.iftrcuBindIndexBufferHandle.argtypes = [vp, vp]
    L.iftrcuBindVertexShaderHandle.argtypes = [vp, C.c_int, vp]
    L.iftrcuCopyBufferHandleToBuffer.argtypes = [vp, vp, C.c_int]
    return L


def test_compat_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "ifrit_b200_compat.h")).read()
    names = sorted(set(re.findall(r"IFRIT_B200_API[^;(]*?\b(\w+)\s*\(", hdr)))
    assert len(names) == 26, names
    L = R.load_library()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ifrit_b200_compat.h but not exported"
    # the 21 names of the reference's own export headers, verbatim
    for n in ("ifcrCreateImageFP32 ifcrDestroyImageFP32 ifcrGetImageRawDataFP32 iftrCreateFrameBuffer iftrDestroyFrameBuffer "
              "iftrSetFrameBufferColorAttachmentFP32 iftrSetFrameBufferDepthAttachmentFP32 iftrCreateVaryingDescriptor "
              "iftrDestroyVaryingDescriptor iftrWriteVaryingDescriptor iftrCreateVertexBuffer iftrDestroyVertexBuffer "
              "iftrSetVertexBufferLayout iftrSetVertexBufferSize iftrAllocateVertexBuffer iftrSetVertexBufferValueFloat4 "
              "ifbufCreateBufferManager ifbufDestroyBufferManager ifbufCreateBuffer ifbufDestroyBuffer ifbufBufferData").split():
        assert n in names


def _image_script(L):
    """The same calls against either library: what a caller can observe of an ImageF32."""
    im = L.ifcrCreateImageFP32(7, 5, 4)
    raw = L.ifcrGetImageRawDataFP32(im)
    a = np.ctypeslib.as_array(C.cast(raw, C.POINTER(C.c_float)), shape=(5, 7, 4))
    first = a.copy()
    a[2, 3, 1] = 42.0
    again = np.ctypeslib.as_array(C.cast(L.ifcrGetImageRawDataFP32(im), C.POINTER(C.c_float)), shape=(5 * 7 * 4,))
    flat_index = int(np.flatnonzero(again == 42.0)[0])
    L.ifcrDestroyImageFP32(im)
    return first, flat_index


def test_image_object_behaves_like_the_reference():
    first, idx = _image_script(_ours())
    assert idx == (2 * 7 + 3) * 4 + 1          # [(y*W + x)*C + c], Image.h:224
    if os.path.exists(REF_SO):
        rfirst, ridx = _image_script(_declare(C.CDLL(REF_SO)))
        assert ridx == idx and rfirst.shape == first.shape
    assert not first.any()                      # zero-initialised (Image.h: calloc-style constructor)


def test_buffer_and_descriptor_objects_accept_the_reference_call_sequences():
    L = _ours()
    man = L.ifbufCreateBufferManager()
    buf = L.ifbufCreateBuffer(man, 12)
    data = np.arange(3, dtype=np.int32)
    L.ifbufBufferData(buf, data.ctypes.data, 0, 12)
    L.ifbufBufferData(buf, data.ctypes.data, 8, 12)        # out of range: dropped, no write past the end
    vb = L.iftrCreateVertexBuffer()
    lay = (C.c_int * 3)(IFTP_FLOAT4, 0, IFTP_FLOAT4)       # IFTP_UNDEFINED is stored as FLOAT4 (EngineCommonExport.cpp:108-109)
    L.iftrSetVertexBufferLayout(vb, lay, 3)
    L.iftrSetVertexBufferSize(vb, 2)
    L.iftrAllocateVertexBuffer(vb)
    v = (C.c_float * 4)(1, 2, 3, 4)
    L.iftrSetVertexBufferValueFloat4(vb, 1, 2, v)
    L.iftrSetVertexBufferValueFloat4(vb, 5, 0, v)          # out of range: ignored
    vd = L.iftrCreateVaryingDescriptor()
    L.iftrWriteVaryingDescriptor(vd, lay, 3)
    # no renderer without a GPU, but the binders validate their handles first
    assert L.iftrcuBindVertexBufferHandle(None, vb) == R.IFRIT_ERR_INVALID
    assert L.iftrcuBindIndexBufferHandle(None, buf) == R.IFRIT_ERR_INVALID
    L.iftrDestroyVaryingDescriptor(vd)
    L.iftrDestroyVertexBuffer(vb)
    L.ifbufDestroyBuffer(buf)
    L.ifbufDestroyBufferManager(man)


def test_binders_instantiate_against_reference_headers():
    """tests/cpp/ref_headers_bind.cpp compiled against /root/reference/include (oracle/Makefile target ref_bind)."""
    if not os.path.exists(REF_BIND):
        pytest.skip("oracle/_ref/ref_bind not built (no /root/reference here)")
    out = subprocess.run([REF_BIND], capture_output=True, text=True)
    assert out.returncode == 0 and "instantiate" in out.stdout, out.stdout + out.stderr


def _scene(W, H):
    vb, idx = scenes.mesh_sphere(n=24)
    u = scenes.uniform_block(scenes.default_mvp(W, H))
    return np.ascontiguousarray(vb, np.float32), np.ascontiguousarray(idx, np.int32), np.ascontiguousarray(u, np.float32)


def _pointer_path(W, H, vb, idx, u, fs):
    r = R.TileRasterRendererCuda(0)
    r.Init()
    color = np.zeros((H, W, 4), np.float32)
    depth = np.zeros((H, W), np.float32)
    r.bindFrameBuffer(color, depth)
    R.setup_scene(r, vb, idx, R.VS_MVP, fs, vb.shape[1] // 4 - 1, u)
    r.SetCullMode(R.IF_CULL_MODE_BACK)
    r.SetDepthFunc(1)
    r.setClearValues([(0, 0, 0, 0)], 255.0)
    r.clear()
    r.drawElements()
    r.synchronize()
    out = r.readColor().copy(), r.readDepth().copy()
    r.close()
    return out


@pytest.mark.gpu
def test_scene_built_through_handles_matches_pointer_path():
    W, H = 200, 120
    vb, idx, u = _scene(W, H)
    want_c, want_d = _pointer_path(W, H, vb, idx, u, R.FS_NORMAL_COLOR)
    L = _ours()
    color, depth = L.ifcrCreateImageFP32(W, H, 4), L.ifcrCreateImageFP32(W, H, 1)
    fb = L.iftrCreateFrameBuffer()
    L.iftrSetFrameBufferColorAttachmentFP32(fb, (vp * 1)(color), 1)
    L.iftrSetFrameBufferDepthAttachmentFP32(fb, depth)
    nattr = vb.shape[1] // 4
    hvb = L.iftrCreateVertexBuffer()
    lay = (C.c_int * nattr)(*([IFTP_FLOAT4] * nattr))
    L.iftrSetVertexBufferLayout(hvb, lay, nattr)
    L.iftrSetVertexBufferSize(hvb, vb.shape[0])
    L.iftrAllocateVertexBuffer(hvb)
    for i in range(vb.shape[0]):
        for a in range(nattr):
            L.iftrSetVertexBufferValueFloat4(hvb, i, a, vb[i, 4 * a:4 * a + 4].ctypes.data)
    man = L.ifbufCreateBufferManager()
    hib = L.ifbufCreateBuffer(man, idx.nbytes)
    L.ifbufBufferData(hib, idx.ctypes.data, 0, idx.nbytes)
    hub = L.ifbufCreateBuffer(man, 128)
    L.ifbufBufferData(hub, u.ctypes.data, 0, 128)
    vd = L.iftrCreateVaryingDescriptor()
    L.iftrWriteVaryingDescriptor(vd, lay, nattr - 1)
    r = R.TileRasterRendererCuda(0)
    r.Init()
    h = r._h
    assert L.iftrcuBindFrameBufferHandle(h, fb, 0) == 0
    r.width, r.height = W, H               # the Python mirror sizes its read-back arrays from these
    assert L.iftrcuBindVertexBufferHandle(h, hvb) == 0
    assert L.iftrcuBindIndexBufferHandle(h, hib) == 0
    assert L.iftrcuBindVertexShaderHandle(h, R.VS_MVP, vd) == 0
    r.bindFragmentShader(R.FS_NORMAL_COLOR)
    r.CreateBuffer(0, 128)
    assert L.iftrcuCopyBufferHandleToBuffer(h, hub, 0) == 0
    r.SetCullMode(R.IF_CULL_MODE_BACK)
    r.SetDepthFunc(1)
    r.setClearValues([(0, 0, 0, 0)], 255.0)
    r.clear()
    r.drawElements()
    r.synchronize()
    got_c = np.ctypeslib.as_array(C.cast(L.ifcrGetImageRawDataFP32(color), C.POINTER(C.c_float)), shape=(H, W, 4))
    assert np.array_equal(got_c.view(np.uint32), want_c.view(np.uint32))     # copied back into the ImageF32 (CU:3899-3909)
    assert np.array_equal(r.readDepth().view(np.uint32), want_d.view(np.uint32))
    assert (got_c[..., 3] > 0).mean() > 0.2
    r.close()
    for d, o in ((L.iftrDestroyVaryingDescriptor, vd), (L.ifbufDestroyBuffer, hub), (L.ifbufDestroyBuffer, hib),
                 (L.ifbufDestroyBufferManager, man), (L.iftrDestroyVertexBuffer, hvb), (L.iftrDestroyFrameBuffer, fb),
                 (L.ifcrDestroyImageFP32, depth), (L.ifcrDestroyImageFP32, color)):
        d(o)


@pytest.mark.gpu
def test_reference_typed_host_code_drives_the_renderer(tmp_path):
    """Host code using the reference's own FrameBuffer / VertexBuffer / VaryingDescriptor classes (unmodified reference
    build) + the C++ mirror's templated binders == the ctypes pointer path, bit for bit."""
    if not os.path.exists(REF_BIND):
        pytest.skip("oracle/_ref/ref_bind not built")
    W, H = 200, 120
    vb, idx, u = _scene(W, H)
    want_c, _ = _pointer_path(W, H, vb, idx, u, R.FS_NORMAL_COLOR)
    scene, out = str(tmp_path / "scene.bin"), str(tmp_path / "color.bin")
    with open(scene, "wb") as f:
        np.array([W, H, vb.shape[0], vb.shape[1] // 4, idx.size, R.FS_NORMAL_COLOR], np.int32).tofile(f)
        vb.tofile(f)
        idx.tofile(f)
        u32 = np.zeros(32, np.float32)
        u32[:u.size] = u.ravel()[:32]
        u32.tofile(f)
    res = subprocess.run([REF_BIND, scene, out], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    got = np.fromfile(out, np.float32).reshape(H, W, 4)
    assert np.array_equal(got.view(np.uint32), want_c.view(np.uint32))
