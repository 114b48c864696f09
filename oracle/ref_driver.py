"""ctypes driver for the UNMODIFIED reference CPU renderer (TEST INFRASTRUCTURE).

Loads oracle/_ref/libifrit_ref.so — built by oracle/Makefile from the sources under
/root/reference — and drives it through the reference's own C ABI
(include/ifrit/softgraphics/engine/export/TileRasterRendererExport.h:33-76,
EngineCommonExport.h:33-70, core/export/CommonExport.h:26-31, BufferManagerExport.h:25-35).
Operating rules from SURVEY.md §8c: forceDeterministic=0, depth LESS, never call
iftrDestroyInstance (UB in the reference)."""
import ctypes as C
import os
import time
import numpy as np
from . import oracle as _orc

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "libifrit_ref.so")
IFTP_FLOAT4 = 4
REF_DEPTH_CLEAR = np.frombuffer(np.uint32(0x7F7F7F7F).tobytes(), dtype=np.float32)[0]
REF_THREADS = 17  # 1 master + 16 workers, compile-time constant (TileRasterContext.h:52)
_LIB = None


def available():
    return os.path.exists(REF_PATH)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(REF_PATH)
        vp = C.c_void_p
        for name in ("ifcrCreateImageFP32", "ifcrGetImageRawDataFP32", "iftrCreateFrameBuffer",
                     "iftrCreateVaryingDescriptor", "iftrCreateVertexBuffer", "ifbufCreateBufferManager",
                     "ifbufCreateBuffer", "iftrCreateInstance"):
            getattr(L, name).restype = vp
        L.ifcrCreateImageFP32.argtypes = [C.c_size_t] * 3
        L.ifcrGetImageRawDataFP32.argtypes = [vp]
        L.iftrSetFrameBufferColorAttachmentFP32.argtypes = [vp, C.POINTER(vp), C.c_size_t]
        L.iftrSetFrameBufferDepthAttachmentFP32.argtypes = [vp, vp]
        L.iftrWriteVaryingDescriptor.argtypes = [vp, vp, C.c_size_t]
        L.iftrSetVertexBufferLayout.argtypes = [vp, vp, C.c_size_t]
        L.iftrSetVertexBufferSize.argtypes = [vp, C.c_size_t]
        L.iftrAllocateVertexBuffer.argtypes = [vp]
        L.iftrSetVertexBufferValueFloat4.argtypes = [vp, C.c_int, C.c_int, vp]
        L.ifbufCreateBuffer.argtypes = [vp, C.c_size_t]
        L.ifbufBufferData.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
        for name in ("iftrInit", "iftrClear"):
            getattr(L, name).argtypes = [vp]
        for name in ("iftrBindFrameBuffer", "iftrBindVertexBuffer", "iftrBindIndexBuffer",
                     "iftrBindFragmentShaderFunc"):
            getattr(L, name).argtypes = [vp, vp]
        L.iftrBindVertexShaderFunc.argtypes = [vp, vp, vp]
        L.iftrSetDepthFunc.argtypes = [vp, C.c_int]
        L.iftrOptsetForceDeterministic.argtypes = [vp, C.c_int]
        L.iftrOptSetDepthTestEnable.argtypes = [vp, C.c_int]
        L.iftrDrawLegacy.argtypes = [vp, C.c_int, C.c_int]
        L.iftrSetBlendFunc.argtypes = [vp, vp]
        _LIB = L
    return _LIB


_VS = {(0, 1): "orc_ref_vs_passthrough1", (0, 2): "orc_ref_vs_passthrough2",
       (1, 1): "orc_ref_vs_mvp1", (1, 2): "orc_ref_vs_mvp2"}
_FS = {_orc.FS_VARYING0: "orc_ref_fs_varying0", _orc.FS_NORMAL_COLOR: "orc_ref_fs_normal_color",
       _orc.FS_TEXTURED_LOD0: "orc_ref_fs_textured_lod0"}


class RefRenderer:
    """One reference `TileRasterRenderer` instance bound to a scene (non-owning bindings, so
    every object handed to the reference is kept alive on this object)."""

    def __init__(self, width, height, vb, idx, vs_id, fs_id, nvary, uniforms, texture=None):
        L, O = lib(), _orc.lib()
        self.width, self.height = width, height
        vb = np.ascontiguousarray(vb, dtype=np.float32)
        self.idx = np.ascontiguousarray(idx, dtype=np.int32)
        nverts, nattr = vb.shape[0], vb.shape[1] // 4
        self.uniforms = np.ascontiguousarray(uniforms, dtype=np.float32)
        O.orc_ref_set_uniforms(self.uniforms.ctypes.data_as(C.c_void_p), self.uniforms.nbytes)
        if texture is not None:
            self.texture = np.ascontiguousarray(texture, dtype=np.float32)
            O.orc_ref_set_texture(self.texture.ctypes.data_as(C.c_void_p), self.texture.shape[1],
                                  self.texture.shape[0])
        self.color_img = L.ifcrCreateImageFP32(width, height, 4)
        self.depth_img = L.ifcrCreateImageFP32(width, height, 1)
        self.fb = L.iftrCreateFrameBuffer()
        arr = (C.c_void_p * 1)(self.color_img)
        L.iftrSetFrameBufferColorAttachmentFP32(self.fb, arr, 1)
        L.iftrSetFrameBufferDepthAttachmentFP32(self.fb, self.depth_img)
        self.vbuf = L.iftrCreateVertexBuffer()
        layout = np.full(nattr, IFTP_FLOAT4, dtype=np.int32)
        L.iftrSetVertexBufferLayout(self.vbuf, layout.ctypes.data_as(C.c_void_p), nattr)
        L.iftrSetVertexBufferSize(self.vbuf, nverts)
        L.iftrAllocateVertexBuffer(self.vbuf)
        v4 = vb.reshape(nverts, nattr, 4)
        for i in range(nverts):
            for a in range(nattr):
                L.iftrSetVertexBufferValueFloat4(self.vbuf, i, a, v4[i, a].ctypes.data)
        self.mgr = L.ifbufCreateBufferManager()
        self.ibuf = L.ifbufCreateBuffer(self.mgr, self.idx.nbytes)
        L.ifbufBufferData(self.ibuf, self.idx.ctypes.data_as(C.c_void_p), 0, self.idx.nbytes)
        self.vdesc = L.iftrCreateVaryingDescriptor()
        vlayout = np.full(nvary, IFTP_FLOAT4, dtype=np.int32)
        L.iftrWriteVaryingDescriptor(self.vdesc, vlayout.ctypes.data_as(C.c_void_p), nvary)
        self.h = L.iftrCreateInstance()
        L.iftrInit(self.h)
        L.iftrBindFrameBuffer(self.h, self.fb)
        L.iftrBindVertexBuffer(self.h, self.vbuf)
        L.iftrBindIndexBuffer(self.h, self.ibuf)
        vs = C.cast(getattr(O, _VS[(vs_id, nvary)]), C.c_void_p)
        fs = C.cast(getattr(O, _FS[fs_id]), C.c_void_p)
        L.iftrBindVertexShaderFunc(self.h, vs, self.vdesc)
        L.iftrBindFragmentShaderFunc(self.h, fs)
        L.iftrSetDepthFunc(self.h, _orc.CMP_LESS)
        L.iftrOptsetForceDeterministic(self.h, 0)
        L.iftrOptSetDepthTestEnable(self.h, 1)

    def set_blend(self, src, dst, src_a, dst_a):
        """iftrSetBlendFunc (TileRasterRendererExport.h; TileRasterRenderer.cpp:212-330): IfritColorAttachmentBlendState =
        {bool blendEnable; IfritBlendFactor src, dst, srcAlpha, dstAlpha}.  Blending switches the reference to its
        ordered (deterministic) mode, TileRasterWorker.cpp:410."""
        class _BS(C.Structure):
            _fields_ = [("blendEnable", C.c_uint8), ("src", C.c_int32), ("dst", C.c_int32), ("srcA", C.c_int32),
                        ("dstA", C.c_int32)]
        self._blend = _BS(1, src, dst, src_a, dst_a)
        lib().iftrSetBlendFunc(self.h, C.byref(self._blend))

    def set_depth_test(self, enable):
        lib().iftrOptSetDepthTestEnable(self.h, int(enable))

    def draw(self, clear=True):
        lib().iftrDrawLegacy(self.h, len(self.idx), int(clear))

    def timed_draws(self, n):
        out = []
        for _ in range(n):
            t0 = time.perf_counter()
            self.draw(True)
            out.append(time.perf_counter() - t0)
        return out

    def color(self):
        p = lib().ifcrGetImageRawDataFP32(self.color_img)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(self.height, self.width, 4))
        return a.copy()

    def depth(self):
        p = lib().ifcrGetImageRawDataFP32(self.depth_img)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(self.height, self.width))
        return a.copy()
