"""Host-side mirror of the reference's ``TileRasterRendererCuda`` class
(include/ifrit/softgraphics/engine/tilerastercuda/TileRasterRendererCuda.h:81-115) over the
C-ABI of ``include/ifrit_b200.h``: same verbs, same argument meaning, same non-owning binding
contract; errors raise ``IfritError`` instead of aborting the process.

There is no CPU fallback: constructing a renderer without the CUDA library or without a GPU fails
loudly."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# IFRIT_B200_LIB: developer override used by tools/gpu_ab.sh to time build variants side by side
LIB_PATH = os.environ.get("IFRIT_B200_LIB") or os.path.join(_HERE, "lib", "libifrit_b200.so")

# enums (engine/base/Constants.h:23-96)
IF_CULL_MODE_NONE, IF_CULL_MODE_FRONT, IF_CULL_MODE_BACK, IF_CULL_MODE_FRONT_AND_BACK = 0, 1, 2, 3
(IF_COMPARE_OP_NEVER, IF_COMPARE_OP_LESS, IF_COMPARE_OP_EQUAL, IF_COMPARE_OP_LESS_OR_EQUAL, IF_COMPARE_OP_GREATER,
 IF_COMPARE_OP_NOT_EQUAL, IF_COMPARE_OP_GREATER_OR_EQUAL, IF_COMPARE_OP_ALWAYS) = range(8)
IF_FILTER_NEAREST, IF_FILTER_LINEAR = 0, 1
(IF_SAMPLER_ADDRESS_MODE_REPEAT, IF_SAMPLER_ADDRESS_MODE_MIRRORED_REPEAT, IF_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE,
 IF_SAMPLER_ADDRESS_MODE_CLAMP_TO_BORDER, IF_SAMPLER_ADDRESS_MODE_MIRROR_CLAMP_TO_EDGE) = range(5)
IF_POLYGON_MODE_FILL, IF_POLYGON_MODE_LINE, IF_POLYGON_MODE_POINT = 0, 1, 2
IF_FORMAT_R32G32B32A32_SFLOAT = 1
IFTP_FLOAT4 = 4
VS_PASSTHROUGH, VS_MVP = 0, 1
FS_VARYING0, FS_NORMAL_COLOR, FS_PRIM_ID, FS_TEXTURED, FS_TEXTURED_LOD0, FS_SKYBOX, FS_DERIV_VIS, FS_DEPTH_MOD = range(8)
CONVENTION_CUDA, CONVENTION_CPU = 0, 1
IFRIT_OK, IFRIT_ERR_INVALID, IFRIT_ERR_UNSUPPORTED, IFRIT_ERR_CUDA, IFRIT_ERR_STATE = range(5)


class IfritError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ifrit_b200 error {code}: {msg}")
        self.code = code


class IfritExtent3D(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("depth", C.c_uint32)]


class IfritOffset3D(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("z", C.c_int32)]


class IfritSamplerT(C.Structure):
    _fields_ = [("filterMode", C.c_int), ("addressModeU", C.c_int), ("addressModeV", C.c_int),
                ("addressModeW", C.c_int), ("borderColor", C.c_int), ("anisotropyEnable", C.c_uint8),
                ("maxAnisotropy", C.c_float)]


class IfritImageCreateInfo(C.Structure):
    _fields_ = [("extent", IfritExtent3D), ("tilingMode", C.c_int), ("mipLevels", C.c_uint32),
                ("arrayLayers", C.c_uint32), ("format", C.c_int)]


class IfritImageSubresourceLayers(C.Structure):
    _fields_ = [("mipLevel", C.c_uint32), ("baseArrayLayer", C.c_uint32)]


class IfritImageBlit(C.Structure):
    _fields_ = [("srcSubresource", IfritImageSubresourceLayers), ("srcExtentSt", IfritExtent3D),
                ("srcExtentEd", IfritExtent3D), ("dstSubresource", IfritImageSubresourceLayers),
                ("dstExtentSt", IfritExtent3D), ("dstExtentEd", IfritExtent3D)]


class IfritBufferImageCopy(C.Structure):
    _fields_ = [("bufferOffset", C.c_uint32), ("imageSubresource", IfritImageSubresourceLayers),
                ("imageOffset", IfritOffset3D), ("imageExtent", IfritExtent3D)]


class IfritColorAttachmentBlendState(C.Structure):
    _fields_ = [("blendEnable", C.c_uint8), ("srcColorBlendFactor", C.c_int), ("dstColorBlendFactor", C.c_int),
                ("srcAlphaBlendFactor", C.c_int), ("dstAlphaBlendFactor", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("triangles_submitted", C.c_uint64), ("triangles_assembled", C.c_uint64), ("tile_pairs", C.c_uint64),
                ("fragments_shaded", C.c_uint64), ("kernel_launches", C.c_uint64), ("queue_overflow", C.c_uint32),
                ("reserved", C.c_uint32)]


ASSEMBLED_DTYPE = np.dtype([("prim", "i4"), ("fan", "i4"), ("clipped", "i4"), ("slot", "i4"), ("X", "i4", 3),
                            ("Y", "i4", 3), ("Z", "f4", 3), ("IW", "f4", 3), ("bary", "f4", (3, 2))])

_LIB = None


def load_library():
    """Load lib/libifrit_b200.so; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise IfritError(IFRIT_ERR_CUDA, f"{LIB_PATH} is missing: run `make -C ifrit-v2_b200/csrc` "
                                             "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.iftrcuCreateFragmentShaderFromSource.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.iftrcuCreateMeshShaderFromSource.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
        L.iftrcuCreateTaskShaderFromSource.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
        L.iftrcuCheckMeshShaderSource.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.iftrcuCheckFragmentShaderSource.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
        L.iftrcuGetLastError.restype = C.c_char_p
        L.iftrcuGetLastError.argtypes = [C.c_void_p]
        L.iftrcuOwnedTileBytes.restype = C.c_size_t
        L.iftrcuOwnedTileBytes.argtypes = [C.c_int] * 4
        L.iftrcuCreateInstance.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.iftrcuSetClearValues.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float]
        L.iftrcuCopyImageToHostBuffer.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.iftrcuPackOwnedTiles.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.iftrcuUnpackOwnedTiles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        for name in ("iftrcuBindFrameBuffer", "iftrcuBindFrameBufferDevice"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + \
                ([C.c_int] if name == "iftrcuBindFrameBuffer" else [])
        for name in ("iftrcuBindVertexBuffer", "iftrcuBindVertexBufferDevice"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        for name in ("iftrcuBindIndexBuffer", "iftrcuBindIndexBufferDevice"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.iftrcuBindVertexShader.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.iftrcuSetStream.argtypes = [C.c_void_p, C.c_void_p]
        L.iftrcuSetScissors.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.iftrcuCopyHostBufferToImage.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.iftrcuCopyHostBufferToBuffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.iftrcuBlitImage.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.iftrcuCreateTexture.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.iftrcuCreateSampler.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.iftrcuSetBlendFunc.argtypes = [C.c_void_p, C.c_void_p]
        L.iftrcuReadAssembled.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        for name in ("iftrcuReadColor", "iftrcuReadDepth", "iftrcuReadPrimitiveIds", "iftrcuGetStats"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        for name in ("iftrcuGetDeviceColorPtr", "iftrcuGetDeviceDepthPtr"):
            getattr(L, name).argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        for name in ("iftrcuDestroyInstance", "iftrcuInit", "iftrcuClear", "iftrcuDrawElements", "iftrcuSynchronize"):
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("iftrcuBindFragmentShader", "iftrcuBindGeometryShader", "iftrcuSetScissorTestEnable",
                     "iftrcuSetMsaaSamples", "iftrcuSetRasterizerPolygonMode", "iftrcuSetDepthFunc",
                     "iftrcuSetDepthTestEnable", "iftrcuSetCullMode", "iftrcuSetSampleConvention",
                     "iftrcuSetDebugPrimitiveIds", "iftrcuSetVertexStageFused", "iftrcuSetSamplerHardwareFilter"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        L.iftrcuSetTilePartition.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.iftrcuExportFrameBuffer.argtypes = [C.c_void_p, C.c_void_p]
        L.iftrcuCopyColorToHostAsync.argtypes = [C.c_void_p, C.c_void_p]
        L.iftrcuCopyOwnedTilesToHost.argtypes = [C.c_void_p, C.c_void_p]
        L.iftrcuBindPeerFrameBuffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.iftrcuSetProfiling.argtypes = [C.c_void_p, C.c_int]
        L.iftrcuGetKernelTimes.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.iftrcuGenerateMipmap.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.iftrcuCreateBuffer.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.iftrcuDrawMeshTasks.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.iftrcuLookupVertexShader.argtypes = [C.c_char_p]
        L.iftrcuLookupFragmentShader.argtypes = [C.c_char_p]
        _LIB = L
    return _LIB


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):   # torch tensor
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class TileRasterRendererCuda:
    """Mirror of ``Ifrit::Graphics::SoftGraphics::TileRaster::CUDA::TileRasterRendererCuda``."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.iftrcuCreateInstance(device, C.byref(h))
        if rc != IFRIT_OK:
            raise IfritError(rc, "iftrcuCreateInstance failed: no usable CUDA device (there is no CPU fallback)")
        self._h = h
        self._keep = {}
        self.width = self.height = 0
        self.msaa = 1

    def _ck(self, rc):
        if rc != IFRIT_OK:
            raise IfritError(rc, self._lib.iftrcuGetLastError(self._h).decode())

    def close(self):
        if self._h:
            self._lib.iftrcuDestroyInstance(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- RC verbs -------------------------------------------------------------------------------
    def Init(self):
        self._ck(self._lib.iftrcuInit(self._h))

    def setStream(self, cuda_stream_ptr):
        """cuda_stream_ptr: cudaStream_t as int (0 = legacy default stream); None = the instance's own stream."""
        v = C.c_void_p(-1) if cuda_stream_ptr is None else C.c_void_p(cuda_stream_ptr)
        self._ck(self._lib.iftrcuSetStream(self._h, v))

    def bindFrameBuffer(self, color, depth, width=None, height=None, useDoubleBuffer=False):
        """color: host float32 [H,W,4] (ImageF32 RGBA) or None; depth: host float32 [H,W] or None.
        useDoubleBuffer (reference default: True, TileRasterRendererCuda.h:88): frame k renders while frame k - 1 is
        copied into `color` on a second stream, i.e. the host image lags one drawElements() behind.  The mirror
        defaults to False so that draw -> synchronize -> look at `color` shows the frame just drawn."""
        if width is None:
            height, width = (color.shape[0], color.shape[1]) if color is not None else depth.shape[:2]
        self._keep["fb"] = (color, depth)
        self.width, self.height = width, height
        self._ck(self._lib.iftrcuBindFrameBuffer(self._h, _ptr(color), _ptr(depth), width, height, int(useDoubleBuffer)))

    def bindFrameBufferDevice(self, color, depth, width, height):
        self._keep["fb"] = (color, depth)
        self.width, self.height = width, height
        self._ck(self._lib.iftrcuBindFrameBufferDevice(self._h, _ptr(color), _ptr(depth), width, height))

    def bindPresentBuffer(self, host_rgba8, fmt=1):
        """host_rgba8: uint8 [H,W,4] (or None for a device-only image); fmt 1 = RGBA8 UNORM, 0 = off.  While bound,
        the per-draw copy-back carries this packed image instead of the RGBA32F one."""
        self._keep["present"] = host_rgba8
        self._ck(self._lib.iftrcuBindPresentBuffer(self._h, _ptr(host_rgba8), int(fmt)))

    def readPresent(self):
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self._ck(self._lib.iftrcuReadPresent(self._h, _ptr(out)))
        return out

    def copyColorToHostAsync(self, host_ptr):
        self._ck(self._lib.iftrcuCopyColorToHostAsync(self._h, C.c_void_p(host_ptr)))

    def exportFrameBuffer(self):
        """128-byte CUDA IPC blob of the renderer-owned device framebuffer (colour + depth)."""
        buf = (C.c_ubyte * 128)()
        self._ck(self._lib.iftrcuExportFrameBuffer(self._h, buf))
        return bytes(buf)

    def bindPeerFrameBuffer(self, handle, width, height):
        """Render straight into another process's / GPU's framebuffer over NVLink (peer memory)."""
        self._keep["fb"] = None
        self.width, self.height = width, height
        buf = (C.c_ubyte * 128).from_buffer_copy(handle)
        self._ck(self._lib.iftrcuBindPeerFrameBuffer(self._h, buf, width, height))

    def bindVertexBuffer(self, data, vertexCount=None, numAttributes=None, device=False):
        """data: float32 [V, 4*nattr] interleaved FLOAT4 attributes (VertexBuffer.h:28-90)."""
        if vertexCount is None:
            vertexCount, numAttributes = data.shape[0], data.shape[1] // 4
        self._keep["vb"] = data
        layout = np.full(numAttributes, IFTP_FLOAT4, dtype=np.int32)
        fn = self._lib.iftrcuBindVertexBufferDevice if device else self._lib.iftrcuBindVertexBuffer
        self._ck(fn(self._h, _ptr(data), vertexCount, _ptr(layout), numAttributes))

    def bindIndexBuffer(self, indices, count=None, device=False):
        if count is None:
            count = int(indices.shape[0])
        self._keep["ib"] = indices
        fn = self._lib.iftrcuBindIndexBufferDevice if device else self._lib.iftrcuBindIndexBuffer
        self._ck(fn(self._h, _ptr(indices), count))

    def bindVertexShader(self, shader, numVaryings):
        sid = shader if isinstance(shader, int) else self._lib.iftrcuLookupVertexShader(shader.encode())
        layout = np.full(max(numVaryings, 1), IFTP_FLOAT4, dtype=np.int32)
        self._ck(self._lib.iftrcuBindVertexShader(self._h, sid, _ptr(layout), numVaryings))

    def bindFragmentShader(self, shader):
        sid = shader if isinstance(shader, int) else self._lib.iftrcuLookupFragmentShader(shader.encode())
        self._ck(self._lib.iftrcuBindFragmentShader(self._h, sid))

    def createFragmentShaderFromSource(self, source, numVaryings, needsDerivatives=False, name="user_fs"):
        """CUDA source defining `__device__ float4 fs_main(const IfritFsInput& in)` -> shader id for bindFragmentShader
        (NVRTC; include/ifrit_b200.h: iftrcuCreateFragmentShaderFromSource)."""
        sid = C.c_int(-1)
        self._ck(self._lib.iftrcuCreateFragmentShaderFromSource(self._h, name.encode(), source.encode(), int(numVaryings),
                                                                int(bool(needsDerivatives)), C.byref(sid)))
        return sid.value

    def createMeshShaderFromSource(self, source, name="user_ms"):
        sid = C.c_int(-1)
        self._ck(self._lib.iftrcuCreateMeshShaderFromSource(self._h, name.encode(), source.encode(), C.byref(sid)))
        return sid.value

    def createTaskShaderFromSource(self, source, name="user_ts"):
        sid = C.c_int(-1)
        self._ck(self._lib.iftrcuCreateTaskShaderFromSource(self._h, name.encode(), source.encode(), C.byref(sid)))
        return sid.value

    def bindGeometryShader(self, shader=0):
        self._ck(self._lib.iftrcuBindGeometryShader(self._h, shader))

    def bindMeshShader(self, shaderId, numVaryings, localSize=(1, 1, 1)):
        d = (C.c_int * max(numVaryings, 1))(*([IFTP_FLOAT4] * max(numVaryings, 1)))
        self._ck(self._lib.iftrcuBindMeshShader(self._h, shaderId, d, numVaryings, *[int(x) for x in localSize]))

    def bindTaskShader(self, shaderId=-1):
        self._ck(self._lib.iftrcuBindTaskShader(self._h, shaderId, None, 0))

    def CreateTexture(self, slotId, width, height, mipLevels=0, arrayLayers=1):
        ci = IfritImageCreateInfo()
        ci.extent.width, ci.extent.height, ci.extent.depth = width, height, 1
        ci.tilingMode, ci.mipLevels, ci.arrayLayers, ci.format = 1, mipLevels, arrayLayers, IF_FORMAT_R32G32B32A32_SFLOAT
        self._ck(self._lib.iftrcuCreateTexture(self._h, slotId, C.byref(ci)))

    def CreateSampler(self, slotId, filterMode=IF_FILTER_LINEAR, addressModeU=0, addressModeV=0, addressModeW=0,
                      borderColor=0, anisotropyEnable=False, maxAnisotropy=1.0):
        s = IfritSamplerT(filterMode, addressModeU, addressModeV, addressModeW, borderColor, int(anisotropyEnable),
                          maxAnisotropy)
        self._ck(self._lib.iftrcuCreateSampler(self._h, slotId, C.byref(s)))

    def generateMipmap(self, slotId, filter=IF_FILTER_LINEAR):
        self._ck(self._lib.iftrcuGenerateMipmap(self._h, slotId, filter))

    def blitImage(self, srcSlotId, dstSlotId, region, filter=IF_FILTER_LINEAR):
        self._ck(self._lib.iftrcuBlitImage(self._h, srcSlotId, dstSlotId, C.byref(region), filter))

    def copyHostBufferToImage(self, src, dstSlot, width, height, layers=(0,)):
        """Upload whole level-0 layers; src: float32 [len(layers), H, W, 4]."""
        self._keep[("img", dstSlot)] = src
        regions = (IfritBufferImageCopy * len(layers))()
        for i, layer in enumerate(layers):
            regions[i].bufferOffset = i * width * height * 16
            regions[i].imageSubresource.mipLevel = 0
            regions[i].imageSubresource.baseArrayLayer = layer
            regions[i].imageExtent.width, regions[i].imageExtent.height, regions[i].imageExtent.depth = width, height, 1
        self._ck(self._lib.iftrcuCopyHostBufferToImage(self._h, _ptr(src), dstSlot, regions, len(layers)))

    def copyImageToHost(self, slot, ntexels):
        out = np.empty(ntexels * 4, dtype=np.float32)
        self._ck(self._lib.iftrcuCopyImageToHostBuffer(self._h, slot, _ptr(out), out.nbytes))
        return out

    def CreateBuffer(self, slotId, bufSize):
        self._ck(self._lib.iftrcuCreateBuffer(self._h, slotId, bufSize))

    def copyHostBufferToBuffer(self, src, dstSlot, size=None):
        self._keep[("buf", dstSlot)] = src
        self._ck(self._lib.iftrcuCopyHostBufferToBuffer(self._h, _ptr(src), dstSlot, src.nbytes if size is None else size))

    def SetScissors(self, scissors):
        a = np.ascontiguousarray(scissors, dtype=np.float32).reshape(-1, 4)
        self._ck(self._lib.iftrcuSetScissors(self._h, _ptr(a), a.shape[0]))

    def setScissorTestEnable(self, option):
        self._ck(self._lib.iftrcuSetScissorTestEnable(self._h, int(option)))

    def setMsaaSamples(self, samples):
        self._ck(self._lib.iftrcuSetMsaaSamples(self._h, samples))
        self.msaa = samples

    def setRasterizerPolygonMode(self, mode):
        self._ck(self._lib.iftrcuSetRasterizerPolygonMode(self._h, mode))

    def setBlendFunc(self, blendEnable=False, src=6, dst=7, srcA=6, dstA=7):
        s = IfritColorAttachmentBlendState(int(blendEnable), src, dst, srcA, dstA)
        self._ck(self._lib.iftrcuSetBlendFunc(self._h, C.byref(s)))

    def SetDepthFunc(self, depthFunc):
        self._ck(self._lib.iftrcuSetDepthFunc(self._h, depthFunc))

    def SetDepthTestEnable(self, option):
        self._ck(self._lib.iftrcuSetDepthTestEnable(self._h, int(option)))

    def SetCullMode(self, cullMode):
        self._ck(self._lib.iftrcuSetCullMode(self._h, cullMode))

    def setClearValues(self, clearColors, clearDepth):
        a = np.ascontiguousarray(clearColors, dtype=np.float32).reshape(-1, 4)
        self._ck(self._lib.iftrcuSetClearValues(self._h, _ptr(a), a.shape[0], clearDepth))

    def clear(self):
        self._ck(self._lib.iftrcuClear(self._h))

    def drawElements(self):
        self._ck(self._lib.iftrcuDrawElements(self._h))

    def DrawMeshTasks(self, numWorkGroups, firstWorkGroup):
        self._ck(self._lib.iftrcuDrawMeshTasks(self._h, numWorkGroups, firstWorkGroup))

    # ---- extensions -----------------------------------------------------------------------------
    def setSampleConvention(self, convention):
        self._ck(self._lib.iftrcuSetSampleConvention(self._h, convention))

    def setTilePartition(self, rank, world):
        self._ck(self._lib.iftrcuSetTilePartition(self._h, rank, world))

    def setDebugPrimitiveIds(self, enable=True):
        self._ck(self._lib.iftrcuSetDebugPrimitiveIds(self._h, int(enable)))

    def setSamplerHardwareFilter(self, enable=True):
        self._ck(self._lib.iftrcuSetSamplerHardwareFilter(self._h, int(enable)))

    def setTmaFramebuffer(self, enable=True):
        self._ck(self._lib.iftrcuSetTmaFramebuffer(self._h, int(enable)))

    def setTileChunk(self, entries=0):
        self._ck(self._lib.iftrcuSetTileChunk(self._h, int(entries)))

    def setTileGridHint(self, enable=True):
        self._ck(self._lib.iftrcuSetTileGridHint(self._h, int(enable)))

    def setFramePipelining(self, enable=True):
        self._ck(self._lib.iftrcuSetFramePipelining(self._h, int(enable)))

    def setVertexStageFused(self, enable=True):
        self._ck(self._lib.iftrcuSetVertexStageFused(self._h, int(enable)))

    def synchronize(self):
        self._ck(self._lib.iftrcuSynchronize(self._h))

    def readColor(self):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self._lib.iftrcuReadColor(self._h, _ptr(out)))
        return out

    def readDepth(self):
        out = np.empty((self.height, self.width), dtype=np.float32)
        self._ck(self._lib.iftrcuReadDepth(self._h, _ptr(out)))
        return out

    def readPrimitiveIds(self):
        out = np.empty((self.height, self.width, self.msaa), dtype=np.int32)
        self._ck(self._lib.iftrcuReadPrimitiveIds(self._h, _ptr(out)))
        return out

    def deviceColorPtr(self):
        p = C.c_void_p()
        self._ck(self._lib.iftrcuGetDeviceColorPtr(self._h, C.byref(p)))
        return p.value

    def deviceDepthPtr(self):
        p = C.c_void_p()
        self._ck(self._lib.iftrcuGetDeviceDepthPtr(self._h, C.byref(p)))
        return p.value

    def stats(self):
        s = Stats()
        self._ck(self._lib.iftrcuGetStats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_ if k != "reserved"}

    KERNEL_NAMES = ("vertex", "setup_count", "scan", "fill", "tile")

    def setProfiling(self, enable=True):
        self._ck(self._lib.iftrcuSetProfiling(self._h, int(enable)))

    def kernelTimes(self):
        """Average per-kernel device time in ms over the profiled draws: (dict, number of draws)."""
        ms = np.zeros(5, dtype=np.float32)
        n = C.c_int()
        self._ck(self._lib.iftrcuGetKernelTimes(self._h, _ptr(ms), C.byref(n)))
        return dict(zip(self.KERNEL_NAMES, (float(x) for x in ms))), n.value

    def readAssembled(self, cap):
        out = np.zeros(max(cap, 1), dtype=ASSEMBLED_DTYPE)
        n = C.c_int()
        self._ck(self._lib.iftrcuReadAssembled(self._h, _ptr(out), cap, C.byref(n)))
        return out[: min(n.value, cap)], n.value

    # sort-first frame flags (peer-memory mode), see include/ifrit_b200.h
    def classifyTriangles(self, first, count, dev_mask_ptr):
        self._ck(self._lib.iftrcuClassifyTriangles(self._h, int(first), int(count), C.c_void_p(dev_mask_ptr)))

    def geometryStream(self):
        out = C.c_void_p()
        self._ck(self._lib.iftrcuGetGeometryStream(self._h, C.byref(out)))
        return out.value or 0

    def setTriangleOwnerMask(self, dev_mask_ptr):
        self._ck(self._lib.iftrcuSetTriangleOwnerMask(self._h, C.c_void_p(dev_mask_ptr) if dev_mask_ptr else None))

    def setFrameRing(self, count):
        self._ck(self._lib.iftrcuSetFrameRing(self._h, int(count)))

    def selectFrame(self, index):
        self._ck(self._lib.iftrcuSelectFrame(self._h, int(index)))

    def frameSignal(self, frame_id):
        self._ck(self._lib.iftrcuFrameSignal(self._h, C.c_uint32(frame_id)))

    def frameWait(self, frame_id):
        self._ck(self._lib.iftrcuFrameWait(self._h, C.c_uint32(frame_id)))

    def frameRelease(self, frame_id, cleared=False):
        fn = self._lib.iftrcuFrameReleaseCleared if cleared else self._lib.iftrcuFrameRelease
        self._ck(fn(self._h, C.c_uint32(frame_id)))

    def frameRetire(self, frame_id, cleared=False):
        self._ck(self._lib.iftrcuFrameRetire(self._h, C.c_uint32(frame_id), int(bool(cleared))))

    def frameFence(self, frame_id, cleared=False):
        fn = self._lib.iftrcuFrameFenceCleared if cleared else self._lib.iftrcuFrameFence
        self._ck(fn(self._h, C.c_uint32(frame_id)))

    def copyOwnedTilesToHost(self, host_ptr):
        """Sort-first distributed read-back: this rank's tiles -> the shared full-frame RGBA32F host image (stream order)."""
        self._ck(self._lib.iftrcuCopyOwnedTilesToHost(self._h, C.c_void_p(host_ptr)))

    def packOwnedTiles(self, dev_ptr, nbytes):
        used = C.c_size_t()
        self._ck(self._lib.iftrcuPackOwnedTiles(self._h, C.c_void_p(dev_ptr), nbytes, C.byref(used)))
        return used.value

    def unpackOwnedTiles(self, dev_ptr, rank, world):
        self._ck(self._lib.iftrcuUnpackOwnedTiles(self._h, C.c_void_p(dev_ptr), rank, world))


def owned_tile_bytes(width, height, rank, world):
    return load_library().iftrcuOwnedTileBytes(width, height, rank, world)


def check_mesh_shader_source(source, is_task=False):
    """Compile-only check of a run-time mesh / task shader (no GPU needed): (status code, compiler log)."""
    L = load_library()
    buf = C.create_string_buffer(1 << 16)
    rc = L.iftrcuCheckMeshShaderSource(source.encode(), int(bool(is_task)), buf, len(buf))
    return rc, buf.value.decode(errors="replace")


def check_fragment_shader_source(source, numVaryings, needsDerivatives=False, msaa=1):
    """Compile-only check of a run-time fragment shader (no GPU needed): (status code, compiler log)."""
    L = load_library()
    buf = C.create_string_buffer(1 << 16)
    rc = L.iftrcuCheckFragmentShaderSource(source.encode(), int(numVaryings), int(bool(needsDerivatives)), int(msaa), buf, len(buf))
    return rc, buf.value.decode(errors="replace")


def setup_scene(r, vb, idx, vs, fs, nvary, uniforms, texture=None, mip_levels=0, sampler=None, layers=1):
    """Convenience used by tests / bench: bind geometry, shaders, uniform block and one texture."""
    r.bindVertexBuffer(np.ascontiguousarray(vb, dtype=np.float32))
    r.bindIndexBuffer(np.ascontiguousarray(idx, dtype=np.int32))
    r.bindVertexShader(vs, nvary)
    r.bindFragmentShader(fs)
    u = np.ascontiguousarray(uniforms, dtype=np.float32)
    r.CreateBuffer(0, u.nbytes)
    r.copyHostBufferToBuffer(u, 0)
    if texture is not None:
        t = np.ascontiguousarray(texture, dtype=np.float32)
        if t.ndim == 3:
            t = t[None]
        h, w = t.shape[1], t.shape[2]
        r.CreateTexture(0, w, h, mip_levels, layers)
        r.copyHostBufferToImage(t, 0, w, h, layers=tuple(range(layers)))
        if mip_levels:
            r.generateMipmap(0, IF_FILTER_LINEAR)
        r.CreateSampler(0, **(sampler or {}))
