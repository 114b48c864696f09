"""GPU debug: the handle path of tests/test_compat_abi.py step by step."""
import sys, os, faulthandler, ctypes as C, numpy as np
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_compat_abi as T
from ifrit_v2_b200 import renderer as R
W, H = 200, 120
vb, idx, u = T._scene(W, H)
want_c, want_d = T._pointer_path(W, H, vb, idx, u, R.FS_NORMAL_COLOR)
print("pointer path ok", want_c.shape, float(want_c.sum()), flush=True)
L = T._ours(); vp = C.c_void_p
color, depth = L.ifcrCreateImageFP32(W, H, 4), L.ifcrCreateImageFP32(W, H, 1)
fb = L.iftrCreateFrameBuffer()
L.iftrSetFrameBufferColorAttachmentFP32(fb, (vp * 1)(color), 1)
L.iftrSetFrameBufferDepthAttachmentFP32(fb, depth)
nattr = vb.shape[1] // 4
hvb = L.iftrCreateVertexBuffer()
lay = (C.c_int * nattr)(*([4] * nattr))
L.iftrSetVertexBufferLayout(hvb, lay, nattr); L.iftrSetVertexBufferSize(hvb, vb.shape[0]); L.iftrAllocateVertexBuffer(hvb)
for i in range(vb.shape[0]):
    for a in range(nattr):
        L.iftrSetVertexBufferValueFloat4(hvb, i, a, vb[i, 4 * a:4 * a + 4].ctypes.data)
man = L.ifbufCreateBufferManager()
hib = L.ifbufCreateBuffer(man, idx.nbytes); L.ifbufBufferData(hib, idx.ctypes.data, 0, idx.nbytes)
hub = L.ifbufCreateBuffer(man, 128); L.ifbufBufferData(hub, u.ctypes.data, 0, 128)
vd = L.iftrCreateVaryingDescriptor(); L.iftrWriteVaryingDescriptor(vd, lay, nattr - 1)
print("handles built", flush=True)
r = R.TileRasterRendererCuda(0); r.Init(); h = r._h
print("fb", L.iftrcuBindFrameBufferHandle(h, fb, 0), flush=True); r.width, r.height = W, H
print("vb", L.iftrcuBindVertexBufferHandle(h, hvb), flush=True)
print("ib", L.iftrcuBindIndexBufferHandle(h, hib), flush=True)
print("vs", L.iftrcuBindVertexShaderHandle(h, R.VS_MVP, vd), flush=True)
r.bindFragmentShader(R.FS_NORMAL_COLOR); r.CreateBuffer(0, 128)
print("ub", L.iftrcuCopyBufferHandleToBuffer(h, hub, 0), flush=True)
r.SetCullMode(R.IF_CULL_MODE_BACK); r.SetDepthFunc(1); r.setClearValues([(0, 0, 0, 0)], 255.0)
r.clear(); r.drawElements(); r.synchronize()
print("drawn", r.stats(), flush=True)
raw = L.ifcrGetImageRawDataFP32(color)
print("raw ptr", hex(raw), flush=True)
got = np.ctypeslib.as_array(C.cast(raw, C.POINTER(C.c_float)), shape=(H, W, 4)).copy()
print("got sum", float(got.sum()), "dev sum", float(r.readColor().sum()), "equal", np.array_equal(got, want_c), flush=True)
d = np.abs(got - want_c).max(-1)
print("max diff", float(d.max()), "n diff", int((d > 0).sum()), flush=True)
