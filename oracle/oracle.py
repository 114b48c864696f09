"""ctypes wrapper of oracle/libifrit_oracle.so (TEST INFRASTRUCTURE — see ifrit_oracle.h)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CONV_CUDA, CONV_CPU = 0, 1
CULL_NONE, CULL_FRONT, CULL_BACK, CULL_FRONT_AND_BACK = 0, 1, 2, 3
CMP_NEVER, CMP_LESS, CMP_EQUAL, CMP_LEQUAL, CMP_GREATER, CMP_NOT_EQUAL, CMP_GEQUAL, CMP_ALWAYS = range(8)
VS_PASSTHROUGH, VS_MVP = 0, 1
FS_VARYING0, FS_NORMAL_COLOR, FS_PRIM_ID, FS_TEXTURED, FS_TEXTURED_LOD0, FS_SKYBOX, FS_DERIV_VIS, FS_DEPTH_MOD = range(8)


class State(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("msaa", C.c_int32), ("cull_mode", C.c_int32),
                ("depth_func", C.c_int32), ("convention", C.c_int32), ("scissor_enable", C.c_int32),
                ("nscissor", C.c_int32), ("scissor", (C.c_float * 4) * 16), ("do_clear", C.c_int32),
                ("clear_color", C.c_float * 4), ("clear_depth", C.c_float),
                ("blend_enable", C.c_int32), ("blend_coef", (C.c_float * 4) * 4)]


# updateBlendStateKernel CU:3436-3537: {x1, x(1 - src.a), x(1 - dst.a), zero flag} per IfritBlendFactor — note that
# the reference's SRC_ALPHA / DST_ALPHA rows select (1 - a) and its ONE_MINUS_* rows select a (sic, replicated)
BLEND_ZERO, BLEND_ONE, BLEND_SRC_ALPHA, BLEND_ONE_MINUS_SRC_ALPHA, BLEND_DST_ALPHA, BLEND_ONE_MINUS_DST_ALPHA = 0, 1, 6, 7, 8, 9
BLEND_COEF = {BLEND_ONE: (1, 0, 0, 0), BLEND_ZERO: (0, 0, 0, 1), BLEND_DST_ALPHA: (0, 0, 1, 0), BLEND_SRC_ALPHA: (0, 1, 0, 0),
              BLEND_ONE_MINUS_DST_ALPHA: (1, 0, -1, 0), BLEND_ONE_MINUS_SRC_ALPHA: (1, -1, 0, 0)}


class Texture(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32),
                ("mip_levels", C.c_int32), ("layers", C.c_int32)]


class Sampler(C.Structure):
    _fields_ = [("filter", C.c_int32), ("addr_u", C.c_int32), ("addr_v", C.c_int32), ("addr_w", C.c_int32),
                ("border", C.c_int32), ("aniso_enable", C.c_int32), ("max_aniso", C.c_float)]


class Tri(C.Structure):
    _fields_ = [("prim", C.c_int32), ("fan", C.c_int32), ("clipped", C.c_int32), ("slot", C.c_int32),
                ("X", C.c_int32 * 3), ("Y", C.c_int32 * 3), ("Z", C.c_float * 3), ("IW", C.c_float * 3),
                ("bary", (C.c_float * 2) * 3)]


TRI_DTYPE = np.dtype([("prim", "i4"), ("fan", "i4"), ("clipped", "i4"), ("slot", "i4"), ("X", "i4", 3),
                      ("Y", "i4", 3), ("Z", "f4", 3), ("IW", "f4", 3), ("bary", "f4", (3, 2))])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libifrit_oracle.so")
        src = os.path.join(_HERE, "ifrit_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "libifrit_oracle.so"], stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(path)
        _LIB.orc_texture_total_texels.restype = C.c_int64
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make_state(width, height, msaa=1, cull=CULL_NONE, depth_func=CMP_LESS, convention=CONV_CUDA, clear=True,
               clear_color=(0, 0, 0, 0), clear_depth=1.0, scissors=None, blend=None):
    st = State()
    st.width, st.height, st.msaa, st.cull_mode = width, height, msaa, cull
    st.depth_func, st.convention, st.do_clear = depth_func, convention, int(clear)
    st.clear_color[:] = clear_color
    st.clear_depth = clear_depth
    if blend:     # (srcColor, dstColor, srcAlpha, dstAlpha) IfritBlendFactor values
        st.blend_enable = 1
        for i, f in enumerate(blend):
            st.blend_coef[i][:] = BLEND_COEF[f]
    else:         # blending off: src x 1 + dst x 0 (updateBlendStateKernel CU:3438-3443) — what the ordered path of a
        for i, c in enumerate(((1, 0, 0, 0), (0, 0, 0, 1), (1, 0, 0, 0), (0, 0, 0, 1))):   # depth-writing shader blends with
            st.blend_coef[i][:] = c
    if scissors:
        st.scissor_enable, st.nscissor = 1, len(scissors)
        for i, s in enumerate(scissors):
            st.scissor[i][:] = s
    return st


def make_texture(data, width, height, mip_levels=0, layers=1):
    t = Texture()
    t.data, t.width, t.height, t.mip_levels, t.layers = data.ctypes.data, width, height, mip_levels, layers
    return t


def make_sampler(filter=1, addr_u=0, addr_v=0, addr_w=0, border=0, aniso=False, max_aniso=1.0):
    s = Sampler()
    s.filter, s.addr_u, s.addr_v, s.addr_w, s.border = filter, addr_u, addr_v, addr_w, border
    s.aniso_enable, s.max_aniso = int(aniso), max_aniso
    return s


def attr_offsets(nattr):
    return np.arange(nattr, dtype=np.int32) * 16


def vertex_stage(vs_id, uniforms, vb, nvary):
    vb = np.ascontiguousarray(vb, dtype=np.float32)
    nverts, nattr = vb.shape[0], vb.shape[1] // 4
    pos = np.empty((nverts, 4), np.float32)
    vary = np.empty((nverts, max(nvary, 1), 4), np.float32)
    offs = attr_offsets(nattr)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float32)
    rc = lib().orc_vertex_stage(vs_id, _p(uniforms), _p(vb), vb.shape[1] * 4, _p(offs), nattr, nvary, nverts,
                                _p(pos), _p(vary))
    assert rc == 0, rc
    return pos, vary[:, :nvary].copy()


def assemble(st, pos, idx):
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    ntri = len(idx) // 3
    out = np.zeros(2 * max(ntri, 1), dtype=TRI_DTYPE)
    n = lib().orc_assemble(C.byref(st), _p(pos), _p(idx), ntri, _p(out))
    return out[:n]


def texture_chain(level0, mip_levels, layers=1, filter=1):
    """level0: [layers, H, W, 4] or [H, W, 4] float32 -> flat chain with generated mips."""
    level0 = np.ascontiguousarray(level0, dtype=np.float32)
    if level0.ndim == 3:
        level0 = level0[None]
    h, w = level0.shape[1], level0.shape[2]
    total = lib().orc_texture_total_texels(w, h, mip_levels, layers)
    data = np.zeros(total * 4, dtype=np.float32)
    data[: layers * h * w * 4] = level0.reshape(-1)
    lib().orc_generate_mipmap(_p(data), w, h, mip_levels, layers, filter)
    return data


def render(st, pos, vary, idx, fs_id, uniforms, textures=None, samplers=None, color=None, depth=None, polygon_mode=0):
    """Returns dict(color [H,W,4], depth [H,W], prim_id [H,W,S], stats).  polygon_mode: 0 fill, 1 line, 2 point."""
    W, H, S = st.width, st.height, st.msaa
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    nvary = vary.shape[1] if vary is not None and vary.size else 0
    vary = np.ascontiguousarray(vary, dtype=np.float32) if nvary else np.zeros((1, 1, 4), np.float32)
    color = np.zeros((H, W, 4), np.float32) if color is None else np.ascontiguousarray(color, np.float32).copy()
    depth = np.zeros((H, W), np.float32) if depth is None else np.ascontiguousarray(depth, np.float32).copy()
    prim = np.full((H, W, S), -1, np.int32)
    stats = np.zeros(4, np.int64)
    texs = (Texture * 8)()
    samps = (Sampler * 8)()
    for i, t in enumerate(textures or []):
        texs[i] = t
    for i, s in enumerate(samplers or []):
        samps[i] = s
    uniforms = np.ascontiguousarray(uniforms if uniforms is not None else np.zeros(32), dtype=np.float32)
    if polygon_mode:
        rc = lib().orc_render_prims(C.byref(st), int(polygon_mode), _p(pos), _p(vary), nvary, _p(idx), len(idx), fs_id,
                                    _p(uniforms), texs, samps, _p(color), _p(depth), _p(prim), _p(stats))
    else:
        rc = lib().orc_render(C.byref(st), _p(pos), _p(vary), nvary, _p(idx), len(idx) // 3, fs_id, _p(uniforms),
                              texs, samps, _p(color), _p(depth), _p(prim), _p(stats))
    assert rc == 0, rc
    return dict(color=color, depth=depth, prim_id=prim, frags=int(stats[0]), covered=int(stats[1]),
                assembled=int(stats[2]))


def draw(st, vb, idx, vs_id, fs_id, nvary, uniforms, **kw):
    pos, vary = vertex_stage(vs_id, uniforms, vb, nvary)
    return render(st, pos, vary, idx, fs_id, uniforms, **kw)


# ---- sampler batch entry points (oracle restatement and, when built, the reference's own source) --------------
_REFSB = None


def ref_sampler_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libifrit_ref_sampler.so"))


def sample_batch(kind, chain, w, h, mips, layers, sampler, uv, args, reference=False):
    """kind / args layout: see orc_sample_batch (ifrit_oracle.c) and oracle/ref_sampler_shim.cu.
    sampler: dict(filter, addr_u, addr_v, border, aniso, max_aniso).  reference=True runs the reference's
    ShaderBuiltinCuda.cuh (host instantiation, oracle/_ref/libifrit_ref_sampler.so) instead of the restatement."""
    global _REFSB
    uv = np.ascontiguousarray(uv, np.float32)
    args = np.ascontiguousarray(args, np.float32)
    n = uv.shape[0]
    assert args.shape == (n, 6) and uv.shape == (n, 2)
    out = np.zeros((n, 4), np.float32)
    chain = np.ascontiguousarray(chain, np.float32)
    if reference:
        if _REFSB is None:
            _REFSB = C.CDLL(os.path.join(_HERE, "_ref", "libifrit_ref_sampler.so"))
        s4 = np.array([sampler.get("filter", 1), sampler.get("addr_u", 0), sampler.get("addr_v", 0), sampler.get("border", 0)], np.int32)
        if kind == 3:
            args = args.copy(); args[:, 5] = sampler.get("max_aniso", 1.0)
        rc = _REFSB.refsb_sample(kind, _p(s4), _p(chain), w, h, layers, n, _p(uv), _p(args), _p(out))
    else:
        t = make_texture(chain, w, h, mips, layers)
        s = make_sampler(filter=sampler.get("filter", 1), addr_u=sampler.get("addr_u", 0), addr_v=sampler.get("addr_v", 0),
                         border=sampler.get("border", 0), aniso=sampler.get("aniso", False), max_aniso=sampler.get("max_aniso", 1.0))
        rc = lib().orc_sample_batch(kind, C.byref(t), C.byref(s), n, _p(uv), _p(args), _p(out))
    assert rc == 0, rc
    return out
