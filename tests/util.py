"""Shared helpers for the parity tests: run the same seeded draw through the oracle (CPU restatement)
and through the product's C-ABI (CUDA), and compare."""
import numpy as np
import ifrit_v2_b200  # noqa: F401
from ifrit_v2_b200 import scenes
from oracle import oracle as O


def oracle_draw(W, H, vb, idx, vs, fs, nvary, uniforms, msaa=1, cull=0, depth_func=1, convention=0, clear=True,
                clear_color=(0, 0, 0, 0), clear_depth=1.0, scissors=None, texture_chain=None, tex_dims=None,
                mip_levels=0, layers=1, sampler=None, color=None, depth=None, blend=None, polygon_mode=0):
    st = O.make_state(W, H, msaa=msaa, cull=cull, depth_func=depth_func, convention=convention, clear=clear,
                      clear_color=clear_color, clear_depth=clear_depth, scissors=scissors, blend=blend)
    texs = samps = None
    if texture_chain is not None:
        texs = [O.make_texture(texture_chain, tex_dims[0], tex_dims[1], mip_levels, layers)]
        s = dict(sampler or {})
        samps = [O.make_sampler(filter=s.get("filterMode", 1), addr_u=s.get("addressModeU", 0),
                                addr_v=s.get("addressModeV", 0), border=s.get("borderColor", 0),
                                aniso=s.get("anisotropyEnable", False), max_aniso=s.get("maxAnisotropy", 1.0))]
    return O.draw(st, vb, idx, vs, fs, nvary, uniforms, textures=texs, samplers=samps, color=color, depth=depth, polygon_mode=polygon_mode)


def gpu_draw(r, W, H, vb, idx, vs, fs, nvary, uniforms, msaa=1, cull=0, depth_func=1, convention=0, clear=True,
             clear_color=(0, 0, 0, 0), clear_depth=1.0, scissors=None, texture=None, mip_levels=0, layers=1,
             sampler=None, color=None, depth=None, depth_test=True, partition=None, fused_vs=True, hw_filter=False,
             blend=None, tma=True, polygon_mode=0):
    from ifrit_v2_b200 import renderer as R
    c0 = np.zeros((H, W, 4), np.float32) if color is None else np.ascontiguousarray(color, np.float32)
    d0 = np.zeros((H, W), np.float32) if depth is None else np.ascontiguousarray(depth, np.float32)
    r.bindFrameBuffer(c0, d0)
    R.setup_scene(r, vb, idx, vs, fs, nvary, uniforms, texture=texture, mip_levels=mip_levels, sampler=sampler,
                  layers=layers)
    r.setMsaaSamples(msaa)
    r.SetCullMode(cull)
    r.SetDepthFunc(depth_func)
    r.SetDepthTestEnable(depth_test)
    r.setSampleConvention(convention)
    r.setClearValues([clear_color], clear_depth)
    r.SetScissors(scissors if scissors else np.zeros((0, 4), np.float32))
    r.setScissorTestEnable(bool(scissors))
    r.setTilePartition(*(partition or (0, 1)))
    r.setDebugPrimitiveIds(True)
    r.setVertexStageFused(fused_vs)
    r.setSamplerHardwareFilter(hw_filter)
    r.setTmaFramebuffer(tma)
    r.setRasterizerPolygonMode(polygon_mode)
    if blend:
        r.setBlendFunc(True, *blend)
    else:
        r.setBlendFunc(False)
    if clear:
        r.clear()
    r.drawElements()
    r.synchronize()
    out = dict(color=r.readColor(), depth=r.readDepth(), prim_id=r.readPrimitiveIds())
    out.update(r.stats())
    return out


def compare(o, g, color_tol=1.0 / 255, color_frac=0.999, exact_color=False):
    """Coverage / primitive id / depth bit-exact; colour within tolerance on >= color_frac of pixels."""
    assert np.array_equal(o["prim_id"], g["prim_id"]), \
        f"primitive-id mismatch on {(o['prim_id'] != g['prim_id']).sum()} samples"
    assert np.array_equal(o["depth"].view(np.uint32), g["depth"].view(np.uint32)), \
        f"depth bits differ on {(o['depth'].view(np.uint32) != g['depth'].view(np.uint32)).sum()} pixels"
    diff = np.abs(o["color"] - g["color"]).max(axis=-1)
    if exact_color:
        assert np.array_equal(o["color"].view(np.uint32), g["color"].view(np.uint32)), f"colour bits differ, max {diff.max()}"
    frac = (diff <= color_tol).mean()
    assert frac >= color_frac, f"colour within {color_tol} on only {frac:.5f} of pixels (max diff {diff.max()})"
    return float(diff.max())
