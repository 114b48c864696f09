#!/usr/bin/env python
"""Where does the tile kernel's time go?  mesh50k at 4K with different fragment shaders / sample counts / cull modes,
frame pipelining off; prints the tile kernel's CUDA-event time per combination (run on the GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ifrit_v2_b200  # noqa
from ifrit_v2_b200 import renderer as R, scenes

W, H = 3840, 2160
vb, idx = scenes.mesh_sphere(n=160, seed=1234)
u = scenes.uniform_block(scenes.default_mvp(W, H))
tex = scenes.texture_checker(size=2048, seed=7)
r = R.TileRasterRendererCuda(0)
r.Init()
r.bindFrameBuffer(None, None, W, H)
R.setup_scene(r, vb, idx, R.VS_MVP, R.FS_TEXTURED, 2, u, texture=tex, mip_levels=11,
              sampler=dict(filterMode=1, anisotropyEnable=True, maxAnisotropy=16.0))
r.setClearValues([(0, 0, 0, 0)], 255.0)
r.setFramePipelining(False)
for msaa, cull in ((1, 0), (8, 2), (1, 2)):
    for fs in ("prim_id", "normal_color", "textured_lod0", "textured"):
        r.bindFragmentShader(fs)
        r.setMsaaSamples(msaa); r.SetCullMode(cull)
        for _ in range(5):
            r.clear(); r.drawElements()
        r.synchronize()
        r.setProfiling(True)
        for _ in range(50):
            r.clear(); r.drawElements()
        k, n = r.kernelTimes()
        r.setProfiling(False)
        st = r.stats()
        print(f"msaa={msaa} cull={cull} fs={fs:14s} tile {k['tile']*1e3:7.1f} us  fill {k['fill']*1e3:5.1f} setup {k['setup_count']*1e3:5.1f}  frags {st['fragments_shaded']}", flush=True)
