"""GPU debug: LOAD-op + scissor draw of test_tma_framebuffer_path_equals_per_thread_stores vs the oracle; dumps both."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import oracle_draw, gpu_draw
from oracle import oracle as O
from ifrit_v2_b200 import renderer as R, scenes
W, H = 400, 230
vb, idx = scenes.mesh_sphere(n=40); u = scenes.uniform_block(scenes.default_mvp(W, H))
rng = np.random.default_rng(5); start = rng.random((H, W, 4), dtype=np.float32)
sc = [(W * 0.25, H * 0.2, W * 0.4, H * 0.5)]
r = R.TileRasterRendererCuda(0); r.Init()
res = {}
for name, kw in (("cull1_scis", dict(cull=1, scissors=sc)), ("cull2_scis", dict(cull=2, scissors=sc)), ("cull1_noscis", dict(cull=1)), ("cull0_scis", dict(cull=0, scissors=sc))):
    kw = dict(kw, msaa=1, clear=False, color=start, depth=np.full((H, W), 0.9, np.float32))
    g = gpu_draw(r, W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, **kw)
    o = oracle_draw(W, H, vb, idx, 1, O.FS_NORMAL_COLOR, 2, u, **kw)
    bad = (g["prim_id"] != o["prim_id"])
    print(name, "id mismatch", int(bad.sum()), "gpu covered", int((g["prim_id"] >= 0).sum()), "oracle covered", int((o["prim_id"] >= 0).sum()),
          "depth mismatch", int((g["depth"].view(np.uint32) != o["depth"].view(np.uint32)).sum()), flush=True)
    if bad.any():
        ys, xs = np.nonzero(bad[..., 0])
        print("  bbox of mismatches x", xs.min(), xs.max(), "y", ys.min(), ys.max(), "sample", [(int(x), int(y), int(g["prim_id"][y, x, 0]), int(o["prim_id"][y, x, 0]), float(g["depth"][y, x]), float(o["depth"][y, x])) for x, y in list(zip(xs, ys))[:4]])
