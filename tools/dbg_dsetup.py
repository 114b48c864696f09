"""GPU debug: owner masks from iftrcuClassifyTriangles on one GPU (partition rank/2) — draw with and without the mask."""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ifrit_v2_b200 import renderer as R, scenes
W, H = 1000, 600
vb, idx = scenes.mesh_sphere(n=48)
u = scenes.uniform_block(scenes.default_mvp(W, H))
ntri = len(idx) // 3
dev = torch.device("cuda", 0)
for rank in (0, 1):
    r = R.TileRasterRendererCuda(0); r.Init()
    r.bindFrameBuffer(None, None, W, H)
    R.setup_scene(r, vb, idx, R.VS_MVP, R.FS_NORMAL_COLOR, 2, u)
    r.SetCullMode(R.IF_CULL_MODE_BACK); r.setClearValues([(0.1, 0.2, 0.3, 1.0)], 255.0)
    r.setTilePartition(rank, 2)
    r.clear(); r.drawElements(); r.synchronize()
    want = r.readColor().copy(); st0 = r.stats()
    mask = torch.zeros(ntri + 64, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    r.classifyTriangles(0, ntri, mask.data_ptr()); r.synchronize()
    m = mask.cpu().numpy()[:ntri]
    print("rank", rank, "mask hist", np.bincount(m, minlength=4)[:4], "assembled(no mask)", st0["triangles_assembled"], "pairs", st0["tile_pairs"], flush=True)
    r.setTriangleOwnerMask(mask.data_ptr())
    r.clear(); r.drawElements(); r.synchronize()
    got = r.readColor().copy(); st1 = r.stats()
    print("   with mask: identical", np.array_equal(got, want), "assembled", st1["triangles_assembled"], "pairs", st1["tile_pairs"], "diff px", int((got != want).any(-1).sum()), flush=True)
    r.close()
