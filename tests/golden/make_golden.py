"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE CPU renderer here
(oracle/_ref/libifrit_ref.so built from /root/reference by oracle/Makefile).
The fixtures pin the oracle on boxes where /root/reference does not exist.

    python tests/golden/make_golden.py

Stored per case: the scene inputs' seed/parameters, the reference's colour (float16-rounded ids are
exact) and depth images.  Reference operating rules: forceDeterministic=0, depth LESS, BACK culling,
depth clear 0x7f7f7f7f (WK:40)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ifrit_v2_b200  # noqa: E402,F401
from ifrit_v2_b200 import scenes  # noqa: E402
from oracle import oracle as O, ref_driver as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # name: (W, H, generator, vs, fs, nvary)
    "kat1": (64, 64, lambda: scenes.kat1(), 0, O.FS_VARYING0, 1),
    "kat2": (64, 64, lambda: scenes.kat2(), 0, O.FS_VARYING0, 1),
    "rand_a": (256, 192, lambda: ids(scenes.random_tris(300, 1, size=0.25)), 0, O.FS_VARYING0, 1),
    "rand_clip": (256, 192, lambda: ids(scenes.random_tris(300, 2, neg_w_frac=0.15, size=0.3)), 0, O.FS_VARYING0, 1),
    "rand_small": (256, 192, lambda: ids(scenes.random_tris(3000, 3, size=0.03)), 0, O.FS_VARYING0, 1),
    "rand_big": (256, 192, lambda: ids(scenes.random_tris(60, 4, neg_w_frac=0.05, size=0.8)), 0, O.FS_VARYING0, 1),
    "mesh_normal": (320, 180, lambda: scenes.mesh_sphere(n=40), 1, O.FS_NORMAL_COLOR, 2),
}


# ordered alpha blending on the reference CPU renderer (iftrSetBlendFunc, depth test off): name -> (blend factors, seed,
# triangles, size); 160x96, alpha in [0.2, 0.8]; only the colour image is kept (ordered-mode depth is unreliable)
BLEND_CASES = {
    "blend_over": ((6, 7, 1, 0), 5, 40, 0.4),
    "blend_add": ((1, 1, 1, 1), 5, 40, 0.4),
    "blend_small": ((6, 7, 1, 0), 6, 300, 0.08),
}
BLEND_W, BLEND_H = 160, 96


def blend_scene(seed, n, size):
    vb, idx = scenes.random_tris(n, seed, size=size)
    vb[:, 7] = 0.2 + 0.6 * vb[:, 7]
    return vb, idx


def ids(scene):
    """flat per-triangle id in varying 0 (the reference exports no id buffer)."""
    vb, idx = scene
    vb[:, 4] = np.repeat(np.arange(len(idx) // 3) + 1, 3)
    vb[:, 5:8] = 0
    return vb, idx


def uniforms_for(name, W, H):
    return scenes.uniform_block(scenes.default_mvp(W, H) if name.startswith("mesh") else np.eye(4))


def main():
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    for name, (W, H, gen, vs, fs, nvary) in CASES.items():
        vb, idx = gen()
        u = uniforms_for(name, W, H)
        ref = R.RefRenderer(W, H, vb, idx, vs, fs, nvary, u)
        ref.draw()
        c1, d1 = ref.color(), ref.depth()
        ref.draw()
        # run-to-run: identical except (rarely) a pixel on an equal-depth shared edge (~5e-7 in colour)
        assert np.array_equal(d1, ref.depth()) and np.abs(c1 - ref.color()).max() < 1e-5, "reference not deterministic"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), color=c1, depth=d1, W=W, H=H)
        print(name, "covered", int((d1 < 1e38).sum()))
    for name, (blend, seed, n, size) in BLEND_CASES.items():
        vb, idx = blend_scene(seed, n, size)
        ref = R.RefRenderer(BLEND_W, BLEND_H, vb, idx, 0, O.FS_VARYING0, 1, scenes.uniform_block(np.eye(4)))
        ref.set_blend(*blend)
        ref.set_depth_test(False)
        ref.draw()
        c1 = ref.color()
        ref.draw()
        assert np.abs(c1 - ref.color()).max() < 1e-5, "reference blending not deterministic"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), color=c1, covered=ref.depth() < 1e38)
        print(name, "covered", int((ref.depth() < 1e38).sum()))


if __name__ == "__main__":
    main()
