"""Asset loaders: ifrit_v2_b200.loaders and the C++ twin include/ifrit_b200/WavefrontLoader.hpp against
the reference's own WavefrontLoader.cpp (oracle/_ref/libifrit_ref_loader.so, built where /root/reference exists) and
against a committed golden fixture produced by it (tests/golden/loader_obj.npz); the PNG decoder against a PNG written
here byte by byte."""
import ctypes as C
import os
import struct
import subprocess
import zlib
import numpy as np
import pytest
import ifrit_v2_b200  # noqa: F401
from ifrit_v2_b200 import loaders, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libifrit_ref_loader.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "loader_obj.npz")


def write_obj(path, seed=5):
    """A small OBJ with everything the loader distinguishes: v/vt/vn corners, v//vn, v/vt, bare v, comments, a quad
    (only its first three corners count), blank lines."""
    vb, idx = scenes.mesh_sphere(n=6)
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        f.write("# test asset\n\no sphere\n")
        for p in vb[:, 0:3]:
            f.write("v %.6f %.6f %.6f\n" % tuple(p))
        for n in vb[:, 4:7] + rng.normal(0, 0.05, (len(vb), 3)):
            f.write("vn %.6f %.6f %.6f\n" % tuple(n))
        for t in vb[:, 8:10]:
            f.write("vt %.6f %.6f\n" % tuple(t))
        tri = idx.reshape(-1, 3) + 1
        for k, (a, b, c) in enumerate(tri):
            n = (int(a) * 7 + k) % len(vb) + 1
            if k % 4 == 0:
                f.write(f"f {a}/{a}/{n} {b}/{b}/{b} {c}/{c}/{c}\n")
            elif k % 4 == 1:
                f.write(f"f {a}//{a} {b}//{n} {c}//{c}\n")
            elif k % 4 == 2:
                f.write(f"f {a}/{b} {b}/{c} {c}/{a}\n")
            else:
                f.write(f"f {a} {b} {c} {a}\n")
    return path


def _ours(path):
    v, n, t, i = loaders.load_obj(path)
    return dict(v=v, n=n, t=t, i=i, rn=loaders.remap_normals(n, i, len(v)), ru=loaders.remap_uvs(t, i, len(v)))


def _reference(path):
    L = C.CDLL(REF)
    counts = (C.c_int * 4)()
    L.refld_load(path.encode(), counts)
    nv, nn, nt, ni = counts
    v, n, t = np.zeros((nv, 3), np.float32), np.zeros((nn, 3), np.float32), np.zeros((nt, 2), np.float32)
    i = np.zeros(ni, np.uint32)
    rn, ru = np.zeros((nv, 3), np.float32), np.zeros((nv, 2), np.float32)
    vp = C.c_void_p
    L.refld_get.argtypes = [vp] * 6
    L.refld_get(v.ctypes.data, n.ctypes.data, t.ctypes.data, i.ctypes.data, rn.ctypes.data, ru.ctypes.data)
    return dict(v=v, n=n, t=t, i=i, rn=rn, ru=ru)


def _same(a, b):
    for k in ("v", "n", "t", "i", "ru"):
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
    # normalisation: the reference divides by its own vector length (rsqrt-free, float); allow the last ulp
    assert np.allclose(a["rn"], b["rn"], rtol=0, atol=2e-7, equal_nan=True)


def test_obj_loader_matches_the_reference_loader(tmp_path):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libifrit_ref_loader.so not built (no /root/reference here)")
    path = write_obj(str(tmp_path / "a.obj"))
    _same(_ours(path), _reference(path))


def test_obj_loader_matches_golden_fixture(tmp_path):
    """tests/golden/loader_obj.npz = the reference loader's output for write_obj(seed=5) (tests/golden/make_golden_loader.py)."""
    g = np.load(GOLDEN)
    path = write_obj(str(tmp_path / "a.obj"))
    _same(_ours(path), {k: g[k] for k in ("v", "n", "t", "i", "rn", "ru")})
    vb, idx = loaders.obj_to_buffers(path)
    assert vb.shape == (len(g["v"]), 12) and idx.dtype == np.int32 and np.array_equal(idx, g["i"][0::3].astype(np.int32))
    assert np.array_equal(vb[:, 3], np.ones(len(vb), np.float32)) and np.array_equal(vb[:, 8:10], g["ru"])


def test_cpp_loader_header_matches_python(tmp_path):
    path = write_obj(str(tmp_path / "a.obj"))
    src = tmp_path / "t.cpp"
    src.write_text('''#include "ifrit_b200/WavefrontLoader.hpp"
#include <cstdio>
int main(int argc, char** argv) {
    std::vector<float> vb; std::vector<int> ib;
    if (!ifrit_b200::Utility::Loader::WavefrontLoader::loadInterleaved(argv[1], vb, ib)) return 1;
    FILE* f = fopen(argv[2], "wb");
    int n[2] = {(int)vb.size(), (int)ib.size()};
    fwrite(n, 4, 2, f); fwrite(vb.data(), 4, vb.size(), f); fwrite(ib.data(), 4, ib.size(), f); fclose(f);
    return 0;
}''')
    exe, out = str(tmp_path / "t"), str(tmp_path / "o.bin")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe])
    subprocess.check_call([exe, path, out])
    raw = open(out, "rb").read()
    nvb, nib = struct.unpack("ii", raw[:8])
    vb = np.frombuffer(raw, np.float32, nvb, 8).reshape(-1, 12)
    ib = np.frombuffer(raw, np.int32, nib, 8 + 4 * nvb)
    pvb, pib = loaders.obj_to_buffers(path)
    assert np.array_equal(ib, pib) and np.array_equal(vb[:, [0, 1, 2, 3, 7, 8, 9, 10, 11]], pvb[:, [0, 1, 2, 3, 7, 8, 9, 10, 11]])
    assert np.allclose(vb[:, 4:7], pvb[:, 4:7], rtol=0, atol=2e-7, equal_nan=True)


def _png(path, img, ctype):
    h, w = img.shape[:2]
    raw = b""
    prev = np.zeros(img.shape[1] * img.shape[2], np.int32)
    ch = img.shape[2]
    for y in range(h):
        line = img[y].reshape(-1).astype(np.int32)
        ft = y % 5
        if ft == 0:
            enc = line
        elif ft == 1:
            enc = line - np.concatenate([np.zeros(ch, np.int32), line[:-ch]])
        elif ft == 2:
            enc = line - prev
        elif ft == 3:
            enc = line - ((np.concatenate([np.zeros(ch, np.int32), line[:-ch]]) + prev) >> 1)
        else:
            a = np.concatenate([np.zeros(ch, np.int32), line[:-ch]]); c = np.concatenate([np.zeros(ch, np.int32), prev[:-ch]])
            enc = line - np.array([loaders._paeth(int(x), int(b), int(z)) for x, b, z in zip(a, prev, c)], np.int32)
        raw += bytes([ft]) + (enc & 255).astype(np.uint8).tobytes()
        prev = line

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body))
    z = zlib.compress(raw)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) +
                chunk(b"IDAT", z[:len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b""))


@pytest.mark.parametrize("ctype,ch", [(6, 4), (2, 3), (0, 1), (4, 2)])
def test_png_rgba_loader(tmp_path, ctype, ch):
    """ImageLoader::loadRGBA (ImageLoader.cpp:25-47): 8-bit pixels expanded to RGBA, value / 255.0f; all five PNG filters."""
    rng = np.random.default_rng(ctype)
    img = rng.integers(0, 256, (11, 7, ch), dtype=np.uint8)
    p = str(tmp_path / "t.png")
    _png(p, img, ctype)
    out, h, w = loaders.load_rgba(p)
    assert (h, w) == (11, 7) and out.dtype == np.float32
    want = np.empty((11, 7, 4), np.uint8)
    if ch == 4:
        want[...] = img
    elif ch == 3:
        want[..., :3] = img; want[..., 3] = 255
    elif ch == 1:
        want[..., :3] = img; want[..., 3] = 255
    else:
        want[..., :3] = img[..., :1]; want[..., 3] = img[..., 1]
    assert np.array_equal(out, want.astype(np.float32) / np.float32(255.0))
