"""Asset loaders: the reference's OBJ loader and RGBA image loader restated for the Python mirror.

load_obj / remap_normals / remap_uvs follow projects/softgraphics/src/utility/loader/WavefrontLoader.cpp:24-133 record for
record (see include/ifrit_b200/WavefrontLoader.hpp, the C++ twin); obj_to_buffers interleaves the way the reference's demos
do by hand (DefaultDemo.cpp:255-267).  load_rgba follows ImageLoader.cpp:25-47 (8-bit RGBA -> float / 255) with a
stdlib-only PNG decoder in place of the reference's third-party stb_image."""
import struct
import zlib
import numpy as np


def load_obj(path):
    """-> (vertices [nv,3], normals [nn,3], uvs [nt,2] float32, indices uint32 [3 per face corner: v, vt, vn; 0-based; an empty field = 0, a missing trailing field repeats the one before it — the reference's parser])."""
    v, vn, vt, idx = [], [], [], []

    def floats(tok, n):
        out = []
        for t in tok[:n]:
            try:
                out.append(float(t))
            except ValueError:
                break
        return out + [0.0] * (n - len(out))

    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                v.append(floats(tok[1:], 3))
            elif tok[0] == "vn":
                vn.append(floats(tok[1:], 3))
            elif tok[0] == "vt":
                vt.append(floats(tok[1:], 2))
            elif tok[0] == "f":
                corner = ""
                for k in range(3):                     # exactly the first three corners (WavefrontLoader.cpp:110)
                    if k + 1 < len(tok):
                        corner = tok[k + 1]            # a missing corner re-reads the previous one (`iss >> vertex` fails)
                    fields = corner.split("/") if corner else []
                    field = ""
                    for j in range(3):
                        # std::getline(vss, index, '/') on an exhausted stream leaves `index` as it was: "7" yields
                        # (7, 7, 7) and "7/3" yields (7, 3, 3); only an EMPTY field ("7//2") is stored as 0
                        if j < len(fields):
                            field = fields[j]
                        idx.append(int(field) - 1 if field else 0)
    return (np.asarray(v, np.float32).reshape(-1, 3), np.asarray(vn, np.float32).reshape(-1, 3),
            np.asarray(vt, np.float32).reshape(-1, 2), np.asarray(idx, np.int64).astype(np.uint32))


def remap_normals(normals, indices, num_vertices):
    out = np.zeros((num_vertices, 3), np.float32)
    idx = np.asarray(indices).reshape(-1, 3)
    for face_node, _, normal_node in idx:              # float32 accumulation in file order, as the reference
        out[face_node] = out[face_node] + normals[normal_node]
    length = np.sqrt((out[:, 0] * out[:, 0] + out[:, 1] * out[:, 1]) + out[:, 2] * out[:, 2]).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (out / length[:, None]).astype(np.float32)


def remap_uvs(uvs, indices, num_vertices):
    out = np.zeros((num_vertices, 2), np.float32)
    for face_node, uv_node, _ in np.asarray(indices).reshape(-1, 3):
        out[face_node] = uvs[uv_node]
    return out


def obj_to_buffers(path):
    """-> (vb [nv, 12] float32: position (w = 1) | normal (w = 0) | uv, idx int32 [3 * faces])."""
    pos, nrm, uv, idx = load_obj(path)
    n = len(pos)
    vb = np.zeros((n, 12), np.float32)
    vb[:, 0:3] = pos
    vb[:, 3] = 1.0
    if len(nrm):
        vb[:, 4:7] = remap_normals(nrm, idx, n)
    if len(uv):
        vb[:, 8:10] = remap_uvs(uv, idx, n)
    return vb, idx[0::3].astype(np.int32)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def load_rgba(path):
    """8-bit PNG (grey, grey+alpha, RGB or RGBA; non-interlaced) -> (float32 [h, w, 4] in [0, 1], h, w)."""
    data = open(path, "rb").read()
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG file")
    pos, idat, w = 8, b"", None
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if kind == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body)
            if depth != 8 or interlace or ctype not in (0, 2, 4, 6):
                raise ValueError("unsupported PNG (need 8-bit, non-interlaced, non-palette)")
        elif kind == b"IDAT":
            idat += body
        elif kind == b"IEND":
            break
    ch = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    raw = zlib.decompress(idat)
    stride = w * ch
    img = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    for y in range(h):
        ft = raw[y * (stride + 1)]
        line = np.frombuffer(raw, np.uint8, stride, y * (stride + 1) + 1).astype(np.int32)
        if ft == 2:
            line = (line + prev) & 255
        elif ft in (1, 3, 4):
            out = line.copy()
            for i in range(stride):
                a = out[i - ch] if i >= ch else 0
                b = prev[i]
                c = prev[i - ch] if i >= ch else 0
                out[i] = (line[i] + (a if ft == 1 else ((a + b) >> 1) if ft == 3 else _paeth(a, b, c))) & 255
            line = out
        img[y] = line
        prev = line
    px = img.reshape(h, w, ch)
    rgba = np.empty((h, w, 4), np.uint8)
    if ch == 1:
        rgba[..., :3] = px; rgba[..., 3] = 255
    elif ch == 2:
        rgba[..., :3] = px[..., :1]; rgba[..., 3] = px[..., 1]
    elif ch == 3:
        rgba[..., :3] = px; rgba[..., 3] = 255
    else:
        rgba[...] = px
    return rgba.astype(np.float32) / np.float32(255.0), h, w
