"""Seeded synthetic scenes of SURVEY.md §8(d): S1 "mesh50k", S2 "soup10M", S3 "scene50M",
and the procedural RGBA32F texture.  Pure numpy, float32 throughout; used by tests and bench.

The camera matrices restate the reference's helpers
(include/ifrit/softgraphics/engine/math/ShaderOpsCuda.cuh:98-182: row-major m[i][j],
``multiply(mat, vec)`` = rows dot vec, left-handed, z in [0,1])."""
import math
import numpy as np

F = np.float32


def perspective(fovy, aspect, z_near, z_far):
    m = np.zeros((4, 4), dtype=F)
    top = z_near * math.tan(fovy / 2.0)
    right = top * aspect
    m[0, 0] = 2 * z_near / (2 * right)
    m[1, 1] = 2 * z_near / (2 * top)
    m[2, 2] = z_far / (z_far - z_near)
    m[3, 2] = 1.0
    m[2, 3] = -(z_far * z_near) / (z_far - z_near)
    return m


def look_at(eye, center, up):
    eye, center, up = (np.asarray(v, dtype=np.float64) for v in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    rot = np.eye(4)
    rot[0, :3], rot[1, :3], rot[2, :3] = s, u, f
    tr = np.eye(4)
    tr[:3, 3] = -eye
    return (rot @ tr).astype(F)


def default_mvp(width, height, eye=(0.013, 0.021, 1.8)):
    p = perspective(60.0 * math.pi / 180.0, width / height, 0.1, 100.0)
    v = look_at(eye, (0, 0, 0), (0, 1, 0))
    return (p.astype(np.float64) @ v.astype(np.float64)).astype(F)


def uniform_block(mvp, light=(0.3, 0.5, -0.8), lod_bias=0.0, deriv_scale=1.0):
    """Uniform buffer slot 0 layout shared by the shader registry:
    floats [0..15] MVP (row-major), [16..19] light direction, [20] lod bias / cube lod,
    [21] derivative visualisation scale; padded to 32 floats."""
    u = np.zeros(32, dtype=F)
    u[:16] = np.asarray(mvp, dtype=F).reshape(-1)
    l = np.asarray(light, dtype=np.float64)
    l = l / np.linalg.norm(l)
    u[16:19] = l.astype(F)
    u[20] = lod_bias
    u[21] = deriv_scale
    return u


def _hash01(ix, iy, seed):
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + iy.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
         + np.uint64((int(seed) * 0x165667B19E3779F9) & 0xFFFFFFFFFFFFFFFF))
    h ^= h >> np.uint64(29)
    h *= np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(32)
    return (h & np.uint64(0xFFFFFF)).astype(np.float64) / float(0x1000000)


def value_noise(x, y, seed, period=None):
    """Bilinear value noise on the integer lattice; x, y float arrays."""
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = x - x0, y - y0
    ix, iy = x0.astype(np.int64), y0.astype(np.int64)

    def h(ax, ay):
        if period is not None:
            ax, ay = ax % period, ay % period
        return _hash01(ax & 0xFFFFFFFF, ay & 0xFFFFFFFF, seed)
    sx, sy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
    a = h(ix, iy) * (1 - sx) + h(ix + 1, iy) * sx
    b = h(ix, iy + 1) * (1 - sx) + h(ix + 1, iy + 1) * sx
    return a * (1 - sy) + b * sy


def fbm(x, y, seed, octaves=4):
    amp, total, norm = 1.0, 0.0, 0.0
    for o in range(octaves):
        total = total + amp * value_noise(x * (2 ** o), y * (2 ** o), seed + o)
        norm += amp
        amp *= 0.5
    return total / norm


def mesh_sphere(n=160, seed=1234):
    """S1 "mesh50k": n x n quad grid on a displaced sphere r = 1 + 0.15 fbm(theta, phi).
    Returns (vertex_buffer [V,12] float32 = pos4|normal4|uv4, indices int32 [T*3])."""
    i = np.arange(n + 1)
    u, v = np.meshgrid(i / n, i / n, indexing="xy")
    theta = u * 2 * np.pi + 0.1   # keep the seam off the exact screen centre line (exact edge ties)
    phi = (v * 0.98 + 0.01) * np.pi
    r = 1.0 + 0.15 * (fbm(u * 6, v * 6, seed) - 0.5) * 2
    x = r * np.sin(phi) * np.cos(theta)
    y = r * np.cos(phi)
    z = r * np.sin(phi) * np.sin(theta)
    p = np.stack([x, y, z], -1)
    nrm = p / np.linalg.norm(p, axis=-1, keepdims=True)
    vb = np.zeros(((n + 1) * (n + 1), 12), dtype=F)
    vb[:, 0:3] = p.reshape(-1, 3)
    vb[:, 3] = 1.0
    vb[:, 4:7] = nrm.reshape(-1, 3)
    vb[:, 8] = (u * 4).reshape(-1)
    vb[:, 9] = (v * 2).reshape(-1)
    a = (np.arange(n)[:, None] * (n + 1) + np.arange(n)[None, :]).reshape(-1)
    b, c, d = a + 1, a + n + 1, a + n + 2
    idx = np.stack([a, b, c, b, d, c], -1).reshape(-1).astype(np.int32)
    return vb, idx


def texture_checker(size=2048, seed=7, checker=8):
    """Procedural RGBA32F level-0 texture: `checker`-px checkerboard x value noise."""
    yy, xx = np.mgrid[0:size, 0:size]
    chk = (((xx // checker) + (yy // checker)) & 1).astype(np.float64)
    n = value_noise(xx / 32.0, yy / 32.0, seed, period=size // 32)
    n2 = value_noise(xx / 5.0, yy / 5.0, seed + 11)
    img = np.empty((size, size, 4), dtype=F)
    img[..., 0] = (0.25 + 0.75 * chk) * (0.5 + 0.5 * n)
    img[..., 1] = (0.9 - 0.6 * chk) * (0.4 + 0.6 * n2)
    img[..., 2] = 0.2 + 0.8 * n * n2
    img[..., 3] = 1.0
    return img


def _soup(ntri, width, height, rng, radius_px, zrange=(0.1, 0.9)):
    cx = rng.random(ntri) * 2 - 1
    cy = rng.random(ntri) * 2 - 1
    z = zrange[0] + rng.random(ntri) * (zrange[1] - zrange[0])
    rot = rng.random(ntri) * 2 * np.pi
    flip = rng.random(ntri) < 0.5
    ang = rot[:, None] + np.array([0.0, 2 * np.pi / 3, 4 * np.pi / 3])[None, :] \
        + (rng.random((ntri, 3)) - 0.5) * 0.6
    ang = np.where(flip[:, None], ang[:, ::-1], ang)
    px = cx[:, None] + radius_px[:, None] * np.cos(ang) * (2.0 / width)
    py = cy[:, None] + radius_px[:, None] * np.sin(ang) * (2.0 / height)
    vb = np.zeros((ntri * 3, 8), dtype=F)
    vb[:, 0] = px.reshape(-1)
    vb[:, 1] = py.reshape(-1)
    vb[:, 2] = np.repeat(z, 3)
    vb[:, 3] = 1.0
    col = rng.random((ntri, 3))
    vb[:, 4:7] = np.repeat(col, 3, axis=0)
    vb[:, 7] = 1.0
    return vb


def soup(ntri=10_000_000, width=3840, height=2160, seed=42, chunk=2_000_000):
    """S2 "soup10M": independent small triangles, pre-transformed clip space (w = 1),
    circumradius 0.5-1.5 px, random rotation and winding; attrs float4 pos | float4 colour.
    Returns (vertex_buffer [3T,8], indices int32 [3T])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parts = []
    for s in range(0, ntri, chunk):
        n = min(chunk, ntri - s)
        parts.append(_soup(n, width, height, rng, 0.5 + rng.random(n)))
    vb = np.concatenate(parts) if len(parts) > 1 else parts[0]
    return vb, np.arange(ntri * 3, dtype=np.int32)


def scene_mix(ntri=50_000_000, width=7680, height=4320, seed=43, chunk=2_000_000):
    """S3 "scene50M": size mixture 90 % radius 0.5-2 px, 9.9 % 8-32 px, 0.1 % 128-512 px."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parts = []
    for s in range(0, ntri, chunk):
        n = min(chunk, ntri - s)
        cls = rng.random(n)
        r = np.where(cls < 0.9, 0.5 + 1.5 * rng.random(n),
                     np.where(cls < 0.999, 8 + 24 * rng.random(n), 128 + 384 * rng.random(n)))
        parts.append(_soup(n, width, height, rng, r))
    vb = np.concatenate(parts) if len(parts) > 1 else parts[0]
    return vb, np.arange(ntri * 3, dtype=np.int32)


def kat1():
    """64x64, two triangles forming the quad (8,8)-(40,40)."""
    def ndc(p):
        return [p[0] / 32.0 - 1, p[1] / 32.0 - 1, 0.5, 1.0]
    pts = [(8, 8), (40, 8), (8, 40), (40, 8), (40, 40), (8, 40)]
    vb = np.zeros((6, 8), dtype=F)
    for i, p in enumerate(pts):
        vb[i, :4] = ndc(p)
        vb[i, 4:8] = [1 + i // 3, 0, 0, 1]
    return vb, np.arange(6, dtype=np.int32)


def kat2():
    """64x64 single triangle (-.5,-.5),(.5,-.5),(0,.5), front-facing index order."""
    vb = np.zeros((3, 8), dtype=F)
    vb[0, :4] = [-0.5, -0.5, 0.5, 1]
    vb[1, :4] = [0.5, -0.5, 0.5, 1]
    vb[2, :4] = [0.0, 0.5, 0.5, 1]
    vb[0, 4:8] = [1, 0, 0, 1]
    vb[1, 4:8] = [0, 1, 0, 1]
    vb[2, 4:8] = [0, 0, 1, 1]
    return vb, np.array([0, 1, 2], dtype=np.int32)


def random_tris(ntri, seed, wrange=(0.5, 2.0), extent=1.3, neg_w_frac=0.0, size=0.25):
    """Random clip-space triangles for parity tests (optionally some vertices behind w=0)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = (rng.random((ntri, 1, 2)) * 2 - 1) * extent
    p = c + (rng.random((ntri, 3, 2)) * 2 - 1) * size
    z = rng.random((ntri, 3))
    w = wrange[0] + rng.random((ntri, 3)) * (wrange[1] - wrange[0])
    if neg_w_frac > 0:
        neg = rng.random((ntri, 3)) < neg_w_frac
        w = np.where(neg, -w * 0.5, w)
    vb = np.zeros((ntri * 3, 8), dtype=F)
    vb[:, 0] = (p[..., 0] * w).reshape(-1)
    vb[:, 1] = (p[..., 1] * w).reshape(-1)
    vb[:, 2] = (z * w).reshape(-1)
    vb[:, 3] = w.reshape(-1)
    vb[:, 4:8] = rng.random((ntri * 3, 4))
    return vb, np.arange(ntri * 3, dtype=np.int32)
