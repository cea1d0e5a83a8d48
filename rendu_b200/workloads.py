"""Synthetic inputs of the benchmark configurations (BASELINE.json configs): host-side numpy only.

Nothing here traces a ray; these functions only build the meshes and ray arrays that are handed to the CUDA path,
to the oracle and to the reference, so that all three see bit-identical inputs.
"""
import numpy as np

F = np.float32


def displaced_grid(n):
    """C2/C5 mesh: (n+1)^2 vertices, 2*n^2 triangles. Vertex (i,j): x=2i/n-1, z=2j/n-1,
    y = 0.1*sin(9x)*cos(7z) + 0.02*sin(53x+31z) in fp32; quad (a,b,c,d) -> triangles (a,c,b), (b,c,d)."""
    i = np.arange(n + 1, dtype=F)
    x = (F(2) * i / F(n) - F(1)).astype(F)
    X, Z = np.meshgrid(x, x, indexing="xy")
    Y = (F(0.1) * np.sin(F(9) * X) * np.cos(F(7) * Z) + F(0.02) * np.sin(F(53) * X + F(31) * Z)).astype(F)
    pos = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(F)
    jj, ii = np.meshgrid(np.arange(n, dtype=np.int64), np.arange(n, dtype=np.int64), indexing="ij")
    a = (jj * (n + 1) + ii).ravel()
    b, c = a + 1, a + n + 1
    d = c + 1
    idx = np.stack([a, c, b, b, c, d], -1).reshape(-1).astype(np.uint32)
    return pos, idx


def coherent_rays(res=2048, rows=None):
    """C2 coherent set: res x res primary rays from (0,1.5,2.2) towards ((x+.5)/res*2-1, 0, (y+.5)/res*2-1), row-major,
    directions left un-normalised (the API normalises, Intersection.cpp:3-5). `rows`=(y0,y1) selects a row band."""
    y0, y1 = (0, res) if rows is None else rows
    xs = ((np.arange(res, dtype=F) + F(0.5)) / F(res) * F(2) - F(1)).astype(F)
    ys = ((np.arange(y0, y1, dtype=F) + F(0.5)) / F(res) * F(2) - F(1)).astype(F)
    TX, TZ = np.meshgrid(xs, ys, indexing="xy")
    org = np.empty((TX.size, 3), F)
    org[:] = np.array([0.0, 1.5, 2.2], F)
    tgt = np.stack([TX.ravel(), np.zeros(TX.size, F), TZ.ravel()], -1).astype(F)
    return org, (tgt - org).astype(F)


def _hash_u32(seed, idx, dim):
    """Counter-based uniform bits: a splitmix64-style finaliser over (seed, index, dimension)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(idx, np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed) * np.uint64(0xBF58476D1CE4E5B9)
             + np.uint64(dim) * np.uint64(0x94D049BB133111EB) + np.uint64(0x2545F4914F6CDD1D))
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(32)).astype(np.uint32)


def uniform01(seed, idx, dim):
    """fp32 uniform in [0,1) with 24 random bits."""
    return ((_hash_u32(seed, idx, dim) >> np.uint32(8)).astype(F) * F(1.0 / 16777216.0)).astype(F)


def incoherent_rays(org, dirs, hit, dist, normals, seed=1234, index_base=0):
    """C2 incoherent set derived from the closest hits of the coherent set: for a hit, a cosine-hemisphere bounce
    around the (ray-facing) geometric normal from p + 1e-3*n; for a miss, a uniform-sphere ray from a point drawn in
    [-1,1] x [0.2,0.6] x [-1,1]. `normals` are per-ray unit geometric normals (any orientation)."""
    n = org.shape[0]
    idx = np.arange(index_base, index_base + n, dtype=np.uint64)
    d = dirs / np.sqrt((dirs * dirs).sum(-1, keepdims=True)).astype(F)
    p = (org + dist[:, None].astype(F) * d).astype(F)
    nn = normals.astype(F).copy()
    flip = (nn * d).sum(-1) > 0
    nn[flip] = -nn[flip]
    u1, u2 = uniform01(seed, idx, 0), uniform01(seed, idx, 1)
    r = np.sqrt(u1).astype(F)
    phi = (F(2 * np.pi) * u2).astype(F)
    lx, ly = (r * np.cos(phi)).astype(F), (r * np.sin(phi)).astype(F)
    lz = np.sqrt(np.maximum(F(0), F(1) - u1)).astype(F)
    # Orthonormal frame around nn.
    helper = np.where(np.abs(nn[:, :1]) > 0.9, np.array([[0, 1, 0]], F), np.array([[1, 0, 0]], F)).astype(F)
    t = np.cross(helper, nn).astype(F)
    t /= np.sqrt((t * t).sum(-1, keepdims=True)).astype(F)
    b = np.cross(nn, t).astype(F)
    bounce_dir = (lx[:, None] * t + ly[:, None] * b + lz[:, None] * nn).astype(F)
    bounce_org = (p + F(1e-3) * nn).astype(F)
    # Misses: uniform point in the slab above the terrain, uniform direction on the sphere.
    u3, u4, u5 = uniform01(seed, idx, 2), uniform01(seed, idx, 3), uniform01(seed, idx, 4)
    mo = np.stack([F(2) * u1 - F(1), F(0.2) + F(0.4) * u2, F(2) * u3 - F(1)], -1).astype(F)
    ct = (F(2) * u4 - F(1)).astype(F)
    st = np.sqrt(np.maximum(F(0), F(1) - ct * ct)).astype(F)
    ph = (F(2 * np.pi) * u5).astype(F)
    md = np.stack([st * np.cos(ph), st * np.sin(ph), ct], -1).astype(F)
    h = hit.astype(bool)[:, None]
    return np.where(h, bounce_org, mo).astype(F), np.where(h, bounce_dir, md).astype(F)


def triangle_normals(pos, idx, local_id):
    """Unit geometric normal of the triangle whose first index sits at `local_id` in the index buffer."""
    lid = np.asarray(local_id, np.int64)
    a, b, c = pos[idx[lid]], pos[idx[lid + 1]], pos[idx[lid + 2]]
    n = np.cross(b - a, c - a).astype(F)
    ln = np.sqrt((n * n).sum(-1, keepdims=True)).astype(F)
    return (n / np.maximum(ln, F(1e-30))).astype(F)


def random_rays(n, seed, lo=(-1.2, -0.3, -1.2), hi=(1.2, 1.6, 1.2)):
    """Uniform origins in a box and uniform directions on the sphere (edge-case and small parity sets)."""
    idx = np.arange(n, dtype=np.uint64)
    lo, hi = np.array(lo, F), np.array(hi, F)
    o = np.stack([uniform01(seed, idx, k) for k in range(3)], -1) * (hi - lo) + lo
    ct = F(2) * uniform01(seed, idx, 3) - F(1)
    st = np.sqrt(np.maximum(F(0), F(1) - ct * ct))
    ph = F(2 * np.pi) * uniform01(seed, idx, 4)
    d = np.stack([st * np.cos(ph), st * np.sin(ph), ct], -1)
    return o.astype(F), d.astype(F)


# ------------------------------------------------------------------------------------------------ scenes (host-side data only)
def camera_shifts(eye, center, up, fov, ratio):
    """Camera::pose (Camera.cpp:10-18) + Camera::pixelShifts (Camera.cpp:42-49) in fp32: what PathTracer::render reads
    from the Camera on the host (PathTracer.cpp:177,197)."""
    eye, center, up = [np.asarray(v, F) for v in (eye, center, up)]

    def nrm(v):
        return (v * (F(1) / np.sqrt(F(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])))).astype(F)

    def cross(x, y):
        return np.array([x[1] * y[2] - y[1] * x[2], x[2] * y[0] - y[2] * x[0], x[0] * y[1] - y[0] * x[1]], F)

    upn = nrm(up)
    view = nrm((center - eye).astype(F))
    right = cross(view, upn)
    up2 = cross(right, view)
    hs = F(np.tan(F(0.5) * F(fov)))
    ws = F(ratio) * hs
    d = (eye - center).astype(F)
    dist = F(np.sqrt(F(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])))
    corner = (center + dist * (-ws * right + hs * up2)).astype(F)
    dx = (F(2) * ws * dist * right).astype(F)
    dy = (F(-2) * hs * dist * up2).astype(F)
    return dict(eye=eye, corner=corner, dx=dx, dy=dy)


def _const_tex(rgba, res=8):
    t = np.empty((res, res, 4), F)
    t[:] = np.asarray(rgba, F)
    return t


def _quad(p0, p1, p2, p3, uv_scale=1.0):
    """Quad (two triangles) with a flat tangent frame: positions, normals, tangents, bitangents, texcoords, indices."""
    P = np.array([p0, p1, p2, p3], F)
    t = P[1] - P[0]
    b = P[3] - P[0]
    n = np.cross(t, b)
    t, b, n = [(v / np.linalg.norm(v)).astype(F) for v in (t, b, n)]
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], F) * F(uv_scale)
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    return dict(positions=P, normals=np.tile(n, (4, 1)), tangents=np.tile(t, (4, 1)), bitangents=np.tile(b, (4, 1)), texcoords=uv, indices=idx)


def _obj(geom, textures, model=None, use_uv=False, emissive=False, masked=False):
    o = dict(geom)
    o.update(model=np.eye(4, dtype=F).reshape(16) if model is None else np.asarray(model, F).reshape(16), use_uv=use_uv,
             material_type=8 if emissive else 1, masked=masked, textures=list(textures))
    return o


def tiny_scene():
    """Floor + back wall + a small emissive panel, one point light, constant textures: enough to exercise every kernel."""
    tex = [_const_tex((0.8, 0.8, 0.8, 1)), _const_tex((0.5, 0.5, 1.0, 1)), _const_tex((0.6, 0.0, 1.0, 1)), _const_tex((2.0, 1.5, 1.0, 1)),
           _const_tex((0.7, 0.2, 0.2, 1))]
    objs = [_obj(_quad((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1)), [0, 1, 2]),
            _obj(_quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1)), [4, 1, 2]),
            _obj(_quad((-0.3, 1.99, -0.3), (0.3, 1.99, -0.3), (0.3, 1.99, 0.3), (-0.3, 1.99, 0.3)), [3, 1, 2], emissive=True)]
    lights = [dict(kind=0, casts_shadow=True, intensity=np.array([4, 4, 3.5], F), position=np.array([0.2, 1.5, 0.3], F), direction=np.zeros(3, F), radius=4.0, inner=0.0, outer=0.0)]
    return dict(objects=objs, lights=lights, textures=tex, background=dict(mode=0, color=np.array([0.05, 0.06, 0.08], F), sun=np.array([0, 1, 0], F), textures=[], sky_lut=-1))


def tiny_camera(ratio=1.0):
    return camera_shifts((0, 1, 2.6), (0, 1, 0), (0, 1, 0), 0.9, ratio)


def _grid_object(n, y0=0.0, amp=1.0, uv_scale=4.0):
    """Displaced grid with analytic normals / tangent frame (the C2 mesh as a shaded floor)."""
    pos, idx = displaced_grid(n)
    pos = pos.copy()
    x, z = pos[:, 0].astype(np.float64), pos[:, 2].astype(np.float64)
    pos[:, 1] = (F(amp) * pos[:, 1] + F(y0)).astype(F)
    dydx = amp * (0.9 * np.cos(9 * x) * np.cos(7 * z) + 0.02 * 53 * np.cos(53 * x + 31 * z))
    dydz = amp * (-0.7 * np.sin(9 * x) * np.sin(7 * z) + 0.02 * 31 * np.cos(53 * x + 31 * z))
    t = np.stack([np.ones_like(x), dydx, np.zeros_like(x)], -1)
    b = np.stack([np.zeros_like(x), dydz, np.ones_like(x)], -1)
    nrm = np.cross(b, t)
    t, b, nrm = [(v / np.linalg.norm(v, axis=1, keepdims=True)).astype(F) for v in (t, b, nrm)]
    uv = (np.stack([x, z], -1) * 0.5 + 0.5).astype(F) * F(uv_scale)
    return dict(positions=pos, normals=nrm, tangents=t, bitangents=b, texcoords=uv, indices=idx)


def _procedural_textures(res):
    """Albedo checker with an alpha channel of 1, a bumpy tangent-space normal map, and a roughness/metalness/ao map."""
    u = (np.arange(res, dtype=F) + F(0.5)) / F(res)
    U, V = np.meshgrid(u, u, indexing="xy")
    checker = ((np.floor(U * 8) + np.floor(V * 8)) % 2).astype(F)
    albedo = np.stack([F(0.35) + F(0.5) * checker, F(0.4) + F(0.35) * checker, F(0.45) + F(0.2) * (1 - checker), np.ones_like(U)], -1).astype(F)
    nx, ny = F(0.25) * np.sin(F(2 * np.pi * 6) * U), F(0.25) * np.cos(F(2 * np.pi * 5) * V)
    nz = np.sqrt(np.maximum(F(0), F(1) - nx * nx - ny * ny))
    normal = np.stack([nx * F(0.5) + F(0.5), ny * F(0.5) + F(0.5), nz * F(0.5) + F(0.5), np.ones_like(U)], -1).astype(F)
    rmao = np.stack([F(0.25) + F(0.5) * checker, F(0.8) * (1 - checker) * (U > 0.5), np.ones_like(U), np.ones_like(U)], -1).astype(F)
    return albedo, normal, rmao


def _procedural_cubemap(res):
    """Six fp32 RGBA faces (px,nx,py,ny,pz,nz) of a sky gradient with a warm sun lobe."""
    u = (np.arange(res, dtype=F) + F(0.5)) / F(res) * F(2) - F(1)
    A, B = np.meshgrid(u, u, indexing="xy")
    one = np.ones_like(A)
    dirs = [np.stack([one, -B, -A], -1), np.stack([-one, -B, A], -1), np.stack([A, one, B], -1), np.stack([A, -one, -B], -1),
            np.stack([A, -B, one], -1), np.stack([-A, -B, -one], -1)]
    sun = np.array([0.4, 0.7, 0.3], F)
    sun /= np.linalg.norm(sun)
    faces = []
    for d in dirs:
        d = d / np.linalg.norm(d, axis=-1, keepdims=True)
        up = np.clip(d[..., 1] * 0.5 + 0.5, 0, 1)
        sky = np.stack([0.25 + 0.35 * up, 0.35 + 0.45 * up, 0.55 + 0.6 * up], -1)
        lobe = np.clip((d * sun).sum(-1), 0, 1) ** 64
        col = sky + lobe[..., None] * np.array([6.0, 5.0, 3.5])
        faces.append(np.concatenate([col, np.ones_like(col[..., :1])], -1).astype(F))
    return faces


def cornell_synthetic(grid_n=708, tex_res=512, cube_res=256):
    """C3/C4 scene built from nothing but numpy: textured GGX+Lambert Cornell-style room whose floor is
    the displaced grid (2*grid_n^2 triangles), open front, cubemap environment, one emissive panel, one shadow-casting
    point light. Object order is the scene->objects order (regular materials first, emissive last, Scene.cpp:101-104)."""
    albedo, normal, rmao = _procedural_textures(tex_res)
    faces = _procedural_cubemap(cube_res)
    tex = [albedo, normal, rmao, _const_tex((0.75, 0.2, 0.2, 1)), _const_tex((0.2, 0.7, 0.25, 1)), _const_tex((0.5, 0.5, 1.0, 1)),
           _const_tex((0.6, 0.0, 1.0, 1)), _const_tex((0.15, 1.0, 1.0, 1)), _const_tex((8.0, 7.0, 5.5, 1))] + faces
    cube0 = 9
    floor = _grid_object(grid_n, y0=0.0, amp=0.5)
    objs = [_obj(floor, [0, 1, 2], use_uv=True),
            _obj(_quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1), 2.0), [0, 1, 2], use_uv=True),          # back wall, textured
            _obj(_quad((-1, 0, 1), (-1, 0, -1), (-1, 2, -1), (-1, 2, 1)), [3, 5, 6]),                             # left, red
            _obj(_quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1)), [4, 5, 6]),                                 # right, green
            _obj(_quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1)), [0, 1, 2], use_uv=True),                  # ceiling
            _obj(_quad((-0.45, 0.35, 0.1), (0.25, 0.35, -0.15), (0.25, 1.05, -0.15), (-0.45, 1.05, 0.1)), [4, 5, 7]),  # metal panel
            _obj(_quad((-0.35, 1.98, -0.35), (0.35, 1.98, -0.35), (0.35, 1.98, 0.35), (-0.35, 1.98, 0.35)), [8, 5, 6], emissive=True)]
    lights = [dict(kind=0, casts_shadow=True, intensity=np.array([5, 4.6, 4.0], F), position=np.array([0.3, 1.6, 0.4], F), direction=np.zeros(3, F), radius=4.0, inner=0.0, outer=0.0)]
    bg = dict(mode=2, color=np.zeros(3, F), sun=np.array([0, 1, 0], F), textures=list(range(cube0, cube0 + 6)), sky_lut=-1)
    return dict(objects=objs, lights=lights, textures=tex, background=bg)


CORNELL_CAMERA = dict(eye=(0.0, 1.0, 3.4), center=(0.0, 0.9, 0.0), up=(0.0, 1.0, 0.0), fov=0.75)


def cornell_camera(ratio):
    c = CORNELL_CAMERA
    return camera_shifts(c["eye"], c["center"], c["up"], c["fov"], ratio)
