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
