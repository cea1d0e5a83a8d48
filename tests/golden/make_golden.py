"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref/librendu_ref.so, i.e. the unmodified Rendu
translation units built by oracle/build_ref.sh). Run here, where /root/reference exists:

    python tests/golden/make_golden.py

The vectors pin the oracle (and through it the CUDA path) on machines that have neither /root/reference nor
oracle/_ref. Inputs are seeded; the file records them next to the reference's outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from rendu_b200 import workloads as wl  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
F = np.float32


def translate(t):
    m = np.eye(4, dtype=F)
    m[3, :3] = t  # column-major: translation lives in the 4th column = row 3 of the numpy view
    return m.reshape(16)


def rot_scale(angle, s, t):
    c, sn = np.cos(angle), np.sin(angle)
    m = np.array([[c * s, 0, -sn * s, 0], [0, s, 0, 0], [sn * s, 0, c * s, 0], [t[0], t[1], t[2], 1]], F)  # rows = glm columns
    return m.reshape(16)


def edge_rays():
    """Rays that exercise the NaN / inf corners of Intersection::box: axis-parallel directions, origins exactly on box
    planes, zero-length and denormal direction components, rays inside flat boxes."""
    o, d = [], []
    for ax in range(3):
        for sgn in (1.0, -1.0):
            for off in (0.0, 0.25, -1.0, 1.0, 0.5):
                dd = np.zeros(3, F)
                dd[ax] = sgn
                oo = np.array([off, 0.5, off], F)
                oo[ax] = -3.0 * sgn
                o.append(oo)
                d.append(dd)
    for k in range(16):
        dd = np.array([1e-41 if k & 1 else 0.3, -1.0, 1e-39 if k & 2 else -0.2], F)
        o.append(np.array([-1.0 + 0.125 * k, 2.0, 1.0 - 0.125 * k], F))
        d.append(dd)
    o.append(np.array([0, 1, 0], F)); d.append(np.array([0, 0, 0], F))  # zero direction -> NaN ray
    o.append(np.array([0, np.inf, 0], F)); d.append(np.array([0, -1, 0], F))
    return np.array(o, F), np.array(d, F)


def raycast_case(name, meshes, nrand, seed, mini=1e-4, maxi=1e8, extra=None):
    r = ref.RefRaycaster()
    for pos, idx, model in meshes:
        r.add_mesh(pos, idx, model)
    r.build()
    dump = r.dump()
    ro, rd = wl.random_rays(nrand, seed)
    co, cd = wl.coherent_rays(48)
    eo, ed = edge_rays()
    org = np.concatenate([ro, co, eo] + ([extra[0]] if extra else [])).astype(F)
    dirs = np.concatenate([rd, cd, ed] + ([extra[1]] if extra else [])).astype(F)
    h = r.intersect(org, dirs, mini, maxi)
    any_hit, _ = r.intersect_any(org, dirs, mini, maxi)
    h2 = r.intersect(org, dirs, 0.05, 1.7)
    any2, _ = r.intersect_any(org, dirs, 0.05, 1.7)
    p1 = (org + dirs).astype(F)
    vis = r.visible(org[:512], p1[:512])
    save = dict(org=org, dirs=dirs, mini=F(mini), maxi=F(maxi), hit=h["hit"], dist=h["dist"], u=h["u"], v=h["v"], w=h["w"],
                local_id=h["local_id"], mesh_id=h["mesh_id"], any=any_hit,
                hit_b=h2["hit"], dist_b=h2["dist"], local_b=h2["local_id"], mesh_b=h2["mesh_id"], any_b=any2, visible=vis,
                node_box=dump["node_box"], node_left=dump["node_left"], node_right=dump["node_right"], node_leaf=dump["node_leaf"],
                tri_v=dump["tri_v"], tri_local=dump["tri_local"], tri_mesh=dump["tri_mesh"], verts=dump["verts"], mesh_count=np.uint32(dump["mesh_count"]))
    for i, (pos, idx, model) in enumerate(meshes):
        save["mesh%d_pos" % i] = pos
        save["mesh%d_idx" % i] = idx
        save["mesh%d_model" % i] = np.eye(4, dtype=F).reshape(16) if model is None else model
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "rays", org.shape[0], "hits", int(h["hit"].sum()), "nodes", dump["node_left"].size)


def degenerate_meshes():
    """Geometry that stresses ties and rejects: 40 copies of one triangle (equal centroids -> the nth_element fallback of
    updateHierarchy, Raycaster.cpp:116-123, and exact distance ties), a grid with every triangle listed twice, zero-area /
    collinear / sub-epsilon triangles between good ones, two coplanar overlapping quads."""
    one = np.array([[-0.3, 0.8, -0.3], [0.4, 0.8, -0.2], [-0.1, 0.8, 0.5]], F)
    m1 = (one, np.tile(np.array([0, 1, 2], np.uint32), 40), None)
    p6, i6 = wl.displaced_grid(6)
    m2 = (p6, np.concatenate([i6, i6]).astype(np.uint32), None)
    v = np.array([[0, 0.5, 0], [0.5, 0.5, 0], [0, 0.5, 0.5], [0.25, 0.5, 0.25], [1e-7, 0.5, 0], [0, 0.5, 1e-7], [0.5, 0.5, 0.5], [-0.5, 0.5, -0.5], [0.5, 0.5, -0.5]], F)
    i3 = np.array([0, 0, 1,  0, 3, 6,  0, 4, 5,  0, 1, 2,  1, 1, 1,  7, 8, 6,  0, 6, 3,  2, 1, 6,  7, 6, 2], np.uint32)
    m3 = (v, i3, translate((0.0, -0.1, 0.0)))
    q = np.array([[-0.8, 0.3, -0.8], [0.8, 0.3, -0.8], [0.8, 0.3, 0.8], [-0.8, 0.3, 0.8], [-0.4, 0.3, -0.4], [0.4, 0.3, -0.4], [0.4, 0.3, 0.4], [-0.4, 0.3, 0.4]], F)
    m4 = (q, np.array([0, 1, 2, 0, 2, 3, 4, 5, 6, 4, 6, 7], np.uint32), None)
    return [m1, m2, m3, m4]


def degenerate_rays():
    """Straight down onto the vertices and edge midpoints of the 6x6 grid and of the coplanar quads (shared edges: two or
    more triangles at exactly the same distance), plus grazing rays inside the quads' plane."""
    p6, _ = wl.displaced_grid(6)
    o = [np.array([x, 2.0, z], F) for x, _, z in p6]
    o += [np.array([(a[0] + b[0]) / 2, 2.0, (a[2] + b[2]) / 2], F) for a, b in zip(p6[:-1], p6[1:])]
    o += [np.array([x, 2.0, z], F) for x in (-0.8, -0.4, 0.0, 0.4, 0.8) for z in (-0.8, -0.4, 0.0, 0.4, 0.8)]
    d = [np.array([0, -1, 0], F)] * len(o)
    o += [np.array([-2.0, 0.3, z], F) for z in (-0.8, -0.4, 0.0, 0.1)]
    d += [np.array([1, 0, 0], F)] * 4
    return np.array(o, F), np.array(d, F)


def case_meshes():
    pos, idx = wl.displaced_grid(37)
    # Several meshes: transformed copies, a 2-triangle quad (its root is a leaf), a single triangle, overlapping geometry.
    p8, i8 = wl.displaced_grid(8)
    quad = (np.array([[-1, 0.6, -1], [1, 0.6, -1], [1, 0.6, 1], [-1, 0.6, 1]], F), np.array([0, 1, 2, 0, 2, 3], np.uint32))
    tri = (np.array([[0, 0.9, 0], [0.5, 0.9, 0], [0, 0.9, 0.5]], F), np.array([0, 1, 2], np.uint32))
    return {"raycast_degenerate": degenerate_meshes(),
            "raycast_grid37": [(pos, idx, None)],
            "raycast_multimesh": [(p8, i8, None), (p8, i8, rot_scale(0.7, 0.5, (0.2, 0.4, -0.1))), (quad[0], quad[1], None),
                                  (tri[0], tri[1], translate((0.1, 0.0, 0.1))), (p8, i8, translate((0.0, 0.0, 0.0)))]}


def levels_goldens():
    """The reference's own RaycasterVisualisation::getRayLevels (unmodified, oracle/_ref/librendu_refvis.so) on rays of
    the two raycast cases: boxes drawn per depth level + the hit it returns; and getAllLevels' per-level node counts."""
    save = {}
    for name, meshes in case_meshes().items():
        r = ref.RefRaycaster()
        for pos, idx, model in meshes:
            r.add_mesh(pos, idx, model)
        r.build()
        case = np.load(os.path.join(OUT, name + ".npz"))
        pick = np.concatenate([np.arange(0, 3000, 250), np.arange(3000, 3000 + 48 * 48, 192), np.arange(3000 + 48 * 48, case["org"].shape[0], 7)])
        if name == "raycast_degenerate":
            pick = np.concatenate([pick, np.arange(case["org"].shape[0] - 60, case["org"].shape[0], 3)])
        counts_all, boxes_all, hits = [], [], []
        for i in pick:
            counts, boxes, hit = r.ray_levels(case["org"][i], case["dirs"][i], float(case["mini"]), float(case["maxi"]))
            row = np.zeros(64, np.uint32)
            row[:counts.size] = counts
            counts_all.append(row)
            boxes_all.append(boxes)
            hits.append([hit["hit"], hit["localId"], hit["meshId"], int(np.float32(hit["dist"]).view(np.uint32))])
        save[name + "_rays"] = pick.astype(np.int64)
        save[name + "_counts"] = np.array(counts_all)
        save[name + "_boxes"] = np.concatenate(boxes_all).astype(F)
        save[name + "_hits"] = np.array(hits, np.int64)
        ac, ab = r.all_levels()
        save[name + "_all_counts"] = ac
        print(name, "rays", pick.size, "boxes", save[name + "_boxes"].shape[0], "levels", ac.size)
    np.savez_compressed(os.path.join(OUT, "ray_levels.npz"), **save)


def main():
    cases = case_meshes()
    raycast_case("raycast_grid37", cases["raycast_grid37"], 3000, 7)
    raycast_case("raycast_multimesh", cases["raycast_multimesh"], 3000, 11)
    raycast_case("raycast_degenerate", cases["raycast_degenerate"], 3000, 19, extra=degenerate_rays())
    # Box test KAT on the edge rays against assorted boxes.
    eo, ed = edge_rays()
    rng = np.random.RandomState(5)
    boxes = []
    for k in range(eo.shape[0]):
        lo = rng.uniform(-1, 0.5, 3).astype(F)
        hi = (lo + rng.uniform(0, 1.5, 3)).astype(F)
        if k % 3 == 0:
            hi[k % 3] = lo[k % 3]  # flat box
        if k % 4 == 0:
            lo[(k // 4) % 3] = eo[k][(k // 4) % 3]  # origin exactly on a slab plane
        boxes.append(np.concatenate([lo, hi]))
    boxes = np.array(boxes, F)
    res, rays = ref.box_test(eo, ed, boxes, 1e-4, 1e8)
    np.savez_compressed(os.path.join(OUT, "box_kat.npz"), org=eo, dirs=ed, boxes=boxes, result=res, rays=rays)
    print("box_kat", res.sum(), "/", res.size)


def pathtracer_goldens():
    from rendu_b200 import pathtracer as rpt
    rng = np.random.RandomState(77)
    # Flattened bundled scenes whose textures are tiny (constant 8x8 / 32x32 images).
    for name in ("cornellbox", "cornellbox-emissive"):
        sc = ref.RefScene(name)
        desc = sc.flatten()
        rpt.save_scene(os.path.join(OUT, "scene_%s.npz" % name), desc)
        cam = sc.camera(1.0)
        # (1) reproducible reference render: one forParallel worker, seeded -> pins the oracle's mt19937 mode bit for bit.
        ref.set_concurrency(1)
        ref.seed(5)
        img, _, rays = sc.render(48, 48, 16, 5)
        ref.set_concurrency(0)
        # (2) converged reference image for the statistical gate on the CUDA path.
        ref.seed(11)
        conv, _, _ = sc.render(64, 64, 2048, 5)
        np.savez_compressed(os.path.join(OUT, "render_%s.npz" % name), cam_eye=cam["eye"], cam_corner=cam["corner"], cam_dx=cam["dx"], cam_dy=cam["dy"],
                            mt_seed=np.uint32(5), mt_image=img, mt_rays=np.uint64(rays), converged=conv, converged_spp=np.uint32(2048))
        print("render", name, "mt mean", img.mean(), "rays", rays, "converged mean", conv.mean())
    # Scalar KATs straight from the reference functions.
    n = 4096
    wo = rng.normal(size=(n, 3)).astype(F)
    wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    wo[: n // 8, 2] *= -1  # some below the surface
    wi = rng.normal(size=(n, 3)).astype(F)
    wi /= np.linalg.norm(wi, axis=1, keepdims=True)
    base = rng.uniform(0, 1, (n, 3)).astype(F)
    rough = rng.uniform(0, 1, n).astype(F)
    rough[:64] = 0.0
    metal = (rng.uniform(0, 1, n) > 0.5).astype(F) * rng.uniform(0.5, 1, n).astype(F)
    ev = ref.ggx_eval(wo, base, rough, metal, wi)
    seeds = np.arange(1000, 1000 + n, dtype=np.uint32)
    u3, swi, sev = ref.ggx_sample(wo, base, rough, metal, seeds)
    img8 = rng.uniform(0, 1, (13, 9, 4)).astype(F)
    uv = rng.uniform(-1.5, 2.5, (n, 2)).astype(F)
    uv[:16] = np.array([[0, 0], [1, 1], [0.5, 0.5], [1, 0]] * 4, F)
    i4, i3 = ref.image_sample(img8, uv)
    faces = rng.uniform(0, 2, (6, 8, 8, 4)).astype(F)
    dirs = rng.normal(size=(n, 3)).astype(F)
    dirs[:6] = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], F)
    dirs[6:12] = np.array([[1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1], [-1, 1, -1], [1, -1, 1]], F)
    cube = ref.cubemap_sample(faces, dirs)
    lights = {}
    for scene_name in ("cornellbox", "cornellbox-cutout", "gallery-materials"):
        sc = ref.RefScene(scene_name)
        d = sc.flatten()
        pos = rng.uniform(-2.5, 2.5, (512, 3)).astype(F)
        for li, l in enumerate(d["lights"]):
            key = "%s_%d" % (scene_name.replace("-", "_"), li)
            lights[key + "_desc"] = np.concatenate([[l["kind"], l["casts_shadow"]], l["intensity"], l["position"], l["direction"], [l["radius"], l["inner"], l["outer"]]]).astype(np.float64)
            lights[key + "_pos"] = pos
            lights[key + "_out"] = sc.light_sample(li, pos)
    np.savez_compressed(os.path.join(OUT, "shading_kat.npz"), wo=wo, wi=wi, base=base, rough=rough, metal=metal, ggx_eval=ev, sample_u3=u3, sample_wi=swi,
                        sample_eval=sev, img=img8, uv=uv, img_rgba=i4, img_rgb=i3, faces=faces, dirs=dirs, cube=cube,
                        mt_floats=ref.random_floats(42, 256), mt_ints=ref.random_ints(42, 0, 24, 256), cos_hemi=ref.cosine_hemisphere(9, 256), **lights)
    print("shading_kat", ev.shape, len(lights) // 3, "lights")


def output_goldens():
    """renderOneShot's output tail run by the reference itself (its tonemap lines, its PNG writer, its PNG loader)."""
    rng = np.random.RandomState(31)
    img = (rng.uniform(0, 2.5, (40, 56, 3)) ** 2).astype(F)
    img[0, 0] = [0.0, 1e-9, 50.0]
    img[0, 1] = [1.0, 0.5, 0.25]
    rgba = np.concatenate([np.clip(img, 0, 1), rng.uniform(0, 1, (40, 56, 1)).astype(F)], -1)
    np.savez_compressed(os.path.join(OUT, "output_kat.npz"), img=img, tonemapped=ref.output_ldr(img, True), clipped=ref.output_ldr(np.clip(img, 0, 1), False),
                        rgba=rgba, rgba_out=ref.output_ldr(rgba, True))
    print("output_kat ok")
    # .exr branch: the reference's EXR writer (half channels) read back by its own loader
    hdr = np.exp(rng.uniform(-22, 12.5, (32, 40, 3))).astype(F)
    hdr[0, :8, 0] = [0.0, -0.0, 1e-8, 5.9604645e-8, 2.9802322e-8, 2.98e-8, 6.1e-5, 6.0975552e-05]
    hdr[0, :8, 1] = [65504.0, 65519.9, 65520.0, 70000.0, np.inf, -np.inf, -1.5, 1.0009766]
    hdr[0, :8, 2] = [1.00048828125, 1.0004883, 1.0014648, 0.99975586, 0.33333334, 2047.5, 2048.5, 1e-5]
    hdr[1] *= -1.0
    hdr4 = np.concatenate([hdr, rng.uniform(0, 1, (32, 40, 1)).astype(F)], -1)
    np.savez_compressed(os.path.join(OUT, "output_hdr_kat.npz"), img=hdr, stored=ref.output_hdr(hdr), rgba=hdr4, rgba_stored=ref.output_hdr(hdr4))
    print("output_hdr_kat ok")


if __name__ == "__main__":
    main()
    pathtracer_goldens()
    output_goldens()
    levels_goldens()
