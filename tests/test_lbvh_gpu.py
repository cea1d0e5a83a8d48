"""GPU tests of the LBVH build, the refit and the hierarchy download (north_star: "GPU LBVH build ... sits alongside
an upload path for the reference's host-built BVH"). An LBVH tree is a different tree than the reference's, so the
bar is: same hit/miss for every ray, the same triangle except where two triangles are hit at exactly the same
distance (shared edges), and bit-identical dist/u/v whenever the triangle is the same (same triangle test)."""
import numpy as np
import pytest

import rendu_b200 as rb
from rendu_b200 import workloads as wl
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def rays_for_tests(n=20000, seed=5):
    ro, rd = wl.random_rays(n, seed)
    co, cd = wl.coherent_rays(96)
    return np.concatenate([ro, co]), np.concatenate([rd, cd])


def compare_with_oracle(h, ho, max_ties):
    assert np.array_equal(h["hit"], ho["hit"])
    hit = h["hit"] == 1
    same = (h["localId"] == ho["localId"]) & (h["meshId"] == ho["meshId"])
    ties = hit & ~same
    assert ties.sum() <= max_ties, ties.sum()
    # A different triangle may only win at (numerically) the same distance.
    if ties.any():
        assert np.allclose(h["dist"][ties], ho["dist"][ties], rtol=1e-5, atol=0)
    ok = hit & same
    for k in ("dist", "u", "v", "w"):
        assert np.array_equal(h[k][ok].view(np.uint32), ho[k][ok].view(np.uint32)), k


@pytest.mark.parametrize("n", [1, 3, 40, 150])
def test_lbvh_hits_match_reference_tree(n):
    pos, idx = wl.displaced_grid(n)
    o = orc.OracleRaycaster()
    o.add_mesh(pos, idx)
    o.build()
    r = rb.Raycaster()
    r.addMesh(pos, idx)
    r.updateHierarchy(rb.BUILD_LBVH)
    org, dirs = rays_for_tests()
    h = r.intersects(org, dirs)
    ho = o.intersect(org, dirs, threads=8)
    compare_with_oracle(h, ho, max_ties=20)
    assert np.array_equal(r.intersectsAny(org, dirs).astype(np.uint8), o.intersect_any(org, dirs, threads=8))
    s = r.sizes()
    # a proper binary tree whose leaves hold one or two triangles (RB_LBVH_LEAF, default 2)
    assert s["tris"] - 1 <= s["nodes"] <= 2 * s["tris"] - 1 and s["nodes"] % 2 == 1 and s["meshes"] == 1 and s["depth"] >= 1


def test_lbvh_degenerate_geometry():
    """Equal Morton codes (40 copies of one triangle, every grid triangle listed twice), zero-area triangles, coplanar
    overlaps: the build must terminate with a proper tree, hit/miss must equal the reference's for every ray, and wherever
    the winning triangle differs (duplicates: any copy may win) the distance must be the same."""
    from util import case_meshes, load_case
    c = load_case("raycast_degenerate")
    r = rb.Raycaster()
    for pos, idx, model in case_meshes(c):
        r.addMesh(pos, idx, model)
    r.updateHierarchy(rb.BUILD_LBVH)
    s = r.sizes()
    assert s["tris"] == c["tri_local"].size and s["depth"] >= 1
    h = r.intersects(c["org"], c["dirs"], float(c["mini"]), float(c["maxi"]))
    assert np.array_equal(h["hit"], c["hit"])
    hit = c["hit"] == 1
    assert np.allclose(h["dist"][hit], c["dist"][hit], rtol=1e-5, atol=0)
    same = hit & (h["localId"] == c["local_id"]) & (h["meshId"] == c["mesh_id"])
    assert np.array_equal(h["dist"][same].view(np.uint32), c["dist"][same].view(np.uint32))
    assert np.array_equal(r.intersectsAny(c["org"], c["dirs"], float(c["mini"]), float(c["maxi"])).astype(np.uint8), c["any"])
    nodes, tris, verts = r.download_bvh()
    assert nodes.size == s["nodes"] and tris.size == s["tris"]


def test_lbvh_multi_mesh_single_triangle_and_download_roundtrip():
    rng = np.random.RandomState(3)
    meshes = []
    for k in range(12):
        p, i = wl.displaced_grid(2 + k % 5)
        m = np.eye(4, dtype=np.float32)
        m[3, :3] = rng.uniform(-0.6, 0.6, 3)
        m[0, 0] = m[1, 1] = m[2, 2] = 0.4
        meshes.append((p, i, m.reshape(16)))
    meshes.append((np.array([[0, 1.2, 0], [0.6, 1.2, 0], [0, 1.2, 0.6]], np.float32), np.array([0, 1, 2], np.uint32), None))
    o = orc.OracleRaycaster()
    r = rb.Raycaster()
    for p, i, m in meshes:
        o.add_mesh(p, i, m)
        r.addMesh(p, i, m)
    o.build()
    r.updateHierarchy(rb.BUILD_LBVH)
    org, dirs = rays_for_tests(30000, 9)
    h = r.intersects(org, dirs)
    compare_with_oracle(h, o.intersect(org, dirs, threads=8), max_ties=30)
    # Download in the reference's format: a proper binary tree over every triangle, parents enclose children.
    nodes, tris, verts = r.download_bvh()
    leaves = nodes[nodes["leaf"] == 1]
    assert nodes.size == 2 * leaves.size - 1 and leaves.size <= tris.size  # leaves hold up to RB_LBVH_LEAF (2) triangles
    assert int(leaves["right"].max()) <= 2
    covered = np.zeros(tris.size, np.int32)
    for lf in leaves:
        covered[int(lf["left"]):int(lf["left"] + lf["right"])] += 1
    assert (covered == 1).all()
    inner = nodes[nodes["leaf"] == 0]
    for side in ("left", "right"):
        ch = nodes[inner[side].astype(np.int64)]
        assert (ch["bmin"] >= inner["bmin"]).all() and (ch["bmax"] <= inner["bmax"]).all()
    assert sorted(zip(tris["meshId"].tolist(), tris["localId"].tolist())) == sorted(zip(o.tris["meshId"].tolist(), o.tris["localId"].tolist()))
    # ... and it goes back in through the upload path, giving the same hits.
    r2 = rb.Raycaster()
    r2.upload_bvh(nodes, tris, verts, r.sizes()["meshes"])
    h2 = r2.intersects(org, dirs)
    assert h2[["hit", "dist", "u", "v", "localId", "meshId"]].tobytes() == h[["hit", "dist", "u", "v", "localId", "meshId"]].tobytes()
    # Single triangle: the root is a leaf.
    r3 = rb.Raycaster()
    r3.addMesh(*meshes[-1][:2])
    r3.updateHierarchy(rb.BUILD_LBVH)
    one = r3.intersects(np.array([0.1, 2, 0.1], np.float32), np.array([0, -1, 0], np.float32))
    assert one["hit"] == 1 and abs(float(one["dist"]) - 0.8) < 1e-6


@pytest.mark.parametrize("mode", [rb.BUILD_REFERENCE, rb.BUILD_LBVH])
def test_refit_after_vertices_moved(mode):
    pos, idx = wl.displaced_grid(64)
    r = rb.Raycaster()
    r.addMesh(pos, idx)
    r.updateHierarchy(mode)
    org, dirs = rays_for_tests(20000, 13)
    before = r.intersects(org, dirs)
    # Refit with unchanged vertices: bottom-up unions reproduce the boxes, results are bit-identical.
    r.refit(pos)
    same = r.intersects(org, dirs)
    assert same.tobytes() == before.tobytes()
    # Re-displace (same topology): results must equal a fresh build on the moved mesh.
    moved = pos.copy()
    moved[:, 1] = (moved[:, 1] * np.float32(2.5) + np.float32(0.05) * np.sin(np.float32(11) * moved[:, 0])).astype(np.float32)
    r.refit(moved)
    h = r.intersects(org, dirs)
    o = orc.OracleRaycaster()
    o.add_mesh(moved, idx)
    o.build()
    compare_with_oracle(h, o.intersect(org, dirs, threads=8), max_ties=20)
    nodes, tris, verts = r.download_bvh()
    assert np.array_equal(verts, moved)
    inner = nodes[nodes["leaf"] == 0]
    ch = nodes[inner["left"].astype(np.int64)]
    assert (ch["bmin"] >= inner["bmin"]).all() and (ch["bmax"] <= inner["bmax"]).all()


def test_lbvh_large_build_depth_and_ordered_mode():
    pos, idx = wl.displaced_grid(500)  # 500k triangles
    r = rb.Raycaster()
    r.addMesh(pos, idx)
    r.updateHierarchy(rb.BUILD_LBVH)
    s = r.sizes()
    assert s["tris"] == 500000 and 18 <= s["depth"] <= 64
    co, cd = wl.coherent_rays(256)
    h = r.intersects(co, cd)
    r.set_trace_mode(rb.TRACE_ORDERED)
    h2 = r.intersects(co, cd)
    assert np.array_equal(h["hit"], h2["hit"])
    assert ((h["localId"] != h2["localId"]) & (h["hit"] == 1)).sum() <= 8
    ref_tree = rb.Raycaster()
    ref_tree.addMesh(pos, idx)
    ref_tree.updateHierarchy(rb.BUILD_REFERENCE)
    h3 = ref_tree.intersects(co, cd)
    assert np.array_equal(h["hit"], h3["hit"]) and ((h["localId"] != h3["localId"]) & (h["hit"] == 1)).sum() <= 8


def test_path_tracer_on_lbvh_tree():
    from rendu_b200 import pathtracer as rpt
    desc = wl.cornell_synthetic(24, 64, 16)
    cam = wl.cornell_camera(1.5)
    imgs = []
    for mode in (rb.BUILD_REFERENCE, rb.BUILD_LBVH):
        pt = rpt.PathTracer(rpt.Scene(desc, build_mode=mode))
        img = np.zeros((48, 72, 3), np.float32)
        pt.render(cam, 16, 5, img, rpt.make_opts(seed=8))
        imgs.append(img)
    off = (np.abs(imgs[0] - imgs[1]).max(axis=2) > 2e-3).mean()
    assert off < 0.01, off


def _model(tx=0.0, ty=0.0, tz=0.0, s=1.0, rot=0.0):
    """Column-major (glm) model matrix flattened: rotation about y, uniform scale, translation."""
    c, sn = np.float32(np.cos(rot)), np.float32(np.sin(rot))
    m = np.zeros((4, 4), np.float32)  # m[col][row]
    m[0] = [s * c, 0, -s * sn, 0]
    m[1] = [0, s, 0, 0]
    m[2] = [s * sn, 0, s * c, 0]
    m[3] = [tx, ty, tz, 1]
    return m.reshape(16)


@pytest.mark.parametrize("mode", [rb.BUILD_REFERENCE, rb.BUILD_LBVH])
def test_update_mesh_then_refit_equals_fresh_add_mesh(mode):
    """Incremental geometry: one mesh of two gets a new model matrix; after update_mesh + refit(None) the vertices are
    the ones addMesh would have produced (bit-exact) and the hits equal a fresh build on the moved scene (different
    tree topology, so the lbvh-style comparison applies)."""
    gpos, gidx = wl.displaced_grid(48)
    bpos, bidx = wl.displaced_grid(12)
    m0, m1 = _model(0.1, 0.6, -0.2, 0.35, 0.3), _model(-0.3, 0.9, 0.25, 0.5, -1.1)
    r = rb.Raycaster()
    r.addMesh(gpos, gidx)
    r.addMesh(bpos, bidx, m0)
    r.updateHierarchy(mode)
    org, dirs = rays_for_tests(20000, 21)
    r.intersects(org, dirs)
    r.update_mesh(1, bpos, m1)
    r.refit()
    h = r.intersects(org, dirs)
    o = orc.OracleRaycaster()
    o.add_mesh(gpos, gidx)
    o.add_mesh(bpos, bidx, m1)
    o.build()
    compare_with_oracle(h, o.intersect(org, dirs, threads=8), max_ties=20)
    assert np.array_equal(r.intersectsAny(org, dirs).astype(np.uint8), o.intersect_any(org, dirs, threads=8))
    fresh = rb.Raycaster()
    fresh.addMesh(gpos, gidx)
    fresh.addMesh(bpos, bidx, m1)
    fresh.updateHierarchy(rb.BUILD_REFERENCE)
    assert np.array_equal(r.download_bvh()[2], fresh.download_bvh()[2])
    # Error behaviour: unknown mesh, wrong vertex count.
    with pytest.raises(rb.RenduError):
        r.update_mesh(2, bpos, m1)
    with pytest.raises(rb.RenduError):
        r.update_mesh(1, bpos[:-1], m1)
    # An uploaded reference-order hierarchy gets its per-mesh ranges back (rb_raycaster_upload_bvh reconstructs them from
    # the triangle records), so its meshes can be moved like those of a handle filled through add_mesh.
    up = rb.Raycaster()
    nodes, tris, verts = fresh.download_bvh()
    up.upload_bvh(nodes, tris, verts, fresh.sizes()["meshes"])
    up.update_mesh(1, bpos, m1)
    up.refit()
    assert np.array_equal(up.download_bvh()[2], fresh.download_bvh()[2])
    for f in ("hit", "dist", "localId", "meshId"):
        assert np.array_equal(up.intersects(org, dirs)[f], fresh.intersects(org, dirs)[f]), f
    with pytest.raises(rb.RenduError):
        up.update_mesh(1, bpos[:-1], m1)


def test_rebuild_after_refit_uses_the_moved_vertices():
    """A rebuild after a refit starts from triangle boxes recomputed on the moved vertices. (Like the reference's, a
    second updateHierarchy starts from the triangle order the first one left, so the tree is not the one of a fresh
    handle; the bar is the lbvh one: same hits up to exact-distance ties, and boxes that contain their children.)"""
    pos, idx = wl.displaced_grid(40)
    moved = pos.copy()
    moved[:, 1] = (moved[:, 1] * np.float32(-3.0) + np.float32(0.2) * np.cos(np.float32(5) * moved[:, 2])).astype(np.float32)
    o = orc.OracleRaycaster()
    o.add_mesh(moved, idx)
    o.build()
    org, dirs = rays_for_tests(20000, 22)
    ho = o.intersect(org, dirs, threads=8)
    for mode in (rb.BUILD_REFERENCE, rb.BUILD_LBVH):
        r = rb.Raycaster()
        r.addMesh(pos, idx)
        r.updateHierarchy(rb.BUILD_REFERENCE)
        r.refit(moved)
        r.updateHierarchy(mode)
        compare_with_oracle(r.intersects(org, dirs), ho, max_ties=20)
        nodes, tris, verts = r.download_bvh()
        assert np.array_equal(verts, moved)
        tv = verts[np.stack([tris["v0"], tris["v1"], tris["v2"]], 1).astype(np.int64)]
        assert np.array_equal(tris["bmin"], tv.min(axis=1)) and np.array_equal(tris["bmax"], tv.max(axis=1))
        inner = nodes[nodes["leaf"] == 0]
        ch = nodes[inner["right"].astype(np.int64)]
        assert (ch["bmin"] >= inner["bmin"]).all() and (ch["bmax"] <= inner["bmax"]).all()
