"""CPU tests (no GPU): the C oracle against (a) golden vectors generated from the real reference and (b) the real
reference itself when oracle/_ref is present; plus the product's host-side BVH builder against both."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref
from rendu_b200 import workloads as wl
from util import case_meshes, load_case

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["raycast_grid37", "raycast_multimesh", "raycast_degenerate"]


def build_oracle(case):
    o = orc.OracleRaycaster()
    for pos, idx, model in case_meshes(case):
        o.add_mesh(pos, idx, model)
    o.build()
    return o


@pytest.mark.parametrize("name", CASES)
def test_oracle_build_matches_golden(name):
    c = load_case(name)
    o = build_oracle(c)
    assert o.nodes.size == c["node_left"].size
    assert np.array_equal(o.nodes["left"], c["node_left"])
    assert np.array_equal(o.nodes["right"], c["node_right"])
    assert np.array_equal(o.nodes["leaf"], c["node_leaf"])
    assert np.array_equal(np.concatenate([o.nodes["bmin"], o.nodes["bmax"]], 1).view(np.uint32), c["node_box"].view(np.uint32))
    assert np.array_equal(o.tris["localId"], c["tri_local"])
    assert np.array_equal(o.tris["meshId"], c["tri_mesh"])
    assert np.array_equal(np.stack([o.tris["v0"], o.tris["v1"], o.tris["v2"]], 1), c["tri_v"])
    assert np.array_equal(o.verts.view(np.uint32), c["verts"].view(np.uint32))


@pytest.mark.parametrize("name", CASES)
def test_oracle_hits_match_golden(name):
    c = load_case(name)
    o = build_oracle(c)
    h = o.intersect(c["org"], c["dirs"], float(c["mini"]), float(c["maxi"]))
    assert np.array_equal(h["hit"], c["hit"])
    for k in ("dist", "u", "v", "w"):
        assert np.array_equal(h[k].view(np.uint32), c[k].view(np.uint32)), k
    assert np.array_equal(h["localId"], c["local_id"]) and np.array_equal(h["meshId"], c["mesh_id"])
    assert np.array_equal(o.intersect_any(c["org"], c["dirs"], float(c["mini"]), float(c["maxi"])), c["any"])
    hb = o.intersect(c["org"], c["dirs"], 0.05, 1.7)
    assert np.array_equal(hb["hit"], c["hit_b"]) and np.array_equal(hb["dist"].view(np.uint32), c["dist_b"].view(np.uint32))
    assert np.array_equal(hb["localId"], c["local_b"]) and np.array_equal(hb["meshId"], c["mesh_b"])
    assert np.array_equal(o.intersect_any(c["org"], c["dirs"], 0.05, 1.7), c["any_b"])
    p1 = (c["org"] + c["dirs"]).astype(np.float32)
    assert np.array_equal(o.visible(c["org"][:512], p1[:512]), c["visible"])


def levels_grouped(boxes, depths, nlevels=64):
    """Visiting-order records -> what the viewer draws: boxes grouped by depth (order kept inside a level)."""
    order = np.argsort(depths, kind="stable")
    return np.bincount(depths, minlength=nlevels).astype(np.uint32), boxes[order]


@pytest.mark.parametrize("name", CASES)
def test_oracle_ray_levels_match_reference_visualisation(name):
    """RaycasterVisualisation::getRayLevels (RaycasterVisualisation.cpp:41-77): the oracle's visited nodes, grouped by
    depth, are the boxes the reference's own (unmodified) visualisation draws, level by level, and the hit it returns;
    the breadth-first getAllLevels counts follow from the hierarchy."""
    g = np.load(os.path.join(GOLDEN, "ray_levels.npz"))
    c = load_case(name)
    o = build_oracle(c)
    at = 0
    for k, i in enumerate(g[name + "_rays"]):
        ids, depths, boxes, total, hit = o.ray_levels(c["org"][i], c["dirs"][i], float(c["mini"]), float(c["maxi"]))
        counts, grouped = levels_grouped(boxes, depths)
        assert total == ids.size and np.array_equal(counts, g[name + "_counts"][k])
        want = g[name + "_boxes"][at:at + total]
        at += total
        assert np.array_equal(grouped.view(np.uint32), want.view(np.uint32))
        hw = g[name + "_hits"][k]
        assert int(hit["hit"]) == hw[0] and int(hit["localId"]) == hw[1] and int(hit["meshId"]) == hw[2]
        assert int(np.float32(hit["dist"]).view(np.uint32)) == hw[3]
        single = o.intersect(c["org"][i:i + 1], c["dirs"][i:i + 1], float(c["mini"]), float(c["maxi"]))[0]
        assert single["internalId"] == hit["internalId"] and single["dist"] == hit["dist"]
    assert at == g[name + "_boxes"].shape[0]
    # getAllLevels: node count per depth from a breadth-first walk of the oracle's hierarchy
    depth = np.zeros(o.nodes.size, np.int64)
    frontier, level_counts = list(range(o.mesh_count)), []
    while frontier:
        level_counts.append(len(frontier))
        nxt = []
        for n in frontier:
            if not o.nodes["leaf"][n]:
                nxt += [int(o.nodes["left"][n]), int(o.nodes["right"][n])]
        frontier = nxt
    assert np.array_equal(np.array(level_counts, np.uint32), g[name + "_all_counts"])
    if ref.vis_available():
        r = ref.RefRaycaster()
        for pos, idx, model in case_meshes(c):
            r.add_mesh(pos, idx, model)
        r.build()
        counts, boxes, _ = r.ray_levels(c["org"][5], c["dirs"][5])
        ids, depths, ob, total, _ = o.ray_levels(c["org"][5], c["dirs"][5])
        oc, og = levels_grouped(ob, depths, counts.size)
        assert np.array_equal(oc[:counts.size], counts) and np.array_equal(og.view(np.uint32), boxes.view(np.uint32))


def test_oracle_box_kat():
    c = load_case("box_kat")
    res, rays = orc.box_test(c["org"], c["dirs"], c["boxes"], 1e-4, 1e8)
    assert np.array_equal(res, c["result"])
    assert np.array_equal(rays.view(np.uint32), c["rays"].view(np.uint32))  # Ray::Ray incl. NaN/inf patterns


def test_oracle_threads_agree():
    pos, idx = wl.displaced_grid(24)
    o = orc.OracleRaycaster()
    o.add_mesh(pos, idx)
    o.build()
    ro, rd = wl.random_rays(2000, 3)
    a = o.intersect(ro, rd, threads=1)
    b = o.intersect(ro, rd, threads=5)
    assert a.tobytes() == b.tobytes()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_live_reference_large():
    """Oracle vs the real reference on a larger mesh and fresh seeds (not only the committed vectors)."""
    pos, idx = wl.displaced_grid(150)
    r = ref.RefRaycaster()
    r.add_mesh(pos, idx)
    r.build()
    d = r.dump()
    o = orc.OracleRaycaster()
    o.add_mesh(pos, idx)
    o.build()
    assert np.array_equal(o.nodes["left"], d["node_left"]) and np.array_equal(o.nodes["right"], d["node_right"])
    assert np.array_equal(o.tris["localId"], d["tri_local"])
    ro, rd = wl.random_rays(20000, 99)
    co, cd = wl.coherent_rays(96)
    org, dirs = np.concatenate([ro, co]), np.concatenate([rd, cd])
    hr = r.intersect(org, dirs, threads=4)
    ho = o.intersect(org, dirs, threads=4)
    assert np.array_equal(ho["hit"], hr["hit"])
    assert np.array_equal(ho["dist"].view(np.uint32), hr["dist"].view(np.uint32))
    assert np.array_equal(ho["u"].view(np.uint32), hr["u"].view(np.uint32))
    assert np.array_equal(ho["localId"], hr["local_id"])
    any_r, _ = r.intersect_any(org, dirs, threads=4)
    assert np.array_equal(o.intersect_any(org, dirs, threads=4), any_r)
    assert hr["hit"].sum() > 5000


def test_root_block_skipping_visits_the_same_roots_in_the_same_order():
    """Model of rb_next_root (rendu_b200/csrc/rb_trace.cuh) on the oracle's bit-exact box test: with union boxes over
    aligned blocks of 2^k consecutive roots, walking the roots from the last to the first and skipping a block whose
    union the ray misses enters exactly the roots — in exactly the order — of the reference's plain loop
    (Raycaster.cpp:149-155 pushes every root whose box passes, the stack pops them in reverse). Holds because the slab
    test is monotone in the box for rays without NaNs; rays with zero / denormal direction components take the plain
    loop in the product and are left out here."""
    rng = np.random.RandomState(12)
    n = 333
    side = int(np.ceil(np.sqrt(n)))
    lo = np.zeros((n, 3), np.float32)
    for k in range(n):
        lo[k] = [-0.9 + 1.8 * (k % side) / side, 0.2 * rng.uniform(), -0.9 + 1.8 * (k // side) / side]
    hi = (lo + rng.uniform(0.02, 0.12, (n, 3))).astype(np.float32)
    levels = int(np.floor(np.log2(n)))
    blocks = {0: (lo, hi)}
    for k in range(1, levels + 1):
        plo, phi = blocks[k - 1]
        nb = (plo.shape[0] + 1) // 2
        blo, bhi = np.zeros((nb, 3), np.float32), np.zeros((nb, 3), np.float32)
        for b in range(nb):
            sl = slice(2 * b, min(2 * b + 2, plo.shape[0]))
            blo[b], bhi[b] = plo[sl].min(0), phi[sl].max(0)
        blocks[k] = (blo, bhi)
    org, dirs = wl.random_rays(120, 31)
    co, cd = wl.coherent_rays(8)
    org, dirs = np.concatenate([org, co]), np.concatenate([dirs, cd])
    mini, maxi = 1e-4, 1e8

    def passes(k, idx, i):
        b = np.concatenate([blocks[k][0][idx], blocks[k][1][idx]])[None]
        return bool(orc.box_test(org[i:i + 1], dirs[i:i + 1], b, mini, maxi)[0][0])

    skipped_total = 0
    for i in range(org.shape[0]):
        plain = [r for r in range(n - 1, -1, -1) if passes(0, r, i)]
        entered, root, tests = [], n, 0
        while root > 0:
            r = root - 1
            k = min(((r + 1) & -(r + 1)).bit_length() - 1, levels)
            skipped = False
            while k >= 1:
                tests += 1
                if not passes(k, r >> k, i):
                    root = (r >> k) << k
                    skipped = True
                    break
                k -= 1
            if skipped:
                continue
            root = r
            tests += 1
            if passes(0, r, i):
                entered.append(r)
        assert entered == plain, "ray %d" % i
        skipped_total += n - tests
    assert skipped_total > 0.5 * n * org.shape[0]  # most root tests are saved on a row-ordered scene


@pytest.mark.parametrize("n", [9, 52, 64, 65, 443])
def test_root_tree_walk_enters_the_same_roots_as_the_root_loop(n):
    """Model of rb_roots_tree_pretest (rendu_b200/csrc/rb_trace.cuh) and of the tree it walks (rb_root_tree_split / rb_k_root_tree_fit,
    rb_lbvh.cu) on the oracle's bit-exact box test: a median-split tree over the root boxes whose nodes carry the exact fp32
    unions of their children; a ray descends into a child only if it passes the child's box and collects the roots it reaches.
    That set equals the set of the reference's loop over every root (Raycaster.cpp:149-155), so visiting it from the highest
    index down is the reference's order. Overlapping, nested, flat (zero-height) and duplicate boxes; rays without NaNs (the
    others test every root in the product)."""
    rng = np.random.RandomState(n)
    lo = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    hi = (lo + rng.uniform(0.0, 0.5, (n, 3)) * (rng.uniform(size=(n, 3)) > 0.15)).astype(np.float32)  # some zero extents
    lo[n // 2], hi[n // 2] = lo[0], hi[0]  # a duplicate
    lo[n // 3], hi[n // 3] = np.float32([-2, -2, -2]), np.float32([2, 2, 2])  # one that contains everything

    def build(ids):
        if len(ids) == 1:
            return ~ids[0], lo[ids[0]], hi[ids[0]]
        c = 0.5 * (lo[ids] + hi[ids])
        axis = int(np.argmax(c.max(0) - c.min(0)))
        order = sorted(ids, key=lambda r: (0.5 * (lo[r, axis] + hi[r, axis]), r))
        mid = len(order) // 2
        left, right = build(order[:mid]), build(order[mid:])
        node = (left, right)
        return node, np.minimum(left[1], right[1]), np.maximum(left[2], right[2])

    tree = build(list(range(n)))
    org, dirs = wl.random_rays(150, n)
    co, cd = wl.coherent_rays(8)
    org, dirs = np.concatenate([org, co]), np.concatenate([dirs, cd])

    def passes(i, blo, bhi, mini, maxi):
        return bool(orc.box_test(org[i:i + 1], dirs[i:i + 1], np.concatenate([blo, bhi])[None].astype(np.float32), mini, maxi)[0][0])

    saved = 0
    for mini, maxi in ((1e-4, 1e8), (1e-3, 0.7)):
        for i in range(org.shape[0]):
            plain = {r for r in range(n) if passes(i, lo[r], hi[r], mini, maxi)}
            entered, tests, stack = set(), 0, [tree[0]]
            while stack:
                node = stack.pop()
                for child in node:
                    tests += 1
                    if passes(i, child[1], child[2], mini, maxi):
                        if isinstance(child[0], tuple):
                            stack.append(child[0])
                        else:
                            entered.add(~child[0])
            assert entered == plain, "ray %d" % i
            saved += n - tests
    assert n < 16 or saved > 0
