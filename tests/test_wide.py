"""Eight-wide quantised hierarchy (RB_TRACE_WIDE, rendu_b200/csrc/rb_wide.*): the host collapse is checked on the CPU —
structure, outward rounding of every stored box, and the numpy restatement of the kernel's traversal (tests/wide_model.py)
against the oracle; on the GPU the kernel is compared with the exact mode record for record."""
import os

import numpy as np
import pytest

import rendu_b200 as rb
import wide_model as wm
from rendu_b200 import workloads as wl
from util import ROOT, case_meshes, load_case


def _true_boxes(T):
    """Per wide node and slot: the exact box of all triangles below that child (float64)."""
    n = T.nodes.shape[0]
    lo = np.full((n, 8, 3), np.inf)
    hi = np.full((n, 8, 3), -np.inf)
    v0 = T.tris[:, 0:3].astype(np.float64)
    corners = np.stack([v0, v0 + T.tris[:, 4:7], v0 + T.tris[:, 8:11]], 1)  # fp32 sums, as the kernel's inputs
    for node in range(n - 1, -1, -1):  # children have larger indices than their parent
        for s in range(8):
            m = int(T.meta[node, s])
            if m == 0:
                continue
            if (m & 31) >= 24:
                c = int(T.child_base[node]) + (m & 31) - 24
                assert c > node
                lo[node, s], hi[node, s] = lo[c].min(0), hi[c].max(0)
            else:
                first, count = int(T.tri_base[node]) + (m & 31), bin(m >> 5).count("1")
                pts = corners[first:first + count].reshape(-1, 3)
                lo[node, s], hi[node, s] = pts.min(0), pts.max(0)
    return lo, hi


@pytest.mark.parametrize("case", ["raycast_grid37", "raycast_multimesh", "grid96"])
def test_collapse_structure_and_outward_rounding(case):
    if case == "grid96":
        meshes = [wl.displaced_grid(96)]
    else:
        meshes = [m for m in case_meshes(load_case(case))]
    r, T = wm.build(meshes)
    s = r._wide_sizes
    assert s["tris"] == r.sizes()["tris"] and s["roots"] == r.sizes()["meshes"]
    # every triangle of the hierarchy exactly once, and the records are the binary layout's
    assert np.array_equal(np.sort(T.hier), np.arange(s["tris"]))
    kids = (T.meta != 0).sum(1)
    assert kids.min() >= 1 and kids.max() <= 8 and np.array_equal(kids, (T.nodes[:, 3] >> 24))
    # interior children are consecutive nodes in slot order, leaf triangles consecutive offsets in slot order
    for node in range(T.nodes.shape[0]):
        interior = [int(m & 31) - 24 for m in T.meta[node] if m and (m & 31) >= 24]
        assert interior == list(range(len(interior)))
        off = 0
        for m in T.meta[node]:
            if m and (m & 31) < 24:
                assert (m & 31) == off and (m >> 5) in (1, 3, 7)
                off += bin(m >> 5).count("1")
        empty = T.meta[node] == 0
        assert (T.qlo[node][:, empty] == 255).all() and (T.qhi[node][:, empty] == 0).all()
    # outward rounding: the decoded box contains everything below the child, with half the builder's margin to spare
    lo, hi = _true_boxes(T)
    for node in range(T.nodes.shape[0]):
        dlo, dhi = T.child_boxes(node)
        q = np.ldexp(1.0, T.e[node])[:, None]
        used = T.meta[node] != 0
        assert (dlo[:, used] <= lo[node, used].T - 0.03 * q).all() and (dhi[:, used] >= hi[node, used].T + 0.03 * q).all()


@pytest.mark.parametrize("case", ["raycast_multimesh", "raycast_degenerate"])
def test_model_traversal_returns_the_reference_hits(case):
    """The numpy restatement of the kernel against the hits the reference itself returned (tests/golden): overlapping
    meshes with exact-distance ties, and degenerate / flat / empty geometry."""
    c = load_case(case)
    meshes = list(case_meshes(c))
    _, T = wm.build(meshes)
    regular = np.isfinite(c["org"]).all(1) & np.isfinite(c["dirs"]).all(1) & (np.abs(c["dirs"]) > 1e-12).all(1)
    sel = np.flatnonzero(regular)[:1500]
    mini, maxi = float(c["mini"]), float(c["maxi"])
    hits = 0
    for k in sel:
        tri, t, u, v, _, _, _ = wm.trace(T, c["org"][k], c["dirs"][k], mini, maxi)
        assert (tri >= 0) == bool(c["hit"][k]), k
        if tri >= 0:
            hits += 1
            assert np.float32(t).view(np.uint32) == c["dist"][k].view(np.uint32) and np.float32(u) == c["u"][k] and np.float32(v) == c["v"][k], k
            assert int(T.tris[np.flatnonzero(T.hier == tri)[0], 3].view(np.int32)) == int(c["local_id"][k])
        a = wm.trace(T, c["org"][k], c["dirs"][k], mini, maxi, any_hit=True)[0] >= 0
        assert a == bool(c["any"][k]), k
    assert hits > 100


def test_sah_host_build_is_a_valid_tree_with_the_same_hits():
    """RB_BUILD_SAH: same framework and records as the reference-order build, other splits — every triangle once, boxes
    that contain their subtrees, and (through the wide collapse and its model) the oracle's closest hits."""
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    n = 4000
    c = np.concatenate([rng.normal(0, 0.2, (n // 2, 3)), rng.uniform(-2, 2, (n // 2, 3))]).astype(np.float32)  # a dense clump in a sparse cloud
    e1, e2 = rng.normal(0, 0.03, (n, 3)).astype(np.float32), rng.normal(0, 0.03, (n, 3)).astype(np.float32)
    pos = np.stack([c, c + e1, c + e2], 1).reshape(-1, 3).astype(np.float32)
    idx = np.arange(3 * n, dtype=np.uint32)
    o = orc.OracleRaycaster()
    o.add_mesh(pos, idx)
    o.build()
    r = rb.Raycaster()
    r.addMesh(pos, idx)
    r.build_host_only(rb.BUILD_SAH)
    nodes, tris, _ = r.download_bvh()
    assert sorted(tris["localId"].tolist()) == list(range(0, 3 * n, 3)) and r.sizes()["depth"] < 64
    leaf = nodes["leaf"].astype(bool)
    assert nodes["right"][leaf].max() <= 2 and nodes["right"][leaf].sum() == n
    inner = np.flatnonzero(~leaf)
    for child in ("left", "right"):
        ch = nodes[child][inner].astype(np.int64)
        assert (nodes["bmin"][ch] >= nodes["bmin"][inner]).all() and (nodes["bmax"][ch] <= nodes["bmax"][inner]).all()
    r.build_wide()
    wn, wt, wr = r.download_wide()
    T = wm.WideTree(wn, wt, wr, [(nodes["bmin"][0].astype(np.float32), nodes["bmax"][0].astype(np.float32))])
    local = tris["localId"].astype(np.int64)
    org = rng.uniform(-3, 3, (400, 3)).astype(np.float32)
    dirs = (rng.normal(0, 0.5, (400, 3)).astype(np.float32) - org).astype(np.float32)
    h = o.intersect(org, dirs, threads=4)
    assert int(h["hit"].sum()) > 50
    for k in range(org.shape[0]):
        tri, t, _, _, _, _, _ = wm.trace(T, org[k], dirs[k])
        assert (tri >= 0) == bool(h["hit"][k]), k
        assert tri < 0 or (local[tri] == int(h["localId"][k]) and np.float32(t).view(np.uint32) == h["dist"][k].view(np.uint32)), k


def test_soup_far_from_the_origin_keeps_margins_and_hits():
    """Tiny and huge triangles around (4000, 4000, 4000): fp32 spacing there is 2.4e-4, coarser than the quanta of small
    nodes — the builder has to grow their scale until the outward margin survives; hits must stay the oracle's."""
    from oracle import oracle as orc
    rng = np.random.default_rng(7)
    n = 3000
    c = rng.uniform(-1, 1, (n, 3)).astype(np.float32) * np.float32(5.0) + np.float32(4000.0)
    e1, e2 = rng.normal(0, 0.05, (n, 3)).astype(np.float32), rng.normal(0, 0.05, (n, 3)).astype(np.float32)
    e1[:20] *= 40
    e2[:20] *= 40
    pos = np.stack([c, c + e1, c + e2], 1).reshape(-1, 3).astype(np.float32)
    idx = np.arange(3 * n, dtype=np.uint32)
    o = orc.OracleRaycaster()
    o.add_mesh(pos, idx)
    o.build()
    _, T = wm.build([(pos, idx)])
    lo, hi = _true_boxes(T)
    for node in range(T.nodes.shape[0]):
        dlo, dhi = T.child_boxes(node)
        q = np.ldexp(1.0, T.e[node])[:, None]
        used = T.meta[node] != 0
        assert (dlo[:, used] <= lo[node, used].T - 0.03 * q).all() and (dhi[:, used] >= hi[node, used].T + 0.03 * q).all()
    org = (rng.uniform(-1, 1, (600, 3)) * 12 + 4000).astype(np.float32)
    dirs = ((rng.uniform(-1, 1, (600, 3)) * 5 + 4000).astype(np.float32) - org).astype(np.float32)
    h = o.intersect(org, dirs, threads=4)
    assert int(h["hit"].sum()) > 100
    for k in range(org.shape[0]):
        tri, t, _, _, _, _, _ = wm.trace(T, org[k], dirs[k])
        assert (tri >= 0) == bool(h["hit"][k]), k
        assert tri < 0 or (tri == int(h["internalId"][k]) and np.float32(t).view(np.uint32) == h["dist"][k].view(np.uint32)), k


def test_warp_model_of_the_wide_kernels_stays_above_perfect_packing():
    pos, idx = wl.displaced_grid(24)
    _, T = wm.build([(pos, idx)])
    org, dirs = wl.coherent_rays(12)
    seqs = []
    for k in range(org.shape[0]):
        st = []
        _, _, _, _, nodes, tris, _ = wm.trace(T, org[k], dirs[k], steps=st)
        assert st.count(0) == nodes and st.count(2) == tris
        seqs.append(st)
    ideal = sum(s.count(0) * wm.C_ENTER + s.count(2) * wm.C_TRI for s in seqs) / 32.0 / len(seqs)
    for phased in (False, True):
        cost, lanes = wm.simulate_warp(seqs, 12, phased)
        assert cost > ideal and 1.0 <= lanes <= 32.0


def _host_check():
    """tests/host_check: rb_wide.cuh itself, compiled by g++ (the PTX primitives and intrinsics shimmed), as a ctypes library."""
    import ctypes as C
    import os
    import subprocess
    here = os.path.join(ROOT, "tests", "host_check")
    out = os.path.join(here, "libwide_host_check.so")
    deps = [os.path.join(here, f) for f in ("wide_host_check.cpp", "cuda_shim.h")] + [os.path.join(ROOT, "rendu_b200", "csrc", f) for f in ("rb_wide.cuh", "rb_wide.h", "rb_trace.cuh", "rb_math.cuh", "rb_types.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I/usr/local/cuda/include", deps[0], "-o", out])
    L = C.CDLL(out)
    L.rb_hc_wide_trace.restype = C.c_int
    L.rb_hc_wide_trace.argtypes = [C.c_void_p] + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 3
    return L


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("case", ["raycast_multimesh", "raycast_degenerate", "raycast_grid37"])
def test_device_source_stepped_on_the_host_returns_the_reference_hits(case, mode):
    """The kernels' own traversal source (rb_wide_step; rb_wide_phase_node / _triangle as the phase-voting kernel orders them)
    on every regular ray of the golden cases, closest and any hit, against what the reference returned."""
    from rendu_b200._capi import ptr
    L = _host_check()
    c = load_case(case)
    r, T = wm.build(list(case_meshes(c)))
    nm = len(T.roots)
    roots = np.zeros((nm, 8), np.float32)
    for m in range(nm):
        if T.roots[m] < 0:
            roots[m] = [1, 1, 1, 0, -1, -1, -1, 0]  # empty mesh: the inverted box the library stores
        else:
            roots[m, 0:3], roots[m, 4:7] = T.root_boxes[m]
    org, dirs = np.ascontiguousarray(c["org"], np.float32), np.ascontiguousarray(c["dirs"], np.float32)
    n = org.shape[0]
    nodes, tris, wroots = np.ascontiguousarray(T.nodes), np.ascontiguousarray(T.tris), np.ascontiguousarray(T.roots)
    local = np.ascontiguousarray(tris[:, 3]).view(np.int32)[np.argsort(T.hier)]  # localId by hierarchy index
    for any_hit in (0, 1):
        tri, tuv = np.zeros(n, np.int32), np.zeros((n, 3), np.float32)
        skipped = L.rb_hc_wide_trace(ptr(roots), nm, ptr(nodes), ptr(tris), ptr(wroots), r._wide_sizes["depth"], ptr(org), ptr(dirs), n, float(c["mini"]), float(c["maxi"]), any_hit, mode, 0, ptr(tri), ptr(tuv), None)
        took = tri != -2
        assert skipped == int((~took).sum()) and took.sum() > 0.9 * n
        if any_hit:
            assert np.array_equal(tri[took] >= 0, c["any"][took].astype(bool))
            continue
        assert np.array_equal(tri[took] >= 0, c["hit"][took].astype(bool))
        h = took & (tri >= 0)
        assert h.sum() > 100
        for k, name in enumerate(("dist", "u", "v")):
            assert np.array_equal(tuv[h, k].view(np.uint32), c[name][h].view(np.uint32)), name
        assert np.array_equal(local[tri[h]], c["local_id"][h].astype(np.int32))


@pytest.mark.parametrize("case", ["raycast_multimesh", "raycast_degenerate", "raycast_grid37"])
def test_slots_by_direction_keep_the_hits(case):
    """The second copy of the nodes (children in slots by direction, visited front to back): a permutation of the first
    copy's slots node by node, the same any-hit flags, and the same closest hit — distance bit for bit; the triangle may be
    another one only where two are hit at exactly the same distance."""
    from rendu_b200._capi import ptr
    L = _host_check()
    c = load_case(case)
    r, T = wm.build(list(case_meshes(c)))
    octn = np.ascontiguousarray(r.download_wide_by_octant())
    assert octn.shape == T.nodes.shape and np.array_equal(octn[:, 0:6], T.nodes[:, 0:6])
    meta_o = np.ascontiguousarray(octn[:, 6:8]).view(np.uint8).reshape(-1, 8)
    assert np.array_equal(np.sort(meta_o, 1), np.sort(T.meta, 1))
    nm = len(T.roots)
    roots = np.zeros((nm, 8), np.float32)
    for m in range(nm):
        if T.roots[m] < 0:
            roots[m] = [1, 1, 1, 0, -1, -1, -1, 0]
        else:
            roots[m, 0:3], roots[m, 4:7] = T.root_boxes[m]
    org, dirs = np.ascontiguousarray(c["org"], np.float32), np.ascontiguousarray(c["dirs"], np.float32)
    n = org.shape[0]
    tris, wroots = np.ascontiguousarray(T.tris), np.ascontiguousarray(T.roots)
    local = np.ascontiguousarray(tris[:, 3]).view(np.int32)[np.argsort(T.hier)]
    for mode in (0, 1):
        for any_hit in (0, 1):
            tri, tuv = np.zeros(n, np.int32), np.zeros((n, 3), np.float32)
            L.rb_hc_wide_trace(ptr(roots), nm, ptr(octn), ptr(tris), ptr(wroots), r._wide_sizes["depth"], ptr(org), ptr(dirs), n, float(c["mini"]), float(c["maxi"]), any_hit, mode, 1, ptr(tri), ptr(tuv), None)
            took = tri != -2
            if any_hit:
                assert np.array_equal(tri[took] >= 0, c["any"][took].astype(bool))
                continue
            assert np.array_equal(tri[took] >= 0, c["hit"][took].astype(bool))
            h = took & (tri >= 0)
            assert np.array_equal(tuv[h, 0].view(np.uint32), c["dist"][h].view(np.uint32))
            other = local[tri[h]] != c["local_id"][h].astype(np.int32)
            # another triangle only at exactly the same distance (asserted above): the multi-mesh and degenerate cases are full
            # of exact ties (overlapping copies, repeated triangles); the plain grid has none to speak of
            assert case != "raycast_grid37" or other.sum() <= 2


def test_wide_needs_the_host_built_tree():
    r = rb.Raycaster()
    with pytest.raises(rb.RenduError):
        r.build_wide()  # nothing built
    pos, idx = wl.displaced_grid(8)
    r.addMesh(pos, idx)
    r.build_host_only()
    assert r.build_wide()["nodes"] > 0
    if rb.device_count() == 0:
        with pytest.raises(rb.RenduError):
            r.set_trace_mode(rb.TRACE_WIDE)  # no device copy without a GPU


# The wide kernels ran on a B200 in round 2 (tools/wide_measure.sh, profiles/r02_wide_*): these tests are no longer gated.


@pytest.mark.gpu
def test_wide_kernel_equals_exact_mode():
    """The kernel against the exact mode on the device: random, coherent and the golden edge rays (NaN / inf / axis-parallel
    directions take the binary tree inside the wide kernel), closest and any hit; then a rebuild discards the wide copy."""
    pos, idx = wl.displaced_grid(96)
    c = load_case("raycast_grid37")
    ro, rd = wl.random_rays(200000, 11)
    co, cd = wl.coherent_rays(256)
    org, dirs = np.concatenate([ro, co, c["org"][-60:]]), np.concatenate([rd, cd, c["dirs"][-60:]])
    r = rb.Raycaster()
    r.addMesh(pos, idx)
    r.addMesh(pos * np.float32(0.5) + np.float32(0.25), idx)
    r.updateHierarchy()
    exact, exact_any = r.intersects(org, dirs), r.intersectsAny(org, dirs)
    r.build_wide()
    r.set_trace_mode(rb.TRACE_WIDE)
    wide, wide_any = r.intersects(org, dirs), r.intersectsAny(org, dirs)
    assert np.array_equal(wide_any, exact_any)
    differ = np.flatnonzero([a.tobytes() != b.tobytes() for a, b in zip(wide, exact)])
    # Only a triangle accepted although the reference's rounded slab test rejects one of its boxes may differ.
    assert differ.size <= 2, (differ.size, wide[differ[:4]], exact[differ[:4]])
    assert int(exact["hit"].sum()) > 50000
    r.updateHierarchy()  # the wide copy goes with the rebuild, and the handle is back in exact mode (rb_raycaster_free_device)
    again = r.intersects(org[:64], dirs[:64])
    with pytest.raises(rb.RenduError):
        r.set_trace_mode(rb.TRACE_WIDE)  # RB_TRACE_WIDE without a new rb_raycaster_wide_build fails loudly
    # the rebuilt tree numbers its triangles differently (internalId): compare the hit itself
    for f in ("hit", "dist", "u", "v", "localId", "meshId"):
        assert np.array_equal(again[f], exact[:64][f]), f


@pytest.mark.gpu
def test_wide_phased_kernel_equals_exact_mode(monkeypatch):
    monkeypatch.setenv("RB_WIDE_PHASED", "1")
    pos, idx = wl.displaced_grid(96)
    ro, rd = wl.random_rays(200000, 12)
    co, cd = wl.coherent_rays(256)
    c = load_case("raycast_grid37")
    org, dirs = np.concatenate([ro, co, c["org"][-60:]]), np.concatenate([rd, cd, c["dirs"][-60:]])
    r = rb.Raycaster()
    r.addMesh(pos, idx)
    r.updateHierarchy()
    exact, exact_any = r.intersects(org, dirs), r.intersectsAny(org, dirs)
    r.build_wide()
    r.set_trace_mode(rb.TRACE_WIDE)
    wide, wide_any = r.intersects(org, dirs), r.intersectsAny(org, dirs)
    assert np.array_equal(wide_any, exact_any)
    differ = np.flatnonzero([a.tobytes() != b.tobytes() for a, b in zip(wide, exact)])
    assert differ.size <= 2, (differ.size, wide[differ[:4]], exact[differ[:4]])


@pytest.mark.gpu
def test_path_tracer_on_the_wide_tree_renders_the_same_image():
    """Same counter-based random numbers, same hits: the frame traced through the wide tree equals the exact-mode frame
    except in pixels whose path met one of the (rare) differing hits."""
    from rendu_b200 import pathtracer as rpt
    desc, cam = wl.cornell_synthetic(64, 64, 32), wl.cornell_camera(1.0)
    scene = rpt.Scene(desc)
    pt = rpt.PathTracer(scene)
    a, b = np.zeros((96, 96, 3), np.float32), np.zeros((96, 96, 3), np.float32)
    pt.render(cam, 16, 5, a, rpt.make_opts(seed=5))
    scene.raycaster().build_wide()
    pt.render(cam, 16, 5, b, rpt.make_opts(seed=5, trace_mode=rb.TRACE_WIDE))
    assert np.isfinite(b).all() and b.mean() > 0.01
    assert (np.abs(a - b).max(axis=2) > 1e-6).mean() < 0.002
