"""tests/host_check on the DEFAULT traversal: the source the binary kernels inline (rb_trace.cuh: rb_step_exact / rb_step_ordered,
rb_phase_node / rb_phase_leaf, the NaN-faithful box test for irregular rays) compiled by g++ and stepped one ray at a time on
the CPU, against the hits the reference itself returned (tests/golden). A regression net for the kernels' per-ray logic that
needs no GPU; the kernels themselves are checked on the device by tests/test_raycaster_gpu.py."""
import ctypes as C

import numpy as np
import pytest

import rendu_b200 as rb
from rendu_b200._capi import ptr
from test_wide import _host_check
from util import case_meshes, load_case

NONE = 0x7fffffff


def flatten(r):
    """The device layout of rb_host_flatten (rb_bvh_host.cpp) rebuilt from the hierarchy in the reference's format."""
    nodes, tris, verts = r.download_bvh()
    nm = r.sizes()["meshes"]
    leaf = nodes["leaf"].astype(bool)
    left, right = nodes["left"].astype(np.int64), nodes["right"].astype(np.int64)
    index = np.cumsum(~leaf) - 1
    ref = np.where(~leaf, index, np.where(right == 0, NONE, ~((left << 3) | np.maximum(right - 1, 0)))).astype(np.int32)
    inner = np.flatnonzero(~leaf)
    out = np.zeros((inner.size, 16), np.float32)
    L, R = left[inner], right[inner]
    out[:, 0:3], out[:, 3:6], out[:, 6:9], out[:, 9:12] = nodes["bmin"][L], nodes["bmax"][L], nodes["bmin"][R], nodes["bmax"][R]
    out[:, 12], out[:, 13] = ref[L].view(np.float32), ref[R].view(np.float32)
    roots = np.zeros((nm, 8), np.float32)
    roots[:, 0:3], roots[:, 4:7] = nodes["bmin"][:nm], nodes["bmax"][:nm]
    roots[:, 3] = ref[:nm].view(np.float32)
    empty = leaf[:nm] & (right[:nm] == 0)
    roots[empty, 0:3], roots[empty, 4:7] = 1.0, -1.0
    a, b, c = verts[tris["v0"].astype(np.int64)], verts[tris["v1"].astype(np.int64)], verts[tris["v2"].astype(np.int64)]
    t = np.zeros((tris.size, 12), np.float32)
    t[:, 0:3], t[:, 4:7], t[:, 8:11] = a, (b - a).astype(np.float32), (c - a).astype(np.float32)
    t[:, 3] = tris["localId"].astype(np.int32).view(np.float32)
    t[:, 7] = tris["meshId"].astype(np.int32).view(np.float32)
    return np.ascontiguousarray(out), t, roots, r.sizes()["depth"] + 1, tris


@pytest.mark.parametrize("mode,steps", [(0, 1), (1, 1), (1, 6), (2, 4)])  # 2: roots box-tested up front (RB_ROOTS_MASK, the path tracer's closest-hit wavefront)
@pytest.mark.parametrize("case", ["raycast_grid37", "raycast_multimesh", "raycast_degenerate"])
def test_binary_traversal_source_returns_the_reference_hits(case, mode, steps):
    L = _host_check()
    L.rb_hc_binary_trace.restype = C.c_int
    L.rb_hc_binary_trace.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 2 + [C.c_int, C.c_float, C.c_float] + [C.c_int] * 4 + [C.c_void_p] * 2
    c = load_case(case)
    r = rb.Raycaster()
    for m in case_meshes(c):
        r.addMesh(*m)
    r.build_host_only()
    nodes, tris, roots, depth, records = flatten(r)
    org, dirs = np.ascontiguousarray(c["org"], np.float32), np.ascontiguousarray(c["dirs"], np.float32)
    n = org.shape[0]
    mini, maxi = float(c["mini"]), float(c["maxi"])

    def run(any_hit, ordered):
        tri, tuv = np.zeros(n, np.int32), np.zeros((n, 3), np.float32)
        assert L.rb_hc_binary_trace(ptr(nodes), ptr(tris), ptr(roots), roots.shape[0], depth, ptr(org), ptr(dirs), n, mini, maxi, any_hit, ordered, mode, steps, ptr(tri), ptr(tuv)) == 0
        return tri, tuv

    # exact order: every ray (the NaN / inf / axis-parallel ones too), record for record
    tri, tuv = run(0, 0)
    assert np.array_equal(tri >= 0, c["hit"].astype(bool))
    h = tri >= 0
    for k, name in enumerate(("dist", "u", "v")):
        assert np.array_equal(tuv[h, k].view(np.uint32), c[name][h].view(np.uint32)), name
    assert np.array_equal(records["localId"][tri[h]].astype(np.int64), c["local_id"][h].astype(np.int64))
    assert np.array_equal(records["meshId"][tri[h]].astype(np.int64), c["mesh_id"][h].astype(np.int64))
    if mode == 2:  # the up-front root tests exist in the reference's order only
        tri_any, _ = run(1, 0)
        assert np.array_equal(tri_any >= 0, c["any"].astype(bool))
        return
    # any hit: exact order and near child first give the reference's flags
    for ordered in (0, 1):
        tri_any, _ = run(1, ordered)
        assert np.array_equal(tri_any >= 0, c["any"].astype(bool)), ordered
    # near child first, closest hit: the same distance except where the reference's order decides (ties, grazing)
    tri_o, tuv_o = run(0, 1)
    same = (tri_o >= 0) == c["hit"].astype(bool)
    ho = (tri_o >= 0) & c["hit"].astype(bool)
    same[ho] &= tuv_o[ho, 0].view(np.uint32) == c["dist"][ho].view(np.uint32)
    assert (~same).sum() <= 2, (~same).sum()
