"""CPU tests (no GPU) of the product's host side: the C ABI library loads and exports every symbol the header declares,
the reference-order BVH builder reproduces the reference hierarchy, argument checking, and the no-fallback rule."""
import ctypes as C
import os

import numpy as np
import pytest

import rendu_b200 as rb
from rendu_b200 import _capi
from rendu_b200 import workloads as wl
from oracle import oracle as orc
from util import ROOT, case_meshes, header_symbols, load_case


def test_library_exports_every_header_symbol():
    syms = header_symbols()
    assert len(syms) >= 40
    L = C.CDLL(rb.lib_path())
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    rb.lib()  # prototypes bind
    assert b"sm_100a" in rb.lib().rb_version()


def test_record_layouts_match_reference_structs():
    # sizeof(Raycaster::Hit)=48, Node=48, TriangleInfos=64 (verified against the reference build).
    assert rb.HIT_DTYPE.itemsize == 48 and rb.NODE_DTYPE.itemsize == 48 and rb.TRI_DTYPE.itemsize == 64
    assert rb.HIT_DTYPE.fields["localId"][1] == 24 and rb.NODE_DTYPE.fields["leaf"][1] == 40 and rb.TRI_DTYPE.fields["meshId"][1] == 56


@pytest.mark.parametrize("name", ["raycast_grid37", "raycast_multimesh", "raycast_degenerate"])
def test_host_builder_reproduces_reference_hierarchy(name):
    c = load_case(name)
    r = rb.Raycaster()
    for pos, idx, model in case_meshes(c):
        r.addMesh(pos, idx, model)
    st = r.build_host_only()
    assert st in (_capi.RB_OK, _capi.RB_ERR_NO_DEVICE)
    nodes, tris, verts = r.download_bvh()
    assert nodes.size == c["node_left"].size
    assert np.array_equal(nodes["left"], c["node_left"]) and np.array_equal(nodes["right"], c["node_right"])
    assert np.array_equal(nodes["leaf"], c["node_leaf"])
    assert np.array_equal(np.concatenate([nodes["bmin"], nodes["bmax"]], 1).view(np.uint32), c["node_box"].view(np.uint32))
    assert np.array_equal(tris["localId"], c["tri_local"]) and np.array_equal(tris["meshId"], c["tri_mesh"])
    assert np.array_equal(np.stack([tris["v0"], tris["v1"], tris["v2"]], 1), c["tri_v"])
    assert np.array_equal(verts.view(np.uint32), c["verts"].view(np.uint32))
    assert r.sizes()["meshes"] == int(c["mesh_count"])


def test_host_builder_matches_oracle_on_100k_triangles():
    pos, idx = wl.displaced_grid(224)
    o = orc.OracleRaycaster()
    o.add_mesh(pos, idx)
    o.build()
    r = rb.Raycaster()
    r.addMesh(pos, idx)
    r.build_host_only()
    nodes, tris, _ = r.download_bvh()
    assert nodes.size == o.nodes.size == 106239
    assert nodes.tobytes() == o.nodes.tobytes()
    assert np.array_equal(tris["localId"], o.tris["localId"])
    # A rebuild starts from the already reordered triangles (as the reference would): same size, still a valid tree.
    r.build_host_only()
    nodes2, tris2, _ = r.download_bvh()
    assert abs(int(nodes2.size) - int(nodes.size)) < 1000  # the fp32 centroid mean depends on triangle order
    assert sorted(tris2["localId"].tolist()) == sorted(tris["localId"].tolist())


def test_host_builder_parallel_subtrees_keep_the_reference_numbering():
    """Above 64 K triangles the builder hands sub-trees to worker threads and renumbers them afterwards; the result must
    still be the reference's arrays node for node — several meshes of different sizes, including ones too small to split."""
    r, o = rb.Raycaster(), orc.OracleRaycaster()
    for n, shift in ((150, 0.0), (3, 0.5), (200, 1.0), (40, 1.5), (1, 2.0)):
        p, i = wl.displaced_grid(n)
        m = np.eye(4, dtype=np.float32)
        m[3, 1] = shift
        r.addMesh(p, i, m.reshape(16))
        o.add_mesh(p, i, m.reshape(16))
    r.build_host_only()
    o.build()
    nodes, tris, _ = r.download_bvh()
    assert tris.size > 65536 and nodes.tobytes() == o.nodes.tobytes()
    assert np.array_equal(tris["localId"], o.tris["localId"]) and np.array_equal(tris["meshId"], o.tris["meshId"])


def test_host_builder_parallel_partition_matches_std_partition_on_large_and_degenerate_sets():
    """Sets of 128 K triangles and more are split with a multi-threaded restatement of libstdc++'s std::partition (closed
    form of its two-pointer sweep); the oracle runs the real std::partition. 150 K copies of one triangle take the
    median fallback at the top level, a planar grid has thousands of equal centroid coordinates."""
    one = np.array([[-0.3, 0.8, -0.3], [0.4, 0.8, -0.2], [-0.1, 0.8, 0.5]], np.float32)
    p2, i2 = wl.displaced_grid(300)
    flat = p2.copy()
    flat[:, 1] = 0.0
    r, o = rb.Raycaster(), orc.OracleRaycaster()
    for p, i in ((one, np.tile(np.array([0, 1, 2], np.uint32), 150000)), (p2, i2), (flat, i2)):
        r.addMesh(p, i)
        o.add_mesh(p, i)
    r.build_host_only()
    o.build()
    nodes, tris, _ = r.download_bvh()
    assert tris.size == 510000 and nodes.tobytes() == o.nodes.tobytes()
    assert np.array_equal(tris["localId"], o.tris["localId"]) and np.array_equal(tris["meshId"], o.tris["meshId"])


def test_add_mesh_rejects_bad_indices():
    r = rb.Raycaster()
    pos = np.zeros((3, 3), np.float32)
    with pytest.raises(rb.RenduError):
        r.addMesh(pos, np.array([0, 1, 3], np.uint32))
    assert r.sizes()["tris"] == 0


def test_update_mesh_validates_before_touching_the_device():
    """rb_raycaster_update_mesh: unknown mesh and a vertex count that differs from addMesh's are argument errors on any
    machine; the copy itself needs the device (no CPU fallback)."""
    pos, idx = wl.displaced_grid(4)
    r = rb.Raycaster()
    r.addMesh(pos, idx)
    for call in (lambda: r.update_mesh(1, pos), lambda: r.update_mesh(0, pos[:-1])):
        with pytest.raises(rb.RenduError) as e:
            call()
        assert e.value.status == _capi.RB_ERR_INVALID
    with pytest.raises(rb.RenduError) as e:
        r.update_mesh(0, pos)  # not built yet / no device
    assert e.value.status in (_capi.RB_ERR_INVALID, _capi.RB_ERR_NO_DEVICE)


@pytest.mark.skipif(rb.device_count() > 0, reason="checks the behaviour of a machine without a GPU")
def test_no_cpu_fallback_without_gpu():
    pos, idx = wl.displaced_grid(4)
    r = rb.Raycaster()
    r.addMesh(pos, idx)
    with pytest.raises(rb.RenduError) as e:
        r.updateHierarchy()
    assert e.value.status == _capi.RB_ERR_NO_DEVICE
    with pytest.raises(rb.RenduError):
        r.intersects(np.zeros(3, np.float32), np.array([0, -1, 0], np.float32))
    # every compute entry point refuses to run, none computes on the host
    from rendu_b200 import pathtracer as rpt
    for call in (lambda: r.intersectsAny(np.zeros(3, np.float32), np.array([0, -1, 0], np.float32)),
                 lambda: r.visible(np.zeros(3, np.float32), np.ones(3, np.float32)),
                 lambda: r.rayLevels(np.zeros(3, np.float32), np.array([0, -1, 0], np.float32)),
                 lambda: r.updateHierarchy(rb.BUILD_LBVH),
                 lambda: rpt.output_ldr(np.zeros((4, 4, 3), np.float32)),
                 lambda: rpt.output_hdr(np.zeros((4, 4, 3), np.float32)),
                 lambda: rpt.Scene(wl.tiny_scene())):
        with pytest.raises(rb.RenduError) as e:
            call()
        assert e.value.status in (_capi.RB_ERR_NO_DEVICE, _capi.RB_ERR_INVALID)


def test_product_never_touches_the_oracle():
    """The shipped path must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "rendu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "rendu_oracle", "rendu_ref", "oracle/"):
                    assert needle not in text, "%s mentions %s" % (os.path.join(dirpath, f), needle)
    out = os.popen("ldd %s" % rb.lib_path()).read()
    assert "oracle" not in out and "rendu_ref" not in out


def test_workload_shapes():
    pos, idx = wl.displaced_grid(708)
    assert pos.shape == (502681, 3) and idx.size == 3 * 1002528
    o, d = wl.coherent_rays(64, rows=(3, 5))
    assert o.shape == (128, 3) and d.dtype == np.float32


def test_multiply_high_division_identity():
    """The identity behind rb_div (rb_math.cuh): floor(n / d) = floor(n * ceil(2^64 / d) / 2^64) for all 32-bit n and d > 1
    (the device test runs the kernel's own code; this states the arithmetic with Python integers, edge values included)."""
    import random
    rnd = random.Random(7)
    edge = [1, 2, 3, 255, 256, 257, 1920, 65535, 65536, 2**31 - 1, 2**31, 2**31 + 1, 2**32 - 2, 2**32 - 1]
    pairs = [(n, d) for n in edge + [0] for d in edge] + [(rnd.randrange(2**32), rnd.randrange(1, 2**32)) for _ in range(200000)]
    pairs += [(rnd.randrange(2**32), rnd.randrange(1, 5000)) for _ in range(100000)]
    for n, d in pairs:
        m = (2**64 - 1) // d + 1 if d > 1 else 0  # rb_div_magic
        q = (m * n) >> 64 if m else n  # rb_div
        assert q == n // d, (n, d)
