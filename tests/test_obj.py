"""OBJ ingestion (rb_obj_parse) against the reference's own loader — Mesh(std::istream &, Mesh::Load, name)
from the unmodified Mesh.cpp inside oracle/_ref — on every bundled mesh and on hand-made texts that exercise the loader's
corner cases. Host code only: no GPU needed."""
import glob
import os

import numpy as np
import pytest

import rendu_b200 as rb
from rendu_b200 import mesh as rm
from oracle import ref
from util import ROOT

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
BUNDLED = sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "resources", "**", "*.obj"), recursive=True))


def _same(a, b, what):
    for name in ("positions", "normals", "texcoords"):
        x, y = getattr(a, name), getattr(b, name)
        assert x.shape == y.shape, (what, name, x.shape, y.shape)
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (what, name)
    assert np.array_equal(a.indices, b.indices), what
    assert np.array_equal(a.bbox.view(np.uint32), b.bbox.view(np.uint32)), what


@needs_ref
@pytest.mark.parametrize("mode", [rm.EXPANDED, rm.POINTS, rm.INDEXED])
def test_bundled_meshes_load_like_the_reference(mode):
    assert len(BUNDLED) >= 20
    faces = 0
    for path in BUNDLED:
        text = open(path, "rb").read()
        got, want = rm.Mesh(text, mode), ref.RefMesh(text, mode)
        _same(got, want, os.path.basename(path))
        faces += want.indices.size // 3
    assert mode == rm.POINTS or faces > 10000


@needs_ref
def test_compute_normals_like_the_reference():
    for path in BUNDLED:
        text = open(path, "rb").read()
        got = rm.Mesh(text, rm.INDEXED)
        got.computeNormals()
        want = ref.RefMesh(text, rm.INDEXED, compute_normals=True)
        # NaNs (degenerate faces normalise 0/0) must sit in the same places with the same payload-free pattern
        assert np.array_equal(np.isnan(got.normals), np.isnan(want.normals)), path
        ok = ~np.isnan(want.normals)
        assert np.array_equal(got.normals[ok].view(np.uint32), want.normals[ok].view(np.uint32)), path


NASTY = "\r\n".join([
    "# comment", "", "v", "o thing", "v 0 0 0", "v 1.5e0 +2 -3.25 7 8", "v\t0.1\t0.2   0.3", "v 1 2", "vn 0 0 1", "vn 0 1 0 ", "vt 0.25 0.75", "vt 1 0 0",
    "vt 0.5", "g grp", "s off", "usemtl m", "f 1/1/1 2/2/2 3/1/2", "f 3/1/2 2/2/2 1/1/1 2/1/1", "f 1/1/1 02/2/2 3/1/2", "f 1 2", "f 1/2 2/1 3/2", "f 3/1/2 3/1/2 3/1/2", "x",
    "#f 1/1/1 1/1/1 1/1/1", "v 9 9 9"]) + "\n"


@needs_ref
@pytest.mark.parametrize("mode", [rm.EXPANDED, rm.POINTS, rm.INDEXED])
def test_corner_cases_of_the_loader(mode):
    """CRLF line ends, tabs, extra coordinates, short lines, a quad (fourth corner dropped), the same corner spelled
    differently ("2/2/2" and "02/2/2" are two vertices in Indexed mode, as in the reference), "a/b" corners (normal index
    = texcoord index), vertices after the faces, ignored statements."""
    text = NASTY
    got, want = rm.Mesh(text, mode), ref.RefMesh(text, mode)
    _same(got, want, "nasty")
    if mode == rm.INDEXED:
        assert got.indices.size == 15 and got.positions.shape[0] == 7
    no_attr = "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nf 1 2 3\nf 2 4 3\n"
    _same(rm.Mesh(no_attr, mode), ref.RefMesh(no_attr, mode), "positions only")
    _same(rm.Mesh("", mode), ref.RefMesh("", mode), "empty")
    _same(rm.Mesh("vn 0 0 1\nf 1 1 1\n", mode), ref.RefMesh("vn 0 0 1\nf 1 1 1\n", mode), "no positions")


def test_what_the_reference_would_crash_on_is_an_error():
    for bad in ("v 0 0 zero\n", "v 0 0 0\nf 1 2 3\n", "v 0 0 0\nf -1 -1 -1\n", "v 0 0 0\nvt 0 0\nf 1//1 1//1 1//1\n", "v 1e99 0 0\n"):
        with pytest.raises(rb.RenduError):
            rm.Mesh(bad, rm.INDEXED)


@needs_ref
def test_large_file_spans_many_chunks_and_feeds_the_raycaster():
    """A 300k-triangle displaced grid written as OBJ text (≈25 MB: many parser chunks on many threads) loads like the
    reference, and the host builder turns it into the same hierarchy as the arrays it was written from."""
    from rendu_b200 import workloads as wl
    pos, idx = wl.displaced_grid(388)
    lines = ["v %r %r %r" % (float(p[0]), float(p[1]), float(p[2])) for p in pos]
    tri = idx.reshape(-1, 3) + 1
    lines += ["f %d %d %d" % (a, b, c) for a, b, c in tri]
    text = ("\n".join(lines) + "\n").encode()
    got, want = rm.Mesh(text, rm.INDEXED), ref.RefMesh(text, rm.INDEXED)
    _same(got, want, "grid")
    assert got.indices.size == idx.size
    a, b = rb.Raycaster(), rb.Raycaster()
    a.addMesh(got.positions, got.indices)
    b.addMesh(pos[idx], np.arange(idx.size, dtype=np.uint32))  # the same triangles, vertex for vertex
    a.build_host_only()
    b.build_host_only()
    na, ta, _ = a.download_bvh()
    nb, tb, _ = b.download_bvh()
    assert np.array_equal(na["bmin"], nb["bmin"]) and np.array_equal(na["left"], nb["left"]) and np.array_equal(ta["localId"], tb["localId"])
