"""Python mirror of the OBJ side of the reference's ``Mesh`` (src/engine/resources/Mesh.hpp:13-123) over the C ABI:
``Mesh(text, mode)`` parses a Wavefront OBJ exactly as ``Mesh(std::istream &, Mesh::Load, name)`` does."""
import ctypes as C

import numpy as np

from ._capi import check, lib, ptr

EXPANDED, POINTS, INDEXED = 0, 1, 2


class Mesh:
    def __init__(self, text, mode=INDEXED):
        data = text if isinstance(text, (bytes, bytearray)) else text.encode("utf-8")
        h = C.c_void_p()
        check(lib().rb_obj_parse(bytes(data), len(data), int(mode), C.byref(h)))
        self._h = h
        self._read()

    def _read(self):
        n = np.zeros(4, np.uint64)
        check(lib().rb_mesh_sizes(self._h, ptr(n)))
        self.positions = np.zeros((int(n[0]), 3), np.float32)
        self.normals = np.zeros((int(n[1]), 3), np.float32)
        self.texcoords = np.zeros((int(n[2]), 2), np.float32)
        self.indices = np.zeros(int(n[3]), np.uint32)
        self.bbox = np.zeros(6, np.float32)
        check(lib().rb_mesh_read(self._h, ptr(self.positions), ptr(self.normals), ptr(self.texcoords), ptr(self.indices), ptr(self.bbox)))

    def computeNormals(self):
        check(lib().rb_mesh_compute_normals(self._h))
        self._read()

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().rb_mesh_destroy(self._h)
            except Exception:
                pass
            self._h = None
