"""TEST INFRASTRUCTURE — ctypes binding of oracle/librendu_oracle.so, the plain-C CPU restatement of the reference's
hot path (oracle/rendu_oracle*.c). Checker only: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by rendu_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librendu_oracle.so")

NODE = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("left", "<u8"), ("right", "<u8"), ("leaf", "u1"), ("pad", "u1", 7)])
TRI = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("v0", "<u8"), ("v1", "<u8"), ("v2", "<u8"), ("localId", "<u8"), ("meshId", "<u4"), ("pad", "<u4")])
HIT = np.dtype([("hit", "<u4"), ("dist", "<f4"), ("u", "<f4"), ("v", "<f4"), ("w", "<f4"), ("pad", "<u4"), ("localId", "<u8"), ("meshId", "<u8"), ("internalId", "<u8")])


class Bvh(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("nnodes", C.c_size_t), ("tris", C.c_void_p), ("ntris", C.c_size_t),
                ("verts", C.c_void_p), ("nverts", C.c_size_t), ("meshCount", C.c_uint32)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_add_mesh.restype = C.c_size_t
        L.orc_build.restype = C.c_size_t
        L.orc_box.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleRaycaster:
    """Raycaster (src/engine/raycaster/Raycaster.cpp) restated: addMesh, updateHierarchy, intersects, intersectsAny, visible."""

    def __init__(self):
        self.verts = np.zeros((0, 3), np.float32)
        self.tris = np.zeros(0, TRI)
        self.nodes = np.zeros(0, NODE)
        self.mesh_first, self.mesh_tris = [], []

    @classmethod
    def from_arrays(cls, nodes, tris, verts, mesh_count):
        o = cls()
        o.nodes = np.ascontiguousarray(nodes).view(NODE) if nodes.dtype != NODE else np.ascontiguousarray(nodes)
        o.tris = np.ascontiguousarray(tris).view(TRI) if tris.dtype != TRI else np.ascontiguousarray(tris)
        o.verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
        o.mesh_first = [0] * mesh_count
        o.mesh_tris = [0] * mesh_count
        return o

    @property
    def mesh_count(self):
        return len(self.mesh_first)

    def add_mesh(self, positions, indices, model=None):
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        m = None if model is None else np.ascontiguousarray(model, np.float32).reshape(16)
        off = self.verts.shape[0]
        verts = np.zeros((off + pos.shape[0], 3), np.float32)
        verts[:off] = self.verts
        new_tris = np.zeros(idx.size // 3, TRI)
        lib().orc_add_mesh(_p(pos), C.c_size_t(pos.shape[0]), _p(idx), C.c_size_t(idx.size), _p(m), _p(verts), C.c_size_t(off), _p(new_tris), C.c_uint32(self.mesh_count))
        self.mesh_first.append(self.tris.size)
        self.mesh_tris.append(new_tris.size)
        self.verts = verts
        self.tris = np.concatenate([self.tris, new_tris])

    def build(self):
        nodes = np.zeros(2 * self.tris.size + self.mesh_count + 1, NODE)
        mf = np.array(self.mesh_first, np.uint64)
        mt = np.array(self.mesh_tris, np.uint64)
        self.tris = np.ascontiguousarray(self.tris)
        n = lib().orc_build(_p(self.tris), C.c_size_t(self.tris.size), _p(mf), _p(mt), C.c_uint32(self.mesh_count), _p(nodes))
        self.nodes = nodes[:n].copy()

    def _bvh(self):
        return Bvh(_p(self.nodes), self.nodes.size, _p(self.tris), self.tris.size, _p(self.verts), self.verts.shape[0], self.mesh_count)

    def intersect(self, org, dirs, mini=1e-4, maxi=1e8, threads=1):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        out = np.zeros(org.shape[0], HIT)
        b = self._bvh()
        lib().orc_intersect_batch(C.byref(b), _p(org), _p(dirs), C.c_float(mini), C.c_float(maxi), C.c_size_t(org.shape[0]), int(threads), _p(out))
        return out

    def intersect_any(self, org, dirs, mini=1e-4, maxi=1e8, threads=1):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        out = np.zeros(org.shape[0], np.uint8)
        b = self._bvh()
        lib().orc_intersect_any_batch(C.byref(b), _p(org), _p(dirs), C.c_float(mini), C.c_float(maxi), C.c_size_t(org.shape[0]), int(threads), _p(out))
        return out

    def visible(self, p0, p1):
        p0 = np.ascontiguousarray(p0, np.float32).reshape(-1, 3)
        p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 3)
        b = self._bvh()
        L = lib()
        L.orc_visible.restype = C.c_int
        return np.array([L.orc_visible(C.byref(b), _p(p0[i:i + 1]), _p(p1[i:i + 1])) for i in range(p0.shape[0])], np.uint8)


def box_test(org, dirs, boxes, mini, maxi):
    org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    out = np.zeros(org.shape[0], np.uint8)
    rays = np.zeros((org.shape[0], 9), np.float32)
    L = lib()
    for i in range(org.shape[0]):
        L.orc_ray(_p(org[i:i + 1]), _p(dirs[i:i + 1]), _p(rays[i:i + 1]))
        out[i] = L.orc_box(_p(rays[i:i + 1]), _p(boxes[i, 0:3].copy()), _p(boxes[i, 3:6].copy()), C.c_float(mini), C.c_float(maxi))
    return out, rays
