"""Python mirror of the reference ``Raycaster`` class (src/engine/raycaster/Raycaster.hpp:10-156) over the C ABI.

Method names follow the reference: addMesh / updateHierarchy / intersects / intersectsAny / visible /
interpolateAttribute. The batched forms take (n,3) float32 arrays; a single (3,) origin/direction is also accepted and
returns a single record, like the reference call.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import HIT_DTYPE, NODE_DTYPE, TRI_DTYPE, check, lib, ptr

BUILD_REFERENCE, BUILD_LBVH, BUILD_SAH = 0, 1, 2
TRACE_EXACT, TRACE_ORDERED, TRACE_WIDE = 0, 1, 2


def _camera(c):
    cam = _capi.Camera()
    for k in ("eye", "corner", "dx", "dy"):
        setattr(cam, k, (C.c_float * 3)(*[float(x) for x in c[k]]))
    return cam


def _rays(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    single = a.ndim == 1
    return a.reshape(-1, 3), single


class Raycaster:
    def __init__(self, _handle=None, _owner=None):
        self._owner = _owner
        if _handle is None:
            h = C.c_void_p()
            check(lib().rb_raycaster_create(C.byref(h)))
            self._h = h
            self._own = True
        else:
            self._h = C.c_void_p(_handle)
            self._own = False

    def __del__(self):
        if getattr(self, "_own", False) and getattr(self, "_h", None):
            try:
                lib().rb_raycaster_destroy(self._h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ build (Raycaster.cpp:14-142)
    def addMesh(self, positions, indices, model=None):
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        m = None if model is None else np.ascontiguousarray(model, np.float32).reshape(16)
        check(lib().rb_raycaster_add_mesh(self._h, ptr(pos), pos.shape[0], ptr(idx), idx.size, ptr(m)))

    def updateHierarchy(self, mode=BUILD_REFERENCE):
        check(lib().rb_raycaster_build(self._h, int(mode)))

    def build_host_only(self, mode=BUILD_REFERENCE):
        """Run a host build (reference order, or BUILD_SAH); tolerate the missing GPU (used by CPU-side tests of the builder)."""
        st = lib().rb_raycaster_build(self._h, int(mode))
        if st not in (_capi.RB_OK, _capi.RB_ERR_NO_DEVICE):
            check(st)
        return st

    def upload_bvh(self, nodes, tris, vertices, mesh_count):
        nodes = np.ascontiguousarray(nodes, NODE_DTYPE)
        tris = np.ascontiguousarray(tris, TRI_DTYPE)
        verts = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        check(lib().rb_raycaster_upload_bvh(self._h, ptr(nodes), nodes.size, ptr(tris), tris.size, ptr(verts), verts.shape[0], int(mesh_count)))

    def sizes(self):
        out = np.zeros(5, np.uint64)
        check(lib().rb_raycaster_sizes(self._h, ptr(out)))
        return dict(nodes=int(out[0]), tris=int(out[1]), verts=int(out[2]), meshes=int(out[3]), depth=int(out[4]))

    def download_bvh(self):
        s = self.sizes()
        nodes = np.zeros(s["nodes"], NODE_DTYPE)
        tris = np.zeros(s["tris"], TRI_DTYPE)
        verts = np.zeros((s["verts"], 3), np.float32)
        check(lib().rb_raycaster_download_bvh(self._h, ptr(nodes), ptr(tris), ptr(verts)))
        return nodes, tris, verts

    def refit(self, world_positions=None):
        """Bottom-up refit on the device. None: fit the vertices the handle holds (after update_mesh)."""
        if world_positions is None:
            check(lib().rb_raycaster_refit(self._h, None, 0))
            return
        pos = np.ascontiguousarray(world_positions, np.float32).reshape(-1, 3)
        check(lib().rb_raycaster_refit(self._h, ptr(pos), pos.shape[0]))

    def update_mesh(self, mesh, positions, model=None):
        """New model-space positions / model matrix for one mesh (same topology); call refit() afterwards."""
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        m = None if model is None else np.ascontiguousarray(np.asarray(model, np.float32).reshape(16))
        check(lib().rb_raycaster_update_mesh(self._h, int(mesh), ptr(pos), pos.shape[0], ptr(m) if m is not None else None))

    def tree_cost(self):
        """Surface-area cost of the hierarchy as built and after the last refit (rebuild when the ratio gets large)."""
        out = np.zeros(2, np.float64)
        check(lib().rb_raycaster_tree_cost(self._h, ptr(out)))
        return dict(at_build=float(out[0]), now=float(out[1]), ratio=float(out[1] / out[0]) if out[0] > 0 else 1.0)

    def build_times(self):
        out = np.zeros(2, np.float64)
        check(lib().rb_raycaster_build_times(self._h, ptr(out)))
        return dict(build_device_ms=float(out[0]), refit_device_ms=float(out[1]))

    def set_trace_mode(self, mode):
        check(lib().rb_raycaster_set_trace_mode(self._h, int(mode)))

    def build_wide(self):
        """Collapse the host-built hierarchy into the eight-wide quantised layout (RB_TRACE_WIDE) and upload it when the
        hierarchy is on a device; returns its sizes. The collapse itself runs without a GPU."""
        out = np.zeros(4, np.uint64)
        check(lib().rb_raycaster_wide_build(self._h, ptr(out)))
        self._wide_sizes = dict(nodes=int(out[0]), tris=int(out[1]), roots=int(out[2]), depth=int(out[3]))
        return dict(self._wide_sizes)

    def download_wide(self):
        """(nodes (n, 20) uint32, tris (m, 12) float32 with word 11 = hierarchy index, root_nodes (r,) int32)."""
        s = self._wide_sizes
        nodes = np.zeros((s["nodes"], 20), np.uint32)
        tris = np.zeros((s["tris"], 12), np.float32)
        roots = np.zeros(s["roots"], np.int32)
        check(lib().rb_raycaster_wide_download(self._h, ptr(nodes), ptr(tris), ptr(roots)))
        return nodes, tris, roots

    def download_wide_by_octant(self):
        """The second copy of the wide nodes (n, 20) uint32: children in slots by direction."""
        nodes = np.zeros((self._wide_sizes["nodes"], 20), np.uint32)
        check(lib().rb_raycaster_wide_download_by_octant(self._h, ptr(nodes)))
        return nodes

    # ------------------------------------------------------------------ queries (Raycaster.cpp:144-230)
    def intersects(self, origin, direction, mini=0.0001, maxi=1e8):
        o, single = _rays(origin)
        d, _ = _rays(direction)
        assert o.shape == d.shape
        out = np.zeros(o.shape[0], HIT_DTYPE)
        check(lib().rb_raycaster_intersect(self._h, ptr(o), ptr(d), mini, maxi, o.shape[0], ptr(out)))
        return out[0] if single else out

    def intersectsAny(self, origin, direction, mini=0.0001, maxi=1e8):
        o, single = _rays(origin)
        d, _ = _rays(direction)
        out = np.zeros(o.shape[0], np.uint8)
        check(lib().rb_raycaster_intersect_any(self._h, ptr(o), ptr(d), mini, maxi, o.shape[0], ptr(out)))
        return bool(out[0]) if single else out.astype(bool)

    def visible(self, p0, p1):
        a, single = _rays(p0)
        b, _ = _rays(p1)
        out = np.zeros(a.shape[0], np.uint8)
        check(lib().rb_raycaster_visible(self._h, ptr(a), ptr(b), a.shape[0], ptr(out)))
        return bool(out[0]) if single else out.astype(bool)

    def intersects_primary(self, camera, width, height, mini=0.0001, maxi=1e8, out=None):
        """Closest hits of the primary rays of a pinhole camera (eye, corner, dx, dy as Camera::pixelShifts gives them), generated on
        the device through pixel centres: (height * width,) Hit records, row-major."""
        cam = camera if isinstance(camera, _capi.Camera) else _camera(camera)
        out = np.zeros(width * height, HIT_DTYPE) if out is None else out
        check(lib().rb_raycaster_intersect_primary(self._h, C.byref(cam), width, height, mini, maxi, ptr(out)))
        return out

    def intersectsAny_primary(self, camera, width, height, mini=0.0001, maxi=1e8, out=None):
        cam = camera if isinstance(camera, _capi.Camera) else _camera(camera)
        mine = out is None
        out = np.zeros(width * height, np.uint8) if mine else out
        check(lib().rb_raycaster_intersect_any_primary(self._h, C.byref(cam), width, height, mini, maxi, ptr(out)))
        return out.astype(bool) if mine else out

    def rayLevels(self, origin, direction, mini=1e-4, maxi=1e8, capacity=4096):
        """RaycasterVisualisation::getRayLevels (RaycasterVisualisation.cpp:41-77) from the device tree: boxes (n, 6) and
        depths (n,) of the nodes the ray visited, in visiting order, the total number visited and the Hit."""
        o = np.ascontiguousarray(origin, np.float32).reshape(3)
        d = np.ascontiguousarray(direction, np.float32).reshape(3)
        boxes, depths = np.zeros((capacity, 6), np.float32), np.zeros(capacity, np.uint32)
        count, hit = C.c_uint32(0), np.zeros(1, HIT_DTYPE)
        check(lib().rb_raycaster_ray_levels(self._h, ptr(o), ptr(d), mini, maxi, capacity, ptr(boxes), ptr(depths), C.byref(count), ptr(hit)))
        n = min(count.value, capacity)
        return boxes[:n], depths[:n], count.value, hit[0]

    @staticmethod
    def interpolateAttribute(hit, indices, attribute):
        """Raycaster::interpolateAttribute (Raycaster.hpp:90-97): w*a[i0] + u*a[i1] + v*a[i2]."""
        indices = np.asarray(indices)
        attribute = np.asarray(attribute, np.float32)
        t = int(hit["localId"])
        i0, i1, i2 = int(indices[t]), int(indices[t + 1]), int(indices[t + 2])
        return np.float32(hit["w"]) * attribute[i0] + np.float32(hit["u"]) * attribute[i1] + np.float32(hit["v"]) * attribute[i2]

    # ------------------------------------------------------------------ device-pointer forms (bench / multi-GPU plumbing)
    def intersects_device(self, d_origins, d_directions, n, d_out, mini=0.0001, maxi=1e8, stream=0):
        check(lib().rb_raycaster_intersect_device(self._h, d_origins, d_directions, mini, maxi, n, d_out, stream))

    def intersectsAny_device(self, d_origins, d_directions, n, d_out, mini=0.0001, maxi=1e8, stream=0):
        check(lib().rb_raycaster_intersect_any_device(self._h, d_origins, d_directions, mini, maxi, n, d_out, stream))

    def count_work_device(self, d_origins, d_directions, n, any_hit=False, mini=0.0001, maxi=1e8, stream=0):
        out = np.zeros(4, np.uint64)
        check(lib().rb_raycaster_count_work_device(self._h, d_origins, d_directions, mini, maxi, n, 1 if any_hit else 0, ptr(out), stream))
        return dict(interior=int(out[0]), tris=int(out[1]), roots=int(out[2]), max_stack=int(out[3]))

    def launch_count(self):
        return int(lib().rb_raycaster_launch_count(self._h))
