"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/librendu_ref.so (the real reference, built by
oracle/build_ref.sh from /root/reference). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under rendu_b200/ does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "librendu_ref.so")
RES_ROOT = os.path.join(_HERE, "_ref", "resources")

_lib = None
_resources_added = False

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("reference library missing: run oracle/build_ref.sh (needs /root/reference)")
        L = C.CDLL(LIB_PATH)
        L.ref_raycaster_create.restype = C.c_void_p
        L.ref_raycaster_build.restype = C.c_double
        L.ref_raycaster_intersect.restype = C.c_double
        L.ref_raycaster_intersect_any.restype = C.c_double
        L.ref_scene_load.restype = C.c_void_p
        L.ref_scene_raycaster.restype = C.c_void_p
        L.ref_scene_texture.restype = C.c_void_p
        L.ref_scene_render.restype = C.c_double
        L.ref_ray_counter.restype = C.c_uint64
        _lib = L
        L.ref_set_verbose(0)
    return _lib


_vislib = None


def vis_available():
    return available() and os.path.exists(os.path.join(os.path.dirname(LIB_PATH), "librendu_refvis.so"))


def vislib():
    """Second checker library: the reference's unmodified RaycasterVisualisation (oracle/ref_harness/refvis_api.cpp)."""
    global _vislib
    if _vislib is None:
        path = os.path.join(os.path.dirname(LIB_PATH), "librendu_refvis.so")
        if not os.path.exists(path):
            raise RuntimeError("reference visualisation library missing: run oracle/build_ref.sh")
        lib()  # the raycaster handles come from the main checker library
        _vislib = C.CDLL(path)
        _vislib.refvis_ray_levels.restype = C.c_long
        _vislib.refvis_all_levels.restype = C.c_long
    return _vislib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def threads_default():
    return int(lib().ref_forparallel_threads())


def seed(s):
    lib().ref_seed(C.c_uint(s))


def set_concurrency(n):
    """Override what std::thread::hardware_concurrency() reports to the reference (0 restores the real value).
    set_concurrency(1) makes System::forParallel use one worker, so a seeded render is reproducible."""
    lib().ref_set_concurrency(C.c_uint(n))


class RefRaycaster:
    """The reference Raycaster (src/engine/raycaster/Raycaster.cpp) behind plain arrays."""

    def __init__(self, handle=None, owner=None):
        self._own = handle is None
        self._h = C.c_void_p(lib().ref_raycaster_create()) if handle is None else C.c_void_p(handle)
        self._owner = owner

    def __del__(self):
        if getattr(self, "_own", False) and self._h:
            lib().ref_raycaster_destroy(self._h)
            self._h = None

    def add_mesh(self, positions, indices, model=None):
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        m = None if model is None else np.ascontiguousarray(model, np.float32).reshape(16)
        lib().ref_raycaster_add_mesh(self._h, _p(pos), C.c_size_t(pos.shape[0]), _p(idx), C.c_size_t(idx.size), _p(m))

    def build(self):
        return float(lib().ref_raycaster_build(self._h))

    def sizes(self):
        out = np.zeros(4, np.uint64)
        lib().ref_raycaster_sizes(self._h, _p(out))
        return dict(nodes=int(out[0]), tris=int(out[1]), verts=int(out[2]), meshes=int(out[3]))

    def dump(self):
        s = self.sizes()
        nn, nt, nv = s["nodes"], s["tris"], s["verts"]
        d = dict(
            node_box=np.zeros((nn, 6), np.float32), node_left=np.zeros(nn, np.uint64), node_right=np.zeros(nn, np.uint64),
            node_leaf=np.zeros(nn, np.uint8), tri_box=np.zeros((nt, 6), np.float32), tri_v=np.zeros((nt, 3), np.uint64),
            tri_local=np.zeros(nt, np.uint64), tri_mesh=np.zeros(nt, np.uint32), verts=np.zeros((nv, 3), np.float32),
        )
        lib().ref_raycaster_dump(self._h, _p(d["node_box"]), _p(d["node_left"]), _p(d["node_right"]), _p(d["node_leaf"]),
                                 _p(d["tri_box"]), _p(d["tri_v"]), _p(d["tri_local"]), _p(d["tri_mesh"]), _p(d["verts"]))
        d["mesh_count"] = s["meshes"]
        return d

    def intersect(self, org, dirs, mini=1e-4, maxi=1e8, threads=1):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = org.shape[0]
        outf = np.zeros((n, 4), np.float32)
        outi = np.zeros((n, 3), np.uint64)
        dt = lib().ref_raycaster_intersect(self._h, _p(org), _p(dirs), C.c_float(mini), C.c_float(maxi), C.c_size_t(n), int(threads), _p(outf), _p(outi))
        return dict(hit=outi[:, 0].astype(np.uint32), dist=outf[:, 0].copy(), u=outf[:, 1].copy(), v=outf[:, 2].copy(), w=outf[:, 3].copy(),
                    local_id=outi[:, 1].copy(), mesh_id=outi[:, 2].copy(), seconds=float(dt))

    def intersect_any(self, org, dirs, mini=1e-4, maxi=1e8, threads=1):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = org.shape[0]
        out = np.zeros(n, np.uint8)
        dt = lib().ref_raycaster_intersect_any(self._h, _p(org), _p(dirs), C.c_float(mini), C.c_float(maxi), C.c_size_t(n), int(threads), _p(out))
        return out, float(dt)

    def visible(self, p0, p1):
        p0 = np.ascontiguousarray(p0, np.float32).reshape(-1, 3)
        p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 3)
        out = np.zeros(p0.shape[0], np.uint8)
        lib().ref_raycaster_visible(self._h, _p(p0), _p(p1), C.c_size_t(p0.shape[0]), _p(out))
        return out

    def _levels(self, fn, *head):
        cap_levels, cap_nodes = 256, 1 << 22
        counts, boxes, nlev = np.zeros(cap_levels, np.uint32), np.zeros((cap_nodes, 6), np.float32), C.c_uint32(0)
        total = fn(self._h, *head, C.c_uint32(cap_levels), C.c_uint64(cap_nodes), _p(counts), _p(boxes), C.byref(nlev))
        if total < 0:
            raise RuntimeError("reference visualisation exceeds the harness capacities")
        return counts[:nlev.value].copy(), boxes[:total].copy()

    def ray_levels(self, origin, direction, mini=1e-4, maxi=1e8):
        """The reference's own RaycasterVisualisation::getRayLevels (librendu_refvis.so): per depth level the boxes it draws
        (levels concatenated), and the hit it returns."""
        o = np.ascontiguousarray(origin, np.float32).reshape(3)
        d = np.ascontiguousarray(direction, np.float32).reshape(3)
        hf, hi = np.zeros(4, np.float32), np.zeros(3, np.uint64)
        L = vislib()

        def fn(h, *rest):
            return L.refvis_ray_levels(h, _p(o), _p(d), C.c_float(mini), C.c_float(maxi), *rest, _p(hf), _p(hi))
        counts, boxes = self._levels(fn)
        return counts, boxes, dict(hit=int(hi[0]), dist=hf[0], u=hf[1], v=hf[2], w=hf[3], localId=int(hi[1]), meshId=int(hi[2]))

    def all_levels(self):
        """RaycasterVisualisation::getAllLevels: every node box, breadth first, grouped by depth."""
        return self._levels(vislib().refvis_all_levels)

    def counters(self, org, dirs, mini=1e-4, maxi=1e8, any_hit=False):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        out = np.zeros(5, np.uint64)
        lib().ref_raycaster_counters(self._h, _p(org), _p(dirs), C.c_float(mini), C.c_float(maxi), C.c_size_t(org.shape[0]), 1 if any_hit else 0, _p(out))
        return dict(popped=int(out[0]), box_tests=int(out[1]), tri_tests=int(out[2]), interior=int(out[3]), max_stack=int(out[4]))


def box_test(org, dirs, boxes, mini, maxi):
    org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    n = org.shape[0]
    out = np.zeros(n, np.uint8)
    ray = np.zeros((n, 9), np.float32)
    lib().ref_box_test(_p(org), _p(dirs), _p(boxes), C.c_float(mini), C.c_float(maxi), C.c_size_t(n), _p(out), _p(ray))
    return out, ray


def ggx_eval(wo, base, rough, metal, wi):
    wo, base, wi = [np.ascontiguousarray(a, np.float32).reshape(-1, 3) for a in (wo, base, wi)]
    rough, metal = [np.ascontiguousarray(a, np.float32).reshape(-1) for a in (rough, metal)]
    out = np.zeros_like(wo)
    lib().ref_ggx_eval(_p(wo), _p(base), _p(rough), _p(metal), _p(wi), C.c_size_t(wo.shape[0]), _p(out))
    return out


def ggx_sample(wo, base, rough, metal, seeds):
    wo, base = [np.ascontiguousarray(a, np.float32).reshape(-1, 3) for a in (wo, base)]
    rough, metal = [np.ascontiguousarray(a, np.float32).reshape(-1) for a in (rough, metal)]
    seeds = np.ascontiguousarray(seeds, np.uint32)
    n = wo.shape[0]
    u, wi, ev = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    lib().ref_ggx_sample(_p(wo), _p(base), _p(rough), _p(metal), _p(seeds), C.c_size_t(n), _p(u), _p(wi), _p(ev))
    return u, wi, ev


def random_floats(s, n):
    out = np.zeros(n, np.float32)
    lib().ref_random_floats(C.c_uint(s), C.c_size_t(n), _p(out))
    return out


def random_ints(s, lo, hi, n):
    out = np.zeros(n, np.int32)
    lib().ref_random_ints(C.c_uint(s), int(lo), int(hi), C.c_size_t(n), _p(out))
    return out


def cosine_hemisphere(s, n):
    out = np.zeros((n, 3), np.float32)
    lib().ref_cosine_hemisphere(C.c_uint(s), C.c_size_t(n), _p(out))
    return out


def image_sample(rgba, uv):
    rgba = np.ascontiguousarray(rgba, np.float32)
    h, w = rgba.shape[:2]
    uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
    o4, o3 = np.zeros((uv.shape[0], 4), np.float32), np.zeros((uv.shape[0], 3), np.float32)
    lib().ref_image_sample(_p(rgba), w, h, _p(uv), C.c_size_t(uv.shape[0]), _p(o4), _p(o3))
    return o4, o3


def cubemap_sample(faces, dirs):
    faces = np.ascontiguousarray(faces, np.float32)  # (6,h,w,4)
    h, w = faces.shape[1:3]
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    out = np.zeros((dirs.shape[0], 3), np.float32)
    lib().ref_cubemap_sample(_p(faces), w, h, _p(dirs), C.c_size_t(dirs.shape[0]), _p(out))
    return out


def camera_shifts(eye, center, up, fov, ratio, near=0.01, far=100.0):
    cam = np.array(list(eye) + list(center) + list(up) + [fov, ratio, near, far], np.float32)
    out = np.zeros(12, np.float32)
    lib().ref_camera_shifts(_p(cam), _p(out))
    return dict(corner=out[0:3].copy(), dx=out[3:6].copy(), dy=out[6:9].copy(), eye=out[9:12].copy())


def add_default_resources():
    global _resources_added
    if not _resources_added:
        for d in ("pbrdemo", "common"):
            lib().ref_add_resources(os.path.join(RES_ROOT, d).encode())
        _resources_added = True


MATERIAL_EMISSIVE = 8  # Material::Type::Emissive (src/engine/scene/Material.hpp)
BG_NAMES = {0: "COLOR", 1: "IMAGE", 2: "SKYBOX", 3: "ATMOSPHERE"}


class RefScene:
    """A bundled .scene loaded by the reference's own loader, with its PathTracer (PathTracer.cpp:11-21)."""

    def __init__(self, name):
        add_default_resources()
        h = lib().ref_scene_load(name.encode())
        if not h:
            raise RuntimeError("reference failed to load scene %s" % name)
        self._h = C.c_void_p(h)
        self.name = name

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_scene_destroy(self._h)
            self._h = None

    def raycaster(self):
        return RefRaycaster(lib().ref_scene_raycaster(self._h), owner=self)

    def camera(self, ratio):
        out = np.zeros(12, np.float32)
        lib().ref_scene_camera(self._h, C.c_float(ratio), _p(out))
        return dict(corner=out[0:3].copy(), dx=out[3:6].copy(), dy=out[6:9].copy(), eye=out[9:12].copy())

    def render(self, w, h, samples, depth, components=3):
        px = np.zeros((h, w, components), np.float32)
        lib().ref_ray_counter_reset()
        dt = lib().ref_scene_render(self._h, w, h, samples, depth, components, _p(px))
        return px, float(dt), int(lib().ref_ray_counter())

    def light_sample(self, i, pos):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        out = np.zeros((pos.shape[0], 5), np.float32)
        lib().ref_scene_light_sample(self._h, C.c_size_t(i), _p(pos), C.c_size_t(pos.shape[0]), _p(out))
        return out

    def flatten(self):
        """Plain-array description of the scene: exactly what the product's C ABI (include/rendu_b200.h) takes."""
        L = lib()
        cnt = np.zeros(3, np.uint64)
        L.ref_scene_counts(self._h, _p(cnt))
        nobj, nlights = int(cnt[0]), int(cnt[1])
        textures, tex_index = [], {}

        def tex(i, slot, img=0):
            info = np.zeros(4, np.uint64)
            ident = L.ref_scene_texture(self._h, C.c_size_t(i), C.c_size_t(slot), C.c_size_t(img), _p(info), None)
            if not ident:
                return -1
            key = (ident, img)
            if key not in tex_index:
                w, h, c = int(info[0]), int(info[1]), int(info[2])
                px = np.zeros((h, w, c), np.float32)
                L.ref_scene_texture(self._h, C.c_size_t(i), C.c_size_t(slot), C.c_size_t(img), _p(info), _p(px))
                tex_index[key] = len(textures)
                textures.append(px)
            return tex_index[key]

        objects = []
        for i in range(nobj):
            info = np.zeros(6, np.uint64)
            model = np.zeros(16, np.float32)
            L.ref_scene_object_info(self._h, C.c_size_t(i), _p(info), _p(model))
            nv, ni = int(info[0]), int(info[1])
            counts = np.zeros(5, np.uint64)
            L.ref_scene_object_mesh(self._h, C.c_size_t(i), None, None, None, None, None, None, _p(counts))
            pos = np.zeros((nv, 3), np.float32)
            nor = np.zeros((int(counts[0]), 3), np.float32)
            tan = np.zeros((int(counts[1]), 3), np.float32)
            bit = np.zeros((int(counts[2]), 3), np.float32)
            uv = np.zeros((int(counts[3]), 2), np.float32)
            idx = np.zeros(ni, np.uint32)
            L.ref_scene_object_mesh(self._h, C.c_size_t(i), _p(pos), _p(nor), _p(tan), _p(bit), _p(uv), _p(idx), None)
            objects.append(dict(positions=pos, normals=nor, tangents=tan, bitangents=bit, texcoords=uv, indices=idx,
                                model=model.copy(), use_uv=bool(info[2]), material_type=int(info[3]), masked=bool(info[4]),
                                textures=[tex(i, s) for s in range(min(int(info[5]), 3))]))
        lights = []
        for i in range(nlights):
            ks = np.zeros(2, np.int32)
            f = np.zeros(12, np.float32)
            L.ref_scene_light(self._h, C.c_size_t(i), _p(ks), _p(f))
            lights.append(dict(kind=int(ks[0]), casts_shadow=bool(ks[1]), intensity=f[0:3].copy(), position=f[3:6].copy(),
                               direction=f[6:9].copy(), radius=float(f[9]), inner=float(f[10]), outer=float(f[11])))
        mode = np.zeros(1, np.uint64)
        col = np.zeros(4, np.float32)
        sun = np.zeros(3, np.float32)
        L.ref_scene_background(self._h, _p(mode), _p(col), _p(sun))
        bg = dict(mode=int(mode[0]), color=col[:3].copy(), sun=sun.copy(), textures=[])
        SIZE_MAX = (1 << 64) - 1
        bg["sky_lut"] = -1
        if bg["mode"] == 3:
            bg["sky_lut"] = len(textures)
            textures.append(sky_lut())
        if bg["mode"] == 1:
            bg["textures"] = [tex(SIZE_MAX, 0, 0)]
        elif bg["mode"] == 2:
            bg["textures"] = [tex(SIZE_MAX, 0, f) for f in range(6)]
        return dict(objects=objects, lights=lights, background=bg, textures=textures)


class RefBuiltScene:
    """The reference's own Scene / PathTracer filled from a plain-array scene description (the dict layout of
    rendu_b200.workloads / RefScene.flatten) through oracle/ref_harness/ref_build_scene.cpp: lets the UNMODIFIED
    PathTracer::render be timed on the synthetic C3/C4 scene, which no .scene file describes. Point lights and colour /
    cubemap backgrounds only (what that scene uses)."""

    def __init__(self, desc):
        L = lib()
        L.ref_built_scene_create.restype = C.c_void_p
        L.ref_built_scene_finish.restype = C.c_double
        L.ref_built_scene_render.restype = C.c_double
        self._h = C.c_void_p(L.ref_built_scene_create())
        bg = desc["background"]
        cube = [int(i) for i in bg.get("textures", [])] if int(bg["mode"]) == 2 else []
        tex_id = {}
        for i, t in enumerate(desc["textures"]):
            if i in cube:
                continue
            t = np.ascontiguousarray(t, np.float32)
            assert t.ndim == 3 and t.shape[2] == 4, "scene textures are RGBA on the host (Codable.cpp:143-161)"
            tex_id[i] = L.ref_built_scene_add_texture(self._h, _p(t), t.shape[1], t.shape[0], 1)
        bg_tex = -1
        if cube:
            faces = np.ascontiguousarray(np.stack([desc["textures"][i] for i in cube]), np.float32)
            bg_tex = L.ref_built_scene_add_texture(self._h, _p(faces), faces.shape[2], faces.shape[1], 6)
        elif int(bg["mode"]) == 1:
            bg_tex = tex_id[int(bg["textures"][0])]
        for o in desc["objects"]:
            a = {k: np.ascontiguousarray(o[k], np.float32) for k in ("positions", "normals", "tangents", "bitangents", "texcoords")}
            idx = np.ascontiguousarray(o["indices"], np.uint32).reshape(-1)
            nv = a["positions"].reshape(-1, 3).shape[0]
            uv = a["texcoords"] if a["texcoords"].size == 2 * nv else None
            tex3 = np.array([tex_id.get(int(t), -1) if t is not None and int(t) >= 0 else -1 for t in (list(o["textures"]) + [-1, -1, -1])[:3]], np.int32)
            model = np.ascontiguousarray(np.asarray(o["model"], np.float32).reshape(16))
            L.ref_built_scene_add_object(self._h, _p(a["positions"]), _p(a["normals"]), _p(a["tangents"]), _p(a["bitangents"]), _p(uv), C.c_size_t(nv),
                                         _p(idx), C.c_size_t(idx.size), _p(model), _p(tex3), int(o["material_type"]), int(bool(o["masked"])), int(bool(o["use_uv"])))
        for l in desc["lights"]:
            assert int(l["kind"]) == 0, "RefBuiltScene: point lights only"
            L.ref_built_scene_add_point_light(self._h, _p(np.asarray(l["position"], np.float32)), _p(np.asarray(l["intensity"], np.float32)),
                                              C.c_float(float(l["radius"])), int(bool(l["casts_shadow"])))
        L.ref_built_scene_set_background(self._h, int(bg["mode"]), _p(np.asarray(bg.get("color", (0, 0, 0)), np.float32)), int(bg_tex))
        self.build_seconds = float(L.ref_built_scene_finish(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_built_scene_destroy(self._h)
            self._h = None

    def render(self, eye, center, up, fov, w, h, samples, depth, components=3):
        """PathTracer::render with the reference's own Camera (pose + projection(ratio=w/h, fov)); image, seconds, rays."""
        cam = np.array(list(eye) + list(center) + list(up) + [fov], np.float32)
        px = np.zeros((h, w, components), np.float32)
        lib().ref_ray_counter_reset()
        dt = lib().ref_built_scene_render(self._h, _p(cam), w, h, samples, depth, components, _p(px))
        return px, float(dt), int(lib().ref_ray_counter())


def sky_lut():
    add_default_resources()
    L = lib()
    L.ref_sky_lut.restype = C.c_void_p
    info = np.zeros(3, np.uint64)
    if not L.ref_sky_lut(_p(info), None):
        raise RuntimeError("scattering-precomputed table not found in resources/common")
    px = np.zeros((int(info[1]), int(info[0]), int(info[2])), np.float32)
    L.ref_sky_lut(_p(info), _p(px))
    return px


def sky_eval(org, dirs, sun):
    add_default_resources()
    org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    sun = np.ascontiguousarray(sun, np.float32).reshape(3)
    out = np.zeros_like(org)
    lib().ref_sky_eval(_p(org), _p(dirs), _p(sun), C.c_size_t(org.shape[0]), _p(out))
    return out


def output_ldr(image, tonemap=True):
    """renderOneShot's output tail run by the reference itself: tonemap, Image::save (its PNG writer), reload."""
    import tempfile
    image = np.ascontiguousarray(image, np.float32)
    h, w, c = image.shape
    out = np.zeros((h, w, c), np.uint8)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.png")
        rc = lib().ref_output_ldr(_p(image), w, h, c, 1 if tonemap else 0, path.encode(), _p(out))
    if rc != 0:
        raise RuntimeError("reference Image::save/load failed (%d)" % rc)
    return out


def output_hdr(image):
    """The .exr branch of renderOneShot run by the reference itself: Image::save -> saveAsHDR (half channels), reload
    with its own EXR loader. Returns the stored half values as float32, interleaved like the input."""
    import tempfile
    image = np.ascontiguousarray(image, np.float32)
    h, w, c = image.shape
    out = np.zeros((h, w, c), np.float32)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.exr")
        rc = lib().ref_output_hdr(_p(image), w, h, c, path.encode(), _p(out))
    if rc != 0:
        raise RuntimeError("reference Image::save/load (EXR) failed (%d)" % rc)
    return out


class RefMesh:
    """The reference's own OBJ loader (Mesh.cpp:100-266) on a text in memory, optionally followed by computeNormals."""

    def __init__(self, text, mode=2, compute_normals=False):
        data = text if isinstance(text, (bytes, bytearray)) else text.encode("utf-8")
        L = lib()
        L.ref_mesh_load.restype = C.c_void_p
        L.ref_mesh_load.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int]
        h = C.c_void_p(L.ref_mesh_load(bytes(data), len(data), int(mode), 1 if compute_normals else 0))
        n = np.zeros(4, np.uint64)
        L.ref_mesh_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_mesh_sizes(h, _p(n))
        self.positions = np.zeros((int(n[0]), 3), np.float32)
        self.normals = np.zeros((int(n[1]), 3), np.float32)
        self.texcoords = np.zeros((int(n[2]), 2), np.float32)
        self.indices = np.zeros(int(n[3]), np.uint32)
        self.bbox = np.zeros(6, np.float32)
        L.ref_mesh_read.argtypes = [C.c_void_p] * 6
        L.ref_mesh_read(h, _p(self.positions), _p(self.normals), _p(self.texcoords), _p(self.indices), _p(self.bbox))
        L.ref_mesh_free.argtypes = [C.c_void_p]
        L.ref_mesh_free(h)
