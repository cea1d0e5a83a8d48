"""Python mirror of the reference ``PathTracer`` (src/apps/pathtracer/PathTracer.hpp:10-75) and of the read-only scene
data it consumes, over the C ABI. ``Scene`` takes the flattened description (plain numpy arrays: objects in
``scene->objects`` order, textures as float images, lights, background) and uploads it; ``PathTracer.render`` follows
the reference signature ``render(camera, samples, depth, image)`` and accumulates into the caller's image.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import Camera, LightDesc, ObjectDesc, RenderOpts, RenderStats, check, lib, ptr
from .raycaster import BUILD_REFERENCE, Raycaster

BG_COLOR, BG_IMAGE, BG_SKYBOX, BG_ATMOSPHERE = 0, 1, 2, 3
LIGHT_POINT, LIGHT_SPOT, LIGHT_DIRECTIONAL = 0, 1, 2
MATERIAL_EMISSIVE = 8  # Material::Type::Emissive


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Scene:
    """Scene::objects / lights / background flattened."""

    def __init__(self, description, build_mode=BUILD_REFERENCE):
        h = C.c_void_p()
        check(lib().rb_scene_create(C.byref(h)))
        self._h = h
        self._keep = []
        L = lib()
        tex_ids = []
        for t in description["textures"]:
            t = np.ascontiguousarray(t, np.float32)
            out = C.c_int32(-1)
            check(L.rb_scene_add_texture(self._h, ptr(t), t.shape[1], t.shape[0], t.shape[2] if t.ndim == 3 else 1, C.byref(out)))
            tex_ids.append(out.value)

        def tid(i):
            return tex_ids[i] if i is not None and i >= 0 else -1

        for o in description["objects"]:
            d = ObjectDesc()
            arrs = {}
            for key, width in (("positions", 3), ("normals", 3), ("tangents", 3), ("bitangents", 3), ("texcoords", 2)):
                a = np.ascontiguousarray(o[key], np.float32).reshape(-1, width)
                arrs[key] = a
            idx = np.ascontiguousarray(o["indices"], np.uint32).reshape(-1)
            nverts = arrs["positions"].shape[0]
            for key in ("normals", "tangents", "bitangents"):
                if arrs[key].shape[0] != nverts:
                    raise _capi.RenduError(1, "object lacks per-vertex %s (the path tracer requires local tangent frames)" % key)
            d.positions, d.normals, d.tangents, d.bitangents = [arrs[k].ctypes.data for k in ("positions", "normals", "tangents", "bitangents")]
            d.texcoords = arrs["texcoords"].ctypes.data if arrs["texcoords"].shape[0] == nverts else None
            d.nverts, d.indices, d.nindices = nverts, idx.ctypes.data, idx.size
            d.model = (C.c_float * 16)(*[float(x) for x in np.asarray(o["model"], np.float32).reshape(16)])
            texs = list(o["textures"]) + [-1, -1, -1]
            d.tex_albedo, d.tex_normal, d.tex_rmao = tid(texs[0]), tid(texs[1]), tid(texs[2])
            d.emissive = 1 if o["material_type"] == MATERIAL_EMISSIVE else 0
            d.masked = 1 if o["masked"] else 0
            d.use_texcoords = 1 if o["use_uv"] else 0
            self._keep.append((arrs, idx))
            check(L.rb_scene_add_object(self._h, C.byref(d)))
        self._tex_ids = tex_ids
        self._add_lights(description["lights"])
        self._set_background(description["background"])
        check(L.rb_scene_commit(self._h, int(build_mode)))
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().rb_scene_destroy(self._h)
            except Exception:
                pass
            self._h = None

    def _tid(self, i):
        return self._tex_ids[i] if i is not None and i >= 0 else -1

    def _add_lights(self, lights):
        for l in lights:
            d = LightDesc()
            d.kind, d.casts_shadow = int(l["kind"]), 1 if l["casts_shadow"] else 0
            d.intensity, d.position, d.direction = _f3(l["intensity"]), _f3(l["position"]), _f3(l["direction"])
            d.radius, d.inner_angle, d.outer_angle = float(l["radius"]), float(l["inner"]), float(l["outer"])
            check(lib().rb_scene_add_light(self._h, C.byref(d)))

    def _set_background(self, bg):
        texs = np.array([self._tid(i) for i in bg.get("textures", [])], np.int32)
        color = np.asarray(bg.get("color", (0, 0, 0)), np.float32)
        sun = np.asarray(bg.get("sun", (0, 1, 0)), np.float32)
        lut = self._tid(bg.get("sky_lut", -1))
        check(lib().rb_scene_set_background(self._h, int(bg["mode"]), ptr(color), ptr(texs) if texs.size else None, texs.size, ptr(sun), lut))

    def set_lights(self, lights):
        """Replace the lights of a committed scene (the reference reads scene->lights at render time, PathTracer.cpp:232)."""
        check(lib().rb_scene_clear_lights(self._h))
        self._add_lights(lights)
        check(lib().rb_scene_sync(self._h))

    def set_background(self, bg):
        """Replace the background of a committed scene (texture ids refer to the textures uploaded at construction)."""
        self._set_background(bg)
        check(lib().rb_scene_sync(self._h))

    def set_object_transform(self, index, model, sync=True):
        """New model matrix (column-major, as description["objects"][i]["model"]) for one object of a committed scene:
        its vertices are re-transformed and the hierarchy refitted on the device at the next sync / render."""
        m = np.ascontiguousarray(np.asarray(model, np.float32).reshape(16))
        check(lib().rb_scene_set_object_transform(self._h, int(index), ptr(m)))
        if sync:
            check(lib().rb_scene_sync(self._h))

    def raycaster(self):
        return Raycaster(_handle=lib().rb_scene_raycaster(self._h), _owner=self)


def make_camera(eye, corner, dx, dy):
    c = Camera()
    c.eye, c.corner, c.dx, c.dy = _f3(eye), _f3(corner), _f3(dx), _f3(dy)
    return c


SPLIT_TILE, SPLIT_SAMPLE = 0, 1


def make_opts(seed=0, tile=None, samples=None, trace_mode=0, max_paths=0, accumulate=False, gpus=0, split=SPLIT_TILE, tile_rows=0):
    o = RenderOpts()
    o.seed = int(seed)
    if tile is not None:
        o.tile_rows, o.tile_index, o.tile_count = [int(x) for x in tile]
    if samples is not None:
        o.sample_index, o.sample_count = [int(x) for x in samples]
    o.trace_mode = int(trace_mode)
    o.max_paths_in_flight = int(max_paths)
    o.accumulate = 1 if accumulate else 0
    o.gpus, o.split = int(gpus), int(split)  # gpus > 1: render() fans the frame out over that many devices of this process
    if gpus and gpus > 1 and tile_rows:
        o.tile_rows = int(tile_rows)
    return o


class PathTracer:
    """PathTracer(scene); render(camera, samples, depth, image); raycaster()."""

    def __init__(self, scene):
        self._scene = scene
        h = C.c_void_p()
        check(lib().rb_pathtracer_create(scene._h, C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().rb_pathtracer_destroy(self._h)
            except Exception:
                pass
            self._h = None

    def raycaster(self):
        return self._scene.raycaster()

    def render(self, camera, samples, depth, image, opts=None):
        """`image` is the caller's float32 (H, W, C>=3) array; as in the reference the result is accumulated into it and
        divided by the sample count, so pass zeros (PathTracer.cpp:288-292). `camera` is a dict/struct with eye, corner,
        dx, dy — what Camera::pixelShifts / position give on the host."""
        if not (isinstance(image, np.ndarray) and image.dtype == np.float32 and image.ndim == 3 and image.flags["C_CONTIGUOUS"]):
            raise ValueError("image must be a C-contiguous float32 array of shape (H, W, C)")
        cam = camera if isinstance(camera, Camera) else make_camera(camera["eye"], camera["corner"], camera["dx"], camera["dy"])
        h, w, c = image.shape
        check(lib().rb_pathtracer_render(self._h, C.byref(cam), w, h, int(samples), int(depth), c, ptr(image), C.byref(opts) if opts is not None else None))
        return image

    def render_device(self, camera, width, height, samples, depth, d_sum, opts=None, stream=0):
        cam = camera if isinstance(camera, Camera) else make_camera(camera["eye"], camera["corner"], camera["dx"], camera["dy"])
        check(lib().rb_pathtracer_render_device(self._h, C.byref(cam), width, height, int(samples), int(depth), d_sum, C.byref(opts) if opts is not None else None, stream))

    def resolve_device(self, d_sum, width, height, samples, components, d_pixels, opts=None, stream=0):
        check(lib().rb_pathtracer_resolve_device(self._h, d_sum, width, height, int(samples), components, d_pixels, C.byref(opts) if opts is not None else None, stream))

    # ------------------------------------------------------------------ live rendering (PathTracerApp.cpp:57-98,127-136)
    def live_begin(self, width, height):
        check(lib().rb_pathtracer_live_begin(self._h, int(width), int(height)))
        self._live = (int(width), int(height))

    def live_pass(self, camera, samples, depth, pass_count, opts=None):
        cam = camera if isinstance(camera, Camera) else make_camera(camera["eye"], camera["corner"], camera["dx"], camera["dy"])
        check(lib().rb_pathtracer_live_pass(self._h, C.byref(cam), int(samples), int(depth), int(pass_count), C.byref(opts) if opts is not None else None))

    def live_display(self, components=4, tonemap=True, srgb=True):
        """(H, W, components) uint8 of the frame as refined so far, and the number of passes it holds."""
        w, h = self._live
        out = np.zeros((h, w, components), np.uint8)
        done = C.c_uint32(0)
        check(lib().rb_pathtracer_live_display(self._h, components, 1 if tonemap else 0, 1 if srgb else 0, ptr(out), C.byref(done)))
        return out, int(done.value)

    def multi_stats(self):
        """Slowest GPU's render time and the combine step (band copies / reduce + resolve) of the last gpus > 1 render."""
        m = _capi.MultiStats()
        check(lib().rb_pathtracer_multi_stats(self._h, C.byref(m)))
        return dict(gpus=int(m.gpus), split=int(m.split), used_nccl=bool(m.used_nccl), peer_access=bool(m.peer_access), render_ms=float(m.render_ms), combine_ms=float(m.combine_ms))

    def profile_frame(self, camera, width, height, samples, depth, opts=None):
        """One frame taken apart (bench roofline): device ms and launches per kernel class on a single stream, and the
        nodes / triangles / roots tested by the two traversal wavefronts (counting copies of the kernels)."""
        cam = camera if isinstance(camera, Camera) else make_camera(camera["eye"], camera["corner"], camera["dx"], camera["dy"])
        p = _capi.PtProfile()
        check(lib().rb_pathtracer_profile_frame(self._h, C.byref(cam), width, height, int(samples), int(depth), C.byref(opts) if opts is not None else None, C.byref(p)))
        names = ("raygen", "trace", "shade", "shadow", "accumulate")
        out = {"ms": {n: float(p.ms[i]) for i, n in enumerate(names)}, "launches": {n: int(p.launches[i]) for i, n in enumerate(names)}}
        for k in ("rays_closest", "rays_shadow", "shaded_hits", "samples", "trace_interior", "trace_tris", "trace_roots", "shadow_interior", "shadow_tris", "shadow_roots"):
            out[k] = int(getattr(p, k))
        return out

    def stats(self):
        s = RenderStats()
        check(lib().rb_pathtracer_stats(self._h, C.byref(s)))
        return dict(rays_closest=int(s.rays_closest), rays_shadow=int(s.rays_shadow), samples=int(s.samples), shaded=int(s.shaded_hits),
                    kernel_launches=int(s.kernel_launches), device_seconds=float(s.device_seconds))


def output_ldr(image, tonemap=True, srgb=True):
    """The tail of the reference's headless render (main.cpp:109-119 + Image::saveAsLDR): float image -> bytes for the PNG
    encoder, computed on the GPU."""
    image = np.ascontiguousarray(image, np.float32)
    h, w, c = image.shape
    out = np.zeros((h, w, c), np.uint8)
    check(lib().rb_output_ldr(ptr(image), w, h, c, 1 if tonemap else 0, 1 if srgb else 0, ptr(out)))
    return out


def output_hdr(image, ignore_alpha=True, flip=False):
    """The .exr branch of the headless render (Image::saveAsHDR + the EXR writer's half conversion): float image ->
    planar half layers (uint16, shape (components, H, W)) in the file's channel order (A)BGR, computed on the GPU."""
    image = np.ascontiguousarray(image, np.float32)
    h, w, c = image.shape
    out = np.zeros((c, h, w), np.uint16)
    check(lib().rb_output_hdr(ptr(image), w, h, c, 1 if ignore_alpha else 0, 1 if flip else 0, ptr(out)))
    return out


def save_scene(path, description):
    """Flattened scene -> one compressed .npz (used for golden fixtures)."""
    out = {"n_objects": len(description["objects"]), "n_lights": len(description["lights"]), "n_textures": len(description["textures"])}
    for i, t in enumerate(description["textures"]):
        out["tex%d" % i] = t
    for i, o in enumerate(description["objects"]):
        for k in ("positions", "normals", "tangents", "bitangents", "texcoords", "indices", "model"):
            out["obj%d_%s" % (i, k)] = o[k]
        out["obj%d_meta" % i] = np.array([o["use_uv"], o["material_type"], o["masked"]] + list(o["textures"]) + [-1] * (3 - len(o["textures"])), np.int64)
    for i, l in enumerate(description["lights"]):
        out["light%d" % i] = np.concatenate([[l["kind"], l["casts_shadow"]], l["intensity"], l["position"], l["direction"], [l["radius"], l["inner"], l["outer"]]]).astype(np.float64)
    bg = description["background"]
    out["bg_meta"] = np.array([bg["mode"], bg.get("sky_lut", -1)] + list(bg.get("textures", [])), np.int64)
    out["bg_color"], out["bg_sun"] = np.asarray(bg["color"], np.float32), np.asarray(bg["sun"], np.float32)
    np.savez_compressed(path, **out)


def load_scene(path):
    z = np.load(path)
    textures = [z["tex%d" % i] for i in range(int(z["n_textures"]))]
    objects = []
    for i in range(int(z["n_objects"])):
        meta = z["obj%d_meta" % i]
        o = {k: z["obj%d_%s" % (i, k)] for k in ("positions", "normals", "tangents", "bitangents", "texcoords", "indices", "model")}
        o.update(use_uv=bool(meta[0]), material_type=int(meta[1]), masked=bool(meta[2]), textures=[int(x) for x in meta[3:6]])
        objects.append(o)
    lights = []
    for i in range(int(z["n_lights"])):
        a = z["light%d" % i]
        lights.append(dict(kind=int(a[0]), casts_shadow=bool(a[1]), intensity=a[2:5].astype(np.float32), position=a[5:8].astype(np.float32),
                           direction=a[8:11].astype(np.float32), radius=float(np.float32(a[11])), inner=float(np.float32(a[12])), outer=float(np.float32(a[13]))))
    m = z["bg_meta"]
    bg = dict(mode=int(m[0]), sky_lut=int(m[1]), textures=[int(x) for x in m[2:]], color=z["bg_color"], sun=z["bg_sun"])
    return dict(objects=objects, lights=lights, background=bg, textures=textures)


def smoke():
    """Tiny render used by __graft_entry__.smoke(): a two-quad scene with one point light, checked for sanity here and
    against the oracle in the tests."""
    from . import workloads as wl
    desc = wl.tiny_scene()
    sc = Scene(desc)
    pt = PathTracer(sc)
    img = np.zeros((32, 32, 3), np.float32)
    pt.render(wl.tiny_camera(), 16, 4, img, make_opts(seed=3))
    if not (np.isfinite(img).all() and img.max() > 0.0):
        raise RuntimeError("path tracer smoke render is empty or non-finite")
    st = pt.stats()
    print("path tracer smoke ok: mean %.4f, %d closest + %d shadow rays, %d launches" % (float(img.mean()), st["rays_closest"], st["rays_shadow"], st["kernel_launches"]))
