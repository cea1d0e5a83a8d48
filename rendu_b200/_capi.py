"""ctypes binding of librendu_b200.so — one prototype per entry point of include/rendu_b200.h."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "librendu_b200.so"


class RenduError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("rendu_b200 status %d: %s" % (status, message))
        self.status = status


RB_OK, RB_ERR_INVALID, RB_ERR_NO_DEVICE, RB_ERR_CUDA, RB_ERR_UNSUPPORTED = 0, 1, 2, 3, 4

# Bit-compatible with Raycaster::Hit / Node / TriangleInfos (see include/rendu_b200.h).
HIT_DTYPE = np.dtype([("hit", "<u4"), ("dist", "<f4"), ("u", "<f4"), ("v", "<f4"), ("w", "<f4"), ("_pad", "<u4"),
                      ("localId", "<u8"), ("meshId", "<u8"), ("internalId", "<u8")])
NODE_DTYPE = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("left", "<u8"), ("right", "<u8"), ("leaf", "u1"), ("_pad", "u1", 7)])
TRI_DTYPE = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("v0", "<u8"), ("v1", "<u8"), ("v2", "<u8"), ("localId", "<u8"),
                      ("meshId", "<u4"), ("_pad", "<u4")])
assert HIT_DTYPE.itemsize == 48 and NODE_DTYPE.itemsize == 48 and TRI_DTYPE.itemsize == 64


class ObjectDesc(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("normals", C.c_void_p), ("tangents", C.c_void_p), ("bitangents", C.c_void_p),
                ("texcoords", C.c_void_p), ("nverts", C.c_size_t), ("indices", C.c_void_p), ("nindices", C.c_size_t),
                ("model", C.c_float * 16), ("tex_albedo", C.c_int32), ("tex_normal", C.c_int32), ("tex_rmao", C.c_int32),
                ("emissive", C.c_uint32), ("masked", C.c_uint32), ("use_texcoords", C.c_uint32)]


class LightDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("casts_shadow", C.c_uint32), ("intensity", C.c_float * 3), ("position", C.c_float * 3),
                ("direction", C.c_float * 3), ("radius", C.c_float), ("inner_angle", C.c_float), ("outer_angle", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("corner", C.c_float * 3), ("dx", C.c_float * 3), ("dy", C.c_float * 3)]


class RenderOpts(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("tile_rows", C.c_uint32), ("tile_index", C.c_uint32), ("tile_count", C.c_uint32),
                ("sample_index", C.c_uint32), ("sample_count", C.c_uint32), ("trace_mode", C.c_uint32),
                ("max_paths_in_flight", C.c_uint32), ("accumulate", C.c_uint32), ("gpus", C.c_uint32), ("split", C.c_uint32),
                ("_reserved", C.c_uint32 * 2)]


class RenderStats(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("samples", C.c_uint64), ("shaded_hits", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("device_seconds", C.c_double)]


class MultiStats(C.Structure):
    _fields_ = [("gpus", C.c_uint32), ("split", C.c_uint32), ("used_nccl", C.c_uint32), ("peer_access", C.c_uint32),
                ("render_ms", C.c_float), ("combine_ms", C.c_float)]


class PtProfile(C.Structure):
    _fields_ = [("ms", C.c_double * 5), ("launches", C.c_uint64 * 5), ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("shaded_hits", C.c_uint64), ("samples", C.c_uint64), ("trace_interior", C.c_uint64), ("trace_tris", C.c_uint64),
                ("trace_roots", C.c_uint64), ("shadow_interior", C.c_uint64), ("shadow_tris", C.c_uint64), ("shadow_roots", C.c_uint64)]


_lib = None


def lib_path():
    """In-tree library; RENDU_B200_LIB names another build of the same ABI (kernel experiments)."""
    return os.environ.get("RENDU_B200_LIB") or os.path.join(_HERE, _LIB_NAME)


def lib():
    """Load the CUDA library. Fails loudly if it has not been built (python __graft_entry__.py / make -C rendu_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RenduError(-1, "%s not built: run `make -C rendu_b200/csrc` (there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp, sz, f32, u32, i32, u64 = C.c_void_p, C.c_size_t, C.c_float, C.c_uint32, C.c_int32, C.c_uint64
    protos = {
        "rb_last_error": (C.c_char_p, []),
        "rb_version": (C.c_char_p, []),
        "rb_device_count": (C.c_int, []),
        "rb_set_device": (C.c_int, [C.c_int]),
        "rb_raycaster_create": (C.c_int, [C.POINTER(vp)]),
        "rb_raycaster_destroy": (None, [vp]),
        "rb_raycaster_add_mesh": (C.c_int, [vp, vp, sz, vp, sz, vp]),
        "rb_raycaster_build": (C.c_int, [vp, C.c_int]),
        "rb_raycaster_upload_bvh": (C.c_int, [vp, vp, sz, vp, sz, vp, sz, u32]),
        "rb_raycaster_sizes": (C.c_int, [vp, vp]),
        "rb_raycaster_download_bvh": (C.c_int, [vp, vp, vp, vp]),
        "rb_raycaster_refit": (C.c_int, [vp, vp, sz]),
        "rb_raycaster_update_mesh": (C.c_int, [vp, u32, vp, sz, vp]),
        "rb_raycaster_set_trace_mode": (C.c_int, [vp, C.c_int]),
        "rb_raycaster_wide_build": (C.c_int, [vp, vp]),
        "rb_obj_parse": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, vp]),
        "rb_mesh_destroy": (None, [vp]),
        "rb_mesh_sizes": (C.c_int, [vp, vp]),
        "rb_mesh_read": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "rb_mesh_compute_normals": (C.c_int, [vp]),
        "rb_raycaster_wide_download": (C.c_int, [vp, vp, vp, vp]),
        "rb_raycaster_wide_download_by_octant": (C.c_int, [vp, vp]),
        "rb_raycaster_build_times": (C.c_int, [vp, vp]),
        "rb_raycaster_tree_cost": (C.c_int, [vp, vp]),
        "rb_raycaster_intersect": (C.c_int, [vp, vp, vp, f32, f32, sz, vp]),
        "rb_raycaster_intersect_any": (C.c_int, [vp, vp, vp, f32, f32, sz, vp]),
        "rb_raycaster_visible": (C.c_int, [vp, vp, vp, sz, vp]),
        "rb_raycaster_intersect_async": (C.c_int, [vp, vp, vp, f32, f32, sz, vp]),
        "rb_raycaster_intersect_any_async": (C.c_int, [vp, vp, vp, f32, f32, sz, vp]),
        "rb_raycaster_sync": (C.c_int, [vp]),
        "rb_raycaster_intersect_primary": (C.c_int, [vp, C.POINTER(Camera), u32, u32, f32, f32, vp]),
        "rb_raycaster_intersect_any_primary": (C.c_int, [vp, C.POINTER(Camera), u32, u32, f32, f32, vp]),
        "rb_raycaster_ray_levels": (C.c_int, [vp, vp, vp, f32, f32, u32, vp, vp, vp, vp]),
        "rb_raycaster_intersect_device": (C.c_int, [vp, vp, vp, f32, f32, sz, vp, vp]),
        "rb_raycaster_intersect_any_device": (C.c_int, [vp, vp, vp, f32, f32, sz, vp, vp]),
        "rb_raycaster_count_work_device": (C.c_int, [vp, vp, vp, f32, f32, sz, C.c_int, vp, vp]),
        "rb_raycaster_launch_count": (u64, [vp]),
        "rb_scene_create": (C.c_int, [C.POINTER(vp)]),
        "rb_scene_destroy": (None, [vp]),
        "rb_scene_add_texture": (C.c_int, [vp, vp, u32, u32, u32, C.POINTER(i32)]),
        "rb_scene_add_object": (C.c_int, [vp, C.POINTER(ObjectDesc)]),
        "rb_scene_add_light": (C.c_int, [vp, C.POINTER(LightDesc)]),
        "rb_scene_set_background": (C.c_int, [vp, u32, vp, vp, u32, vp, i32]),
        "rb_scene_commit": (C.c_int, [vp, C.c_int]),
        "rb_scene_clear_lights": (C.c_int, [vp]),
        "rb_scene_sync": (C.c_int, [vp]),
        "rb_scene_set_object_transform": (C.c_int, [vp, u32, vp]),
        "rb_scene_raycaster": (vp, [vp]),
        "rb_pathtracer_create": (C.c_int, [vp, C.POINTER(vp)]),
        "rb_pathtracer_destroy": (None, [vp]),
        "rb_pathtracer_render": (C.c_int, [vp, C.POINTER(Camera), u32, u32, u32, u32, u32, vp, C.POINTER(RenderOpts)]),
        "rb_pathtracer_render_device": (C.c_int, [vp, C.POINTER(Camera), u32, u32, u32, u32, vp, C.POINTER(RenderOpts), vp]),
        "rb_pathtracer_resolve_device": (C.c_int, [vp, vp, u32, u32, u32, u32, vp, C.POINTER(RenderOpts), vp]),
        "rb_pathtracer_stats": (C.c_int, [vp, C.POINTER(RenderStats)]),
        "rb_pathtracer_multi_stats": (C.c_int, [vp, C.POINTER(MultiStats)]),
        "rb_pathtracer_live_begin": (C.c_int, [vp, u32, u32]),
        "rb_pathtracer_live_pass": (C.c_int, [vp, C.POINTER(Camera), u32, u32, u32, C.POINTER(RenderOpts)]),
        "rb_pathtracer_live_display": (C.c_int, [vp, u32, C.c_int, C.c_int, vp, C.POINTER(u32)]),
        "rb_pathtracer_profile_frame": (C.c_int, [vp, C.POINTER(Camera), u32, u32, u32, u32, C.POINTER(RenderOpts), C.POINTER(PtProfile)]),
        "rb_output_ldr": (C.c_int, [vp, u32, u32, u32, C.c_int, C.c_int, vp]),
        "rb_output_ldr_device": (C.c_int, [vp, u32, u32, u32, C.c_int, C.c_int, vp, vp]),
        "rb_output_hdr": (C.c_int, [vp, u32, u32, u32, C.c_int, C.c_int, vp]),
        "rb_output_hdr_device": (C.c_int, [vp, u32, u32, u32, C.c_int, C.c_int, vp, vp]),
        "rb_kat_ggx_eval": (C.c_int, [vp, vp, vp, vp, vp, sz, vp]),
        "rb_kat_ggx_sample": (C.c_int, [vp, vp, vp, vp, vp, sz, vp, vp]),
        "rb_kat_image_sample": (C.c_int, [vp, u32, u32, vp, sz, vp]),
        "rb_kat_cubemap_sample": (C.c_int, [vp, u32, u32, vp, sz, vp]),
        "rb_kat_light_sample": (C.c_int, [C.POINTER(LightDesc), vp, sz, vp]),
        "rb_kat_sky_eval": (C.c_int, [vp, u32, u32, vp, vp, vp, sz, vp]),
        "rb_kat_box_test": (C.c_int, [vp, vp, vp, f32, f32, sz, vp]),
        "rb_kat_triangle_test": (C.c_int, [vp, vp, vp, f32, f32, sz, vp]),
        "rb_kat_div": (C.c_int, [vp, vp, sz, vp]),
    }
    missing = []
    for name, (res, args) in protos.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:
        raise RenduError(-1, "librendu_b200.so lacks symbols declared in include/rendu_b200.h: %s" % ", ".join(missing))
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by tests from the header


def check(status):
    if status != RB_OK:
        raise RenduError(status, lib().rb_last_error().decode("utf-8", "replace"))


def device_count():
    return int(lib().rb_device_count())


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
