// Entry points not implemented yet in this build step: they fail loudly (never a CPU path).
#include "rb_internal.h"
#define RB_TODO(name, ...) extern "C" rb_status name(__VA_ARGS__) { return rb_fail(RB_ERR_UNSUPPORTED, #name ": not implemented yet"); }
RB_TODO(rb_scene_create, rb_scene **)
extern "C" void rb_scene_destroy(rb_scene *) {}
RB_TODO(rb_scene_add_texture, rb_scene *, const float *, uint32_t, uint32_t, uint32_t, int32_t *)
RB_TODO(rb_scene_add_object, rb_scene *, const rb_object_desc *)
RB_TODO(rb_scene_add_light, rb_scene *, const rb_light_desc *)
RB_TODO(rb_scene_set_background, rb_scene *, uint32_t, const float *, const int32_t *, uint32_t, const float *, int32_t)
RB_TODO(rb_scene_commit, rb_scene *, rb_build_mode)
extern "C" rb_raycaster * rb_scene_raycaster(rb_scene *) { return nullptr; }
RB_TODO(rb_pathtracer_create, rb_scene *, rb_pathtracer **)
extern "C" void rb_pathtracer_destroy(rb_pathtracer *) {}
RB_TODO(rb_pathtracer_render, rb_pathtracer *, const rb_camera *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, float *, const rb_render_opts *)
RB_TODO(rb_pathtracer_render_device, rb_pathtracer *, const rb_camera *, uint32_t, uint32_t, uint32_t, uint32_t, float *, const rb_render_opts *, void *)
RB_TODO(rb_pathtracer_resolve_device, rb_pathtracer *, const float *, uint32_t, uint32_t, uint32_t, uint32_t, float *, const rb_render_opts *, void *)
RB_TODO(rb_pathtracer_stats, const rb_pathtracer *, rb_render_stats *)
RB_TODO(rb_kat_ggx_eval, const float *, const float *, const float *, const float *, const float *, size_t, float *)
RB_TODO(rb_kat_ggx_sample, const float *, const float *, const float *, const float *, const float *, size_t, float *, float *)
RB_TODO(rb_kat_image_sample, const float *, uint32_t, uint32_t, const float *, size_t, float *)
RB_TODO(rb_kat_cubemap_sample, const float *, uint32_t, uint32_t, const float *, size_t, float *)
RB_TODO(rb_kat_light_sample, const rb_light_desc *, const float *, size_t, float *)
RB_TODO(rb_kat_sky_eval, const float *, uint32_t, uint32_t, const float *, const float *, const float *, size_t, float *)
