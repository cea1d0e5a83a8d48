/*
 * rendu_b200.h — C ABI of the B200-native replacement for Rendu's offline path-tracer hot path.
 *
 * This is the drop-in boundary: the entry points below are exactly what a C++ `Raycaster` / `PathTracer`
 * (reference: src/engine/raycaster/Raycaster.hpp, src/apps/pathtracer/PathTracer.hpp) forwards to once its CPU
 * bodies are removed. Plain pointers and sizes only; no torch, glm or STL types cross this line.
 *
 * Conventions
 *  - Every function returns rb_status (0 = RB_OK). On failure rb_last_error() gives a thread-local message.
 *    The reference has no error channel on this path (it logs and returns, PathTracer.cpp:162-173); the C++
 *    wrappers translate a non-zero status into Log::Error() and keep their void signatures.
 *  - "host" entry points take host pointers and are blocking: they copy in, launch, copy out, and never keep a
 *    caller pointer. "_device" entry points take device pointers plus a cudaStream_t (passed as void*) and only
 *    enqueue work on that stream.
 *  - There is no CPU fallback. Without a CUDA device every compute entry point returns RB_ERR_NO_DEVICE.
 *  - Query functions are safe to call concurrently from several host threads on one handle (the reference calls
 *    Raycaster::intersects from many threads, PathTracer.cpp:186,204); build / add / upload are not.
 *  - Matrices are column-major 4x4 (glm::mat4 memory order). Vectors are tightly packed float[3].
 */
#ifndef RENDU_B200_H
#define RENDU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RB_API __declspec(dllexport)
#else
#define RB_API __attribute__((visibility("default")))
#endif

typedef int rb_status;
enum {
	RB_OK				= 0,
	RB_ERR_INVALID		= 1, /* bad argument / call order */
	RB_ERR_NO_DEVICE	= 2, /* no CUDA device: this library never computes on the CPU */
	RB_ERR_CUDA			= 3, /* CUDA runtime failure, message in rb_last_error() */
	RB_ERR_UNSUPPORTED	= 4
};

/* ------------------------------------------------------------------------------------------------ records
 * The three records below are bit-compatible with the reference's own structs so that a C++ wrapper can pass
 * `_hierarchy.data()` / `_triangles.data()` and receive `Raycaster::Hit` arrays without repacking. */

/* Raycaster::Hit (Raycaster.hpp:16-42): {bool hit; float dist,u,v,w; ulong localId, meshId; ulong internalId}.
 * Miss: hit=0, dist=FLT_MAX, everything else 0 (Raycaster.cpp:6-8). Hit: w = 1-u-v (Raycaster.cpp:10-12).
 * internalId (always 0 in the reference) carries the index of the triangle in the raycaster's own order. */
typedef struct rb_hit {
	uint32_t hit; /* 0 or 1; byte 0 is the reference's bool */
	float dist, u, v, w;
	uint32_t _pad;
	uint64_t localId;
	uint64_t meshId;
	uint64_t internalId;
} rb_hit; /* 48 bytes */

/* Raycaster::Node (Raycaster.hpp:127-132): leaf ? (left = first triangle, right = count) : child node ids. */
typedef struct rb_node {
	float bmin[3], bmax[3];
	uint64_t left, right;
	uint8_t leaf;
	uint8_t _pad[7];
} rb_node; /* 48 bytes */

/* Raycaster::TriangleInfos (Raycaster.hpp:117-124). v0..v2 index the merged world-space vertex array. */
typedef struct rb_tri {
	float bmin[3], bmax[3];
	uint64_t v0, v1, v2;
	uint64_t localId;
	uint32_t meshId;
	uint32_t _pad;
} rb_tri; /* 64 bytes */

/* ------------------------------------------------------------------------------------------------ library */

RB_API const char * rb_last_error(void);
RB_API const char * rb_version(void);
/* Number of visible CUDA devices (0 when none, never an error). */
RB_API int rb_device_count(void);
/* Select the CUDA device used by handles created afterwards on this host thread (default: current device).
 * A handle stays bound to the device its hierarchy / scene was uploaded to; every later call on it makes that device
 * current for the calling thread (as cudaSetDevice would) — device pointers and streams passed to the *_device
 * entry points must belong to it. One process per GPU (the multi-GPU model of this library) never notices. */
RB_API rb_status rb_set_device(int device);

/* ------------------------------------------------------------------------------------------------ raycaster
 * Replaces class Raycaster (src/engine/raycaster/Raycaster.{hpp,cpp}). */

typedef struct rb_raycaster rb_raycaster;

typedef enum rb_build_mode {
	/* Raycaster::updateHierarchy (Raycaster.cpp:54-142) restated on the host: same split rules, same libstdc++
	 * partition / nth_element calls, hence the same node numbering and triangle order as the reference; then
	 * uploaded. This is the mode in which hits are bit-identical to the reference. */
	RB_BUILD_REFERENCE = 0,
	/* GPU LBVH: Morton codes on triangle-box centroids, radix sort, Karras hierarchy, bottom-up AABB fit.
	 * One tree over all meshes. Same closest hits except exact-distance ties / box-edge grazing (documented). */
	RB_BUILD_LBVH = 1,
	/* Host build in the reference's framework (one root per mesh, same leaf rule, same node records) with the split of
	 * every set chosen by the binned surface-area heuristic instead of the mean of the longest axis: a better tree for
	 * traversal, not the reference's. Same closest hits except exact-distance ties / box-edge grazing. */
	RB_BUILD_SAH = 2
} rb_build_mode;

typedef enum rb_trace_mode {
	/* Reference traversal order (LIFO, right child first, roots in reverse) and the reference's arithmetic,
	 * operation for operation, no FMA contraction: results are bit-identical to Raycaster::intersects. */
	RB_TRACE_EXACT = 0,
	/* Near-child-first ordered traversal on the same tree, same triangle test. Same hit except where the
	 * reference's non-watertight box test makes the result depend on visiting order. Any-hit queries
	 * (intersect_any, visible) always use this order: they never shrink maxi, so their boolean cannot depend on it. */
	RB_TRACE_ORDERED = 1,
	/* Experiment: the reference-order tree collapsed into eight-wide nodes with 8-bit child boxes rounded outwards
	 * (rb_raycaster_wide_build), children visited in the reference's order. Enters a superset of the reference's
	 * leaves in the same relative order: same closest hit (also on exact ties) except where a triangle is accepted
	 * whose own boxes the reference's rounded slab test rejects. Raycaster queries only (not the path tracer). */
	RB_TRACE_WIDE = 2
} rb_trace_mode;

RB_API rb_status rb_raycaster_create(rb_raycaster ** out);
RB_API void rb_raycaster_destroy(rb_raycaster * rc);

/* Raycaster::addMesh (Raycaster.cpp:14-52). positions: nverts*3 floats, indices: nindices (multiple of 3) u32,
 * model: 16 floats or NULL for identity. Vertices are transformed once, as `vec3(model * vec4(v,1))`, unless the
 * matrix is exactly identity. The mesh index (meshId of its triangles) is the number of earlier calls. */
RB_API rb_status rb_raycaster_add_mesh(rb_raycaster * rc, const float * positions, size_t nverts, const uint32_t * indices, size_t nindices, const float * model);

/* Raycaster::updateHierarchy (Raycaster.cpp:54-142). */
RB_API rb_status rb_raycaster_build(rb_raycaster * rc, rb_build_mode mode);

/* Upload path for a hierarchy built elsewhere (e.g. by the reference itself): arrays exactly as the reference
 * holds them in _hierarchy / _triangles / _vertices / _meshCount (Raycaster.hpp:152-156). Replaces any geometry
 * previously added. */
RB_API rb_status rb_raycaster_upload_bvh(rb_raycaster * rc, const rb_node * nodes, size_t nnodes, const rb_tri * tris, size_t ntris, const float * vertices, size_t nverts, uint32_t mesh_count);

/* Sizes of the hierarchy in the reference format: out[0]=nodes, out[1]=triangles, out[2]=vertices, out[3]=roots (one per
 * mesh for RB_BUILD_REFERENCE, a single one for RB_BUILD_LBVH; pass it as mesh_count to rb_raycaster_upload_bvh),
 * out[4]=max depth. */
RB_API rb_status rb_raycaster_sizes(const rb_raycaster * rc, uint64_t out[5]);

/* Read the hierarchy back in the reference's format (any pointer may be NULL). After an LBVH build this
 * downloads the device tree; it is what RaycasterVisualisation (RaycasterVisualisation.cpp:15-30) needs. */
RB_API rb_status rb_raycaster_download_bvh(const rb_raycaster * rc, rb_node * nodes, rb_tri * tris, float * vertices);

/* Re-fit all boxes after the vertices moved (same topology): positions replace the merged world-space vertex
 * array (nverts must match). Device-side bottom-up refit; the tree order is kept. world_positions == NULL (nverts 0)
 * fits the vertices the handle already holds, i.e. after one or more rb_raycaster_update_mesh calls. */
RB_API rb_status rb_raycaster_refit(rb_raycaster * rc, const float * world_positions, size_t nverts);

/* Incremental geometry (no reference counterpart: Raycaster bakes `model` at addMesh, Raycaster.cpp:22-28, and an
 * animated Object never reaches it). Mesh `mesh` (addMesh order) gets new model-space positions and / or a new model
 * matrix with the same topology; its world-space vertices are recomputed with addMesh's own arithmetic and only that
 * range is uploaded. Boxes are stale until the next rb_raycaster_refit(rc, NULL, 0). Needs a hierarchy built from
 * rb_raycaster_add_mesh calls (not rb_raycaster_upload_bvh, which carries no per-mesh vertex ranges). */
RB_API rb_status rb_raycaster_update_mesh(rb_raycaster * rc, uint32_t mesh, const float * positions, size_t nverts, const float * model);

/* When to stop re-fitting: out[0] = surface-area cost of the hierarchy as built (sum over interior nodes of the areas of
 * their two child boxes — proportional to the number of child boxes a random line enters), out[1] = the same after the last
 * refit. A refit keeps the topology, so out[1] / out[0] grows as objects move apart from where the tree put them;
 * rb_scene_sync rebuilds instead of re-fitting once the ratio passes 2 (RB_REBUILD_RATIO overrides). */
RB_API rb_status rb_raycaster_tree_cost(rb_raycaster * rc, double out[2]);

/* Device time (CUDA events, milliseconds) of the last hierarchy build (out[0]; 0 for RB_BUILD_REFERENCE, which builds on
 * the host) and of the last refit (out[1]), excluding host packing and host-device copies. */
RB_API rb_status rb_raycaster_build_times(const rb_raycaster * rc, double out[2]);

RB_API rb_status rb_raycaster_set_trace_mode(rb_raycaster * rc, rb_trace_mode mode);

/* Collapse the host-built reference-order hierarchy (Raycaster.cpp:58-142) into the eight-wide quantised layout of
 * RB_TRACE_WIDE and, when the hierarchy is on a device, upload it. The collapse itself needs no GPU. out (may be NULL):
 * out[0]=wide nodes, out[1]=triangles, out[2]=roots, out[3]=depth. Fails for LBVH-built or re-fitted hierarchies and for
 * leaves of more than three triangles. A later rebuild or refit discards the wide copy. */
RB_API rb_status rb_raycaster_wide_build(rb_raycaster * rc, uint64_t out[4]);
/* Host copy of the wide hierarchy: nodes = out[0]*20 words, tris = out[1]*12 floats (word 11 = triangle index in hierarchy
 * order), root_nodes = out[2] ints (wide node of each mesh, -1 for an empty mesh). Any pointer may be NULL. */
RB_API rb_status rb_raycaster_wide_download(const rb_raycaster * rc, uint32_t * nodes, float * tris, int32_t * root_nodes);
/* The second copy of the wide nodes (out[0]*20 words), children placed in slots by direction for front-to-back visiting
 * where the reference's order is not promised (environment RB_WIDE_OCTANT=1: closest-hit queries of RB_TRACE_WIDE; same hits
 * except exact-distance ties). Same indices, offsets and boxes as the first copy. */
RB_API rb_status rb_raycaster_wide_download_by_octant(const rb_raycaster * rc, uint32_t * nodes);

/* Raycaster::intersects (Raycaster.cpp:144-185), batched: origins/directions are nrays*3 floats (directions need
 * not be normalised, Intersection.cpp:3-5). Reference defaults: mini = 1e-4f, maxi = 1e8f (Raycaster.hpp:65). */
RB_API rb_status rb_raycaster_intersect(const rb_raycaster * rc, const float * origins, const float * directions, float mini, float maxi, size_t nrays, rb_hit * out);

/* Raycaster::intersectsAny (Raycaster.cpp:187-224), batched: out[i] = 1 if any triangle is hit. */
RB_API rb_status rb_raycaster_intersect_any(const rb_raycaster * rc, const float * origins, const float * directions, float mini, float maxi, size_t nrays, uint8_t * out);

/* Raycaster::visible (Raycaster.cpp:226-230), batched: out[i] = 1 if the segment p0[i]..p1[i] is free. */
RB_API rb_status rb_raycaster_visible(const rb_raycaster * rc, const float * p0, const float * p1, size_t n, uint8_t * out);

/* RaycasterVisualisation::getRayLevels (src/engine/raycaster/RaycasterVisualisation.cpp:41-77) on the device tree: one
 * closest-hit query in the reference's order that also reports every node it visited, in visiting order — its box
 * (6 floats: min, max) and its depth (roots are depth 0) — for BVH-level displays and click-to-cast ray debugging
 * (BVHRenderer::castRay). At most `capacity` records are written; *count receives the number of visited nodes. Works
 * on both tree kinds (for an LBVH tree the levels are those of that tree). hit = what rb_raycaster_intersect returns. */
RB_API rb_status rb_raycaster_ray_levels(const rb_raycaster * h, const float * origin, const float * direction, float mini, float maxi, uint32_t capacity,
										 float * boxes6, uint32_t * depths, uint32_t * count, rb_hit * hit);

/* What render() derives from the Camera on the host (PathTracer.cpp:177,197; Camera.cpp:42-49). */
typedef struct rb_camera {
	float eye[3];	 /* Camera::position() */
	float corner[3]; /* Camera::pixelShifts */
	float dx[3];
	float dy[3];
} rb_camera;

/* Primary rays of a pinhole camera, generated on the device (an opt-in wire format of the batch API: nothing but the
 * camera goes up the host link, where rb_raycaster_intersect uploads 24 bytes per ray). Ray of pixel (x, y), row-major,
 * y * width + x: from `eye` through the centre of the pixel with the reference's arithmetic (PathTracer.cpp:193-198:
 * ndc = (x + 0.5, y + 0.5) / (width, height); worldPos = corner + ndc.x * dx + ndc.y * dy; direction = worldPos - eye,
 * normalised by Ray::Ray) — bit for bit the rays a host loop with the same formula would pass to rb_raycaster_intersect,
 * hence the same Raycaster::Hit records. out holds width * height records / flags. */
RB_API rb_status rb_raycaster_intersect_primary(const rb_raycaster * rc, const rb_camera * camera, uint32_t width, uint32_t height, float mini, float maxi, rb_hit * out);
RB_API rb_status rb_raycaster_intersect_any_primary(const rb_raycaster * rc, const rb_camera * camera, uint32_t width, uint32_t height, float mini, float maxi, uint8_t * out);

/* Asynchronous batch queries: as rb_raycaster_intersect / _intersect_any, but the call returns as soon as its chunks are
 * queued on the calling thread's ring; the results are complete after rb_raycaster_sync called from the same thread.
 * Consecutive asynchronous calls of a thread queue up in call order and overlap their uploads, traversals and downloads;
 * issue the download-heavy closest-hit queries first. The host buffers must stay valid until the sync, and should be
 * page-locked (otherwise the copies are synchronous and nothing overlaps). */
RB_API rb_status rb_raycaster_intersect_async(const rb_raycaster * rc, const float * origins, const float * directions, float mini, float maxi, size_t nrays, rb_hit * out);
RB_API rb_status rb_raycaster_intersect_any_async(const rb_raycaster * rc, const float * origins, const float * directions, float mini, float maxi, size_t nrays, uint8_t * out);
RB_API rb_status rb_raycaster_sync(const rb_raycaster * rc);

/* Device-pointer variants: enqueue on `stream` (a cudaStream_t), no host synchronisation. */
RB_API rb_status rb_raycaster_intersect_device(const rb_raycaster * rc, const float * d_origins, const float * d_directions, float mini, float maxi, size_t nrays, rb_hit * d_out, void * stream);
RB_API rb_status rb_raycaster_intersect_any_device(const rb_raycaster * rc, const float * d_origins, const float * d_directions, float mini, float maxi, size_t nrays, uint8_t * d_out, void * stream);

/* Work counters of a batch, measured by an instrumented copy of the traversal kernel (not the timed one):
 * counters[0] = interior nodes visited (both children box-tested, 64 B fetched each),
 * counters[1] = triangle tests (48 B fetched each), counters[2] = root boxes tested, counters[3] = max stack depth.
 * These define the algorithmic bytes per ray used for the roofline. any_hit selects the loop. */
RB_API rb_status rb_raycaster_count_work_device(const rb_raycaster * rc, const float * d_origins, const float * d_directions, float mini, float maxi, size_t nrays, int any_hit, uint64_t counters[4], void * stream);

/* Kernel launches issued by this handle so far (for bench.py's gpu_launches). */
RB_API uint64_t rb_raycaster_launch_count(const rb_raycaster * rc);

/* ------------------------------------------------------------------------------------------------ scene
 * Read-only data PathTracer::render touches: Scene::objects / lights / background, Object::mesh /
 * model / material / useTexCoords, Material::textures / type / masked, Mesh attribute arrays. Flattened by the
 * caller into plain arrays; copied to the device at rb_scene_commit. */

typedef struct rb_scene rb_scene;

typedef enum rb_light_kind { RB_LIGHT_POINT = 0, RB_LIGHT_SPOT = 1, RB_LIGHT_DIRECTIONAL = 2 } rb_light_kind;
/* Scene::Background (src/engine/scene/Scene.hpp): COLOR, IMAGE, SKYBOX, ATMOSPHERE. */
typedef enum rb_background_mode { RB_BG_COLOR = 0, RB_BG_IMAGE = 1, RB_BG_SKYBOX = 2, RB_BG_ATMOSPHERE = 3 } rb_background_mode;

typedef struct rb_object_desc {
	const float * positions;  /* nverts*3, model space (Mesh::positions) */
	const float * normals;	  /* nverts*3 */
	const float * tangents;	  /* nverts*3 */
	const float * bitangents; /* nverts*3 */
	const float * texcoords;  /* nverts*2 or NULL */
	size_t nverts;
	const uint32_t * indices; /* Mesh::indices */
	size_t nindices;
	float model[16];		  /* Object::model() */
	int32_t tex_albedo;		  /* Material::textures()[0]: albedo / emissive colour + alpha */
	int32_t tex_normal;		  /* [1]: normal map */
	int32_t tex_rmao;		  /* [2]: roughness, metalness, ao */
	uint32_t emissive;		  /* Material::type() == Emissive */
	uint32_t masked;		  /* Material::masked() */
	uint32_t use_texcoords;	  /* Object::useTexCoords() */
} rb_object_desc;

typedef struct rb_light_desc {
	uint32_t kind;		   /* rb_light_kind */
	uint32_t casts_shadow; /* Light::castsShadow() */
	float intensity[3];	   /* Light::intensity() */
	float position[3];	   /* point, spot */
	float direction[3];	   /* spot, directional (normalised as the reference stores it) */
	float radius;		   /* point, spot */
	float inner_angle, outer_angle; /* spot: SpotLight::angles() (half-angles, radians) */
} rb_light_desc;

RB_API rb_status rb_scene_create(rb_scene ** out);
RB_API void rb_scene_destroy(rb_scene * sc);
/* Image as the reference holds it on the CPU: float pixels, row-major, `components` (3 or 4) per texel. The texel pool
 * goes to the device at rb_scene_commit: a texture (or object) added to a committed scene un-commits it, and rendering,
 * rb_scene_sync and rb_pathtracer_create fail with RB_ERR_INVALID until rb_scene_commit runs again. */
RB_API rb_status rb_scene_add_texture(rb_scene * sc, const float * pixels, uint32_t width, uint32_t height, uint32_t components, int32_t * out_id);
/* Objects must be added in scene->objects order (PathTracer.cpp:13-18): the call index is the meshId. */
RB_API rb_status rb_scene_add_object(rb_scene * sc, const rb_object_desc * desc);
RB_API rb_status rb_scene_add_light(rb_scene * sc, const rb_light_desc * desc);
/* color: background Material::parameters()[0] (COLOR). textures: 1 id (IMAGE) or 6 ids px,nx,py,ny,pz,nz (SKYBOX).
 * sun_direction: Sky::direction() (ATMOSPHERE). sky_lut: id of the "scattering-precomputed" table or -1. */
RB_API rb_status rb_scene_set_background(rb_scene * sc, uint32_t mode, const float color[3], const int32_t * textures, uint32_t ntextures, const float sun_direction[3], int32_t sky_lut);
/* Upload everything and build the raycaster (addMesh per object + updateHierarchy, PathTracer.cpp:11-21). */
RB_API rb_status rb_scene_commit(rb_scene * sc, rb_build_mode mode);
/* Lights and background may be edited after the commit — the reference reads them through the Scene at render time
 * (PathTracer.cpp:23-55, :232-255), so whatever its caller changed between two render() calls shows in the next frame:
 * rb_scene_clear_lights + rb_scene_add_light and / or rb_scene_set_background, then rb_scene_sync (the next render
 * also does it). Only the two small tables are re-uploaded; geometry, textures and the hierarchy stay. */
RB_API rb_status rb_scene_clear_lights(rb_scene * sc);
RB_API rb_status rb_scene_sync(rb_scene * sc);
/* New model matrix (16 floats, column-major) for object `object` (rb_scene_add_object order) — what Object::update
 * leaves in Object::model() for an animated object (Object.cpp, Animation.cpp). Before the commit it simply replaces
 * the matrix; on a committed scene the next rb_scene_sync (or render) re-transforms that object's vertices, refits
 * the hierarchy on the device (same topology, rb_raycaster_update_mesh + rb_raycaster_refit) and rewrites the object's
 * normal matrix; textures, attributes and the other objects are untouched. */
RB_API rb_status rb_scene_set_object_transform(rb_scene * sc, uint32_t object, const float model[16]);
/* The scene's raycaster (owned by the scene) — PathTracer::raycaster(). Queries, rb_raycaster_wide_build and refits are
 * fine; a rb_raycaster_build on it reorders the triangles under the per-triangle shading records gathered at commit, and
 * the next render then fails with RB_ERR_INVALID until the scene is committed again. */
RB_API rb_raycaster * rb_scene_raycaster(rb_scene * sc);

/* ------------------------------------------------------------------------------------------------ path tracer
 * Replaces PathTracer::render (src/apps/pathtracer/PathTracer.cpp:159-299). */

typedef struct rb_pathtracer rb_pathtracer;

typedef struct rb_render_opts {
	uint64_t seed;		 /* counter-based RNG key; same seed => same image, whatever the GPU count */
	/* Tile split: only rows y with (y / tile_rows) % tile_count == tile_index are rendered (others untouched).
	 * tile_count <= 1 renders everything. */
	uint32_t tile_rows, tile_index, tile_count;
	/* Sample split: only stratification cells sid with sid % sample_count == sample_index are traced; the sum is
	 * still normalised by the full `samples`, so per-GPU images add up to the full estimate. */
	uint32_t sample_index, sample_count;
	uint32_t trace_mode; /* rb_trace_mode used for primary, bounce and shadow rays (RB_TRACE_WIDE, experiment: closest-hit
	                      * wavefront — and the shadow wavefront of scenes without masked materials — on the wide tree, after
	                      * rb_raycaster_wide_build on rb_scene_raycaster()) */
	uint32_t max_paths_in_flight; /* paths per wavefront batch; 0 = default (32 Mi; two batches are in flight: 176 B x 2 x 32 Mi = 11.8 GB of the B200's 180 GB) */
	/* Progressive rendering: with accumulate != 0 rb_pathtracer_render_device ADDS this call's samples
	 * to d_sum instead of overwriting it, so a viewer can trace one slice of the stratification grid per frame
	 * (sample_index = frame % sample_count) and resolve what it has so far with samples = cells traced so far. */
	uint32_t accumulate;
	/* All GPUs of the box from one call (rb_pathtracer_render only; replaces System::forParallel, System.hpp:72-104,
	 * PathTracer.cpp:186): gpus > 1 replicates the scene on that many devices at the first such render (one host thread
	 * per GPU) and splits every frame among them — RB_SPLIT_TILE: bands of tile_rows rows (0 = 2) dealt round-robin,
	 * gathered on the path tracer's own device with strided peer copies, bit-identical to one GPU; RB_SPLIT_SAMPLE:
	 * stratification cells dealt round-robin, per-pixel sums combined with one ncclReduce over NVLink (NCCL is loaded at
	 * run time; RB_SAMPLE_REDUCE=p2p sums over peer memory inside the resolve kernel instead), equal to one GPU up to
	 * fp32 summation order. tile_count / sample_count must stay 0 then. 0 or 1: this device only. */
	uint32_t gpus;
	uint32_t split; /* rb_split */
	uint32_t _reserved[2];
} rb_render_opts;

typedef enum rb_split { RB_SPLIT_TILE = 0, RB_SPLIT_SAMPLE = 1 } rb_split;

/* Timing of the last multi-GPU render of a path tracer: the slowest GPU's share (CUDA events on its stream) and the
 * combine step on the root device (band copies or reduce, plus the resolve). */
typedef struct rb_multi_stats {
	uint32_t gpus, split, used_nccl, peer_access;
	float render_ms, combine_ms;
} rb_multi_stats;

typedef struct rb_render_stats {
	uint64_t rays_closest; /* primary + bounce queries (Raycaster::intersects from render) */
	uint64_t rays_shadow;  /* queries issued by checkVisibility (PathTracer.cpp:128) */
	uint64_t samples;	   /* pixel samples traced */
	uint64_t shaded_hits;
	uint64_t kernel_launches;
	double device_seconds; /* CUDA-event time of the device work of the last render */
} rb_render_stats;

RB_API rb_status rb_pathtracer_create(rb_scene * scene, rb_pathtracer ** out);
RB_API void rb_pathtracer_destroy(rb_pathtracer * pt);

/* PathTracer::render: `pixels` is the caller's Image::pixels (width*height*components floats, components >= 3).
 * As in the reference the result is accumulated into it (`+=` then `/= samples`, PathTracer.cpp:288-292), so the
 * caller zero-fills; samples is rounded to the nearest power of two (PathTracer.cpp:169-173). opts may be NULL. */
RB_API rb_status rb_pathtracer_render(rb_pathtracer * pt, const rb_camera * camera, uint32_t width, uint32_t height, uint32_t samples, uint32_t depth, uint32_t components, float * pixels, const rb_render_opts * opts);

/* Device variant for multi-GPU plumbing: writes per-pixel SUMS of clamped sample radiance (not yet divided) into
 * d_sum (width*height*3 floats, overwritten) on `stream`; rb_pathtracer_resolve_device then applies the reference's
 * `pixels = (pixels + sum) / samples` to a device image. Between the two a caller may reduce d_sum across ranks. */
RB_API rb_status rb_pathtracer_render_device(rb_pathtracer * pt, const rb_camera * camera, uint32_t width, uint32_t height, uint32_t samples, uint32_t depth, float * d_sum, const rb_render_opts * opts, void * stream);
RB_API rb_status rb_pathtracer_resolve_device(rb_pathtracer * pt, const float * d_sum, uint32_t width, uint32_t height, uint32_t samples, uint32_t components, float * d_pixels, const rb_render_opts * opts, void * stream);
RB_API rb_status rb_pathtracer_stats(const rb_pathtracer * pt, rb_render_stats * out);
RB_API rb_status rb_pathtracer_multi_stats(rb_pathtracer * pt, rb_multi_stats * out);

/* Live rendering (PathTracerApp's "Live render", src/apps/pathtracer/PathTracerApp.cpp:57-98,127-136, which
 * calls render() for the whole frame every UI frame and uploads the float image again). A frame is refined pass by pass:
 * _live_begin sizes the session, every _live_pass traces the next slice of the stratification grid (cells sid % pass_count
 * == passes done) and adds it to per-pixel sums that never leave the device — a changed camera, sample count, depth or
 * pass_count starts over, a finished frame makes the call a no-op — and _live_display returns the mean so far as 8-bit
 * pixels (tonemap 1 - exp(-c), sRGB, quantisation as in rb_output_ldr, alpha 255 for 4 components): `components` bytes per
 * pixel come back per displayed frame, nothing goes up but the camera. After pass_count passes the bytes equal
 * rb_output_ldr of one rb_pathtracer_render with the same samples and seed. */
RB_API rb_status rb_pathtracer_live_begin(rb_pathtracer * pt, uint32_t width, uint32_t height);
RB_API rb_status rb_pathtracer_live_pass(rb_pathtracer * pt, const rb_camera * camera, uint32_t samples, uint32_t depth, uint32_t pass_count, const rb_render_opts * opts);
RB_API rb_status rb_pathtracer_live_display(rb_pathtracer * pt, uint32_t components, int tonemap, int srgb, uint8_t * out, uint32_t * passes_done);

/* Measurement aid (the reference has no counterpart; bench.py's roofline uses it outside its timed regions): one frame
 * rendered twice — once on a single stream with a CUDA event before every kernel, giving device milliseconds and launch
 * counts per kernel class (0 ray generation, 1 closest-hit wavefront, 2 shading, 3 shadow wavefront, 4 accumulation),
 * and once with counting copies of the two traversal kernels, giving the interior nodes / triangles / roots they test
 * (the per-ray figures of DESIGN.md section 5). The image is discarded. */
typedef struct rb_pt_profile {
	double ms[5];
	uint64_t launches[5];
	uint64_t rays_closest, rays_shadow, shaded_hits, samples;
	uint64_t trace_interior, trace_tris, trace_roots;
	uint64_t shadow_interior, shadow_tris, shadow_roots;
} rb_pt_profile;
RB_API rb_status rb_pathtracer_profile_frame(rb_pathtracer * pt, const rb_camera * camera, uint32_t width, uint32_t height, uint32_t samples, uint32_t depth, const rb_render_opts * opts, rb_pt_profile * out);

/* ------------------------------------------------------------------------------------------------ output stage
 * The headless `--render` tail: tonemap `1 - exp(-c)` (src/apps/pathtracer/main.cpp:109-116, only when
 * the target is not EXR), then Image::saveAsLDR's conversion (src/engine/resources/Image.cpp:138-152): `pow(v, 1/2.2)` on
 * every channel but alpha, `min(255, max(0, 255*v))`, truncation to a byte, alpha forced to 255 (IGNORE_ALPHA). Fused in
 * one kernel; the PNG encoder stays on the host. `components` is 3 or 4; `out` holds width*height*components bytes. */
RB_API rb_status rb_output_ldr(const float * pixels, uint32_t width, uint32_t height, uint32_t components, int tonemap, int srgb, uint8_t * out);
RB_API rb_status rb_output_ldr_device(const float * d_pixels, uint32_t width, uint32_t height, uint32_t components, int tonemap, int srgb, uint8_t * d_out, void * stream);
/* The .exr branch (main.cpp:108,119 -> Image::saveAsHDR, src/engine/resources/Image.cpp:160-289): the layer split
 * and the writer's float -> half conversion (src/libs/tinyexr/tinyexr.h:971-1008) in one pass. planes receives
 * `components` layers of width*height halves in the file's channel order: B,G,R for 3 components, A,B,G,R for 4
 * (A = 1.0 when ignore_alpha, as main.cpp:119 asks); flip reverses the rows (Image::Save::FLIP). The encoder is then
 * given HALF input layers. components must be 3 or 4. */
RB_API rb_status rb_output_hdr(const float * pixels, uint32_t width, uint32_t height, uint32_t components, int ignore_alpha, int flip, uint16_t * planes);
RB_API rb_status rb_output_hdr_device(const float * d_pixels, uint32_t width, uint32_t height, uint32_t components, int ignore_alpha, int flip, uint16_t * d_planes, void * stream);

/* ------------------------------------------------------------------------------------------------ OBJ ingestion
 * Host side of SURVEY §8f rank 2: what Mesh(std::istream &, Mesh::Load, name) (src/engine/resources/Mesh.cpp:100-266, the
 * loader behind Resources::getMesh, ResourcesManager.cpp:306-340) makes of a Wavefront OBJ text — the same vertices in the
 * same order, the same indices, the same floats — parsed on all host threads from memory. Only `v`, `vn`, `vt` and the
 * first three corners of `f` lines count, `vt` is stored as (u, 1 - v), Indexed mode creates one vertex per distinct
 * corner text in first-use order. Where the reference would throw or read out of bounds (a number that does not parse,
 * an index outside the file's arrays, relative indices) this returns RB_ERR_INVALID. Tangent frames are not computed
 * (the reference delegates them to MikkTSpace). No GPU involved; the result feeds rb_raycaster_add_mesh. */
typedef struct rb_mesh rb_mesh;
typedef enum rb_obj_mode { /* Mesh::Load (Mesh.hpp:19-23) */
	RB_OBJ_EXPANDED = 0, /* one vertex per face corner */
	RB_OBJ_POINTS	= 1, /* vertices as listed, no faces */
	RB_OBJ_INDEXED	= 2	 /* vertices shared between corners with identical attribute indices (what getMesh uses) */
} rb_obj_mode;
RB_API rb_status rb_obj_parse(const char * text, size_t nbytes, rb_obj_mode mode, rb_mesh ** out);
RB_API void rb_mesh_destroy(rb_mesh * mesh);
/* out[0]=positions, out[1]=normals, out[2]=texcoords (each a count of vertices, 0 when the file has none), out[3]=indices. */
RB_API rb_status rb_mesh_sizes(const rb_mesh * mesh, uint64_t out[4]);
/* Copies the arrays out (any pointer may be NULL); bbox receives Mesh::computeBoundingBox (Mesh.cpp:306-324): min xyz, max xyz. */
RB_API rb_status rb_mesh_read(const rb_mesh * mesh, float * positions, float * normals, float * texcoords, uint32_t * indices, float bbox[6]);
/* Mesh::computeNormals (Mesh.cpp:326-352): replaces the normals by area-independent face-normal averages. */
RB_API rb_status rb_mesh_compute_normals(rb_mesh * mesh);

/* ------------------------------------------------------------------------------------------------ device KATs
 * Known-answer hooks: evaluate the device versions of the reference's scalar building blocks on arrays (host
 * pointers). Used only by the parity tests; they launch real kernels and have no CPU path. */
RB_API rb_status rb_kat_ggx_eval(const float * wo, const float * base, const float * rough, const float * metal, const float * wi, size_t n, float * out3);
RB_API rb_status rb_kat_ggx_sample(const float * wo, const float * base, const float * rough, const float * metal, const float * u3, size_t n, float * out_wi3, float * out_eval3);
RB_API rb_status rb_kat_image_sample(const float * rgba, uint32_t width, uint32_t height, const float * uv, size_t n, float * out4);
RB_API rb_status rb_kat_cubemap_sample(const float * faces, uint32_t width, uint32_t height, const float * dirs, size_t n, float * out3);
RB_API rb_status rb_kat_light_sample(const rb_light_desc * light, const float * positions, size_t n, float * out5);
RB_API rb_status rb_kat_sky_eval(const float * lut_rgba, uint32_t width, uint32_t height, const float * origins, const float * dirs, const float sun[3], size_t n, float * out3);
/* The exact multiply-high division of the path-tracer kernels (slot -> pixel, sample; replaces the u32 divisions the
 * reference's nested loops never needed): out2[2i] = that division, out2[2i+1] = the device's own n / d. */
RB_API rb_status rb_kat_div(const uint32_t * numerators, const uint32_t * denominators, size_t n, uint32_t * out2);
RB_API rb_status rb_kat_box_test(const float * origins, const float * dirs, const float * boxes6, float mini, float maxi, size_t n, uint8_t * out);
/* Raycaster::intersects(const Ray &, const TriangleInfos &, mini, maxi) (Raycaster.cpp:232-262; private, reached by the friend
 * RaycasterVisualisation): n (ray, triangle) pairs. positions / dirs are a Ray's `pos` and already-normalised `dir`, used as
 * given; vertices9 holds v0, v1, v2 per pair; out4 = (hit, dist, u, v) per pair. */
RB_API rb_status rb_kat_triangle_test(const float * positions, const float * dirs, const float * vertices9, float mini, float maxi, size_t n, float * out4);

#ifdef __cplusplus
}
#endif
#endif /* RENDU_B200_H */
