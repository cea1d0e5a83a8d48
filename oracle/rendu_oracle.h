/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * rendu_oracle: a plain-C CPU restatement of the reference's path-tracer hot path (kosua20/Rendu), used ONLY as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs. Nothing under
 * rendu_b200/ includes, links or loads it. Every function cites the reference file:line it follows.
 *
 * Pinning: the reference ships no tests, golden vectors or fixtures for this path, so the
 * oracle is pinned against the reference ITSELF: oracle/_ref/librendu_ref.so is the reference's own translation
 * units compiled unmodified (oracle/build_ref.sh); tests/test_oracle_raycaster.py and tests/test_oracle_pathtracer.py check every function here
 * against it bit for bit (rays, hierarchy, scalar kernels, and whole renders with the same mt19937 stream), and
 * tests/golden/ holds vectors generated from it (tests/golden/make_golden.py) for machines without /root/reference.
 */
#ifndef RENDU_ORACLE_H
#define RENDU_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Same memory layout as the reference's Raycaster::Node / TriangleInfos (Raycaster.hpp:117-132). */
typedef struct { float bmin[3], bmax[3]; uint64_t left, right; uint8_t leaf; uint8_t pad[7]; } orc_node;
typedef struct { float bmin[3], bmax[3]; uint64_t v0, v1, v2, localId; uint32_t meshId; uint32_t pad; } orc_tri;
/* Raycaster::Hit (Raycaster.hpp:16-42). */
typedef struct { uint32_t hit; float dist, u, v, w; uint32_t pad; uint64_t localId, meshId, internalId; } orc_hit;

typedef struct {
	const orc_node * nodes; size_t nnodes;
	const orc_tri * tris; size_t ntris;
	const float * verts; size_t nverts;
	uint32_t meshCount;
} orc_bvh;

/* Ray::Ray (Intersection.cpp:3-5): out9 = pos, dir, invdir. */
void orc_ray(const float * origin, const float * direction, float * out9);
/* Intersection::box (Intersection.cpp:22-32). */
int orc_box(const float * ray9, const float * bmin, const float * bmax, float mini, float maxi);
/* Raycaster::intersects(ray, tri, mini, maxi) (Raycaster.cpp:232-262). Returns 1 on hit and fills t,u,v. */
int orc_tri_test(const orc_bvh * b, const float * ray9, const orc_tri * tri, float mini, float maxi, float * t, float * u, float * v);
/* RaycasterVisualisation::getRayLevels (RaycasterVisualisation.cpp:41-77): visited nodes (id, depth) in pop order. */
void orc_ray_levels(const orc_bvh * b, const float * origin, const float * direction, float mini, float maxi, size_t cap, uint64_t * nodeIds, uint32_t * depths, size_t * count, orc_hit * hit);
/* Raycaster::intersects (Raycaster.cpp:144-185). */
orc_hit orc_intersect(const orc_bvh * b, const float * origin, const float * direction, float mini, float maxi);
/* Raycaster::intersectsAny (Raycaster.cpp:187-224). */
int orc_intersect_any(const orc_bvh * b, const float * origin, const float * direction, float mini, float maxi);
/* Raycaster::visible (Raycaster.cpp:226-230). */
int orc_visible(const orc_bvh * b, const float * p0, const float * p1);
/* Batches, optionally over `threads` pthreads with static contiguous blocks (System.hpp:72-104). */
void orc_intersect_batch(const orc_bvh * b, const float * org, const float * dir, float mini, float maxi, size_t n, int threads, orc_hit * out);
void orc_intersect_any_batch(const orc_bvh * b, const float * org, const float * dir, float mini, float maxi, size_t n, int threads, uint8_t * out);

/* Raycaster::addMesh (Raycaster.cpp:14-52) on plain arrays: appends to caller-sized outputs; returns triangles added. */
size_t orc_add_mesh(const float * positions, size_t nverts, const uint32_t * indices, size_t nindices, const float * model16,
	float * vertsOut, size_t vertOffset, orc_tri * trisOut, uint32_t meshId);
/* Raycaster::updateHierarchy (Raycaster.cpp:54-142); C++ because the tree is defined by libstdc++'s std::partition /
 * std::nth_element (rendu_oracle_build.cpp). tris is reordered in place; nodesOut must hold 2*ntris+meshCount entries;
 * meshFirst/meshTris give each mesh's triangle range. Returns the node count. */
size_t orc_build(orc_tri * tris, size_t ntris, const uint64_t * meshFirst, const uint64_t * meshTris, uint32_t meshCount, orc_node * nodesOut);

/* ------------------------------------------------------------------------------------------------ path tracer
 * Plain-array scene: what PathTracer::render reads from Scene / Object / Material / Mesh / Light. */
typedef struct { const float * pixels; uint32_t width, height, components; } orc_image; /* Image (Image.hpp:143-146) */
typedef struct {
	const float * positions; const float * normals; const float * tangents; const float * bitangents; const float * texcoords;
	size_t nverts; const uint32_t * indices; size_t nindices; float model[16];
	int32_t tex[3]; /* Material::textures()[0..2] */
	uint32_t emissive, masked, useTexCoords;
} orc_object;
typedef struct { uint32_t kind, castsShadow; float intensity[3], position[3], direction[3], radius, inner, outer; } orc_light;
typedef struct {
	const orc_object * objects; size_t nobjects;
	const orc_light * lights; size_t nlights;
	const orc_image * images; size_t nimages;
	uint32_t bgMode; float bgColor[3]; int32_t bgTex[6]; float sunDir[3]; int32_t skyLut;
	orc_bvh bvh;
} orc_scene;
typedef struct {
	float eye[3], corner[3], dx[3], dy[3];
	uint32_t width, height, samples, depth, components;
	uint64_t seed;
	uint32_t sampleIndex, sampleCount; /* sampleCount <= 1: all stratification cells */
	uint32_t window[4];                /* x0, y0, x1, y1 of the pixels to render (all zero: the whole image) */
} orc_render_args;
/* Sequential random source (the reference's per-thread mt19937, Random.cpp:21-27,72-80). */
typedef struct { float (*next_float)(void *); int (*next_int)(void *, int lo, int hi); } orc_rng_stream;

/* Counter-based generator shared with the CUDA path (rendu_b200/csrc/rb_shade.cuh). */
uint64_t orc_rng_key(uint64_t seed, uint32_t pixel, uint32_t sid);
float orc_rng_float(uint64_t key, uint32_t dim);

/* PathTracer::render (PathTracer.cpp:159-299): accumulates into `pixels` (+=, then / samples). stream == NULL selects
 * the counter-based generator (any thread count); otherwise draws are sequential and one thread is used. Returns the
 * number of Raycaster::intersects queries issued. */
uint64_t orc_render(const orc_scene * scene, const orc_render_args * args, const orc_rng_stream * stream, void * streamState, int threads, float * pixels);

/* Scalar building blocks on arrays (KATs). */
void orc_image_rgbal(const orc_image * im, const float * uv, size_t n, float * out4);          /* Image.cpp:86-106 */
void orc_cubemap(const orc_image * faces6, const float * dirs, size_t n, float * out3);         /* Texture.cpp:141-191 */
void orc_light_sample(const orc_light * l, const float * pos, size_t n, float * out5);          /* PointLight/SpotLight/DirectionalLight::sample */
void orc_ggx_eval(const float * wo, const float * base, const float * rough, const float * metal, const float * wi, size_t n, float * out3); /* MaterialGGX.cpp:94-101 */
void orc_ggx_sample(const float * wo, const float * base, const float * rough, const float * metal, const float * u3, size_t n, float * outWi, float * outEval); /* :54-92 */
void orc_sky_eval(const orc_image * lut, const float * org, const float * dirs, const float * sun, size_t n, float * out3); /* MaterialSky.cpp:6-98 */

/* renderOneShot's tonemap + Image::saveAsLDR's byte conversion (main.cpp:109-119, Image.cpp:138-152). */
uint16_t orc_float_to_half(float value);
float orc_half_to_float(uint16_t h);
void orc_output_hdr(const float * pixels, uint32_t width, uint32_t height, uint32_t components, int ignore_alpha, int flip, uint16_t * planes);
void orc_output_ldr(const float * pixels, uint32_t width, uint32_t height, uint32_t components, int tonemap, int srgb, uint8_t * out);

/* The reference's random stream (rendu_oracle_std.cpp, libstdc++ <random>): state after Random::seed(seed) as seen by the
 * single worker thread System::forParallel spawns when one CPU is available. */
void * orc_mt_create(unsigned int seed);
void orc_mt_destroy(void * state);
const orc_rng_stream * orc_mt_stream(void);
void orc_mt_floats(unsigned int seed, size_t n, float * out); /* Random::seed(seed) then n x Random::Float() on the seeding thread */

#ifdef __cplusplus
}
#endif
#endif
