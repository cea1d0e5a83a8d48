/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * C-callable harness around the *unmodified* reference translation units (kosua20/Rendu, compiled from
 * /root/reference by oracle/build_ref.sh into oracle/_ref/librendu_ref.so). It exposes the reference's
 * Raycaster and PathTracer so that tests/ and bench.py's cpu_baseline / --impl reference legs can
 *   (1) build the reference BVH and read it back (private members reached through the
 *       `friend class RaycasterVisualisation` hook, src/engine/raycaster/Raycaster.hpp:12,18),
 *   (2) trace batches of rays with Raycaster::intersects / intersectsAny (Raycaster.cpp:144-224),
 *   (3) load a bundled .scene, flatten it into plain arrays (the layout the product's C ABI takes) and
 *       render it with PathTracer::render (PathTracer.cpp:159-299),
 *   (4) evaluate the reference's scalar building blocks on arrays for known-answer tests.
 * Nothing under rendu_b200/ may link or load this library.
 */
#include "raycaster/Raycaster.hpp"
#include "raycaster/Intersection.hpp"
#include "resources/ResourcesManager.hpp"
#include "resources/Texture.hpp"
#include "resources/Image.hpp"
#include "resources/Mesh.hpp"
#include "scene/Scene.hpp"
#include "scene/Sky.hpp"
#include "scene/lights/PointLight.hpp"
#include "scene/lights/SpotLight.hpp"
#include "scene/lights/DirectionalLight.hpp"
#include "generation/Random.hpp"
#include "system/System.hpp"
#include "input/Camera.hpp"
#include "PathTracer.hpp"
#include "MaterialGGX.hpp"
#include "MaterialSky.hpp"

#include <atomic>
#include <chrono>
#include <cstring>
#include <sched.h>
#include <sstream>
#include <thread>

/* The reference declares this class a friend of Raycaster and Raycaster::Hit; the harness supplies its own
 * definition (the reference's RaycasterVisualisation.cpp is not part of the oracle build). */
class RaycasterVisualisation {
public:
	static size_t nodeCount(const Raycaster & r) { return r._hierarchy.size(); }
	static size_t triCount(const Raycaster & r) { return r._triangles.size(); }
	static size_t vertCount(const Raycaster & r) { return r._vertices.size(); }
	static unsigned meshCount(const Raycaster & r) { return r._meshCount; }

	/* nodes: per node 6 floats box (min, max) ; left/right as u64 ; leaf as u8. */
	static void dump(const Raycaster & r, float * nodeBox, uint64_t * nodeLeft, uint64_t * nodeRight, uint8_t * nodeLeaf,
		float * triBox, uint64_t * triV, uint64_t * triLocalId, uint32_t * triMeshId, float * verts) {
		for(size_t i = 0; i < r._hierarchy.size(); ++i) {
			const auto & n = r._hierarchy[i];
			if(nodeBox) {
				for(int k = 0; k < 3; ++k) {
					nodeBox[6 * i + k]	   = n.box.minis[k];
					nodeBox[6 * i + 3 + k] = n.box.maxis[k];
				}
			}
			if(nodeLeft) nodeLeft[i] = n.left;
			if(nodeRight) nodeRight[i] = n.right;
			if(nodeLeaf) nodeLeaf[i] = n.leaf ? 1 : 0;
		}
		for(size_t i = 0; i < r._triangles.size(); ++i) {
			const auto & t = r._triangles[i];
			if(triBox) {
				for(int k = 0; k < 3; ++k) {
					triBox[6 * i + k]	  = t.box.minis[k];
					triBox[6 * i + 3 + k] = t.box.maxis[k];
				}
			}
			if(triV) {
				triV[3 * i + 0] = t.v0;
				triV[3 * i + 1] = t.v1;
				triV[3 * i + 2] = t.v2;
			}
			if(triLocalId) triLocalId[i] = t.localId;
			if(triMeshId) triMeshId[i] = t.meshId;
		}
		if(verts) {
			std::memcpy(verts, r._vertices.data(), r._vertices.size() * sizeof(glm::vec3));
		}
	}

	/* Replay of the closest-hit / any-hit loops (Raycaster.cpp:144-224) with work counters. The loop structure
	 * follows the reference; the per-triangle and per-box tests call the reference functions themselves. */
	static void replayCounters(const Raycaster & r, const glm::vec3 & origin, const glm::vec3 & direction, float mini, float maxi,
		bool any, uint64_t * counters /* popped, boxTests, triTests, interiorVisited, maxStack */) {
		const Ray ray(origin, direction);
		std::vector<size_t> stack;
		size_t maxStack = 0;
		for(size_t nid = 0; nid < r._meshCount; ++nid) {
			counters[1]++;
			if(Intersection::box(ray, r._hierarchy[nid].box, mini, maxi)) {
				stack.push_back(nid);
			}
		}
		float best = std::numeric_limits<float>::max();
		while(!stack.empty()) {
			maxStack = std::max(maxStack, stack.size());
			const Raycaster::Node & node = r._hierarchy[stack.back()];
			stack.pop_back();
			counters[0]++;
			if(node.leaf) {
				for(size_t tid = 0; tid < node.right; ++tid) {
					counters[2]++;
					const Raycaster::Hit h = r.intersects(ray, r._triangles[node.left + tid], mini, maxi);
					if(h.hit && any) {
						counters[4] = std::max<uint64_t>(counters[4], maxStack);
						return;
					}
					if(h.hit && h.dist < best) {
						best = h.dist;
						maxi = best;
					}
				}
				continue;
			}
			counters[3]++;
			counters[1] += 2;
			if(Intersection::box(ray, r._hierarchy[node.left].box, mini, maxi)) stack.push_back(node.left);
			if(Intersection::box(ray, r._hierarchy[node.right].box, mini, maxi)) stack.push_back(node.right);
		}
		counters[4] = std::max<uint64_t>(counters[4], maxStack);
	}
};

namespace {

double nowSeconds() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/* Plain row-block fan-out used by the batch entry points; mirrors the reference's policy of static contiguous
 * blocks (System.hpp:72-104) but with an explicit thread count so that benchmarks can state it. */
template<typename F>
void parallelBlocks(size_t n, int threads, F f) {
	if(threads <= 1 || n < 2) {
		f(0, n);
		return;
	}
	std::vector<std::thread> pool;
	const size_t span = (n + threads - 1) / threads;
	for(int t = 0; t < threads; ++t) {
		const size_t a = std::min(n, t * span), b = std::min(n, (t + 1) * span);
		if(a < b) pool.emplace_back(f, a, b);
	}
	for(auto & t : pool) t.join();
}

struct SceneHolder {
	std::shared_ptr<Scene> scene;
	std::unique_ptr<PathTracer> tracer;
};

std::atomic<uint64_t> g_rayCount(0);

int lightKind(const Light * l) {
	if(dynamic_cast<const PointLight *>(l)) return 0;
	if(dynamic_cast<const SpotLight *>(l)) return 1;
	if(dynamic_cast<const DirectionalLight *>(l)) return 2;
	return -1;
}

} // namespace

extern "C" {

/* ---------------------------------------------------------------- misc */

unsigned int __wrap__ZNSt6thread20hardware_concurrencyEv();
int ref_hardware_threads() {
	return int(__wrap__ZNSt6thread20hardware_concurrencyEv());
}

/* Number of worker threads System::forParallel will use (System.hpp:82). */
int ref_forparallel_threads() {
	return std::max(int(__wrap__ZNSt6thread20hardware_concurrencyEv()) - 1, 1);
}

void ref_seed(unsigned int s) {
	/* Touch the calling thread's generator first: its thread_local construction draws from the shared generator
	 * (Random.cpp:72-80), and if that happened INSIDE Random::seed the number of draws consumed before the worker
	 * thread is seeded would depend on whether this thread had used Random before. */
	(void)Random::Float();
	Random::seed(s);
}

void ref_set_verbose(int on) {
	Log::setDefaultVerbose(on != 0);
}

/* std::thread::hardware_concurrency() as seen by the reference objects (linked with --wrap): 0 = the real value,
 * otherwise the override. With an override of 1 or 2 System::forParallel (System.hpp:82) spawns a single worker, which
 * makes a seeded reference render reproducible (Random.cpp:72-80 hands out per-thread seeds in lock order). */
static unsigned int g_concurrencyOverride = 0;
unsigned int __real__ZNSt6thread20hardware_concurrencyEv();
unsigned int __wrap__ZNSt6thread20hardware_concurrencyEv() {
	return g_concurrencyOverride ? g_concurrencyOverride : __real__ZNSt6thread20hardware_concurrencyEv();
}
void ref_set_concurrency(unsigned int n) {
	g_concurrencyOverride = n;
}

/* Restrict this process to one CPU (kept for experiments; hardware_concurrency ignores affinity on this libstdc++). */
int ref_pin_single_cpu(int on) {
	static cpu_set_t saved;
	static bool haveSaved = false;
	if(on) {
		if(!haveSaved) {
			sched_getaffinity(0, sizeof(saved), &saved);
			haveSaved = true;
		}
		cpu_set_t one;
		CPU_ZERO(&one);
		for(int c = 0; c < CPU_SETSIZE; ++c) {
			if(CPU_ISSET(c, &saved)) {
				CPU_SET(c, &one);
				break;
			}
		}
		return sched_setaffinity(0, sizeof(one), &one);
	}
	if(haveSaved) return sched_setaffinity(0, sizeof(saved), &saved);
	return 0;
}

/* ---------------------------------------------------------------- raycaster */

void * ref_raycaster_create() {
	return new Raycaster();
}

void ref_raycaster_destroy(void * h) {
	delete static_cast<Raycaster *>(h);
}

/* model: 16 floats, column-major (glm layout). */
void ref_raycaster_add_mesh(void * h, const float * pos, size_t nverts, const uint32_t * idx, size_t nidx, const float * model) {
	Mesh mesh("harness");
	mesh.positions.resize(nverts);
	std::memcpy(mesh.positions.data(), pos, nverts * sizeof(glm::vec3));
	mesh.indices.assign(idx, idx + nidx);
	glm::mat4 m(1.0f);
	if(model) std::memcpy(&m[0][0], model, 16 * sizeof(float));
	static_cast<Raycaster *>(h)->addMesh(mesh, m);
}

double ref_raycaster_build(void * h) {
	const double t0 = nowSeconds();
	static_cast<Raycaster *>(h)->updateHierarchy();
	return nowSeconds() - t0;
}

void ref_raycaster_sizes(void * h, uint64_t * out4) {
	const Raycaster & r = *static_cast<Raycaster *>(h);
	out4[0] = RaycasterVisualisation::nodeCount(r);
	out4[1] = RaycasterVisualisation::triCount(r);
	out4[2] = RaycasterVisualisation::vertCount(r);
	out4[3] = RaycasterVisualisation::meshCount(r);
}

void ref_raycaster_dump(void * h, float * nodeBox, uint64_t * nodeLeft, uint64_t * nodeRight, uint8_t * nodeLeaf,
	float * triBox, uint64_t * triV, uint64_t * triLocalId, uint32_t * triMeshId, float * verts) {
	RaycasterVisualisation::dump(*static_cast<Raycaster *>(h), nodeBox, nodeLeft, nodeRight, nodeLeaf, triBox, triV, triLocalId, triMeshId, verts);
}

/* Batch closest hit. outF: 4 floats/ray (dist,u,v,w); outI: 3 u64/ray (hit, localId, meshId). Returns seconds. */
double ref_raycaster_intersect(void * h, const float * org, const float * dir, float mini, float maxi, size_t n, int threads, float * outF, uint64_t * outI) {
	const Raycaster & r = *static_cast<Raycaster *>(h);
	const double t0 = nowSeconds();
	parallelBlocks(n, threads, [&](size_t a, size_t b) {
		for(size_t i = a; i < b; ++i) {
			const Raycaster::Hit hit = r.intersects(glm::vec3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), glm::vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), mini, maxi);
			if(outF) {
				outF[4 * i + 0] = hit.dist;
				outF[4 * i + 1] = hit.u;
				outF[4 * i + 2] = hit.v;
				outF[4 * i + 3] = hit.w;
			}
			if(outI) {
				outI[3 * i + 0] = hit.hit ? 1 : 0;
				outI[3 * i + 1] = hit.localId;
				outI[3 * i + 2] = hit.meshId;
			}
		}
	});
	return nowSeconds() - t0;
}

double ref_raycaster_intersect_any(void * h, const float * org, const float * dir, float mini, float maxi, size_t n, int threads, uint8_t * out) {
	const Raycaster & r = *static_cast<Raycaster *>(h);
	const double t0 = nowSeconds();
	parallelBlocks(n, threads, [&](size_t a, size_t b) {
		for(size_t i = a; i < b; ++i) {
			const bool any = r.intersectsAny(glm::vec3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), glm::vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), mini, maxi);
			if(out) out[i] = any ? 1 : 0;
		}
	});
	return nowSeconds() - t0;
}

void ref_raycaster_visible(void * h, const float * p0, const float * p1, size_t n, uint8_t * out) {
	const Raycaster & r = *static_cast<Raycaster *>(h);
	for(size_t i = 0; i < n; ++i) {
		out[i] = r.visible(glm::vec3(p0[3 * i], p0[3 * i + 1], p0[3 * i + 2]), glm::vec3(p1[3 * i], p1[3 * i + 1], p1[3 * i + 2])) ? 1 : 0;
	}
}

/* Sum of traversal counters over a batch: popped nodes, box tests, triangle tests, interior visits, max stack. */
void ref_raycaster_counters(void * h, const float * org, const float * dir, float mini, float maxi, size_t n, int any, uint64_t * counters5) {
	const Raycaster & r = *static_cast<Raycaster *>(h);
	for(int k = 0; k < 5; ++k) counters5[k] = 0;
	for(size_t i = 0; i < n; ++i) {
		RaycasterVisualisation::replayCounters(r, glm::vec3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), glm::vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), mini, maxi, any != 0, counters5);
	}
}

/* Ray::Ray and Intersection::box on arrays (KAT). rays: org(3) dir(3); boxes: min(3) max(3). */
void ref_box_test(const float * org, const float * dir, const float * boxes, float mini, float maxi, size_t n, uint8_t * out, float * outRay9) {
	for(size_t i = 0; i < n; ++i) {
		const Ray ray(glm::vec3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), glm::vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
		BoundingBox box;
		box.minis = glm::vec3(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]);
		box.maxis = glm::vec3(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]);
		out[i] = Intersection::box(ray, box, mini, maxi) ? 1 : 0;
		if(outRay9) {
			for(int k = 0; k < 3; ++k) {
				outRay9[9 * i + k]	   = ray.pos[k];
				outRay9[9 * i + 3 + k] = ray.dir[k];
				outRay9[9 * i + 6 + k] = ray.invdir[k];
			}
		}
	}
}

/* ---------------------------------------------------------------- scalar KATs */

void ref_ggx_eval(const float * wo, const float * base, const float * rough, const float * metal, const float * wi, size_t n, float * out3) {
	for(size_t i = 0; i < n; ++i) {
		const glm::vec3 r = MaterialGGX::eval(glm::vec3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), glm::vec3(base[3 * i], base[3 * i + 1], base[3 * i + 2]), rough[i], metal[i], glm::vec3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]));
		out3[3 * i] = r.x; out3[3 * i + 1] = r.y; out3[3 * i + 2] = r.z;
	}
}

/* For each item: seed the generator with seeds[i], record the three uniforms the sampler will consume, re-seed and call
 * MaterialGGX::sampleAndEval. outU: 3 floats, outWi: 3 floats, outEval: 3 floats. */
void ref_ggx_sample(const float * wo, const float * base, const float * rough, const float * metal, const uint32_t * seeds, size_t n, float * outU, float * outWi, float * outEval) {
	for(size_t i = 0; i < n; ++i) {
		(void)Random::Float();
		Random::seed(seeds[i]);
		for(int k = 0; k < 3; ++k) outU[3 * i + k] = Random::Float();
		Random::seed(seeds[i]);
		glm::vec3 wi(0.0f);
		const glm::vec3 r = MaterialGGX::sampleAndEval(glm::vec3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), glm::vec3(base[3 * i], base[3 * i + 1], base[3 * i + 2]), rough[i], metal[i], wi);
		for(int k = 0; k < 3; ++k) {
			outWi[3 * i + k]   = wi[k];
			outEval[3 * i + k] = r[k];
		}
	}
}

/* Draw n uniforms / ints after seeding (pins the oracle's mt19937 + distribution restatement). */
void ref_random_floats(unsigned int seed, size_t n, float * out) {
	(void)Random::Float();
	Random::seed(seed);
	for(size_t i = 0; i < n; ++i) out[i] = Random::Float();
}

void ref_random_ints(unsigned int seed, int lo, int hi, size_t n, int * out) {
	(void)Random::Float();
	Random::seed(seed);
	for(size_t i = 0; i < n; ++i) out[i] = Random::Int(lo, hi);
}

void ref_cosine_hemisphere(unsigned int seed, size_t n, float * out3) {
	(void)Random::Float();
	Random::seed(seed);
	for(size_t i = 0; i < n; ++i) {
		const glm::vec3 d = Random::sampleCosineHemisphere();
		out3[3 * i] = d.x; out3[3 * i + 1] = d.y; out3[3 * i + 2] = d.z;
	}
}

/* Image::rgbal / rgbl on a caller-provided RGBA float image. */
void ref_image_sample(const float * rgba, uint32_t w, uint32_t h, const float * uv, size_t n, float * out4, float * out3) {
	Image img(w, h, 4);
	std::memcpy(img.pixels.data(), rgba, size_t(w) * h * 4 * sizeof(float));
	for(size_t i = 0; i < n; ++i) {
		const glm::vec4 a = img.rgbal(uv[2 * i], uv[2 * i + 1]);
		const glm::vec3 b = img.rgbl(uv[2 * i], uv[2 * i + 1]);
		for(int k = 0; k < 4; ++k) out4[4 * i + k] = a[k];
		for(int k = 0; k < 3; ++k) out3[3 * i + k] = b[k];
	}
}

/* Texture::sampleCubemap on 6 caller-provided RGBA float faces of size w x h. */
void ref_cubemap_sample(const float * faces, uint32_t w, uint32_t h, const float * dirs, size_t n, float * out3) {
	Texture tex("harness-cube");
	tex.shape = TextureShape::Cube;
	tex.width = w; tex.height = h; tex.depth = 6; tex.levels = 1;
	for(int f = 0; f < 6; ++f) {
		tex.images.emplace_back(w, h, 4);
		std::memcpy(tex.images.back().pixels.data(), faces + size_t(f) * w * h * 4, size_t(w) * h * 4 * sizeof(float));
	}
	for(size_t i = 0; i < n; ++i) {
		const glm::vec3 c = tex.sampleCubemap(glm::vec3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
		out3[3 * i] = c.x; out3[3 * i + 1] = c.y; out3[3 * i + 2] = c.z;
	}
}

/* Camera::pose + projection + pixelShifts. cam: eye(3) center(3) up(3) fov ratio near far. out: corner dx dy position. */
void ref_camera_shifts(const float * cam13, float * out12) {
	Camera c;
	c.pose(glm::vec3(cam13[0], cam13[1], cam13[2]), glm::vec3(cam13[3], cam13[4], cam13[5]), glm::vec3(cam13[6], cam13[7], cam13[8]));
	c.projection(cam13[10], cam13[9], cam13[11], cam13[12]);
	glm::vec3 corner, dx, dy;
	c.pixelShifts(corner, dx, dy);
	for(int k = 0; k < 3; ++k) {
		out12[k] = corner[k]; out12[3 + k] = dx[k]; out12[6 + k] = dy[k]; out12[9 + k] = c.position()[k];
	}
}

/* ---------------------------------------------------------------- scenes */

void ref_add_resources(const char * path) {
	Resources::manager().addResources(path);
}

void * ref_scene_load(const char * name) {
	SceneHolder * s = new SceneHolder();
	s->scene.reset(new Scene(name));
	if(!s->scene->init(Storage::CPU | Storage::FORCE_FRAME)) {
		delete s;
		return nullptr;
	}
	s->tracer.reset(new PathTracer(s->scene));
	return s;
}

void ref_scene_destroy(void * h) {
	delete static_cast<SceneHolder *>(h);
}

void * ref_scene_raycaster(void * h) {
	return const_cast<Raycaster *>(&static_cast<SceneHolder *>(h)->tracer->raycaster());
}

/* counts: objects, lights, backgroundMode, distinct textures are resolved by the caller through pointers. */
void ref_scene_counts(void * h, uint64_t * out3) {
	const Scene & s = *static_cast<SceneHolder *>(h)->scene;
	out3[0] = s.objects.size();
	out3[1] = s.lights.size();
	out3[2] = uint64_t(s.backgroundMode);
}

/* Object i: sizes = nverts, nidx, hasUV(useTexCoords), materialType, masked, ntextures. model = 16 floats. */
void ref_scene_object_info(void * h, size_t i, uint64_t * out6, float * model16) {
	const Object & o = static_cast<SceneHolder *>(h)->scene->objects[i];
	const Mesh & m	 = *o.mesh();
	out6[0] = m.positions.size();
	out6[1] = m.indices.size();
	out6[2] = o.useTexCoords() ? 1 : 0;
	out6[3] = uint64_t(o.material().type());
	out6[4] = o.material().masked() ? 1 : 0;
	out6[5] = o.material().textures().size();
	std::memcpy(model16, &o.model()[0][0], 16 * sizeof(float));
}

/* Any pointer may be null. Sizes of attribute arrays are reported through counts5 (normals,tangents,bitangents,texcoords,positions). */
void ref_scene_object_mesh(void * h, size_t i, float * pos, float * nor, float * tan, float * bit, float * uv, uint32_t * idx, uint64_t * counts5) {
	const Mesh & m = *static_cast<SceneHolder *>(h)->scene->objects[i].mesh();
	if(counts5) {
		counts5[0] = m.normals.size(); counts5[1] = m.tangents.size(); counts5[2] = m.bitangents.size(); counts5[3] = m.texcoords.size(); counts5[4] = m.positions.size();
	}
	if(pos) std::memcpy(pos, m.positions.data(), m.positions.size() * sizeof(glm::vec3));
	if(nor) std::memcpy(nor, m.normals.data(), m.normals.size() * sizeof(glm::vec3));
	if(tan) std::memcpy(tan, m.tangents.data(), m.tangents.size() * sizeof(glm::vec3));
	if(bit) std::memcpy(bit, m.bitangents.data(), m.bitangents.size() * sizeof(glm::vec3));
	if(uv) std::memcpy(uv, m.texcoords.data(), m.texcoords.size() * sizeof(glm::vec2));
	if(idx) std::memcpy(idx, m.indices.data(), m.indices.size() * sizeof(uint32_t));
}

/* Texture slot `slot` of object i (or of the background object when i == SIZE_MAX): identity pointer (for de-duplication),
 * image index `img` of texture->images. out: width,height,components,imagesCount. pixels may be null. */
const void * ref_scene_texture(void * h, size_t i, size_t slot, size_t img, uint64_t * out4, float * pixels) {
	const Scene & s		 = *static_cast<SceneHolder *>(h)->scene;
	const Material & mat = (i == SIZE_MAX) ? s.background->material() : s.objects[i].material();
	if(slot >= mat.textures().size()) return nullptr;
	const Texture * t = mat.textures()[slot];
	if(img >= t->images.size()) return nullptr;
	const Image & im = t->images[img];
	out4[0] = im.width; out4[1] = im.height; out4[2] = im.components; out4[3] = t->images.size();
	if(pixels) std::memcpy(pixels, im.pixels.data(), im.pixels.size() * sizeof(float));
	return t;
}

/* Background: mode (Scene::Background), colour parameter (if any), sun direction (if Sky). */
void ref_scene_background(void * h, uint64_t * mode, float * color4, float * sunDir3) {
	const Scene & s = *static_cast<SceneHolder *>(h)->scene;
	*mode = uint64_t(s.backgroundMode);
	const Material & mat = s.background->material();
	for(int k = 0; k < 4; ++k) color4[k] = mat.parameters().empty() ? 0.0f : mat.parameters()[0][k];
	const Sky * sky = dynamic_cast<const Sky *>(s.background.get());
	for(int k = 0; k < 3; ++k) sunDir3[k] = sky ? sky->direction()[k] : 0.0f;
}

/* The 512x512 "scattering-precomputed" table MaterialSky::eval fetches through Resources (MaterialSky.cpp:10). */
const void * ref_sky_lut(uint64_t * out3, float * pixels) {
	const Texture * t = Resources::manager().getTexture("scattering-precomputed", Layout::RGBA32F, Storage::CPU);
	if(!t || t->images.empty()) return nullptr;
	const Image & im = t->images[0];
	out3[0] = im.width; out3[1] = im.height; out3[2] = im.components;
	if(pixels) std::memcpy(pixels, im.pixels.data(), im.pixels.size() * sizeof(float));
	return t;
}

/* Light i: kind (0 point, 1 spot, 2 directional), castsShadow, intensity(3); parameters recovered through the encode() text
 * are avoided — instead the harness probes the public getters. out: kind, shadow | floats: intensity(3), position(3), direction(3), radius, innerAngle, outerAngle. */
void ref_scene_light(void * h, size_t i, int * kindShadow2, float * out12) {
	const Light * l = static_cast<SceneHolder *>(h)->scene->lights[i].get();
	kindShadow2[0] = lightKind(l);
	kindShadow2[1] = l->castsShadow() ? 1 : 0;
	for(int k = 0; k < 12; ++k) out12[k] = 0.0f;
	for(int k = 0; k < 3; ++k) out12[k] = l->intensity()[k];
	if(const PointLight * p = dynamic_cast<const PointLight *>(l)) {
		for(int k = 0; k < 3; ++k) out12[3 + k] = p->position()[k];
		out12[9] = p->radius();
	} else if(const SpotLight * sp = dynamic_cast<const SpotLight *>(l)) {
		for(int k = 0; k < 3; ++k) { out12[3 + k] = sp->position()[k]; out12[6 + k] = sp->direction()[k]; }
		out12[9] = sp->radius(); out12[10] = sp->angles()[0]; out12[11] = sp->angles()[1];
	} else if(const DirectionalLight * d = dynamic_cast<const DirectionalLight *>(l)) {
		for(int k = 0; k < 3; ++k) out12[6 + k] = d->direction()[k];
	}
}

/* Light::sample on arrays (KAT): out: dir(3), dist, falloff. */
void ref_scene_light_sample(void * h, size_t i, const float * pos, size_t n, float * out5) {
	const Light * l = static_cast<SceneHolder *>(h)->scene->lights[i].get();
	for(size_t k = 0; k < n; ++k) {
		float dist = 0.0f, att = 0.0f;
		const glm::vec3 d = l->sample(glm::vec3(pos[3 * k], pos[3 * k + 1], pos[3 * k + 2]), dist, att);
		out5[5 * k] = d.x; out5[5 * k + 1] = d.y; out5[5 * k + 2] = d.z; out5[5 * k + 3] = dist; out5[5 * k + 4] = att;
	}
}

/* Scene viewpoint at aspect `ratio`: corner, dx, dy, position (what render() derives at PathTracer.cpp:177,197). */
void ref_scene_camera(void * h, float ratio, float * out12) {
	Camera camera = static_cast<SceneHolder *>(h)->scene->viewpoint();
	camera.ratio(ratio);
	glm::vec3 corner, dx, dy;
	camera.pixelShifts(corner, dx, dy);
	for(int k = 0; k < 3; ++k) {
		out12[k] = corner[k]; out12[3 + k] = dx[k]; out12[6 + k] = dy[k]; out12[9 + k] = camera.position()[k];
	}
}

/* MaterialSky::eval on arrays (needs the scattering LUT from resources/common). */
void ref_sky_eval(const float * org, const float * dir, const float * sun3, size_t n, float * out3) {
	const glm::vec3 sun(sun3[0], sun3[1], sun3[2]);
	for(size_t i = 0; i < n; ++i) {
		const glm::vec3 c = MaterialSky::eval(glm::vec3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), glm::vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), sun);
		out3[3 * i] = c.x; out3[3 * i + 1] = c.y; out3[3 * i + 2] = c.z;
	}
}

/* PathTracer::render on the scene's own viewpoint. Returns seconds (steady clock around render()). `components` 3 or 4.
 * Ray count since the last ref_ray_counter_reset() is available when the library was linked with --wrap. */
double ref_scene_render(void * h, uint32_t w, uint32_t hgt, uint32_t samples, uint32_t depth, uint32_t components, float * pixels) {
	SceneHolder * s = static_cast<SceneHolder *>(h);
	Image render(w, hgt, components);
	Camera camera = s->scene->viewpoint();
	camera.ratio(float(w) / float(hgt));
	const double t0 = nowSeconds();
	s->tracer->render(camera, samples, depth, render);
	const double dt = nowSeconds() - t0;
	std::memcpy(pixels, render.pixels.data(), render.pixels.size() * sizeof(float));
	return dt;
}

/* The tail of renderOneShot (main.cpp:106-119): optional tonemap, then Image::save(path, SRGB_LDR | IGNORE_ALPHA) — the
 * reference's own PNG writer — and the file read back with the reference's own loader. out: width*height*components bytes. */
int ref_output_ldr(const float * pixels, uint32_t w, uint32_t h, uint32_t components, int tonemap, const char * tmpPath, uint8_t * out) {
	Image render(w, h, components);
	std::memcpy(render.pixels.data(), pixels, render.pixels.size() * sizeof(float));
	if(tonemap) {
		System::forParallel(0, render.height, [&render](size_t y) {
			for(uint x = 0; x < render.width; ++x) {
				const glm::vec3 & color	   = render.rgb(int(x), int(y));
				render.rgb(int(x), int(y)) = glm::vec3(1.0f) - glm::exp(-color * 1.f);
			}
		});
	}
	if(render.save(tmpPath, Image::Save::SRGB_LDR | Image::Save::IGNORE_ALPHA) != 0) return 1;
	Image back;
	if(back.load(tmpPath, components, false, true) != 0) return 2;
	for(size_t i = 0; i < back.pixels.size(); ++i) out[i] = uint8_t(back.pixels[i] * 255.0f + 0.5f);
	return 0;
}

/* The .exr branch of renderOneShot (main.cpp:108,119: no tonemap, Image::save -> saveAsHDR, Image.cpp:160-289): the
 * reference's own EXR writer (half channels) and the file read back with its own loader (half -> float, Image.cpp:335+).
 * out: width*height*components floats holding the stored half values. */
int ref_output_hdr(const float * pixels, uint32_t w, uint32_t h, uint32_t components, const char * tmpPath, float * out) {
	Image render(w, h, components);
	std::memcpy(render.pixels.data(), pixels, render.pixels.size() * sizeof(float));
	if(render.save(tmpPath, Image::Save::SRGB_LDR | Image::Save::IGNORE_ALPHA) != 0) return 1;
	Image back;
	if(back.load(tmpPath, components, false, true) != 0) return 2;
	if(back.pixels.size() != render.pixels.size()) return 3;
	std::memcpy(out, back.pixels.data(), back.pixels.size() * sizeof(float));
	return 0;
}

/* The reference's OBJ loader as Resources::getMesh drives it (ResourcesManager.cpp:316-328), on a text in memory:
 * Mesh(std::istream &, Mesh::Load, name), optionally computeNormals(), then computeBoundingBox(). mode follows
 * Mesh::Load's order (Expanded, Points, Indexed). */
void * ref_mesh_load(const char * text, size_t nbytes, int mode, int computeNormals) {
	std::stringstream in(std::string(text, nbytes));
	Mesh * m = new Mesh(in, static_cast<Mesh::Load>(mode), "harness");
	if(computeNormals) m->computeNormals();
	m->computeBoundingBox();
	return m;
}
void ref_mesh_sizes(void * h, uint64_t out[4]) {
	const Mesh * m = static_cast<const Mesh *>(h);
	out[0] = m->positions.size();
	out[1] = m->normals.size();
	out[2] = m->texcoords.size();
	out[3] = m->indices.size();
}
void ref_mesh_read(void * h, float * positions, float * normals, float * texcoords, uint32_t * indices, float * bbox) {
	const Mesh * m = static_cast<const Mesh *>(h);
	if(!m->positions.empty()) std::memcpy(positions, m->positions.data(), m->positions.size() * sizeof(glm::vec3));
	if(!m->normals.empty()) std::memcpy(normals, m->normals.data(), m->normals.size() * sizeof(glm::vec3));
	if(!m->texcoords.empty()) std::memcpy(texcoords, m->texcoords.data(), m->texcoords.size() * sizeof(glm::vec2));
	for(size_t i = 0; i < m->indices.size(); ++i) indices[i] = m->indices[i];
	for(int k = 0; k < 3; ++k) {
		bbox[k]		= m->bbox.minis[k];
		bbox[3 + k] = m->bbox.maxis[k];
	}
}
void ref_mesh_free(void * h) {
	delete static_cast<Mesh *>(h);
}

uint64_t ref_ray_counter() {
	return g_rayCount.load();
}

void ref_ray_counter_reset() {
	g_rayCount.store(0);
}

/* Linked with -Wl,--wrap=<mangled Raycaster::intersects(vec3,vec3,float,float)>: counts every closest-hit query the
 * reference path tracer issues (including the ones inside checkVisibility). */
Raycaster::Hit __real__ZNK9Raycaster10intersectsERKN3glm3vecILi3EfLNS0_9qualifierE0EEES5_ff(const Raycaster *, const glm::vec3 &, const glm::vec3 &, float, float);
Raycaster::Hit __wrap__ZNK9Raycaster10intersectsERKN3glm3vecILi3EfLNS0_9qualifierE0EEES5_ff(const Raycaster * self, const glm::vec3 & o, const glm::vec3 & d, float mini, float maxi) {
	/* Per-thread tally flushed on thread exit (forParallel joins its workers before render() returns). */
	struct Tally {
		uint64_t n = 0;
		~Tally() { g_rayCount.fetch_add(n, std::memory_order_relaxed); }
	};
	static thread_local Tally tally;
	++tally.n;
	return __real__ZNK9Raycaster10intersectsERKN3glm3vecILi3EfLNS0_9qualifierE0EEES5_ff(self, o, d, mini, maxi);
}

} // extern "C"
