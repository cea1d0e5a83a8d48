// Drop-in for src/engine/raycaster/Raycaster.cpp: forwards to the CUDA library (include/rendu_b200.h). No CPU tracing.
#include "Raycaster.hpp"
#include "rendu_b200.h"
#include "system/Logger.hpp"

#include <cstdlib>
#include <cstring>

static_assert(sizeof(Raycaster::Hit) == sizeof(rb_hit), "Raycaster::Hit must stay bit-compatible with rb_hit");

static bool rbCheck(rb_status st, const char * what) {
	if(st == RB_OK) return true;
	Log::Error() << "[Raycaster] " << what << ": " << rb_last_error() << std::endl; // no exceptions on this path, like the reference
	return false;
}

Raycaster::Hit::Hit() : hit(false), dist(std::numeric_limits<float>::max()), u(0.0f), v(0.0f), w(0.0f), localId(0), meshId(0), internalId(0) {}
Raycaster::Hit::Hit(float distance, float uu, float vv, unsigned long lid, unsigned long mid) : hit(true), dist(distance), u(uu), v(vv), w(1.0f - uu - vv), localId(lid), meshId(mid), internalId(0) {}

int Raycaster::defaultBuildMode() {
	const char * e = std::getenv("RENDU_B200_BUILD");
	return (e && std::strcmp(e, "lbvh") == 0) ? RB_BUILD_LBVH : RB_BUILD_REFERENCE;
}

Raycaster::Raycaster() {
	rbCheck(rb_raycaster_create(&_device), "create");
}

Raycaster::~Raycaster() {
	if(_ownsDevice && _device) rb_raycaster_destroy(_device);
}

void Raycaster::adopt(rb_raycaster * external) {
	if(_ownsDevice && _device) rb_raycaster_destroy(_device);
	_device		= external;
	_ownsDevice = false;
	refreshMirror();
}

void Raycaster::addMesh(const Mesh & mesh, const glm::mat4 & model) {
	if(!_device) return;
	const size_t before = _vertices.size();
	if(!rbCheck(rb_raycaster_add_mesh(_device, mesh.positions.empty() ? nullptr : &mesh.positions[0][0], mesh.positions.size(), mesh.indices.data(), mesh.indices.size(), &model[0][0]), "addMesh")) return;
	_vertices.resize(before + mesh.positions.size()); // sized now, filled by refreshMirror
	Log::Info() << "[Raycaster] Mesh " << _meshCount << " added, " << mesh.indices.size() / 3 << " triangles, " << mesh.positions.size() << " vertices." << std::endl;
	++_meshCount;
}

void Raycaster::updateHierarchy() {
	if(!_device) return;
	uint64_t sz[5] = {0, 0, 0, 0, 0};
	rb_raycaster_sizes(_device, sz);
	Log::Info() << "[Raycaster] Building hierarchy for " << sz[1] << " triangles... " << std::flush;
	if(!rbCheck(rb_raycaster_build(_device, rb_build_mode(defaultBuildMode())), "updateHierarchy")) return;
	refreshMirror();
	// Traversal mode (extension): RENDU_B200_TRACE=ordered | wide; default is the reference's own order and bits. The
	// wide mode (experiment) collapses the tree just built into eight-wide quantised nodes first.
	if(const char * mode = std::getenv("RENDU_B200_TRACE")) {
		if(std::strcmp(mode, "ordered") == 0) rbCheck(rb_raycaster_set_trace_mode(_device, RB_TRACE_ORDERED), "trace mode");
		else if(std::strcmp(mode, "wide") == 0 && rbCheck(rb_raycaster_wide_build(_device, nullptr), "wide hierarchy")) rbCheck(rb_raycaster_set_trace_mode(_device, RB_TRACE_WIDE), "trace mode");
	}
	Log::Info() << "Done: " << _hierarchy.size() << " nodes created." << std::endl;
}

void Raycaster::refreshMirror() {
	uint64_t sz[5] = {0, 0, 0, 0, 0};
	if(!_device || !rbCheck(rb_raycaster_sizes(_device, sz), "sizes")) return;
	static_assert(sizeof(Node) == sizeof(rb_node) && sizeof(TriangleInfos) == sizeof(rb_tri), "mirror records must match the C ABI");
	_hierarchy.resize(sz[0]);
	_triangles.resize(sz[1]);
	_vertices.resize(sz[2]);
	_meshCount = (unsigned int)sz[3];
	rbCheck(rb_raycaster_download_bvh(_device, reinterpret_cast<rb_node *>(_hierarchy.data()), reinterpret_cast<rb_tri *>(_triangles.data()), _vertices.empty() ? nullptr : &_vertices[0][0]), "download");
}

Raycaster::Hit Raycaster::intersects(const glm::vec3 & origin, const glm::vec3 & direction, float mini, float maxi) const {
	Hit h;
	if(_device) rbCheck(rb_raycaster_intersect(_device, &origin[0], &direction[0], mini, maxi, 1, reinterpret_cast<rb_hit *>(&h)), "intersects");
	return h;
}

bool Raycaster::intersectsAny(const glm::vec3 & origin, const glm::vec3 & direction, float mini, float maxi) const {
	uint8_t r = 0;
	if(_device) rbCheck(rb_raycaster_intersect_any(_device, &origin[0], &direction[0], mini, maxi, 1, &r), "intersectsAny");
	return r != 0;
}

bool Raycaster::visible(const glm::vec3 & p0, const glm::vec3 & p1) const {
	uint8_t r = 0;
	if(_device) rbCheck(rb_raycaster_visible(_device, &p0[0], &p1[0], 1, &r), "visible");
	return r != 0;
}

void Raycaster::intersects(const glm::vec3 * origins, const glm::vec3 * directions, size_t n, Hit * out, float mini, float maxi) const {
	if(_device && n) rbCheck(rb_raycaster_intersect(_device, &origins[0][0], &directions[0][0], mini, maxi, n, reinterpret_cast<rb_hit *>(out)), "intersects(batch)");
}

void Raycaster::intersectsAny(const glm::vec3 * origins, const glm::vec3 * directions, size_t n, unsigned char * out, float mini, float maxi) const {
	if(_device && n) rbCheck(rb_raycaster_intersect_any(_device, &origins[0][0], &directions[0][0], mini, maxi, n, out), "intersectsAny(batch)");
}

// The private per-triangle test (Raycaster.cpp:232-262). Only the friend RaycasterVisualisation::getRayLevels calls it — on
// the triangles of the leaves ONE ray visits (RaycasterVisualisation.cpp:61-66) — so a one-pair launch per call is fine. The
// semantics are the reference's: this triangle alone against the ray and (mini, maxi), whatever else the ray would hit.
Raycaster::Hit Raycaster::intersects(const Ray & ray, const TriangleInfos & tri, float mini, float maxi) const {
	if(tri.v0 >= _vertices.size() || tri.v1 >= _vertices.size() || tri.v2 >= _vertices.size()) return Hit();
	const float verts[9] = {_vertices[tri.v0].x, _vertices[tri.v0].y, _vertices[tri.v0].z, _vertices[tri.v1].x, _vertices[tri.v1].y, _vertices[tri.v1].z,
							_vertices[tri.v2].x, _vertices[tri.v2].y, _vertices[tri.v2].z};
	float out[4] = {0.0f, 0.0f, 0.0f, 0.0f};
	if(!rbCheck(rb_kat_triangle_test(&ray.pos[0], &ray.dir[0], verts, mini, maxi, 1, out), "intersects(ray, triangle)") || out[0] == 0.0f) return Hit();
	return Hit(out[1], out[2], out[3], tri.localId, tri.meshId);
}
