/* TEST INFRASTRUCTURE — second checker library (oracle/_ref/librendu_refvis.so): the reference's own, unmodified
 * RaycasterVisualisation (src/engine/raycaster/RaycasterVisualisation.cpp, compiled where it lies) run on a reference
 * Raycaster created by librendu_ref.so. It lives in its own library because ref_api.cpp defines a stand-in class of the
 * same name to reach Raycaster's private members (the friend hook). Nothing under rendu_b200/ may link or load it.
 *
 * getRayLevels / getAllLevels return display meshes: one Mesh per depth level, 8 corners per node box
 * (BoundingBox::getCorners order: corner 0 = minis, corner 7 = maxis). The harness hands the corners back as boxes. */
#include "raycaster/RaycasterVisualisation.hpp"

#include <cstdint>
#include <cstring>

namespace {
/* boxes: 6 floats per node, levels concatenated; levelCounts[l] = nodes drawn at depth l. Returns total nodes, or -1 if
 * a capacity is too small. */
long collect(const std::vector<Mesh> & meshes, uint32_t capLevels, uint64_t capNodes, uint32_t * levelCounts, float * boxes, uint32_t * nLevels) {
	if(meshes.size() > capLevels) return -1;
	*nLevels = uint32_t(meshes.size());
	uint64_t total = 0;
	for(size_t l = 0; l < meshes.size(); ++l) {
		const size_t n = meshes[l].positions.size() / 8;
		levelCounts[l] = uint32_t(n);
		if(total + n > capNodes) return -1;
		for(size_t k = 0; k < n; ++k) {
			const glm::vec3 & lo = meshes[l].positions[8 * k];
			const glm::vec3 & hi = meshes[l].positions[8 * k + 7];
			float * b			 = boxes + 6 * (total + k);
			b[0] = lo.x; b[1] = lo.y; b[2] = lo.z;
			b[3] = hi.x; b[4] = hi.y; b[5] = hi.z;
		}
		total += n;
	}
	return long(total);
}
} // namespace

extern "C" {

long refvis_ray_levels(const void * raycaster, const float * origin, const float * direction, float mini, float maxi, uint32_t capLevels, uint64_t capNodes,
	uint32_t * levelCounts, float * boxes, uint32_t * nLevels, float * hitFloats /* dist,u,v,w */, uint64_t * hitIds /* hit, localId, meshId */) {
	const Raycaster & rc = *static_cast<const Raycaster *>(raycaster);
	RaycasterVisualisation vis(rc);
	std::vector<Mesh> meshes;
	const Raycaster::Hit hit = vis.getRayLevels(glm::vec3(origin[0], origin[1], origin[2]), glm::vec3(direction[0], direction[1], direction[2]), meshes, mini, maxi);
	hitFloats[0] = hit.dist; hitFloats[1] = hit.u; hitFloats[2] = hit.v; hitFloats[3] = hit.w;
	hitIds[0] = hit.hit ? 1 : 0; hitIds[1] = hit.localId; hitIds[2] = hit.meshId;
	return collect(meshes, capLevels, capNodes, levelCounts, boxes, nLevels);
}

long refvis_all_levels(const void * raycaster, uint32_t capLevels, uint64_t capNodes, uint32_t * levelCounts, float * boxes, uint32_t * nLevels) {
	const Raycaster & rc = *static_cast<const Raycaster *>(raycaster);
	RaycasterVisualisation vis(rc);
	std::vector<Mesh> meshes;
	vis.getAllLevels(meshes);
	return collect(meshes, capLevels, capNodes, levelCounts, boxes, nLevels);
}

}
