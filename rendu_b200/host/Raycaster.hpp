#pragma once
// Drop-in for the reference's src/engine/raycaster/Raycaster.hpp: same class name, same public surface, same private
// mirror members (RaycasterVisualisation, a friend, reads them directly), but every query runs on the GPU through the
// C ABI of include/rendu_b200.h. Put this directory ahead of src/engine on the include path and compile
// Raycaster.cpp from here instead of the reference's.
#include "resources/Mesh.hpp"
#include "resources/Bounds.hpp"
#include "Common.hpp"
#include "raycaster/Intersection.hpp"

struct rb_raycaster;

class Raycaster {
	friend class RaycasterVisualisation;
	friend class PathTracer;

public:
	// Result of a query; layout and meaning as in the reference (hit flag, distance, barycentrics, ids).
	struct Hit {
		friend class RaycasterVisualisation;
		bool hit;
		float dist;
		float u, v, w;
		unsigned long localId; // first index of the triangle in its mesh's index buffer
		unsigned long meshId;  // order of the addMesh call
		Hit();
		Hit(float distance, float uu, float vv, unsigned long lid, unsigned long mid);
	private:
		unsigned long internalId;
	};

	Raycaster();
	~Raycaster();
	Raycaster(const Raycaster &) = delete;
	Raycaster & operator=(const Raycaster &) = delete;
	Raycaster(Raycaster &&) = delete;
	Raycaster & operator=(Raycaster &&) = delete;

	void addMesh(const Mesh & mesh, const glm::mat4 & model);
	void updateHierarchy();
	Hit intersects(const glm::vec3 & origin, const glm::vec3 & direction, float mini = 0.0001f, float maxi = 1e8f) const;
	bool intersectsAny(const glm::vec3 & origin, const glm::vec3 & direction, float mini = 0.0001f, float maxi = 1e8f) const;
	bool visible(const glm::vec3 & p0, const glm::vec3 & p1) const;

	// Batch extensions (not in the reference): n rays per call, arrays of n elements.
	void intersects(const glm::vec3 * origins, const glm::vec3 * directions, size_t n, Hit * out, float mini = 0.0001f, float maxi = 1e8f) const;
	void intersectsAny(const glm::vec3 * origins, const glm::vec3 * directions, size_t n, unsigned char * out, float mini = 0.0001f, float maxi = 1e8f) const;

	template<typename T>
	static T interpolateAttribute(const Hit & hit, const Mesh & geometry, const std::vector<T> & attribute) {
		const unsigned long first = hit.localId;
		const unsigned long a = geometry.indices[first], b = geometry.indices[first + 1], c = geometry.indices[first + 2];
		return hit.w * attribute[a] + hit.u * attribute[b] + hit.v * attribute[c];
	}

	// Build mode used by updateHierarchy: 0 = reference-identical tree (default), 1 = GPU LBVH. RENDU_B200_BUILD overrides.
	static int defaultBuildMode();

private:
	// Same records as the reference keeps; they mirror the device tree for RaycasterVisualisation.
	struct TriangleInfos {
		BoundingBox box;
		unsigned long v0 = 0, v1 = 0, v2 = 0;
		unsigned long localId = 0;
		unsigned int meshId = 0;
	};
	struct Node {
		BoundingBox box;
		size_t left = 0, right = 0;
		bool leaf = true;
	};
	Hit intersects(const Ray & ray, const TriangleInfos & tri, float mini, float maxi) const; // kept for RaycasterVisualisation

	void adopt(rb_raycaster * external); // PathTracer: use the scene's raycaster instead of an own one
	void refreshMirror();

	std::vector<TriangleInfos> _triangles;
	std::vector<glm::vec3> _vertices;
	std::vector<Node> _hierarchy;
	unsigned int _meshCount = 0;

	rb_raycaster * _device = nullptr;
	bool _ownsDevice = true;
};
