// rb_scene.h — host-side scene and path-tracer handles (internal).
#pragma once
#include "rb_internal.h"
#include "rb_shade.cuh"

struct RbHostObject {
	std::vector<float> positions; // model space, handed to the raycaster at commit
	float model[16];
	size_t nverts = 0, nindices = 0;
	uint32_t vertOffset = 0, indexOffset = 0;
	RbObject dev;
	bool transformDirty = false; // model changed after the commit: vertices and normal matrix follow at the next sync
};

struct rb_scene {
	std::vector<float> texels; // RGBA fp32 pool
	std::vector<RbTexInfo> textures;
	std::vector<RbHostObject> objects;
	std::vector<RbLight> lights;
	std::vector<float> vertexAttr; // 12 floats per vertex
	std::vector<uint32_t> indices;
	uint32_t bgMode	 = RB_BG_COLOR;
	float bgColor[3] = {0, 0, 0};
	float sunDir[3]	 = {0, 1, 0};
	int32_t bgTex[6] = {-1, -1, -1, -1, -1, -1};
	int32_t skyLut	 = -1;
	bool committed	 = false;
	bool geometryCommitted = false; // objects/textures are on the device; lights / background may still be stale
	uint64_t rebuilds = 0;			// times rb_scene_sync chose a rebuild over a refit
	uint64_t editSerial = 0;		// bumped by every call that changes the scene (replicas on other GPUs are keyed to it)
	uint64_t triShadeSerial = 0;	// rb_raycaster::buildSerial the per-triangle shading records were gathered for
	bool tablesDirty = false;		// lights or background changed since the last upload (rb_scene_sync)

	rb_raycaster * raycaster = nullptr;
	float * d_texels		 = nullptr;
	RbTexInfo * d_textures	 = nullptr;
	RbObject * d_objects	 = nullptr;
	RbLight * d_lights		 = nullptr;
	float * d_vertexAttr	 = nullptr;
	float4 * d_triShade		 = nullptr; // 9 float4 per triangle, hierarchy order
	uint32_t * d_indices	 = nullptr;
	uint32_t * d_objVertOffset	= nullptr;
	uint32_t * d_objIndexOffset = nullptr;
	RbSceneDev dev = {};
};

struct rb_pathtracer;
// rb_multi.cu
rb_status rb_pt_multi_render(rb_pathtracer * pt, const rb_camera * camera, uint32_t width, uint32_t height, uint32_t samples, uint32_t depth, uint32_t components, float * pixels, const rb_render_opts * opts);
void rb_pt_multi_destroy(void * multi);
// Host side of a scene copied into a new, uncommitted handle (no device state): what a replica on another GPU starts from.
rb_status rb_scene_clone_host(const rb_scene * sc, rb_scene ** out);
