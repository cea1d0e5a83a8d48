// rb_scene.cu — the read-only scene data PathTracer::render touches, flattened into device tables.
// Reference: Scene::objects/lights/background (src/engine/scene/Scene.hpp:65-78), Object::mesh/model/material/
// useTexCoords (Object.hpp:54-82), Material::textures/type/masked (Material.hpp:111-132), Mesh arrays (Mesh.hpp:126-132).
#include "rb_scene.h"

#include <cmath>
#include <cstring>

extern "C" rb_status rb_scene_create(rb_scene ** out) {
	if(!out) return rb_fail(RB_ERR_INVALID, "rb_scene_create: null output");
	*out = new rb_scene();
	return RB_OK;
}

static void rb_scene_free_device(rb_scene * sc) {
	if(!rb_have_device()) return;
	cudaFree(sc->d_texels);
	cudaFree(sc->d_textures);
	cudaFree(sc->d_objects);
	cudaFree(sc->d_lights);
	cudaFree(sc->d_vertexAttr);
	cudaFree(sc->d_triShade);
	cudaFree(sc->d_indices);
	cudaFree(sc->d_objVertOffset);
	cudaFree(sc->d_objIndexOffset);
	sc->d_texels = nullptr;
	sc->d_textures = nullptr;
	sc->d_objects = nullptr;
	sc->d_lights = nullptr;
	sc->d_vertexAttr = nullptr;
	sc->d_triShade = nullptr;
	sc->d_indices = nullptr;
	sc->d_objVertOffset = nullptr;
	sc->d_objIndexOffset = nullptr;
}

extern "C" void rb_scene_destroy(rb_scene * sc) {
	if(!sc) return;
	rb_scene_free_device(sc);
	if(sc->raycaster) rb_raycaster_destroy(sc->raycaster);
	delete sc;
}

extern "C" rb_status rb_scene_add_texture(rb_scene * sc, const float * pixels, uint32_t width, uint32_t height, uint32_t components, int32_t * out_id) {
	if(sc) sc->editSerial++;
	if(!sc || !pixels || !out_id) return rb_fail(RB_ERR_INVALID, "rb_scene_add_texture: null argument");
	if(width == 0 || height == 0 || components < 1 || components > 4) return rb_fail(RB_ERR_INVALID, "rb_scene_add_texture: bad dimensions %ux%ux%u", width, height, components);
	RbTexInfo ti;
	ti.offset = uint32_t(sc->texels.size() / 4);
	ti.width  = width;
	ti.height = height;
	ti._pad	  = 0;
	const size_t n = size_t(width) * height;
	if((sc->texels.size() / 4 + n) > 0xffffffffull) return rb_fail(RB_ERR_UNSUPPORTED, "rb_scene_add_texture: texel pool exceeds 2^32 texels");
	sc->texels.resize(sc->texels.size() + n * 4);
	float * dst = sc->texels.data() + size_t(ti.offset) * 4;
	for(size_t i = 0; i < n; ++i) {
		for(uint32_t c = 0; c < 4; ++c) {
			dst[4 * i + c] = (c < components) ? pixels[i * components + c] : (c == 3 ? 1.0f : 0.0f);
		}
	}
	*out_id = int32_t(sc->textures.size());
	sc->textures.push_back(ti);
	// The texel pool on the device is uploaded by rb_scene_commit only: a texture added afterwards exists on the host alone,
	// so the scene counts as uncommitted until the next commit (rb_scene_sync and rendering refuse it until then).
	sc->committed		  = false;
	sc->geometryCommitted = false;
	return RB_OK;
}

// invtp = glm::inverse(glm::transpose(glm::mat3(model))) (PathTracer.cpp:94), evaluated once with glm's formulas
// (func_matrix.inl:87-105 transpose, :269-291 inverse) so the values equal what the reference recomputes per hit.
static void rb_normal_matrix(const float * model, float * invtp) {
	float m[3][3]; // m[col][row] of the TRANSPOSED upper-left 3x3
	for(int c = 0; c < 3; ++c) {
		for(int r = 0; r < 3; ++r) m[c][r] = model[r * 4 + c]; // transpose: result[c][r] = model[r][c]
	}
	const float oneOverDet = 1.0f / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
	float inv[3][3];
	inv[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * oneOverDet;
	inv[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * oneOverDet;
	inv[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * oneOverDet;
	inv[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * oneOverDet;
	inv[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * oneOverDet;
	inv[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * oneOverDet;
	inv[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * oneOverDet;
	inv[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * oneOverDet;
	inv[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * oneOverDet;
	for(int c = 0; c < 3; ++c) {
		for(int r = 0; r < 3; ++r) invtp[c * 3 + r] = inv[c][r];
	}
}

extern "C" rb_status rb_scene_add_object(rb_scene * sc, const rb_object_desc * d) {
	if(sc) sc->editSerial++;
	if(!sc || !d) return rb_fail(RB_ERR_INVALID, "rb_scene_add_object: null argument");
	if(!d->positions || !d->indices || !d->normals || !d->tangents || !d->bitangents) {
		// PathTracer::PathTracer logs "requires local tangent frames for all meshes" (PathTracer.cpp:14-16); here it is an error.
		return rb_fail(RB_ERR_INVALID, "rb_scene_add_object: positions, indices, normals, tangents and bitangents are required");
	}
	if(d->use_texcoords && !d->texcoords) return rb_fail(RB_ERR_INVALID, "rb_scene_add_object: use_texcoords set without texcoords");
	const int32_t ntex = int32_t(sc->textures.size());
	if(d->tex_albedo < 0 || d->tex_albedo >= ntex) return rb_fail(RB_ERR_INVALID, "rb_scene_add_object: albedo texture id %d out of range", d->tex_albedo);
	if(!d->emissive && (d->tex_rmao < 0 || d->tex_rmao >= ntex)) return rb_fail(RB_ERR_INVALID, "rb_scene_add_object: rmao texture id %d out of range", d->tex_rmao);
	if(!d->emissive && d->use_texcoords && (d->tex_normal < 0 || d->tex_normal >= ntex)) return rb_fail(RB_ERR_INVALID, "rb_scene_add_object: normal texture id %d out of range", d->tex_normal);
	{
		std::atomic<size_t> bad{SIZE_MAX};
		rb_parallel_for(d->nindices, size_t(1) << 18, [&](size_t begin, size_t end) {
			for(size_t i = begin; i < end; ++i) {
				if(d->indices[i] >= d->nverts) {
					size_t seen = bad.load();
					while(i < seen && !bad.compare_exchange_weak(seen, i)) {}
					return;
				}
			}
		});
		if(bad.load() != SIZE_MAX) return rb_fail(RB_ERR_INVALID, "rb_scene_add_object: index %zu out of range", bad.load());
	}
	RbHostObject o;
	o.vertOffset  = uint32_t(sc->vertexAttr.size() / 12);
	o.indexOffset = uint32_t(sc->indices.size());
	o.nverts	  = d->nverts;
	o.nindices	  = d->nindices;
	o.positions.assign(d->positions, d->positions + d->nverts * 3);
	std::memcpy(o.model, d->model, sizeof(o.model));
	sc->indices.insert(sc->indices.end(), d->indices, d->indices + d->nindices);
	sc->vertexAttr.resize(sc->vertexAttr.size() + d->nverts * 12);
	float * va = sc->vertexAttr.data() + size_t(o.vertOffset) * 12;
	rb_parallel_for(d->nverts, size_t(1) << 16, [&](size_t begin, size_t end) {
		for(size_t v = begin; v < end; ++v) {
			float * p = va + v * 12;
			p[0] = d->normals[3 * v]; p[1] = d->normals[3 * v + 1]; p[2] = d->normals[3 * v + 2];
			p[3] = d->tangents[3 * v]; p[4] = d->tangents[3 * v + 1]; p[5] = d->tangents[3 * v + 2];
			p[6] = d->bitangents[3 * v]; p[7] = d->bitangents[3 * v + 1]; p[8] = d->bitangents[3 * v + 2];
			p[9]  = d->texcoords ? d->texcoords[2 * v] : 0.0f;
			p[10] = d->texcoords ? d->texcoords[2 * v + 1] : 0.0f;
			p[11] = 0.0f;
		}
	});
	rb_normal_matrix(d->model, o.dev.invtp);
	o.dev.texAlbedo = d->tex_albedo;
	o.dev.texNormal = d->tex_normal;
	o.dev.texRmao	= d->tex_rmao;
	o.dev.flags		= (d->emissive ? RB_OBJ_EMISSIVE : 0u) | (d->masked ? RB_OBJ_MASKED : 0u) | (d->use_texcoords ? RB_OBJ_USE_UV : 0u);
	sc->objects.push_back(std::move(o));
	sc->committed		  = false;
	sc->geometryCommitted = false;
	return RB_OK;
}

extern "C" rb_status rb_scene_add_light(rb_scene * sc, const rb_light_desc * d) {
	if(sc) sc->editSerial++;
	if(!sc || !d) return rb_fail(RB_ERR_INVALID, "rb_scene_add_light: null argument");
	if(d->kind > RB_LIGHT_DIRECTIONAL) return rb_fail(RB_ERR_INVALID, "rb_scene_add_light: unknown kind %u", d->kind);
	RbLight l;
	std::memset(&l, 0, sizeof(l));
	l.kind		  = d->kind;
	l.castsShadow = d->casts_shadow;
	std::memcpy(l.intensity, d->intensity, sizeof(l.intensity));
	std::memcpy(l.position, d->position, sizeof(l.position));
	std::memcpy(l.direction, d->direction, sizeof(l.direction));
	l.radius = d->radius;
	// SpotLight::sample (SpotLight.cpp:66-73) calls std::cos on the host for every sample; the three values are constants
	// of the light, computed here with the same libm so they carry the same bits.
	l.outerCos	  = std::cos(d->outer_angle);
	l.cosOuterCos = std::cos(l.outerCos);
	l.innerCos	  = std::cos(d->inner_angle);
	sc->lights.push_back(l);
	sc->tablesDirty = true;
	if(!sc->geometryCommitted) sc->committed = false;
	return RB_OK;
}

extern "C" rb_status rb_scene_set_background(rb_scene * sc, uint32_t mode, const float color[3], const int32_t * textures, uint32_t ntextures, const float sun_direction[3], int32_t sky_lut) {
	if(sc) sc->editSerial++;
	if(!sc) return rb_fail(RB_ERR_INVALID, "rb_scene_set_background: null handle");
	if(mode > RB_BG_ATMOSPHERE) return rb_fail(RB_ERR_INVALID, "rb_scene_set_background: unknown mode %u", mode);
	const uint32_t need = mode == RB_BG_IMAGE ? 1u : (mode == RB_BG_SKYBOX ? 6u : 0u);
	if(ntextures < need || (need && !textures)) return rb_fail(RB_ERR_INVALID, "rb_scene_set_background: mode %u needs %u texture ids", mode, need);
	for(uint32_t k = 0; k < need; ++k) {
		if(textures[k] < 0 || textures[k] >= int32_t(sc->textures.size())) return rb_fail(RB_ERR_INVALID, "rb_scene_set_background: texture id %d out of range", textures[k]);
	}
	if(mode == RB_BG_ATMOSPHERE && (sky_lut < 0 || sky_lut >= int32_t(sc->textures.size()))) return rb_fail(RB_ERR_INVALID, "rb_scene_set_background: ATMOSPHERE needs the scattering table texture");
	sc->bgMode = mode;
	for(int k = 0; k < 3; ++k) {
		sc->bgColor[k] = color ? color[k] : 0.0f;
		sc->sunDir[k]  = sun_direction ? sun_direction[k] : 0.0f;
	}
	for(uint32_t k = 0; k < 6; ++k) sc->bgTex[k] = (k < need) ? textures[k] : -1;
	sc->skyLut		= sky_lut;
	sc->tablesDirty = true;
	if(!sc->geometryCommitted) sc->committed = false;
	return RB_OK;
}

// Per triangle, in hierarchy order: the packed attributes of its three vertices, gathered once here so that shading
// reads one contiguous 144-byte record per hit instead of hit.localId -> indices[] -> three vertices
// (Raycaster.hpp:91-95). The mesh id rides in the spare component of the first vertex.
__global__ void rb_k_build_tri_shade(const float4 * __restrict__ tris, size_t ntris, const uint32_t * __restrict__ indices, const uint32_t * __restrict__ objVertOffset, const uint32_t * __restrict__ objIndexOffset,
									 const float4 * __restrict__ vertexAttr, float4 * __restrict__ out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= ntris) return;
	const uint32_t localId = __float_as_uint(tris[3 * i].w);
	const uint32_t meshId  = __float_as_uint(tris[3 * i + 1].w);
	const uint32_t * idx   = indices + objIndexOffset[meshId] + localId;
	const uint32_t vo	   = objVertOffset[meshId];
	for(int k = 0; k < 3; ++k) {
		const float4 * v = vertexAttr + 3 * size_t(vo + idx[k]);
		float4 v2		 = v[2];
		if(k == 0) v2.w = __uint_as_float(meshId);
		out[9 * i + 3 * k + 0] = v[0];
		out[9 * i + 3 * k + 1] = v[1];
		out[9 * i + 3 * k + 2] = v2;
	}
}

template<typename T>
static rb_status rb_upload(T ** dst, const void * src, size_t bytes) {
	RB_CUDA(cudaMalloc(dst, bytes + 64));
	if(bytes) RB_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
	return RB_OK;
}

extern "C" rb_status rb_scene_commit(rb_scene * sc, rb_build_mode mode) {
	if(sc) sc->editSerial++;
	if(!sc) return rb_fail(RB_ERR_INVALID, "rb_scene_commit: null handle");
	RB_REQUIRE_DEVICE();
	if(sc->objects.empty()) return rb_fail(RB_ERR_INVALID, "rb_scene_commit: scene has no object");
	rb_status st;
	// Raycaster: addMesh per object in scene order, then updateHierarchy (PathTracer.cpp:11-21).
	if(sc->raycaster) rb_raycaster_destroy(sc->raycaster);
	sc->raycaster = nullptr;
	st = rb_raycaster_create(&sc->raycaster);
	if(st) return st;
	for(const RbHostObject & o : sc->objects) {
		st = rb_raycaster_add_mesh(sc->raycaster, o.positions.data(), o.nverts, sc->indices.data() + o.indexOffset, o.nindices, o.model);
		if(st) return st;
	}
	st = rb_raycaster_build(sc->raycaster, mode);
	if(st) return st;

	rb_scene_free_device(sc);
	std::vector<RbObject> objs;
	std::vector<uint32_t> vertOff, idxOff;
	bool anyMasked = false;
	for(const RbHostObject & o : sc->objects) {
		objs.push_back(o.dev);
		vertOff.push_back(o.vertOffset);
		idxOff.push_back(o.indexOffset);
		anyMasked = anyMasked || ((o.dev.flags & RB_OBJ_MASKED) && (o.dev.flags & RB_OBJ_USE_UV));
	}
	if((st = rb_upload(&sc->d_texels, sc->texels.data(), sc->texels.size() * sizeof(float)))) return st;
	if((st = rb_upload(&sc->d_textures, sc->textures.data(), sc->textures.size() * sizeof(RbTexInfo)))) return st;
	if((st = rb_upload(&sc->d_objects, objs.data(), objs.size() * sizeof(RbObject)))) return st;
	if((st = rb_upload(&sc->d_lights, sc->lights.data(), sc->lights.size() * sizeof(RbLight)))) return st;
	if((st = rb_upload(&sc->d_vertexAttr, sc->vertexAttr.data(), sc->vertexAttr.size() * sizeof(float)))) return st;
	if((st = rb_upload(&sc->d_indices, sc->indices.data(), sc->indices.size() * sizeof(uint32_t)))) return st;
	if((st = rb_upload(&sc->d_objVertOffset, vertOff.data(), vertOff.size() * sizeof(uint32_t)))) return st;
	if((st = rb_upload(&sc->d_objIndexOffset, idxOff.data(), idxOff.size() * sizeof(uint32_t)))) return st;
	const size_t ntris = sc->raycaster->nTris;
	RB_CUDA(cudaMalloc(&sc->d_triShade, ntris * 9 * sizeof(float4) + 64));
	if(ntris) {
		rb_k_build_tri_shade<<<unsigned((ntris + 255) / 256), 256>>>(sc->raycaster->d_tris, ntris, sc->d_indices, sc->d_objVertOffset, sc->d_objIndexOffset,
																	 reinterpret_cast<const float4 *>(sc->d_vertexAttr), sc->d_triShade);
		RB_CUDA(cudaGetLastError());
		RB_CUDA(cudaDeviceSynchronize());
	}
	RbSceneDev & dv = sc->dev;
	dv.texels	  = reinterpret_cast<const float4 *>(sc->d_texels);
	dv.textures	  = sc->d_textures;
	dv.objects	  = sc->d_objects;
	dv.lights	  = sc->d_lights;
	dv.triShade	  = sc->d_triShade;
	dv.nLights	  = int(sc->lights.size());
	dv.bgMode	  = sc->bgMode;
	for(int k = 0; k < 3; ++k) {
		dv.bgColor[k] = sc->bgColor[k];
		dv.sunDir[k]  = sc->sunDir[k];
	}
	for(int k = 0; k < 6; ++k) dv.bgTex[k] = sc->bgTex[k];
	dv.skyLut	 = sc->skyLut;
	dv.anyMasked = anyMasked ? 1u : 0u;
	sc->triShadeSerial	  = sc->raycaster->buildSerial;
	sc->committed		  = true;
	sc->geometryCommitted = true;
	sc->tablesDirty		  = false;
	return RB_OK;
}

extern "C" rb_status rb_scene_clear_lights(rb_scene * sc) {
	if(sc) sc->editSerial++;
	if(!sc) return rb_fail(RB_ERR_INVALID, "rb_scene_clear_lights: null handle");
	sc->lights.clear();
	sc->tablesDirty = true;
	return RB_OK;
}

extern "C" rb_status rb_scene_set_object_transform(rb_scene * sc, uint32_t object, const float model[16]) {
	if(sc) sc->editSerial++;
	if(!sc || !model) return rb_fail(RB_ERR_INVALID, "rb_scene_set_object_transform: null argument");
	if(object >= sc->objects.size()) return rb_fail(RB_ERR_INVALID, "rb_scene_set_object_transform: object %u out of range (%zu objects)", object, sc->objects.size());
	RbHostObject & o = sc->objects[object];
	std::memcpy(o.model, model, sizeof(o.model));
	rb_normal_matrix(o.model, o.dev.invtp);
	if(sc->geometryCommitted) {
		o.transformDirty = true;
		sc->tablesDirty	 = true;
	}
	return RB_OK;
}

// Lights and background of a committed scene: the reference reads them through the Scene at render time
// (PathTracer.cpp:23-55, :232-255), so a caller may edit them between two renders. Re-uploads the two small tables;
// geometry, textures and the hierarchy are untouched.
extern "C" rb_status rb_scene_sync(rb_scene * sc) {
	if(!sc) return rb_fail(RB_ERR_INVALID, "rb_scene_sync: null handle");
	RB_REQUIRE_DEVICE();
	if(!sc->geometryCommitted) return rb_fail(RB_ERR_INVALID, "rb_scene_sync: call rb_scene_commit first");
	if(!sc->tablesDirty) return RB_OK;
	if(sc->bgMode == RB_BG_IMAGE || sc->bgMode == RB_BG_SKYBOX) {
		const int need = sc->bgMode == RB_BG_IMAGE ? 1 : 6;
		for(int k = 0; k < need; ++k) {
			if(sc->bgTex[k] < 0 || sc->bgTex[k] >= int32_t(sc->textures.size())) return rb_fail(RB_ERR_INVALID, "rb_scene_sync: background texture id %d out of range", sc->bgTex[k]);
		}
	}
	RB_CUDA(cudaSetDevice(sc->raycaster->device));
	RB_CUDA(cudaDeviceSynchronize()); // a render may still be reading the old tables
	// Objects whose model matrix changed: vertices re-transformed, one refit for all of them, normal matrices rewritten.
	bool moved = false;
	for(size_t k = 0; k < sc->objects.size(); ++k) {
		RbHostObject & o = sc->objects[k];
		if(!o.transformDirty) continue;
		rb_status us = rb_raycaster_update_mesh(sc->raycaster, uint32_t(k), o.positions.data(), o.nverts, o.model);
		if(us) return us;
		RB_CUDA(cudaMemcpy(sc->d_objects + k, &o.dev, sizeof(RbObject), cudaMemcpyHostToDevice));
		o.transformDirty = false;
		moved			 = true;
	}
	if(moved) {
		rb_status fs = rb_raycaster_refit(sc->raycaster, nullptr, 0);
		if(fs) return fs;
		// Refit or rebuild? The refit keeps the topology; once the boxes of the re-fitted tree offer twice the surface they
		// did when the tree was built (rb_raycaster_tree_cost), traversal pays more every frame than one rebuild costs: the
		// scene is committed again from the current matrices — the tree a fresh scene with these transforms would get.
		double cost[2] = {0.0, 0.0};
		fs = rb_raycaster_tree_cost(sc->raycaster, cost);
		if(fs) return fs;
		const char * e	   = std::getenv("RB_REBUILD_RATIO");
		const double limit = e ? std::atof(e) : 2.0;
		if(limit > 0.0 && cost[0] > 0.0 && cost[1] > limit * cost[0]) {
			sc->rebuilds++;
			return rb_scene_commit(sc, rb_build_mode(sc->raycaster->buildMode));
		}
	}
	cudaFree(sc->d_lights);
	sc->d_lights = nullptr;
	rb_status st = rb_upload(&sc->d_lights, sc->lights.data(), sc->lights.size() * sizeof(RbLight));
	if(st) return st;
	RbSceneDev & dv = sc->dev;
	dv.lights  = sc->d_lights;
	dv.nLights = int(sc->lights.size());
	dv.bgMode  = sc->bgMode;
	for(int k = 0; k < 3; ++k) {
		dv.bgColor[k] = sc->bgColor[k];
		dv.sunDir[k]  = sc->sunDir[k];
	}
	for(int k = 0; k < 6; ++k) dv.bgTex[k] = sc->bgTex[k];
	dv.skyLut		= sc->skyLut;
	sc->tablesDirty = false;
	sc->committed	= true;
	return RB_OK;
}

rb_status rb_scene_clone_host(const rb_scene * sc, rb_scene ** out) {
	if(!sc || !out) return rb_fail(RB_ERR_INVALID, "rb_scene_clone_host: null argument");
	rb_scene * r = new rb_scene(*sc); // host vectors are copied; everything that names device memory is reset below
	r->raycaster		 = nullptr;
	r->d_texels			 = nullptr;
	r->d_textures		 = nullptr;
	r->d_objects		 = nullptr;
	r->d_lights			 = nullptr;
	r->d_vertexAttr		 = nullptr;
	r->d_triShade		 = nullptr;
	r->d_indices		 = nullptr;
	r->d_objVertOffset	 = nullptr;
	r->d_objIndexOffset	 = nullptr;
	r->dev				 = RbSceneDev{};
	r->committed		 = false;
	r->geometryCommitted = false;
	r->tablesDirty		 = false;
	for(RbHostObject & o : r->objects) o.transformDirty = false; // the clone is committed from the current matrices
	*out = r;
	return RB_OK;
}

extern "C" rb_raycaster * rb_scene_raycaster(rb_scene * sc) {
	return sc ? sc->raycaster : nullptr;
}
