/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Builds a reference `Scene` from plain arrays instead of a .scene file, so that the UNMODIFIED reference
 * PathTracer::render (src/apps/pathtracer/PathTracer.cpp:159-299) can be timed — by bench.py's cpu_baseline and
 * --impl reference legs — on the same synthetic C3/C4 scene the CUDA path renders (rendu_b200/workloads.py,
 * cornell_synthetic), which exists in no file the reference's loaders could read. Every object handed to the reference
 * is one of its own classes (Mesh, Texture, Image, Material, Object, PointLight, Scene, Camera, PathTracer) filled
 * through its public interface; the three members that only the .scene decoder can set (Object::_skipUVs,
 * Material::_masked, Scene::_backgroundMaterial) are reached by compiling THIS file with the access keywords
 * neutralised — the reference translation units themselves are compiled as they are (oracle/build_ref.sh).
 * Nothing under rendu_b200/ may link or load this library.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#define private public
#define protected public
#include "resources/ResourcesManager.hpp"
#include "resources/Texture.hpp"
#include "resources/Image.hpp"
#include "resources/Mesh.hpp"
#include "scene/Scene.hpp"
#include "scene/lights/PointLight.hpp"
#include "input/Camera.hpp"
#include "PathTracer.hpp"
#undef private
#undef protected

namespace {

struct BuiltScene {
	std::shared_ptr<Scene> scene;
	std::unique_ptr<PathTracer> tracer;
	std::deque<Mesh> meshes;
	std::deque<Texture> textures;
	std::deque<Material> materials;
};

double seconds() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

extern "C" {

void * ref_built_scene_create() {
	BuiltScene * b = new BuiltScene();
	b->scene.reset(new Scene("synthetic"));
	return b;
}

void ref_built_scene_destroy(void * h) {
	delete static_cast<BuiltScene *>(h);
}

/* One fp32 RGBA image per texture (what Codable::decodeTexture + loadLDR/loadHDR leave on the CPU); `layers` = 6 makes
 * a cubemap whose faces are images[0..5] (ResourcesManager.cpp:144). Returns the texture index. */
int ref_built_scene_add_texture(void * h, const float * rgba, uint32_t width, uint32_t height, uint32_t layers) {
	BuiltScene * b = static_cast<BuiltScene *>(h);
	b->textures.emplace_back("synthetic_" + std::to_string(b->textures.size()));
	Texture & t = b->textures.back();
	t.width	 = width;
	t.height = height;
	t.depth	 = layers;
	t.levels = 1;
	t.shape	 = layers == 6 ? TextureShape::Cube : TextureShape::D2;
	const size_t texels = size_t(width) * height * 4;
	for(uint32_t l = 0; l < layers; ++l) {
		t.images.emplace_back(width, height, 4);
		std::memcpy(t.images.back().pixels.data(), rgba + l * texels, texels * sizeof(float));
	}
	return int(b->textures.size()) - 1;
}

/* Object: arrays as Mesh holds them (Mesh.hpp:126-132), column-major model matrix, three texture indices (albedo,
 * normal, roughness-metalness-ao; Material.hpp:12-15), Material::Type, masked, useTexCoords. Objects must be added in
 * scene->objects order (Scene::init sorts by material type, Scene.cpp:101-104: the caller lists them that way). */
void ref_built_scene_add_object(void * h, const float * pos, const float * nor, const float * tan, const float * bit, const float * uv, size_t nverts,
	const uint32_t * idx, size_t nidx, const float * model16, const int * tex3, int materialType, int masked, int useUV) {
	BuiltScene * b = static_cast<BuiltScene *>(h);
	b->meshes.emplace_back("synthetic_mesh_" + std::to_string(b->meshes.size()));
	Mesh & m = b->meshes.back();
	m.positions.resize(nverts);
	m.normals.resize(nverts);
	m.tangents.resize(nverts);
	m.bitangents.resize(nverts);
	std::memcpy(m.positions.data(), pos, nverts * sizeof(glm::vec3));
	std::memcpy(m.normals.data(), nor, nverts * sizeof(glm::vec3));
	std::memcpy(m.tangents.data(), tan, nverts * sizeof(glm::vec3));
	std::memcpy(m.bitangents.data(), bit, nverts * sizeof(glm::vec3));
	if(uv) {
		m.texcoords.resize(nverts);
		std::memcpy(m.texcoords.data(), uv, nverts * sizeof(glm::vec2));
	}
	m.indices.assign(idx, idx + nidx);
	m.computeBoundingBox();
	b->materials.emplace_back(Material::Type(materialType));
	Material & mat = b->materials.back();
	for(int k = 0; k < 3; ++k) {
		if(tex3[k] >= 0) mat.addTexture(&b->textures[size_t(tex3[k])]);
	}
	mat._masked = masked != 0;
	b->scene->objects.emplace_back(&m, true);
	Object & o = b->scene->objects.back();
	glm::mat4 model;
	std::memcpy(&model[0][0], model16, 16 * sizeof(float));
	o.set(model);
	o._skipUVs = useUV == 0;
	// std::vector<Object> may have moved the earlier objects; materials live in a deque and stay put.
	o.setMaterial(&mat);
}

void ref_built_scene_add_point_light(void * h, const float * position3, const float * intensity3, float radius, int castsShadow) {
	BuiltScene * b = static_cast<BuiltScene *>(h);
	std::shared_ptr<PointLight> l(new PointLight(glm::vec3(position3[0], position3[1], position3[2]), glm::vec3(intensity3[0], intensity3[1], intensity3[2]), radius));
	l->setCastShadow(castsShadow != 0);
	b->scene->lights.push_back(l);
}

/* mode: Scene::Background (0 COLOR, 1 IMAGE, 2 SKYBOX); texture index for IMAGE / SKYBOX. */
void ref_built_scene_set_background(void * h, int mode, const float * color3, int texture) {
	BuiltScene * b = static_cast<BuiltScene *>(h);
	Scene & s		  = *b->scene;
	s.backgroundMode  = Scene::Background(mode);
	s._backgroundMaterial = Material(Material::Type::None);
	if(texture >= 0) s._backgroundMaterial.addTexture(&b->textures[size_t(texture)]);
	s._backgroundMaterial.addParameter(glm::vec4(color3[0], color3[1], color3[2], 1.0f));
	s.background.reset(new Object());
	s.background->setMaterial(&s._backgroundMaterial);
}

/* PathTracer::PathTracer(scene): addMesh per object + updateHierarchy (PathTracer.cpp:11-21). Returns its seconds. */
double ref_built_scene_finish(void * h) {
	BuiltScene * b = static_cast<BuiltScene *>(h);
	const double t0 = seconds();
	b->tracer.reset(new PathTracer(b->scene));
	return seconds() - t0;
}

/* PathTracer::render with a Camera posed from eye / center / up, vertical fov, ratio = w / h (Camera.cpp:10-49).
 * cam10: eye(3) center(3) up(3) fov. Returns seconds around render(). */
double ref_built_scene_render(void * h, const float * cam10, uint32_t w, uint32_t hgt, uint32_t samples, uint32_t depth, uint32_t components, float * pixels) {
	BuiltScene * b = static_cast<BuiltScene *>(h);
	Camera camera;
	camera.pose(glm::vec3(cam10[0], cam10[1], cam10[2]), glm::vec3(cam10[3], cam10[4], cam10[5]), glm::vec3(cam10[6], cam10[7], cam10[8]));
	camera.projection(float(w) / float(hgt), cam10[9], 0.01f, 100.0f);
	Image render(w, hgt, components);
	const double t0 = seconds();
	b->tracer->render(camera, samples, depth, render);
	const double dt = seconds() - t0;
	if(pixels) std::memcpy(pixels, render.pixels.data(), render.pixels.size() * sizeof(float));
	return dt;
}

} // extern "C"
