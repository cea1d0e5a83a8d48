// Drop-in for src/apps/pathtracer/PathTracer.cpp. The constructor flattens what render() reads from the Scene
// (objects in scene->objects order, their meshes, materials, textures, lights, background) into the plain arrays of
// the C ABI; render() passes the camera quantities the reference derives on the host and gets the image back.
#include "PathTracer.hpp"
#include "rendu_b200.h"

#include "resources/ResourcesManager.hpp"
#include "scene/Sky.hpp"
#include "scene/lights/PointLight.hpp"
#include "scene/lights/SpotLight.hpp"
#include "scene/lights/DirectionalLight.hpp"
#include "system/Query.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

namespace {
bool ok(rb_status st, const char * what) {
	if(st == RB_OK) return true;
	Log::Error() << "[PathTracer] " << what << ": " << rb_last_error() << std::endl;
	return false;
}

// Textures are shared between materials: upload each (texture, image) once.
struct TextureTable {
	rb_scene * scene;
	std::map<std::pair<const Texture *, size_t>, int32_t> ids;
	int32_t get(const Texture * tex, size_t image = 0) {
		if(!tex || image >= tex->images.size()) return -1;
		const auto key = std::make_pair(tex, image);
		const auto it  = ids.find(key);
		if(it != ids.end()) return it->second;
		const Image & im = tex->images[image];
		int32_t id		 = -1;
		if(!ok(rb_scene_add_texture(scene, im.pixels.data(), im.width, im.height, im.components, &id), "texture upload")) return -1;
		ids[key] = id;
		return id;
	}
};
} // namespace

PathTracer::PathTracer(const std::shared_ptr<Scene> & scene) {
	_scene = scene;
	if(const char * e = std::getenv("RENDU_B200_GPUS")) _gpus = unsigned(std::atoi(e));
	if(const char * e = std::getenv("RENDU_B200_SPLIT")) _gpuSampleSplit = std::string(e) == "sample";
	if(!ok(rb_scene_create(&_deviceScene), "scene")) return;
	TextureTable textures{_deviceScene, {}};
	for(const auto & obj : scene->objects) {
		const Mesh & mesh = *obj.mesh();
		if(mesh.tangents.empty()) {
			Log::Error() << "The path tracer requires local tangent frames for all meshes." << std::endl;
		}
		const Material & mat = obj.material();
		rb_object_desc d;
		std::memset(&d, 0, sizeof(d));
		d.positions	 = mesh.positions.empty() ? nullptr : &mesh.positions[0][0];
		d.normals	 = mesh.normals.empty() ? nullptr : &mesh.normals[0][0];
		d.tangents	 = mesh.tangents.empty() ? nullptr : &mesh.tangents[0][0];
		d.bitangents = mesh.bitangents.empty() ? nullptr : &mesh.bitangents[0][0];
		d.texcoords	 = mesh.texcoords.size() == mesh.positions.size() ? &mesh.texcoords[0][0] : nullptr;
		d.nverts	 = mesh.positions.size();
		d.indices	 = mesh.indices.data();
		d.nindices	 = mesh.indices.size();
		std::memcpy(d.model, &obj.model()[0][0], sizeof(d.model));
		const auto & tex = mat.textures();
		d.tex_albedo	 = tex.size() > 0 ? textures.get(tex[0]) : -1;
		d.tex_normal	 = tex.size() > 1 ? textures.get(tex[1]) : -1;
		d.tex_rmao		 = tex.size() > 2 ? textures.get(tex[2]) : -1;
		d.emissive		 = mat.type() == Material::Type::Emissive ? 1u : 0u;
		d.masked		 = mat.masked() ? 1u : 0u;
		d.use_texcoords	 = (obj.useTexCoords() && d.texcoords) ? 1u : 0u;
		if(!ok(rb_scene_add_object(_deviceScene, &d), "object upload")) return;
	}
	if(!uploadLights()) return;
	{
		const Material & bg = scene->background->material();
		float color[3]		= {0.0f, 0.0f, 0.0f}, sun[3] = {0.0f, 1.0f, 0.0f};
		int32_t ids[6]		= {-1, -1, -1, -1, -1, -1};
		uint32_t nids		= 0;
		int32_t lut			= -1;
		const uint32_t mode = uint32_t(scene->backgroundMode);
		if(!bg.parameters().empty()) {
			for(int k = 0; k < 3; ++k) color[k] = bg.parameters()[0][k];
		}
		if(scene->backgroundMode == Scene::Background::IMAGE) {
			ids[0] = textures.get(bg.textures()[0]);
			nids   = 1;
		} else if(scene->backgroundMode == Scene::Background::SKYBOX) {
			for(size_t f = 0; f < 6; ++f) ids[f] = textures.get(bg.textures()[0], f);
			nids = 6;
		} else if(scene->backgroundMode == Scene::Background::ATMOSPHERE) {
			const Sky * sky = dynamic_cast<const Sky *>(scene->background.get());
			if(sky) {
				for(int k = 0; k < 3; ++k) sun[k] = sky->direction()[k];
			}
			lut = textures.get(Resources::manager().getTexture("scattering-precomputed", Layout::RGBA32F, Storage::CPU));
		}
		if(!ok(rb_scene_set_background(_deviceScene, mode, color, ids, nids, sun, lut), "background")) return;
	}
	// Same log lines as Raycaster::addMesh / updateHierarchy (Raycaster.cpp:48-49,56,141): scripts parse them.
	size_t totalTriangles = 0, meshId = 0;
	for(const auto & obj : scene->objects) {
		Log::Info() << "[Raycaster] Mesh " << meshId++ << " added, " << obj.mesh()->indices.size() / 3 << " triangles, " << obj.mesh()->positions.size() << " vertices." << std::endl;
		totalTriangles += obj.mesh()->indices.size() / 3;
	}
	Log::Info() << "[Raycaster] Building hierarchy for " << totalTriangles << " triangles... " << std::flush;
	if(!ok(rb_scene_commit(_deviceScene, rb_build_mode(Raycaster::defaultBuildMode())), "scene commit")) return;
	_raycaster.adopt(rb_scene_raycaster(_deviceScene));
	uint64_t sz[5] = {0, 0, 0, 0, 0};
	rb_raycaster_sizes(rb_scene_raycaster(_deviceScene), sz);
	Log::Info() << "Done: " << sz[0] << " nodes created." << std::endl;
	ok(rb_pathtracer_create(_deviceScene, &_device), "create");
}

PathTracer::~PathTracer() {
	if(_device) rb_pathtracer_destroy(_device);
	if(_deviceScene) rb_scene_destroy(_deviceScene); // also frees the raycaster _raycaster merely adopted
}

// scene->lights flattened for the device. Called at construction and again before every render: the reference reads
// the lights through the shared Scene inside render() (PathTracer.cpp:232-255), so edits made between two frames count.
bool PathTracer::uploadLights() {
	std::vector<rb_light_desc> descs;
	for(const auto & light : _scene->lights) {
		rb_light_desc d;
		std::memset(&d, 0, sizeof(d));
		d.casts_shadow = light->castsShadow() ? 1u : 0u;
		for(int k = 0; k < 3; ++k) d.intensity[k] = light->intensity()[k];
		if(const PointLight * p = dynamic_cast<const PointLight *>(light.get())) {
			d.kind = RB_LIGHT_POINT;
			for(int k = 0; k < 3; ++k) d.position[k] = p->position()[k];
			d.radius = p->radius();
		} else if(const SpotLight * s = dynamic_cast<const SpotLight *>(light.get())) {
			d.kind = RB_LIGHT_SPOT;
			for(int k = 0; k < 3; ++k) {
				d.position[k]  = s->position()[k];
				d.direction[k] = s->direction()[k];
			}
			d.radius	  = s->radius();
			d.inner_angle = s->angles()[0];
			d.outer_angle = s->angles()[1];
		} else if(const DirectionalLight * dl = dynamic_cast<const DirectionalLight *>(light.get())) {
			d.kind = RB_LIGHT_DIRECTIONAL;
			for(int k = 0; k < 3; ++k) d.direction[k] = dl->direction()[k];
		}
		descs.push_back(d);
	}
	// Unchanged since the last frame (the usual case): nothing to send, and the replicas on the other GPUs stay valid.
	const unsigned char * bytes = reinterpret_cast<const unsigned char *>(descs.data());
	std::vector<unsigned char> flat(bytes, bytes + descs.size() * sizeof(rb_light_desc));
	flat.push_back(1); // an empty light list still differs from "never uploaded"
	if(flat == _lightsUploaded) return true;
	if(!ok(rb_scene_clear_lights(_deviceScene), "light upload")) return false;
	for(const rb_light_desc & d : descs) {
		if(!ok(rb_scene_add_light(_deviceScene, &d), "light upload")) return false;
	}
	_lightsUploaded.swap(flat);
	return true;
}

void PathTracer::render(const Camera & camera, size_t samples, size_t depth, Image & render) {
	if(!_scene || !_device) {
		Log::Error() << "[PathTracer] No scene available." << std::endl;
		return;
	}
	if(render.components < 3) {
		Log::Warning() << "[PathTracer] Expected a RGB image." << std::endl;
		return;
	}
	const size_t rounded = size_t(std::pow(2, std::round(std::log2(float(samples)))));
	if(rounded != samples) {
		Log::Warning() << "[PathTracer] Non power-of-2 samples count. Using " << rounded << " instead." << std::endl;
	}
	glm::vec3 corner, dx, dy;
	camera.pixelShifts(corner, dx, dy);
	rb_camera cam;
	for(int k = 0; k < 3; ++k) {
		cam.eye[k]	  = camera.position()[k];
		cam.corner[k] = corner[k];
		cam.dx[k]	  = dx[k];
		cam.dy[k]	  = dy[k];
	}
	rb_render_opts opts;
	std::memset(&opts, 0, sizeof(opts));
	opts.seed		  = _seed;
	opts.tile_rows	  = _tile[0];
	opts.tile_index	  = _tile[1];
	opts.tile_count	  = _tile[2];
	opts.sample_index = _split[0];
	opts.sample_count = _split[1];
	opts.gpus		  = _gpus;
	opts.split		  = _gpuSampleSplit ? RB_SPLIT_SAMPLE : RB_SPLIT_TILE;
	if(!uploadLights()) return;
	Query timer;
	timer.begin();
	ok(rb_pathtracer_render(_device, &cam, render.width, render.height, uint32_t(samples), uint32_t(depth), render.components, render.pixels.data(), &opts), "render");
	timer.end();
	rb_render_stats st;
	std::memset(&st, 0, sizeof(st));
	rb_pathtracer_stats(_device, &st);
	const float seconds = float(timer.value()) / 1000000000.0f;
	Log::Info() << "[PathTracer] Rendering took " << seconds << "s at " << render.width << "x" << render.height << "." << std::endl;
	Log::Info() << "[PathTracer] " << (st.rays_closest + st.rays_shadow) << " rays, " << double(st.rays_closest + st.rays_shadow) / (st.device_seconds > 0.0 ? st.device_seconds : 1.0) / 1e6 << " Mrays/s on the device." << std::endl;
}

void PathTracer::cameraFor(const Camera & camera, rb_camera & cam) const {
	glm::vec3 corner, dx, dy;
	camera.pixelShifts(corner, dx, dy); // PathTracer.cpp:177
	for(int k = 0; k < 3; ++k) {
		cam.eye[k]	  = camera.position()[k];
		cam.corner[k] = corner[k];
		cam.dx[k]	  = dx[k];
		cam.dy[k]	  = dy[k];
	}
}

bool PathTracer::liveBegin(unsigned int width, unsigned int height) {
	return _device && ok(rb_pathtracer_live_begin(_device, width, height), "live begin");
}

bool PathTracer::livePass(const Camera & camera, size_t samples, size_t depth, size_t passCount) {
	if(!_device || !uploadLights()) return false;
	rb_camera cam;
	cameraFor(camera, cam);
	rb_render_opts opts;
	std::memset(&opts, 0, sizeof(opts));
	opts.seed = _seed;
	return ok(rb_pathtracer_live_pass(_device, &cam, uint32_t(samples), uint32_t(depth), uint32_t(passCount), &opts), "live pass");
}

unsigned int PathTracer::liveDisplay(std::vector<unsigned char> & pixels, unsigned int components) {
	if(!_device) return 0;
	uint32_t done = 0;
	// The session knows its own size; the caller's vector is sized by the first successful call through liveBegin's values.
	if(pixels.empty()) return 0;
	if(!ok(rb_pathtracer_live_display(_device, components, 1, 1, pixels.data(), &done), "live display")) return 0;
	return done;
}

