#pragma once
// Drop-in for the reference's src/apps/pathtracer/PathTracer.hpp: PathTracer(scene), render(camera, samples, depth,
// image), raycaster() — rendered by the CUDA wavefront path tracer through include/rendu_b200.h.
#include "Raycaster.hpp"
#include "scene/Scene.hpp"
#include "Common.hpp"

struct rb_scene;
struct rb_pathtracer;
struct rb_camera;

class PathTracer {
public:
	PathTracer() = default;
	explicit PathTracer(const std::shared_ptr<Scene> & scene);
	~PathTracer();
	PathTracer(const PathTracer &) = delete;
	PathTracer & operator=(const PathTracer &) = delete;

	// Accumulates into `render` (caller zero-fills) and divides by the sample count, like the reference.
	void render(const Camera & camera, size_t samples, size_t depth, Image & render);
	const Raycaster & raycaster() const { return _raycaster; }

	// Extensions: RNG seed of the counter-based generator, and the share of the frame this process renders
	// (tile split: rows (y / tileRows) % tileCount == tileIndex; sample split: cells sid % sampleCount == sampleIndex).
	void setSeed(unsigned long long seed) { _seed = seed; }
	void setTile(unsigned int rows, unsigned int index, unsigned int count) { _tile[0] = rows; _tile[1] = index; _tile[2] = count; }
	void setSampleSplit(unsigned int index, unsigned int count) { _split[0] = index; _split[1] = count; }
	// All GPUs of the box from this one render() call — what replaces System::forParallel (System.hpp:72-104): `count`
	// devices, frame split by image tile (sampleSplit = false: bit-identical to one GPU) or by stratification cell
	// (one NCCL reduce of the sums). Defaults come from the environment: RENDU_B200_GPUS, RENDU_B200_SPLIT=tile|sample.
	void setGpus(unsigned int count, bool sampleSplit = false) { _gpus = count; _gpuSampleSplit = sampleSplit; }

	// Live rendering — what PathTracerApp's "Live render" loop (PathTracerApp.cpp:57-98,127-136) would call instead of a full
	// render() + Texture::upload per UI frame: liveBegin sizes the session; every livePass traces the next 1 / passCount of
	// the stratification cells and adds it to sums that stay on the GPU (a moved camera starts over, a converged frame costs
	// nothing); liveDisplay fetches the frame as refined so far as 8-bit pixels (tonemapped, sRGB) ready for an RGBA8
	// texture. Returns the number of passes the image holds (0 on failure).
	bool liveBegin(unsigned int width, unsigned int height);
	bool livePass(const Camera & camera, size_t samples, size_t depth, size_t passCount);
	unsigned int liveDisplay(std::vector<unsigned char> & pixels, unsigned int components = 4);

private:
	bool uploadLights();
	void cameraFor(const Camera & camera, struct rb_camera & out) const;

	Raycaster _raycaster;
	std::shared_ptr<Scene> _scene;
	rb_scene * _deviceScene = nullptr;
	rb_pathtracer * _device = nullptr;
	unsigned long long _seed = 0;
	unsigned int _tile[3]	 = {0, 0, 0};
	unsigned int _split[2]	 = {0, 0};
	unsigned int _gpus		 = 0;
	bool _gpuSampleSplit	 = false;
	std::vector<unsigned char> _lightsUploaded; // flattened scene->lights as last sent to the device
};
