// Headless driver of the drop-in: the reference's own Scene loader + the GPU PathTracer/Raycaster from this directory.
//
// Two forms:
//   pathtracer_b200 --render --scene <name> [--samples N] [--depth D] [--size W H] [--output path.png|.exr]
//                   [--resources-root <dir holding common/ and pbrdemo/>] [--gpus N] [--split tile|sample] [--seed S] [--reference-writer]
//       the reference's `PathTracer --render` (src/apps/pathtracer/main.cpp:21-122): same flags, same defaults (8 spp,
//       depth 5, 1024x1024, ./test_<scene>_... .png), same log lines. The tail of renderOneShot — tonemap 1 - exp(-c) for
//       non-float targets, sRGB, 8-bit quantisation, or the EXR writer's layer split and float->half conversion — runs on
//       the GPU (rb_output_ldr / rb_output_hdr); the bytes go to the SAME encoders the reference uses (stb_image_write,
//       tinyexr, linked from the reference's own Image.o). --reference-writer saves through the reference's Image::save
//       instead (CPU conversions): the two files are byte-identical (tests/test_dropin_gpu.py).
//   pathtracer_b200 <resources-root> <scene> <width> <height> <samples> <depth> <out.f32>
//       test harness form: raw float image, plus single-ray queries and the reference's RaycasterVisualisation.
#include "PathTracer.hpp"
#include "rendu_b200.h"
#include "resources/ResourcesManager.hpp"
#include "raycaster/RaycasterVisualisation.hpp"
#include "system/Config.hpp"
#include "system/System.hpp"
#include <stb_image/stb_image_write.h>
#include <tinyexr/tinyexr.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

namespace {

// System::timestamp's format (System.cpp:83-94; that translation unit pulls in the windowing libraries and is not linked).
std::string timestampNow() {
	const std::time_t now = std::time(nullptr);
	char text[64];
	std::strftime(text, sizeof(text), "%Y_%m_%d_%H_%M_%S", std::localtime(&now));
	return text;
}

// The flags of PathTracerConfig (main.cpp:24-76) plus the ones that only exist here.
struct RenderFlags : public RenderingConfig {
	explicit RenderFlags(const std::vector<std::string> & argv) : RenderingConfig(argv) {
		for(const KeyValues & arg : arguments()) {
			const std::vector<std::string> & v = arg.values;
			if(arg.key == "samples" && !v.empty()) samples = size_t(std::stoi(v[0]));
			else if(arg.key == "depth" && !v.empty()) depth = size_t(std::stoi(v[0]));
			else if(arg.key == "scene" && !v.empty()) scene = v[0];
			else if(arg.key == "output" && !v.empty()) outputPath = v[0];
			else if(arg.key == "size" && v.size() >= 2) { width = std::stoi(v[0]); height = std::stoi(v[1]); }
			else if(arg.key == "render") directRender = true;
			else if(arg.key == "resources-root" && !v.empty()) resources = v[0];
			else if(arg.key == "gpus" && !v.empty()) gpus = unsigned(std::stoi(v[0]));
			else if(arg.key == "split" && !v.empty()) sampleSplit = v[0] == "sample";
			else if(arg.key == "seed" && !v.empty()) seed = std::stoull(v[0]);
			else if(arg.key == "reference-writer") referenceWriter = true;
			else if(arg.key == "frames" && !v.empty()) frames = std::max(1, std::stoi(v[0]));
		}
		const size_t asked = samples;
		samples = size_t(std::pow(2, std::round(std::log2(float(asked)))));
		if(asked != samples) Log::Warning() << "Non power-of-2 samples count. Using " << samples << " instead." << std::endl;
		registerSection("Path tracer");
		registerArgument("size", "", "Dimensions of the image.", std::vector<std::string> {"width", "height"});
		registerArgument("samples", "", "Number of samples per pixel (closest power of 2).", "int");
		registerArgument("depth", "", "Maximum path depth.", "int");
		registerArgument("scene", "", "Name of the scene to load.", "string");
		registerArgument("output", "", "Path for the output image (.png, or .exr for the linear image).", "path");
		registerArgument("render", "", "Run a render immediatly (the only mode of this binary).");
		registerSection("B200");
		registerArgument("resources-root", "", "Directory holding the reference's resources/common and resources/pbrdemo.", "path");
		registerArgument("gpus", "", "Number of GPUs of this box to render on.", "int");
		registerArgument("split", "", "How a frame is split among the GPUs.", "tile|sample");
		registerArgument("seed", "", "Seed of the counter-based random numbers.", "int");
		registerArgument("frames", "", "Render the frame this many times (timing: the first one also builds the replicas on the other GPUs).", "int");
		registerArgument("reference-writer", "", "Save through the reference's own Image::save (CPU conversions) instead of the GPU output stage.");
		if(outputPath.empty()) {
			outputPath = "./test_" + scene + "_" + std::to_string(samples) + "_" + std::to_string(depth) + "_" + std::to_string(width) + "x" + std::to_string(height) + "_" + timestampNow() + ".png";
		}
	}
	int width = 1024, height = 1024;
	size_t samples = 8, depth = 5;
	std::string outputPath, scene, resources = "../../../resources";
	bool directRender = false, sampleSplit = false, referenceWriter = false;
	unsigned gpus = 0;
	int frames = 1;
	unsigned long long seed = 0;
};

bool saveLdrFromDevice(const Image & img, const std::string & path) {
	std::vector<unsigned char> bytes(size_t(img.width) * img.height * img.components);
	// tonemap + sRGB + quantisation fused on the GPU (main.cpp:109-116 + Image.cpp:138-152)
	if(rb_output_ldr(img.pixels.data(), img.width, img.height, img.components, 1, 1, bytes.data()) != RB_OK) {
		Log::Error() << "[PathTracer] output stage: " << rb_last_error() << std::endl;
		return false;
	}
	stbi_flip_vertically_on_write(0);
	return stbi_write_png(path.c_str(), int(img.width), int(img.height), int(img.components), bytes.data(), int(img.width) * int(img.components)) != 0;
}

bool saveHdrFromDevice(const Image & img, const std::string & path) {
	if(img.width < 16 || img.height < 16) return false; // Image.cpp:163-167
	const int channels = int(img.components);
	const size_t plane = size_t(img.width) * img.height;
	std::vector<uint16_t> halves(plane * size_t(channels));
	// layer split (B, G, R order) and round-half-up float -> half conversion on the GPU (Image.cpp:181-233, tinyexr.h:971-1008)
	if(rb_output_hdr(img.pixels.data(), img.width, img.height, img.components, 1, 0, halves.data()) != RB_OK) {
		Log::Error() << "[PathTracer] output stage: " << rb_last_error() << std::endl;
		return false;
	}
	EXRHeader header;
	InitEXRHeader(&header);
	EXRImage image;
	InitEXRImage(&image);
	unsigned char * planes[4] = {nullptr, nullptr, nullptr, nullptr};
	for(int c = 0; c < channels; ++c) planes[c] = reinterpret_cast<unsigned char *>(halves.data() + size_t(c) * plane);
	image.num_channels = channels;
	image.images	   = planes;
	image.width		   = int(img.width);
	image.height	   = int(img.height);
	std::vector<EXRChannelInfo> infos;
	infos.resize(size_t(channels));
	const char * names = channels == 4 ? "ABGR" : "BGR";
	std::vector<int> inTypes(size_t(channels), TINYEXR_PIXELTYPE_HALF), outTypes(size_t(channels), TINYEXR_PIXELTYPE_HALF);
	for(int c = 0; c < channels; ++c) {
		std::memset(&infos[size_t(c)], 0, sizeof(EXRChannelInfo));
		infos[size_t(c)].name[0] = names[c];
	}
	header.num_channels			 = channels;
	header.channels				 = infos.data();
	header.pixel_types			 = inTypes.data();  // the layers already hold halves
	header.requested_pixel_types = outTypes.data();
	unsigned char * data = nullptr;
	const size_t size	 = SaveEXRImageToMemory(&image, &header, &data, nullptr);
	if(size == 0 || !data) return false;
	Resources::saveRawDataToExternalFile(path, reinterpret_cast<char *>(data), size);
	free(data);
	return true;
}

// renderOneShot (main.cpp:85-122) with the GPU classes.
int renderOneShot(const RenderFlags & config) {
	Resources::manager().addResources(config.resources + "/common");
	Resources::manager().addResources(config.resources + "/pbrdemo");
	std::shared_ptr<Scene> scene(new Scene(config.scene));
	if(!scene->init(Storage::CPU | Storage::FORCE_FRAME)) return 1;
	Image render(uint(config.width), uint(config.height), 3);
	Camera camera = scene->viewpoint();
	camera.ratio(float(config.width) / float(config.height));
	PathTracer tracer(scene);
	tracer.setSeed(config.seed);
	if(config.gpus > 1) tracer.setGpus(config.gpus, config.sampleSplit);
	Log::Info() << "[PathTracer] Rendering..." << std::endl;
	for(int frame = 0; frame < config.frames; ++frame) {
		std::fill(render.pixels.begin(), render.pixels.end(), 0.0f); // render() accumulates into the image (PathTracer.cpp:288)
		tracer.render(camera, config.samples, config.depth, render);
	}
	Log::Info() << "[PathTracer] Saving to " << config.outputPath << "." << std::endl;
	bool ok;
	if(config.referenceWriter) {
		// The reference's own tail, on the CPU: for comparing files.
		if(!Image::isFloat(config.outputPath)) {
			for(size_t p = 0; p < render.pixels.size(); ++p) render.pixels[p] = 1.0f - std::exp(-render.pixels[p] * 1.f);
		}
		ok = render.save(config.outputPath, Image::Save::SRGB_LDR | Image::Save::IGNORE_ALPHA) == 0;
	} else {
		ok = Image::isFloat(config.outputPath) ? saveHdrFromDevice(render, config.outputPath) : saveLdrFromDevice(render, config.outputPath);
	}
	if(!ok) {
		Log::Error() << "[PathTracer] Could not write " << config.outputPath << "." << std::endl;
		return 1;
	}
	return 0;
}

} // namespace

int main(int argc, char ** argv) {
	if(argc > 1 && std::strncmp(argv[1], "--", 2) == 0) {
		RenderFlags config(std::vector<std::string>(argv, argv + argc));
		if(config.showHelp()) return 0;
		if(!config.directRender || config.scene.empty()) {
			std::fprintf(stderr, "%s: the interactive viewer is not part of the drop-in: pass --render --scene <name> (see --help)\n", argv[0]);
			return 2;
		}
		return renderOneShot(config);
	}
	if(argc < 8) {
		std::fprintf(stderr, "usage: %s <resources-root> <scene> <width> <height> <samples> <depth> <out.f32>\n", argv[0]);
		return 2;
	}
	const std::string root = argv[1];
	Resources::manager().addResources(root + "/pbrdemo");
	Resources::manager().addResources(root + "/common");
	std::shared_ptr<Scene> scene(new Scene(argv[2]));
	if(!scene->init(Storage::CPU | Storage::FORCE_FRAME)) return 1;
	const int w = std::atoi(argv[3]), h = std::atoi(argv[4]);
	Image render(w, h, 3);
	Camera camera = scene->viewpoint();
	camera.ratio(float(w) / float(h));
	PathTracer tracer(scene);
	tracer.setSeed(1);
	tracer.render(camera, size_t(std::atoi(argv[5])), size_t(std::atoi(argv[6])), render);
	// A few single-ray queries through the unchanged Raycaster API.
	const Raycaster::Hit hit = tracer.raycaster().intersects(camera.position(), scene->boundingBox().getCentroid() - camera.position());
	std::printf("center ray: hit=%d dist=%f mesh=%lu\n", int(hit.hit), hit.dist, hit.meshId);
	// The reference's own debug view on top of the GPU raycaster: BVH levels and the nodes one ray visits.
	{
		RaycasterVisualisation visu(tracer.raycaster());
		std::vector<Mesh> levels, rayLevels;
		visu.getAllLevels(levels);
		const Raycaster::Hit vh = visu.getRayLevels(camera.position(), scene->boundingBox().getCentroid() - camera.position(), rayLevels);
		std::printf("bvh levels: %zu, ray levels: %zu, visualisation hit=%d dist=%f\n", levels.size(), rayLevels.size(), int(vh.hit), vh.dist);
	}
	FILE * f = std::fopen(argv[7], "wb");
	if(!f) return 1;
	std::fwrite(render.pixels.data(), sizeof(float), render.pixels.size(), f);
	std::fclose(f);
	double mean = 0.0;
	for(float p : render.pixels) mean += p;
	std::printf("image mean %f\n", mean / double(render.pixels.size()));
	return 0;
}
