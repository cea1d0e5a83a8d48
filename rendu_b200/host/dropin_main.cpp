// Headless check of the drop-in: the reference's own Scene loader + the GPU PathTracer/Raycaster from this directory.
// Mirrors renderOneShot (src/apps/pathtracer/main.cpp:85-105) minus tonemap/PNG: writes the raw float image.
//   pathtracer_b200 <resources-root> <scene> <width> <height> <samples> <depth> <out.f32>
#include "PathTracer.hpp"
#include "resources/ResourcesManager.hpp"
#include "raycaster/RaycasterVisualisation.hpp"
#include <cstdio>
#include <cstdlib>

int main(int argc, char ** argv) {
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
