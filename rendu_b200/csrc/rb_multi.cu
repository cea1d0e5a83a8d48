// rb_multi.cu — PathTracer::render on all GPUs of one box from ONE host process.
//
// What it replaces: System::forParallel (src/engine/system/System.hpp:72-104), the reference's only fan-out — one
// std::thread per host core over contiguous blocks of image rows (PathTracer.cpp:186). Here the fan-out is one host
// thread and one replica of the scene per GPU:
//   * tile split   — bands of `tile_rows` rows dealt round-robin to the GPUs; every GPU writes the per-pixel sums of its
//                    bands at their place in a full-frame buffer; the other GPUs' bands are copied into GPU 0's buffer
//                    with one strided peer copy each (NVLink), GPU 0 resolves. No reduction: bit-identical to one GPU.
//   * sample split — GPU g traces the stratification cells sid % G == g of every pixel; the per-pixel sums are combined
//                    with ONE ncclReduce(sum, fp32, W*H*3) to GPU 0 (valid because the reference clamps per sample
//                    BEFORE accumulating, PathTracer.cpp:288), or — RB_SAMPLE_REDUCE=p2p — by the resolve kernel itself,
//                    which reads the G buffers over peer memory in a fixed order (deterministic, no NCCL).
// NCCL is resolved at run time (dlopen of libnccl.so.2: whichever copy the process already carries, e.g. PyTorch's),
// so the library has no link-time dependency on it and single-GPU users never load it.
#include "rb_scene.h"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <mutex>
#include <thread>
#include <utility>

rb_scene * rb_pt_scene(rb_pathtracer * pt);
void ** rb_pt_multi_slot(rb_pathtracer * pt);
void rb_pt_set_stats(rb_pathtracer * pt, const rb_render_stats & s);
uint32_t rb_pt_round_samples(uint32_t samples);

namespace {

struct NcclApi {
	void * lib = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *)																	  = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t)																						  = nullptr;
	ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t)				  = nullptr;
	ncclResult_t (*GroupStart)()																								  = nullptr;
	ncclResult_t (*GroupEnd)()																									  = nullptr;
	const char * (*GetErrorString)(ncclResult_t)																				  = nullptr;
	bool load(std::string & why) {
		if(lib) return true;
		for(const char * name : {"libnccl.so.2", "libnccl.so"}) {
			lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if(lib) break;
		}
		if(!lib) {
			why = std::string("cannot load NCCL (") + dlerror() + "); RB_SAMPLE_REDUCE=p2p combines the sums without it";
			return false;
		}
		CommInitAll	   = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
		CommDestroy	   = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
		Reduce		   = reinterpret_cast<decltype(Reduce)>(dlsym(lib, "ncclReduce"));
		GroupStart	   = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
		GroupEnd	   = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
		GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
		if(!CommInitAll || !CommDestroy || !Reduce || !GroupStart || !GroupEnd || !GetErrorString) {
			why = "libnccl.so.2 lacks one of ncclCommInitAll / ncclReduce / ncclGroupStart / ncclGroupEnd";
			return false;
		}
		return true;
	}
};
NcclApi g_nccl;

struct Replica {
	int device		   = -1;
	rb_scene * scene   = nullptr; // replicas 1.. own theirs; replica 0 is the caller's
	rb_pathtracer * pt = nullptr;
	float * d_sum	   = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	rb_status status = RB_OK;
	std::string error;
	float renderMs = 0.0f;
};

struct Multi {
	std::vector<Replica> reps;
	std::vector<ncclComm_t> comms;
	uint64_t sceneEdit = 0;
	size_t sumFloats   = 0;
	float * d_pixels   = nullptr; // on replica 0
	size_t pixelFloats = 0;
	const float ** d_sumTable = nullptr; // on replica 0: the replicas' sum buffers (peer pointers), for the p2p resolve
	bool peerAccess = false;
	rb_multi_stats last = {};
};

void destroyMulti(Multi * m) {
	if(!m) return;
	for(size_t g = 0; g < m->comms.size(); ++g) {
		if(m->comms[g] && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comms[g]);
	}
	for(size_t g = 0; g < m->reps.size(); ++g) {
		Replica & r = m->reps[g];
		if(r.device < 0) continue;
		cudaSetDevice(r.device);
		if(r.d_sum) cudaFree(r.d_sum);
		if(r.e0) cudaEventDestroy(r.e0);
		if(r.e1) cudaEventDestroy(r.e1);
		if(r.stream) cudaStreamDestroy(r.stream);
		if(g > 0) {
			if(r.pt) rb_pathtracer_destroy(r.pt);
			if(r.scene) rb_scene_destroy(r.scene);
		}
	}
	if(!m->reps.empty()) {
		cudaSetDevice(m->reps[0].device);
		cudaFree(m->d_pixels);
		cudaFree(m->d_sumTable);
	}
	delete m;
}

// Sum of the G buffers in rank order, then `(pixels + sum) / samples` (PathTracer.cpp:288-292): the sample-split combine
// and the resolve in one pass over peer memory (loads from the other GPUs travel over NVLink).
__global__ void rb_k_resolve_peers(const float * const * __restrict__ sums, int count, uint32_t npixels, uint32_t components, float samples, float * __restrict__ pixels) {
	for(uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < npixels; p += gridDim.x * blockDim.x) {
		float acc[3] = {0.0f, 0.0f, 0.0f};
		for(int g = 0; g < count; ++g) {
			const float * s = sums[g] + 3 * size_t(p);
			acc[0] += s[0];
			acc[1] += s[1];
			acc[2] += s[2];
		}
		for(int c = 0; c < 3; ++c) pixels[size_t(p) * components + c] = (pixels[size_t(p) * components + c] + acc[c]) / samples;
	}
}

} // namespace

void rb_pt_multi_destroy(void * m) {
	destroyMulti(static_cast<Multi *>(m));
}

// Replicas for `gpus` devices: the caller's path tracer on its own device first, then the lowest-numbered other devices.
static rb_status ensureReplicas(rb_pathtracer * pt, uint32_t gpus, Multi ** out) {
	rb_scene * sc = rb_pt_scene(pt);
	void ** slot  = rb_pt_multi_slot(pt);
	Multi * m	  = static_cast<Multi *>(*slot);
	if(m && (m->reps.size() != gpus || m->sceneEdit != sc->editSerial)) {
		destroyMulti(m);
		m = nullptr;
		*slot = nullptr;
	}
	if(m) {
		*out = m;
		return RB_OK;
	}
	int ndev = 0;
	RB_CUDA(cudaGetDeviceCount(&ndev));
	if(int(gpus) > ndev) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_render: opts.gpus = %u but %d CUDA device(s) are visible", gpus, ndev);
	m = new Multi();
	m->reps.resize(gpus);
	m->sceneEdit = sc->editSerial;
	const int primary = sc->raycaster->device;
	m->reps[0].device = primary;
	m->reps[0].scene  = sc;
	m->reps[0].pt	  = pt;
	for(int g = 1, d = 0; g < int(gpus); ++g, ++d) {
		if(d == primary) ++d;
		m->reps[g].device = d;
	}
	// One host thread per GPU: clone the host side of the scene, commit it on that device (same builder, hence the same
	// tree and triangle order), create its path tracer.
	const rb_build_mode mode = rb_build_mode(sc->raycaster->buildMode);
	std::vector<std::thread> pool;
	for(int g = 0; g < int(gpus); ++g) {
		pool.emplace_back([&, g]() {
			Replica & r = m->reps[g];
			auto fail = [&](rb_status st) {
				r.status = st;
				r.error	 = rb_last_error();
			};
			if(cudaSetDevice(r.device) != cudaSuccess) return fail(rb_fail(RB_ERR_CUDA, "cannot select device %d", r.device));
			if(g > 0) {
				rb_status st = rb_scene_clone_host(sc, &r.scene);
				if(st) return fail(st);
				st = rb_scene_commit(r.scene, mode);
				if(st) return fail(st);
				st = rb_pathtracer_create(r.scene, &r.pt);
				if(st) return fail(st);
			}
			if(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess)
				return fail(rb_fail(RB_ERR_CUDA, "stream setup failed on device %d", r.device));
		});
	}
	for(auto & t : pool) t.join();
	for(const Replica & r : m->reps) {
		if(r.status) {
			const rb_status st = rb_fail(r.status, "rb_pathtracer_render: replica on device %d: %s", r.device, r.error.c_str());
			destroyMulti(m);
			return st;
		}
	}
	// Peer access from device 0's side (the strided band copies and the p2p resolve read the other GPUs' buffers).
	cudaSetDevice(primary);
	m->peerAccess = true;
	for(int g = 1; g < int(gpus); ++g) {
		int can = 0;
		cudaDeviceCanAccessPeer(&can, primary, m->reps[g].device);
		if(!can) {
			m->peerAccess = false;
			continue;
		}
		// Peer access stays enabled for the life of the process: asked for once per (device, peer) pair.
		static std::mutex lock;
		static std::vector<std::pair<int, int>> enabled;
		std::lock_guard<std::mutex> guard(lock);
		const std::pair<int, int> key(primary, m->reps[g].device);
		if(std::find(enabled.begin(), enabled.end(), key) != enabled.end()) continue;
		const cudaError_t e = cudaDeviceEnablePeerAccess(m->reps[g].device, 0);
		if(e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) enabled.push_back(key);
		else m->peerAccess = false;
		if(e != cudaSuccess) cudaGetLastError(); // clear the (non-sticky) error state
	}
	*slot = m;
	*out  = m;
	return RB_OK;
}

rb_status rb_pt_multi_render(rb_pathtracer * pt, const rb_camera * camera, uint32_t width, uint32_t height, uint32_t samples, uint32_t depth, uint32_t components, float * pixels, const rb_render_opts * opts) {
	const uint32_t gpus = opts->gpus;
	if(opts->tile_count > 1 || opts->sample_count > 1) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_render: opts.gpus > 1 does the splitting itself; leave tile_count / sample_count at 0");
	if(opts->split > RB_SPLIT_SAMPLE) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_render: unknown opts.split %u", opts->split);
	if(opts->accumulate) return rb_fail(RB_ERR_UNSUPPORTED, "rb_pathtracer_render: opts.accumulate with opts.gpus > 1");
	Multi * m	 = nullptr;
	rb_status st = ensureReplicas(pt, gpus, &m);
	if(st) return st;
	const bool sampleSplit = opts->split == RB_SPLIT_SAMPLE;
	const char * how	   = std::getenv("RB_SAMPLE_REDUCE");
	const bool p2pReduce   = sampleSplit && how && std::string(how) == "p2p";
	if(sampleSplit && !p2pReduce && m->comms.empty()) {
		std::string why;
		if(!g_nccl.load(why)) return rb_fail(RB_ERR_UNSUPPORTED, "rb_pathtracer_render: sample split: %s", why.c_str());
		std::vector<int> devs;
		for(const Replica & r : m->reps) devs.push_back(r.device);
		m->comms.assign(gpus, nullptr);
		const ncclResult_t nr = g_nccl.CommInitAll(m->comms.data(), int(gpus), devs.data());
		if(nr != ncclSuccess) {
			m->comms.clear();
			return rb_fail(RB_ERR_CUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString(nr));
		}
	}
	if(p2pReduce && !m->peerAccess) return rb_fail(RB_ERR_UNSUPPORTED, "rb_pathtracer_render: RB_SAMPLE_REDUCE=p2p needs peer access from device %d to every other GPU", m->reps[0].device);
	const size_t npx = size_t(width) * height;
	// Buffers: per-GPU sums (full frame), the caller's image on GPU 0.
	if(m->sumFloats < npx * 3) {
		for(Replica & r : m->reps) {
			RB_CUDA(cudaSetDevice(r.device));
			cudaFree(r.d_sum);
			r.d_sum = nullptr;
			RB_CUDA(cudaMalloc(&r.d_sum, npx * 3 * sizeof(float)));
		}
		m->sumFloats = npx * 3;
		RB_CUDA(cudaSetDevice(m->reps[0].device));
		cudaFree(m->d_sumTable);
		m->d_sumTable = nullptr;
		std::vector<const float *> table;
		for(const Replica & r : m->reps) table.push_back(r.d_sum);
		RB_CUDA(cudaMalloc(&m->d_sumTable, table.size() * sizeof(float *)));
		RB_CUDA(cudaMemcpy(m->d_sumTable, table.data(), table.size() * sizeof(float *), cudaMemcpyHostToDevice));
	}
	RB_CUDA(cudaSetDevice(m->reps[0].device));
	if(m->pixelFloats < npx * components) {
		cudaFree(m->d_pixels);
		m->d_pixels = nullptr;
		RB_CUDA(cudaMalloc(&m->d_pixels, npx * components * sizeof(float)));
		m->pixelFloats = npx * components;
	}
	Replica & root = m->reps[0];
	RB_CUDA(cudaMemcpyAsync(m->d_pixels, pixels, npx * components * sizeof(float), cudaMemcpyHostToDevice, root.stream));

	// ---- fan-out: one host thread per GPU renders its share into its own sum buffer
	const uint32_t tileRows = opts->tile_rows ? opts->tile_rows : 2;
	std::vector<std::thread> pool;
	for(int g = 0; g < int(gpus); ++g) {
		pool.emplace_back([&, g]() {
			Replica & r = m->reps[g];
			r.status	= RB_OK;
			if(cudaSetDevice(r.device) != cudaSuccess) {
				r.status = rb_fail(RB_ERR_CUDA, "cannot select device %d", r.device);
				r.error	 = rb_last_error();
				return;
			}
			rb_render_opts o = *opts;
			o.gpus			 = 0;
			if(sampleSplit) {
				o.sample_index = uint32_t(g);
				o.sample_count = gpus;
			} else {
				o.tile_rows	 = tileRows;
				o.tile_index = uint32_t(g);
				o.tile_count = gpus;
			}
			cudaEventRecord(r.e0, r.stream);
			r.status = rb_pathtracer_render_device(r.pt, camera, width, height, samples, depth, r.d_sum, &o, r.stream);
			cudaEventRecord(r.e1, r.stream);
			if(!r.status && cudaStreamSynchronize(r.stream) != cudaSuccess) r.status = rb_fail(RB_ERR_CUDA, "render failed on device %d: %s", r.device, cudaGetErrorString(cudaGetLastError()));
			if(r.status) r.error = rb_last_error();
			else cudaEventElapsedTime(&r.renderMs, r.e0, r.e1);
		});
	}
	for(auto & t : pool) t.join();
	for(const Replica & r : m->reps) {
		if(r.status) return rb_fail(r.status, "rb_pathtracer_render: device %d: %s", r.device, r.error.c_str());
	}

	// ---- combine on GPU 0
	RB_CUDA(cudaSetDevice(root.device));
	cudaEvent_t c0 = root.e0, c1 = root.e1;
	float renderMs = 0.0f;
	for(const Replica & r : m->reps) renderMs = std::max(renderMs, r.renderMs);
	RB_CUDA(cudaEventRecord(c0, root.stream));
	bool resolved = false;
	if(!sampleSplit) {
		// Rank g's bands sit at rows (k * G + g) * tileRows ...: one strided copy per GPU, same offsets on both sides.
		const size_t rowBytes  = size_t(width) * 3 * sizeof(float);
		const uint32_t nbands  = (height + tileRows - 1) / tileRows;
		for(int g = 1; g < int(gpus); ++g) {
			const Replica & r = m->reps[g];
			if(uint32_t(g) >= nbands) continue;
			const uint32_t mine	  = (nbands - uint32_t(g) + gpus - 1) / gpus; // bands g, g + G, ...
			const uint32_t lastBand = uint32_t(g) + (mine - 1) * gpus;
			const bool lastPartial  = (lastBand + 1) * tileRows > height;
			const uint32_t full		= lastPartial ? mine - 1 : mine;
			const size_t first		= size_t(g) * tileRows * rowBytes;
			if(full) {
				const size_t pitch = size_t(gpus) * tileRows * rowBytes;
				RB_CUDA(cudaMemcpy2DAsync(reinterpret_cast<char *>(root.d_sum) + first, pitch, reinterpret_cast<const char *>(r.d_sum) + first, pitch, tileRows * rowBytes, full, cudaMemcpyDefault, root.stream));
			}
			if(lastPartial) {
				const size_t off = size_t(lastBand) * tileRows * rowBytes;
				RB_CUDA(cudaMemcpyPeerAsync(reinterpret_cast<char *>(root.d_sum) + off, root.device, reinterpret_cast<char *>(r.d_sum) + off, r.device, (size_t(height) - size_t(lastBand) * tileRows) * rowBytes, root.stream));
			}
		}
	} else if(p2pReduce) {
		// (the sample count render() rounds to, PathTracer.cpp:169-170 — what the ordinary resolve divides by)
		rb_k_resolve_peers<<<1184, 256, 0, root.stream>>>(m->d_sumTable, int(gpus), uint32_t(npx), components, float(rb_pt_round_samples(samples)), m->d_pixels);
		RB_CUDA(cudaGetLastError());
		resolved = true;
	} else {
		ncclResult_t nr = g_nccl.GroupStart();
		for(int g = 0; g < int(gpus) && nr == ncclSuccess; ++g) {
			Replica & r = m->reps[g];
			nr = g_nccl.Reduce(r.d_sum, g == 0 ? r.d_sum : nullptr, npx * 3, ncclFloat, ncclSum, 0, m->comms[g], r.stream);
		}
		const ncclResult_t ne = g_nccl.GroupEnd();
		if(nr != ncclSuccess || ne != ncclSuccess) return rb_fail(RB_ERR_CUDA, "ncclReduce: %s", g_nccl.GetErrorString(nr != ncclSuccess ? nr : ne));
		for(int g = 1; g < int(gpus); ++g) {
			RB_CUDA(cudaSetDevice(m->reps[g].device));
			RB_CUDA(cudaStreamSynchronize(m->reps[g].stream));
		}
		RB_CUDA(cudaSetDevice(root.device));
	}
	if(!resolved) {
		// The frame is whole on GPU 0 now: the ordinary resolve over every row.
		rb_render_opts full = *opts;
		full.gpus			= 0;
		st = rb_pathtracer_resolve_device(pt, root.d_sum, width, height, samples, components, m->d_pixels, &full, root.stream);
		if(st) return st;
	}
	RB_CUDA(cudaEventRecord(c1, root.stream));
	RB_CUDA(cudaMemcpyAsync(pixels, m->d_pixels, npx * components * sizeof(float), cudaMemcpyDeviceToHost, root.stream));
	RB_CUDA(cudaStreamSynchronize(root.stream));
	float combineMs = 0.0f;
	cudaEventElapsedTime(&combineMs, c0, c1);

	rb_render_stats total = {};
	for(const Replica & r : m->reps) {
		rb_render_stats s = {};
		RB_CUDA(cudaSetDevice(r.device));
		st = rb_pathtracer_stats(r.pt, &s);
		if(st) return st;
		total.rays_closest += s.rays_closest;
		total.rays_shadow += s.rays_shadow;
		total.samples += s.samples;
		total.shaded_hits += s.shaded_hits;
		total.kernel_launches += s.kernel_launches;
	}
	total.kernel_launches += 1;
	total.device_seconds = double(renderMs + combineMs) * 1e-3;
	RB_CUDA(cudaSetDevice(root.device));
	rb_pt_set_stats(pt, total);
	m->last.gpus		= gpus;
	m->last.split		= opts->split;
	m->last.render_ms	= renderMs;
	m->last.combine_ms	= combineMs;
	m->last.used_nccl	= (sampleSplit && !p2pReduce) ? 1u : 0u;
	m->last.peer_access = m->peerAccess ? 1u : 0u;
	return RB_OK;
}

extern "C" rb_status rb_pathtracer_multi_stats(rb_pathtracer * pt, rb_multi_stats * out) {
	if(!pt || !out) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_multi_stats: null argument");
	Multi * m = static_cast<Multi *>(*rb_pt_multi_slot(pt));
	if(!m) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_multi_stats: no multi-GPU render has run on this handle");
	*out = m->last;
	return RB_OK;
}
