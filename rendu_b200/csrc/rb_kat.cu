// rb_kat.cu — known-answer hooks: the device versions of the reference's scalar building blocks evaluated on arrays.
// They exist for the parity tests (tests/test_shading_gpu.py) and launch real kernels; there is no host path.
#include "rb_scene.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {
struct DevMem {
	std::vector<void *> ptrs;
	~DevMem() {
		for(void * p : ptrs) cudaFree(p);
	}
	template<typename T>
	cudaError_t up(T ** d, const T * h, size_t count) {
		void * raw	  = nullptr;
		cudaError_t e = cudaMalloc(&raw, count * sizeof(T) + 64);
		if(e != cudaSuccess) return e;
		ptrs.push_back(raw);
		*d = static_cast<T *>(raw);
		if(h && count) e = cudaMemcpy(raw, h, count * sizeof(T), cudaMemcpyHostToDevice);
		return e;
	}
};
inline unsigned blocks(size_t n) { return unsigned((n + 127) / 128); }
inline V3 ld3(const float * p, size_t i) { return V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
} // namespace

__device__ __forceinline__ V3 kld3(const float * p, size_t i) { return V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
__device__ __forceinline__ void kst3(float * p, size_t i, V3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

__global__ void rb_k_kat_ggx_eval(const float * wo, const float * base, const float * rough, const float * metal, const float * wi, size_t n, float * out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	kst3(out, i, rb_ggx_eval(kld3(wo, i), kld3(base, i), rough[i], metal[i], kld3(wi, i)));
}

__global__ void rb_k_kat_ggx_sample(const float * wo, const float * base, const float * rough, const float * metal, const float * u, size_t n, float * outWi, float * outEval) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	V3 wi = v3(0.0f);
	const V3 e = rb_ggx_sample_and_eval(kld3(wo, i), kld3(base, i), rough[i], metal[i], u[3 * i], u[3 * i + 1], u[3 * i + 2], wi);
	kst3(outWi, i, wi);
	kst3(outEval, i, e);
}

__global__ void rb_k_kat_image(RbSceneDev sc, const float * uv, size_t n, float * out4) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	const float4 r = rb_image_rgbal(sc, 0, uv[2 * i], uv[2 * i + 1]);
	out4[4 * i] = r.x; out4[4 * i + 1] = r.y; out4[4 * i + 2] = r.z; out4[4 * i + 3] = r.w;
}

__global__ void rb_k_kat_cubemap(RbSceneDev sc, const float * dirs, size_t n, float * out3) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	kst3(out3, i, rb_sample_cubemap(sc, sc.bgTex, kld3(dirs, i)));
}

__global__ void rb_k_kat_light(RbLight light, const float * pos, size_t n, float * out5) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	float dist = 0.0f, att = 0.0f;
	const V3 d = rb_light_sample(light, kld3(pos, i), dist, att);
	out5[5 * i] = d.x; out5[5 * i + 1] = d.y; out5[5 * i + 2] = d.z; out5[5 * i + 3] = dist; out5[5 * i + 4] = att;
}

__global__ void rb_k_kat_sky(RbSceneDev sc, const float * org, const float * dirs, V3 sun, size_t n, float * out3) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	kst3(out3, i, rb_sky_eval(sc, kld3(org, i), kld3(dirs, i), sun));
}

#define KAT_BEGIN(name)                                      \
	RB_REQUIRE_DEVICE();                                     \
	if(n == 0) return RB_OK;                                 \
	DevMem mem;

#define KAT_END(dst, src, count)                                                                       \
	RB_CUDA(cudaGetLastError());                                                                       \
	RB_CUDA(cudaDeviceSynchronize());                                                                  \
	RB_CUDA(cudaMemcpy(dst, src, (count) * sizeof(float), cudaMemcpyDeviceToHost));

extern "C" rb_status rb_kat_ggx_eval(const float * wo, const float * base, const float * rough, const float * metal, const float * wi, size_t n, float * out3) {
	KAT_BEGIN(ggx_eval)
	float *d_wo, *d_base, *d_r, *d_m, *d_wi, *d_out;
	RB_CUDA(mem.up(&d_wo, wo, 3 * n)); RB_CUDA(mem.up(&d_base, base, 3 * n)); RB_CUDA(mem.up(&d_r, rough, n));
	RB_CUDA(mem.up(&d_m, metal, n)); RB_CUDA(mem.up(&d_wi, wi, 3 * n)); RB_CUDA(mem.up(&d_out, (const float *)nullptr, 3 * n));
	rb_k_kat_ggx_eval<<<blocks(n), 128>>>(d_wo, d_base, d_r, d_m, d_wi, n, d_out);
	KAT_END(out3, d_out, 3 * n)
	return RB_OK;
}

extern "C" rb_status rb_kat_ggx_sample(const float * wo, const float * base, const float * rough, const float * metal, const float * u3, size_t n, float * out_wi3, float * out_eval3) {
	KAT_BEGIN(ggx_sample)
	float *d_wo, *d_base, *d_r, *d_m, *d_u, *d_wi, *d_ev;
	RB_CUDA(mem.up(&d_wo, wo, 3 * n)); RB_CUDA(mem.up(&d_base, base, 3 * n)); RB_CUDA(mem.up(&d_r, rough, n));
	RB_CUDA(mem.up(&d_m, metal, n)); RB_CUDA(mem.up(&d_u, u3, 3 * n));
	RB_CUDA(mem.up(&d_wi, (const float *)nullptr, 3 * n)); RB_CUDA(mem.up(&d_ev, (const float *)nullptr, 3 * n));
	rb_k_kat_ggx_sample<<<blocks(n), 128>>>(d_wo, d_base, d_r, d_m, d_u, n, d_wi, d_ev);
	KAT_END(out_wi3, d_wi, 3 * n)
	RB_CUDA(cudaMemcpy(out_eval3, d_ev, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
	return RB_OK;
}

static rb_status kat_scene_textures(DevMem & mem, const float * rgba, uint32_t width, uint32_t height, int count, RbSceneDev & sc) {
	float * d_tex;
	RbTexInfo * d_info;
	std::vector<RbTexInfo> infos(count);
	for(int k = 0; k < count; ++k) infos[k] = RbTexInfo{uint32_t(k) * width * height, width, height, 0};
	RB_CUDA(mem.up(&d_tex, rgba, size_t(count) * width * height * 4));
	RB_CUDA(mem.up(&d_info, infos.data(), size_t(count)));
	std::memset(&sc, 0, sizeof(sc));
	sc.texels	= reinterpret_cast<const float4 *>(d_tex);
	sc.textures = d_info;
	for(int k = 0; k < 6; ++k) sc.bgTex[k] = k < count ? k : 0;
	sc.skyLut = 0;
	return RB_OK;
}

extern "C" rb_status rb_kat_image_sample(const float * rgba, uint32_t width, uint32_t height, const float * uv, size_t n, float * out4) {
	KAT_BEGIN(image)
	RbSceneDev sc;
	rb_status st = kat_scene_textures(mem, rgba, width, height, 1, sc);
	if(st) return st;
	float *d_uv, *d_out;
	RB_CUDA(mem.up(&d_uv, uv, 2 * n)); RB_CUDA(mem.up(&d_out, (const float *)nullptr, 4 * n));
	rb_k_kat_image<<<blocks(n), 128>>>(sc, d_uv, n, d_out);
	KAT_END(out4, d_out, 4 * n)
	return RB_OK;
}

extern "C" rb_status rb_kat_cubemap_sample(const float * faces, uint32_t width, uint32_t height, const float * dirs, size_t n, float * out3) {
	KAT_BEGIN(cubemap)
	RbSceneDev sc;
	rb_status st = kat_scene_textures(mem, faces, width, height, 6, sc);
	if(st) return st;
	float *d_dirs, *d_out;
	RB_CUDA(mem.up(&d_dirs, dirs, 3 * n)); RB_CUDA(mem.up(&d_out, (const float *)nullptr, 3 * n));
	rb_k_kat_cubemap<<<blocks(n), 128>>>(sc, d_dirs, n, d_out);
	KAT_END(out3, d_out, 3 * n)
	return RB_OK;
}

extern "C" rb_status rb_kat_light_sample(const rb_light_desc * light, const float * positions, size_t n, float * out5) {
	if(!light) return rb_fail(RB_ERR_INVALID, "rb_kat_light_sample: null light");
	KAT_BEGIN(light)
	RbLight l;
	std::memset(&l, 0, sizeof(l));
	l.kind = light->kind; l.castsShadow = light->casts_shadow;
	std::memcpy(l.intensity, light->intensity, 12); std::memcpy(l.position, light->position, 12); std::memcpy(l.direction, light->direction, 12);
	l.radius = light->radius;
	l.outerCos = std::cos(light->outer_angle); l.cosOuterCos = std::cos(l.outerCos); l.innerCos = std::cos(light->inner_angle);
	float *d_pos, *d_out;
	RB_CUDA(mem.up(&d_pos, positions, 3 * n)); RB_CUDA(mem.up(&d_out, (const float *)nullptr, 5 * n));
	rb_k_kat_light<<<blocks(n), 128>>>(l, d_pos, n, d_out);
	KAT_END(out5, d_out, 5 * n)
	return RB_OK;
}

extern "C" rb_status rb_kat_sky_eval(const float * lut_rgba, uint32_t width, uint32_t height, const float * origins, const float * dirs, const float sun[3], size_t n, float * out3) {
	KAT_BEGIN(sky)
	RbSceneDev sc;
	rb_status st = kat_scene_textures(mem, lut_rgba, width, height, 1, sc);
	if(st) return st;
	float *d_o, *d_d, *d_out;
	RB_CUDA(mem.up(&d_o, origins, 3 * n)); RB_CUDA(mem.up(&d_d, dirs, 3 * n)); RB_CUDA(mem.up(&d_out, (const float *)nullptr, 3 * n));
	rb_k_kat_sky<<<blocks(n), 128>>>(sc, d_o, d_d, V3{sun[0], sun[1], sun[2]}, n, d_out);
	KAT_END(out3, d_out, 3 * n)
	return RB_OK;
}

// rb_div (rb_math.cuh): the exact multiply-high division the path-tracer kernels map slots to pixels and samples with,
// against the device's own u32 division, element by element.
__global__ void rb_k_kat_div(const uint32_t * num, const uint32_t * den, size_t n, uint32_t * out2) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	out2[2 * i]		= rb_div(num[i], rb_div_magic(den[i]));
	out2[2 * i + 1] = num[i] / den[i];
}

extern "C" rb_status rb_kat_div(const uint32_t * numerators, const uint32_t * denominators, size_t n, uint32_t * out2) {
	KAT_BEGIN(div)
	if(!numerators || !denominators || !out2) return rb_fail(RB_ERR_INVALID, "rb_kat_div: null argument");
	for(size_t i = 0; i < n; ++i) {
		if(denominators[i] == 0) return rb_fail(RB_ERR_INVALID, "rb_kat_div: denominator %zu is zero", i);
	}
	uint32_t *d_n, *d_d, *d_o;
	RB_CUDA(mem.up(&d_n, numerators, n)); RB_CUDA(mem.up(&d_d, denominators, n)); RB_CUDA(mem.up(&d_o, (const uint32_t *)nullptr, 2 * n));
	rb_k_kat_div<<<blocks(n), 128>>>(d_n, d_d, n, d_o);
	RB_CUDA(cudaGetLastError());
	RB_CUDA(cudaDeviceSynchronize());
	RB_CUDA(cudaMemcpy(out2, d_o, 2 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	return RB_OK;
}
