// rb_output.cu — the output stage of the headless render: the tonemap of renderOneShot
// (src/apps/pathtracer/main.cpp:109-116) and the float -> 8-bit conversion of Image::saveAsLDR
// (src/engine/resources/Image.cpp:138-152) fused into one pass over the resolved image, so that a frame leaves the GPU
// as bytes ready for the PNG encoder.
#include "rb_internal.h"
#include "rb_math.cuh"

__global__ void rb_k_output_ldr(const float * __restrict__ pixels, size_t n, uint32_t components, int tonemap, int srgb, uint8_t * __restrict__ out) {
	for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const uint32_t cid = uint32_t(i % components);
		float value = pixels[i];
		if(cid == 3) {
			out[i] = 255; // IGNORE_ALPHA (main.cpp:119)
			continue;
		}
		if(tonemap) value = 1.0f - expf(-value * 1.f);			// glm::vec3(1.0f) - glm::exp(-color * 1.f)
		if(srgb) value = powf(value, 1.0f / 2.2f);				// Image.cpp:146
		const float scaled = std_min(255.0f, std_max(0.0f, 255.0f * value)); // Image.cpp:148
		out[i] = static_cast<unsigned char>(scaled);
	}
}

extern "C" rb_status rb_output_ldr_device(const float * d_pixels, uint32_t width, uint32_t height, uint32_t components, int tonemap, int srgb, uint8_t * d_out, void * stream) {
	RB_REQUIRE_DEVICE();
	if(!d_pixels || !d_out) return rb_fail(RB_ERR_INVALID, "rb_output_ldr_device: null argument");
	if(components != 3 && components != 4) return rb_fail(RB_ERR_INVALID, "rb_output_ldr_device: components must be 3 or 4");
	const size_t n = size_t(width) * height * components;
	if(n == 0) return RB_OK;
	int sms = 148;
	int dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	rb_k_output_ldr<<<unsigned(sms * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_pixels, n, components, tonemap, srgb, d_out);
	RB_CUDA(cudaGetLastError());
	return RB_OK;
}

extern "C" rb_status rb_output_ldr(const float * pixels, uint32_t width, uint32_t height, uint32_t components, int tonemap, int srgb, uint8_t * out) {
	RB_REQUIRE_DEVICE();
	if(!pixels || !out) return rb_fail(RB_ERR_INVALID, "rb_output_ldr: null argument");
	const size_t n = size_t(width) * height * components;
	if(n == 0) return RB_OK;
	cudaStream_t s = cudaStreamPerThread;
	float * d_in   = nullptr;
	uint8_t * d_o  = nullptr;
	RB_CUDA(cudaMalloc(&d_in, n * sizeof(float)));
	if(cudaMalloc(&d_o, n) != cudaSuccess) {
		cudaFree(d_in);
		return rb_fail(RB_ERR_CUDA, "rb_output_ldr: cudaMalloc failed");
	}
	rb_status st = RB_OK;
	if(cudaMemcpyAsync(d_in, pixels, n * sizeof(float), cudaMemcpyHostToDevice, s) != cudaSuccess) st = rb_fail(RB_ERR_CUDA, "rb_output_ldr: upload failed");
	if(!st) st = rb_output_ldr_device(d_in, width, height, components, tonemap, srgb, d_o, s);
	if(!st && cudaMemcpyAsync(out, d_o, n, cudaMemcpyDeviceToHost, s) != cudaSuccess) st = rb_fail(RB_ERR_CUDA, "rb_output_ldr: download failed");
	if(!st && cudaStreamSynchronize(s) != cudaSuccess) st = rb_fail(RB_ERR_CUDA, "rb_output_ldr: kernel failed");
	cudaFree(d_in);
	cudaFree(d_o);
	return st;
}
