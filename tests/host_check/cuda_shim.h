// cuda_shim.h — TEST INFRASTRUCTURE (tests/host_check only; nothing in the library includes it).
// Lets g++ compile the device traversal headers of rendu_b200/csrc as ordinary C++ so that their per-ray logic can be
// stepped on the CPU and compared with the reference's golden hits when no GPU is at hand. It supplies the handful of
// CUDA intrinsics those headers use and the six PTX primitives they keep behind RB_HOST_CHECK. This is a checker for
// the kernels' source, not a way to run the product: the library has no CPU path and never defines RB_HOST_CHECK.
#pragma once
#include <cuda_runtime.h> // vector types; __device__ / __forceinline__ expand harmlessly under g++
#include <cmath>
#include <cstdint>
#include <cstring>

#define RB_HOST_CHECK 1
#include <algorithm>
using std::isfinite;
using std::max;
using std::min;

static inline float __uint_as_float(unsigned int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline unsigned int __float_as_uint(float f) { unsigned int v; std::memcpy(&v, &f, 4); return v; }
static inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline int __ffs(unsigned int v) { return __builtin_ffs(int(v)); }
static inline int __clz(unsigned int v) { return v ? __builtin_clz(v) : 32; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
template<typename T> static inline T __ldg(const T * p) { return *p; }
// PRMT: result byte i = byte (selector nibble i) of the eight bytes {a (0-3), b (4-7)} (no sign-replication modes used).
static inline unsigned int __byte_perm(unsigned int a, unsigned int b, unsigned int selector) {
	const uint64_t both = uint64_t(a) | (uint64_t(b) << 32);
	unsigned int r = 0;
	for(int i = 0; i < 4; ++i) r |= (unsigned int)((both >> (8 * ((selector >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
	return r;
}

// The "shared memory" of the one simulated thread; shared-window addresses are byte offsets into it.
static unsigned char rb_hc_shared[1 << 16];
static inline size_t __cvta_generic_to_shared(const void * p) { return size_t(static_cast<const unsigned char *>(p) - rb_hc_shared); }
static inline void rb_sts(unsigned int address, unsigned int value) { std::memcpy(rb_hc_shared + address, &value, 4); }
static inline unsigned int rb_lds(unsigned int address) { unsigned int v; std::memcpy(&v, rb_hc_shared + address, 4); return v; }

static inline unsigned long long rb_pack2(float a, float b) { return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32); }
static inline void rb_unpack2(unsigned long long v, float & a, float & b) { a = __uint_as_float((unsigned int)v); b = __uint_as_float((unsigned int)(v >> 32)); }
static inline unsigned long long rb_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
	float a0, a1, b0, b1, c0, c1;
	rb_unpack2(a, a0, a1); rb_unpack2(b, b0, b1); rb_unpack2(c, c0, c1);
	return rb_pack2(std::fmaf(a0, b0, c0), std::fmaf(a1, b1, c1));
}
static inline unsigned long long rb_submul2(unsigned long long a, unsigned long long b, unsigned long long c) {
	float a0, a1, b0, b1, c0, c1;
	rb_unpack2(a, a0, a1); rb_unpack2(b, b0, b1); rb_unpack2(c, c0, c1);
	volatile float d0 = a0 - b0, d1 = a1 - b1; // two roundings per lane, never contracted
	return rb_pack2(d0 * c0, d1 * c1);
}
// The node fetch of rb_trace.cuh (two 256-bit loads on the device): plain reads here.
static inline void rb_ldg_node(const float4 * n, float4 & n0, float4 & n1, float4 & n2, float4 & n3) { n0 = n[0]; n1 = n[1]; n2 = n[2]; n3 = n[3]; }
