// rb_math.cuh — fp32 vector helpers that follow the reference's arithmetic operation for operation.
//
// The reference is scalar fp32 C++ built for baseline x86-64 (no FMA instructions, IEEE div/sqrt) on top of glm
// 0.9.9.7. Every translation unit of this library is compiled with -fmad=false, so nvcc never contracts a*b+c and
// the expressions below round exactly where glm's do:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z            src/libs/glm/detail/func_geometric.inl:48-56
//   cross(x,y)    = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)   func_geometric.inl:70-80
//   normalize(v)  = v * (1.0f / sqrt(dot(v,v)))                func_geometric.inl:83-92, func_exponential.inl:134-139
//   glm::min(x,y) = (y < x) ? y : x ;  glm::max(x,y) = (x < y) ? y : x          func_common.inl:17-30
//   std::max(a,b) = (a < b) ? b : a ;  std::min(a,b) = (b < a) ? b : a
// Where a fused multiply-add is wanted for speed (shading code whose parity is statistical) it is written out
// explicitly with fmaf().
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <float.h>

#define RB_HD __host__ __device__ __forceinline__
#define RB_D __device__ __forceinline__

struct V3 {
	float x, y, z;
};

RB_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
RB_HD V3 v3(float s) { return V3{s, s, s}; }
RB_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
RB_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
RB_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
RB_HD V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
RB_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
RB_HD V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
RB_HD V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
RB_HD V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
RB_HD V3 & operator+=(V3 & a, V3 b) { a = a + b; return a; }
RB_HD V3 & operator*=(V3 & a, V3 b) { a = a * b; return a; }
RB_HD V3 & operator*=(V3 & a, float s) { a = a * s; return a; }

RB_HD float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RB_HD V3 cross3(V3 x, V3 y) { return V3{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
RB_HD float rsqrt_glm(float x) { return 1.0f / sqrtf(x); }
RB_HD V3 normalize3(V3 v) { return v * rsqrt_glm(dot3(v, v)); }
RB_HD float length3(V3 v) { return sqrtf(dot3(v, v)); }
// Literal comparison forms (NaN-faithful): do not replace with fminf/fmaxf.
RB_HD float glm_min(float x, float y) { return (y < x) ? y : x; }
RB_HD float glm_max(float x, float y) { return (x < y) ? y : x; }
RB_HD float std_max(float a, float b) { return (a < b) ? b : a; }
RB_HD float std_min(float a, float b) { return (b < a) ? b : a; }
RB_HD float glm_clamp(float x, float lo, float hi) { return glm_min(glm_max(x, lo), hi); }
RB_HD float glm_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
RB_HD V3 glm_mix(V3 x, V3 y, float a) { return x * (1.0f - a) + y * a; }
RB_HD V3 glm_min3(V3 a, V3 b) { return V3{glm_min(a.x, b.x), glm_min(a.y, b.y), glm_min(a.z, b.z)}; }
RB_HD V3 glm_max3(V3 a, V3 b) { return V3{glm_max(a.x, b.x), glm_max(a.y, b.y), glm_max(a.z, b.z)}; }

// 3x3 matrix, column-major like glm::mat3 (c[k] is column k).
struct M3 {
	V3 c[3];
};
RB_HD V3 mul(const M3 & m, V3 v) {
	// glm mat3*vec3: m[0][i]*v.x + m[1][i]*v.y + m[2][i]*v.z   (type_mat3x3.inl operator*)
	return V3{m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z,
			  m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z,
			  m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z};
}
RB_HD V3 mulT(const M3 & m, V3 v) {
	// transpose(m) * v : row i of the transpose is column i of m.
	return V3{m.c[0].x * v.x + m.c[0].y * v.y + m.c[0].z * v.z,
			  m.c[1].x * v.x + m.c[1].y * v.y + m.c[1].z * v.z,
			  m.c[2].x * v.x + m.c[2].y * v.y + m.c[2].z * v.z};
}

#include <stdint.h>
// Exact division of a 32-bit number by a run-time constant d: floor(n / d) = floor(m * n / 2^64) with m = ceil(2^64 / d)
// (Lemire, Kaser, Kurz 2019: exact for every 32-bit n and d because 64 >= 32 + 32). m does not fit for d = 1: stored as 0.
// A hardware-free u32 division is ~20 instructions; every kernel that maps a path slot to (pixel, sample) did four.
RB_HD uint64_t rb_div_magic(uint32_t d) { return d > 1u ? 0xFFFFFFFFFFFFFFFFull / d + 1ull : 0ull; }
RB_HD uint32_t rb_div(uint32_t n, uint64_t m) {
#ifdef __CUDA_ARCH__
	return m ? uint32_t(__umul64hi(m, uint64_t(n))) : n;
#else
	return m ? uint32_t((unsigned __int128)m * n >> 64) : n;
#endif
}
