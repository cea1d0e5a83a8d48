// rb_shade.cuh — device restatement of the scalar building blocks PathTracer::render calls per hit.
// Every function names the reference code it follows; arithmetic keeps the reference's operation order (the library
// is built with -fmad=false), libm calls map to CUDA's powf/expf/sinf/cosf/sqrtf.
#pragma once
#include "rb_math.cuh"
#include "rb_types.h"
#include <stdint.h>

// ---------------------------------------------------------------------------------------------- scene tables
struct RbTexInfo {
	uint32_t offset; // first texel in the float4 pool
	uint32_t width, height;
	uint32_t _pad;
};

struct RbObject {
	float invtp[9];	 // inverse(transpose(mat3(model))), column-major, precomputed once (PathTracer.cpp:94 does it per hit)
	int32_t texAlbedo, texNormal, texRmao;
	uint32_t flags;	 // RB_OBJ_*
};
#define RB_OBJ_EMISSIVE 1u
#define RB_OBJ_MASKED 2u
#define RB_OBJ_USE_UV 4u

struct RbLight {
	uint32_t kind, castsShadow;
	float intensity[3];
	float position[3];
	float direction[3];
	float radius;
	float outerCos, cosOuterCos, innerCos; // cos(angles.y), cos(cos(angles.y)) [the SpotLight.cpp:68 quirk], cos(angles.x)
};

struct RbSceneDev {
	const float4 * texels;
	const RbTexInfo * textures;
	const RbObject * objects;
	const RbLight * lights;
	const float4 * triShade; // per triangle in hierarchy order, 9 float4: for each of its three vertices
							 // (n.xyz, t.x) (t.yz, b.xy) (b.z, uv.xy, mesh id bits) — see rb_k_build_tri_shade
	int nLights;
	uint32_t bgMode;
	float bgColor[3];
	int32_t bgTex[6];
	float sunDir[3];
	int32_t skyLut;
	uint32_t anyMasked; // some object is masked and has texture coordinates: shadow rays need the alpha walk
};

// ---------------------------------------------------------------------------------------------- RNG
// Counter-based generator replacing the reference's thread_local mt19937 (Random.cpp:72-80): every uniform is a pure
// function of (seed, pixel, sample id, bounce, dimension), so an image does not depend on how pixels and samples are
// distributed over kernels or GPUs. splitmix64 finaliser; 24 random bits -> float in [0,1).
RB_HD uint64_t rb_mix64(uint64_t z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
RB_HD uint64_t rb_rng_key(uint64_t seed, uint32_t pixel, uint32_t sid) {
	return rb_mix64(rb_mix64(seed) ^ ((uint64_t(pixel) << 32) | uint64_t(sid)));
}
// dimension layout: 0,1 = pixel jitter; 8*(bounce+1) + {0: light pick, 1: lobe pick, 2,3: lobe sample}
RB_HD float rb_rng_float(uint64_t key, uint32_t dim) {
	const uint64_t z = rb_mix64(key + uint64_t(dim) * 0xD1B54A32D192ED03ull);
	return float(uint32_t(z >> 40)) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------------------------- images
RB_HD int rb_mod_pos(int x, int w) { return ((x % w) + w) % w; } // Image.hpp:189-191

// modPos(x, w) and modPos(x + 1, w) — the two columns (rows) of a bilinear tap — without the four integer divisions the
// literal form costs per pair (a signed 32-bit modulo by a run-time value is ~28 instructions; the three textures of a
// shaded hit took 24 of them, 40 % of the shade kernel's instructions). Same values: for a power-of-two size the
// two's-complement AND is the floor modulo; otherwise the second one follows from the first.
RB_D void rb_wrap_pair(int x, int w, int & a, int & b) {
	if((w & (w - 1)) == 0) {
		a = x & (w - 1);
		b = (x + 1) & (w - 1);
	} else {
		a = rb_mod_pos(x, w);
		b = (a + 1 == w) ? 0 : a + 1;
	}
}

// Image::rgbal (Image.cpp:86-106): wrap-around bilinear, texel-corner convention (no half-texel offset).
// (Measured and not kept: this body, rb_ggx, powf and the 1/sqrt of normalize as out-of-line calls shrink the shade kernel
// from 5.7 K to 3.9 K instructions and make it 2-8 % SLOWER each — profiles/r02_shade_diet.txt.)
RB_D float4 rb_image_rgbal_p(const float4 * __restrict__ texels, const RbTexInfo * __restrict__ textures, int tex, float x, float y) {
	const RbTexInfo ti = textures[tex];
	const float xi = x * float(ti.width);
	const float yi = y * float(ti.height);
	const float xb = floorf(xi);
	const float yb = floorf(yi);
	const float dx = xi - xb;
	const float dy = yi - yb;
	int x0, x1, y0, y1;
	rb_wrap_pair(int(xb), int(ti.width), x0, x1);
	rb_wrap_pair(int(yb), int(ti.height), y0, y1);
	const float4 * base = texels + ti.offset;
	const float4 p00 = __ldg(base + size_t(y0) * ti.width + x0);
	const float4 p01 = __ldg(base + size_t(y1) * ti.width + x0);
	const float4 p10 = __ldg(base + size_t(y0) * ti.width + x1);
	const float4 p11 = __ldg(base + size_t(y1) * ti.width + x1);
	const float ax = 1.0f - dx, ay = 1.0f - dy;
	float4 r;
	r.x = ax * (ay * p00.x + dy * p01.x) + dx * (ay * p10.x + dy * p11.x);
	r.y = ax * (ay * p00.y + dy * p01.y) + dx * (ay * p10.y + dy * p11.y);
	r.z = ax * (ay * p00.z + dy * p01.z) + dx * (ay * p10.z + dy * p11.z);
	r.w = ax * (ay * p00.w + dy * p01.w) + dx * (ay * p10.w + dy * p11.w);
	return r;
}
RB_D float4 rb_image_rgbal(const RbSceneDev & sc, int tex, float x, float y) {
	return rb_image_rgbal_p(sc.texels, sc.textures, tex, x, y);
}
// Image::rgbl (Image.cpp:64-84): same taps, rgb only.
RB_D V3 rb_image_rgbl(const RbSceneDev & sc, int tex, float x, float y) {
	const float4 r = rb_image_rgbal(sc, tex, x, y);
	return V3{r.x, r.y, r.z};
}

// Texture::sampleCubemap (Texture.cpp:141-191): faces px,nx,py,ny,pz,nz; one-texel clamp; bilinear on the face.
RB_D V3 rb_sample_cubemap(const RbSceneDev & sc, const int32_t * faces, V3 dir) {
	const float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
	// The face is picked with constant indices: `faces` is an array inside a kernel parameter, and a run-time index into it
	// makes the compiler copy the whole parameter block into local memory at the start of the kernel.
	int tex = faces[0];
	float x = 0.0f, y = 0.0f, denom = 1.0f;
	if(ax >= ay && ax >= az) {
		denom = ax;
		y	  = dir.y;
		if(dir.x >= 0.0f) { tex = faces[0]; x = -dir.z; } else { tex = faces[1]; x = dir.z; }
	} else if(ay >= ax && ay >= az) {
		denom = ay;
		x	  = dir.x;
		if(dir.y >= 0.0f) { tex = faces[2]; y = -dir.z; } else { tex = faces[3]; y = dir.z; }
	} else if(az >= ax && az >= ay) {
		denom = az;
		y	  = dir.y;
		if(dir.z >= 0.0f) { tex = faces[4]; x = dir.x; } else { tex = faces[5]; x = -dir.x; }
	}
	x = 0.5f * (x / denom) + 0.5f;
	y = 0.5f * (-y / denom) + 0.5f;
	const RbTexInfo ti = sc.textures[tex];
	const float eps	   = 1.0f / float(min(ti.width, ti.height));
	x = glm_clamp(x, 0.0f + eps, 1.0f - eps);
	y = glm_clamp(y, 0.0f + eps, 1.0f - eps);
	return rb_image_rgbl(sc, tex, x, y);
}

// ---------------------------------------------------------------------------------------------- lights
// Light::sample for Point (PointLight.cpp:68-87), Spot (SpotLight.cpp:52-81), Directional (DirectionalLight.cpp:40-44).
RB_D V3 rb_light_sample(const RbLight & l, V3 position, float & dist, float & attenuation) {
	if(l.kind == RB_LIGHT_DIRECTIONAL) {
		attenuation = 1.0f;
		dist		= FLT_MAX;
		return V3{-l.direction[0], -l.direction[1], -l.direction[2]};
	}
	V3 direction = V3{l.position[0], l.position[1], l.position[2]} - position;
	dist		 = length3(direction);
	attenuation	 = 0.0f;
	if(dist > l.radius) {
		return V3{0.0f, 0.0f, 0.0f};
	}
	if(dist > 0.0f) {
		direction = direction / dist;
	}
	float angleAttenuation = 1.0f;
	if(l.kind == RB_LIGHT_SPOT) {
		const float currentCos = dot3(-direction, V3{l.direction[0], l.direction[1], l.direction[2]});
		// `currentCos < std::cos(outerCos)`: the reference compares against the cosine OF the cosine (SpotLight.cpp:68).
		if(currentCos < l.cosOuterCos) {
			return V3{0.0f, 0.0f, 0.0f};
		}
		angleAttenuation = glm_clamp((currentCos - l.outerCos) / (l.innerCos - l.outerCos), 0.0f, 1.0f);
	}
	const float radiusRatio	 = dist / l.radius;
	const float radiusRatio2 = radiusRatio * radiusRatio;
	const float attenNum	 = glm_clamp(1.0f - radiusRatio2, 0.0f, 1.0f);
	attenuation = (l.kind == RB_LIGHT_SPOT) ? (angleAttenuation * attenNum * attenNum) : (attenNum * attenNum);
	return direction;
}

// ---------------------------------------------------------------------------------------------- Lambert + GGX
#define RB_ONE_OVER_PI 0.318309886183790671537767526745028724f
#define RB_TWO_PI 6.28318530717958647692528676655900576f
#define RB_PI 3.14159265358979323846264338327950288f
#define RB_QUARTER_PI 0.785398163397448309615660845819875721f
#define RB_HALF_PI 1.57079632679489661923132169163975144f

// MaterialGGX::alphaFromRoughness (MaterialGGX.cpp:24-28)
RB_D float rb_alpha_from_roughness(float roughness) {
	const float roughClamp = std_max(0.045f, roughness);
	return std_max(0.0001f, roughClamp * roughClamp);
}
// MaterialGGX::D (MaterialGGX.cpp:9-13)
RB_D float rb_ggx_D(float NdotH, float alpha) {
	const float halfDenum = NdotH * NdotH * (alpha * alpha - 1.0f) + 1.0f;
	const float halfTerm  = alpha / std_max(0.0001f, halfDenum);
	return halfTerm * halfTerm * RB_ONE_OVER_PI;
}
// MaterialGGX::V (MaterialGGX.cpp:15-21)
RB_D float rb_ggx_V(float NdotL, float NdotV, float alpha) {
	const float alpha2 = alpha * alpha;
	const float visL   = NdotV * sqrtf((-NdotL * alpha2 + NdotL) * NdotL + alpha2);
	const float visV   = NdotL * sqrtf((-NdotV * alpha2 + NdotV) * NdotV + alpha2);
	return 0.5f / std_max(0.0001f, visV + visL);
}
// MaterialGGX::GGX (MaterialGGX.cpp:30-52) including the wo/wi naming swap (NdotL = wo.z, NdotV = wi.z) and the
// multi-scattering adjustment. A shaded hit evaluates it twice with the same (wo, baseColor, roughness, metallic) — once
// towards the light, once along the sampled direction — so everything that does not depend on wi is computed once
// (RbGgxSurface: same expressions, same rounding; it only is not repeated).
struct RbGgxSurface {
	V3 wo, F0, oneMinusF0, diffuse;
	float alpha, alpha2, twoAlpha2, NdotL, rootL;
};
RB_D RbGgxSurface rb_ggx_surface(V3 wo, V3 baseColor, float roughness, float metallic) {
	RbGgxSurface g;
	g.wo		 = wo;
	g.alpha		 = rb_alpha_from_roughness(roughness);
	g.alpha2	 = g.alpha * g.alpha;
	g.twoAlpha2	 = 2.0f * g.alpha * g.alpha;
	g.NdotL		 = std_max(wo.z, 0.0f);
	g.rootL		 = sqrtf((-g.NdotL * g.alpha2 + g.NdotL) * g.NdotL + g.alpha2); // the wi-independent root of MaterialGGX::V
	g.F0		 = glm_mix(v3(0.04f), baseColor, metallic);
	g.oneMinusF0 = v3(1.0f) - g.F0;
	g.diffuse	 = ((1.0f - metallic) * RB_ONE_OVER_PI) * baseColor * g.oneMinusF0;
	return g;
}
// Returns (brdf, pdf of the specular lobe): eval ignores the pdf, sampleAndEval uses it.
RB_D float4 rb_ggx(const RbGgxSurface & g, V3 wi) {
	const V3 h		  = normalize3(wi + g.wo);
	const float NdotH = std_max(h.z, 0.0f);
	const float VdotH = std_max(dot3(wi, h), 0.0f);
	const float NdotV = std_max(wi.z, 0.0f);
	const float Dh	  = rb_ggx_D(NdotH, g.alpha);
	const float hPdf  = Dh * NdotH;
	const float pdf	  = hPdf / (4.0f * std_max(0.0001f, VdotH));
	// pow(1 - VdotH, 5) (MaterialGGX.cpp:6) as x^2 * x^2 * x: three roundings, within 1.5 ulp of the exact power — tighter than
	// the bound of CUDA's powf, which stood here before at ~95 instructions per call, two calls per shaded hit.
	const float fm	  = 1.0f - VdotH;
	const float fm2	  = fm * fm;
	const float fres  = (fm2 * fm2) * fm;
	const V3 F		  = g.F0 + fres * g.oneMinusF0;									   // MaterialGGX::F (:5-7)
	// MaterialGGX::V (:15-21)
	const float visL  = NdotV * g.rootL;
	const float visV  = g.NdotL * sqrtf((-NdotV * g.alpha2 + NdotV) * NdotV + g.alpha2);
	const float Vis	  = 0.5f / std_max(0.0001f, visV + visL);
	const V3 specular = (Dh * Vis) * F;
	const V3 multiAdj = v3(1.0f) + (g.twoAlpha2 * NdotV) * g.F0;
	const V3 brdf	  = (g.diffuse + specular * multiAdj) * NdotV;
	return make_float4(brdf.x, brdf.y, brdf.z, pdf);
}
// MaterialGGX::eval (MaterialGGX.cpp:94-101)
RB_D V3 rb_ggx_eval(const RbGgxSurface & g, V3 wi) {
	if(wi.z < 0.0f) {
		return v3(0.0f);
	}
	const float4 r = rb_ggx(g, wi);
	return V3{r.x, r.y, r.z};
}
RB_D V3 rb_ggx_eval(V3 wo, V3 baseColor, float roughness, float metallic, V3 wi) {
	return rb_ggx_eval(rb_ggx_surface(wo, baseColor, roughness, metallic), wi);
}
// Random::sampleDisk + sampleCosineHemisphere (Random.cpp:41-70) on two given uniforms.
RB_D V3 rb_cosine_hemisphere(float u0, float u1) {
	const float x = 2.0f * u0 - 1.0f;
	const float y = 2.0f * u1 - 1.0f;
	float dx = 0.0f, dy = 0.0f;
	if(!(x == 0.0f && y == 0.0f)) {
		float angle, radius;
		if(fabsf(x) > fabsf(y)) {
			radius = x;
			angle  = RB_QUARTER_PI * y / x;
		} else {
			radius = y;
			angle  = RB_HALF_PI - RB_QUARTER_PI * x / y;
		}
		dx = radius * cosf(angle);
		dy = radius * sinf(angle);
	}
	const float z = sqrtf(std_max(0.0f, 1.0f - dx * dx - dy * dy));
	return V3{dx, dy, z};
}
// sincos(double) out of line: its argument-reduction slow path (never taken for an angle in [0, 2 pi)) is a few hundred
// instructions in the middle of the specular branch otherwise.
static __device__ __noinline__ double2 rb_sincos_call(double x) { double s, c; sincos(x, &s, &c); return make_double2(s, c); }
// MaterialGGX::sampleAndEval (MaterialGGX.cpp:54-92). u0 picks the lobe; (u1,u2) sample it. The reference draws its
// uniforms in this order: lobe pick, then either (x, phi) for the specular lobe or the two disk coordinates.
RB_D V3 rb_ggx_sample_and_eval(const RbGgxSurface & gs, V3 baseColor, float metallic, float u0, float u1, float u2, V3 & wi) {
	const V3 wo = gs.wo;
	const float probaSpecular = glm_mix(1.0f / (dot3(baseColor, v3(1.0f)) / 3.0f + 1.0f), 1.0f, metallic);
	if(u0 < probaSpecular) {
		const float a2			 = gs.alpha2;
		const float x			 = u1;
		const float phiH		 = u2 * RB_TWO_PI;
		const float cosThetaHSqr = std_min((1.0f - x) / ((a2 - 1.0f) * x + 1.0f), 1.0f);
		const float cosThetaH	 = sqrtf(cosThetaHSqr);
		const float sinThetaH	 = sqrtf(1.0f - cosThetaHSqr);
		// The reference calls the unqualified C cos/sin here (MaterialGGX.cpp:69): evaluated in double, multiplied by
		// sinThetaH in double, narrowed to float by the vec3 constructor. One fp64 sincos per bounce is noise on B200.
		double sphi, cphi;
		const double2 sc2 = rb_sincos_call((double)phiH);
		sphi = sc2.x; cphi = sc2.y;
		const V3 lh = V3{(float)((double)sinThetaH * cphi), (float)((double)sinThetaH * sphi), cosThetaH};
		wi = (2.0f * dot3(wo, lh)) * lh - wo;
	} else {
		wi = rb_cosine_hemisphere(u1, u2);
		if(wo.z < 0.0f) {
			wi.z *= -1.0f;
		}
	}
	if(wi.z < 0.0f) {
		return v3(0.0f);
	}
	const float4 g = rb_ggx(gs, wi);
	const V3 brdf  = V3{g.x, g.y, g.z};
	const float pdf = glm_mix(RB_ONE_OVER_PI * std_max(wi.z, 0.0f), g.w, probaSpecular);
	if(pdf == 0.0f) {
		return v3(0.0f);
	}
	return brdf / pdf;
}

RB_D V3 rb_ggx_sample_and_eval(V3 wo, V3 baseColor, float roughness, float metallic, float u0, float u1, float u2, V3 & wi) {
	return rb_ggx_sample_and_eval(rb_ggx_surface(wo, baseColor, roughness, metallic), baseColor, metallic, u0, u1, u2, wi);
}

// ---------------------------------------------------------------------------------------------- sky
// Intersection::sphere (Intersection.cpp:7-20)
RB_D bool rb_sphere(V3 rayOrigin, V3 rayDir, float radius, float & r0, float & r1) {
	const float a	  = dot3(rayDir, rayDir);
	const float b	  = dot3(rayOrigin, rayDir);
	const float c	  = dot3(rayOrigin, rayOrigin) - radius * radius;
	const float delta = b * b - a * c;
	if(delta < 0.0f) {
		return false;
	}
	const float dsqrt = sqrtf(delta);
	r0 = (-b - dsqrt) / a;
	r1 = (-b + dsqrt) / a;
	return true;
}
// MaterialSky::eval (MaterialSky.cpp:6-87) with Sky::AtmosphereParameters defaults (Sky.hpp:53-65).
// Out of line: 550 instructions that only scenes with an atmosphere background execute, and then only for rays that
// leave the scene; kept out of the shade kernel's straight-line code (together with the fp64 sincos below: shade -1.5 %).
static __device__ __noinline__ V3 rb_sky_eval_p(const float4 * __restrict__ texels, const RbTexInfo * __restrict__ textures, int skyLut, V3 rayOrigin, V3 rayDir, V3 sunDir);
RB_D V3 rb_sky_eval(const RbSceneDev & sc, V3 rayOrigin, V3 rayDir, V3 sunDir) { return rb_sky_eval_p(sc.texels, sc.textures, sc.skyLut, rayOrigin, rayDir, sunDir); }
static __device__ __noinline__ V3 rb_sky_eval_p(const float4 * __restrict__ texels, const RbTexInfo * __restrict__ textures, int skyLut, V3 rayOrigin, V3 rayDir, V3 sunDir) {
	const V3 sunColor	   = V3{1.474f, 1.8504f, 1.91198f};
	const V3 kRayleigh	   = V3{5.5e-6f, 13.0e-6f, 22.4e-6f};
	const float groundRadius = 6371e3f, topRadius = 6471e3f, sunIntensity = 20.0f, kMie = 21e-6f;
	const float heightRayleigh = 8000.0f, heightMie = 1200.0f, gMie = 0.758f, sunRadius = 0.04675f, sunRadiusCos = 0.998f;
	const int samplesCount = 16;

	const V3 planetPos = rayOrigin + V3{0.0f, groundRadius, 0.0f} + V3{0.0f, 1.0f, 0.0f};
	float top0 = 0.0f, top1 = 0.0f, gr0 = 0.0f, gr1 = 0.0f;
	if(!rb_sphere(planetPos, rayDir, topRadius, top0, top1)) {
		return v3(0.0f);
	}
	const bool didHitGround		= rb_sphere(planetPos, rayDir, groundRadius, gr0, gr1);
	const float distanceToInter = std_min(top1, didHitGround ? gr0 : 0.0f);
	const float stepSize		= (distanceToInter - top0) / float(samplesCount);
	const float cosViewSun		= dot3(rayDir, sunDir);
	float rayleighDist = 0.0f, mieDist = 0.0f;
	V3 rayleighScatt = v3(0.0f), mieScatt = v3(0.0f), transmittance = v3(0.0f);
	for(int i = 0; i < samplesCount; ++i) {
		const V3 currPos = planetPos + ((float(i) + 0.5f) * stepSize) * rayDir;
		float currHeight = length3(currPos) - groundRadius;
		if(i == (samplesCount - 1) && currHeight < 0.0f) {
			currHeight = 0.0f;
		}
		const float rayleighStep = expf(-currHeight / heightRayleigh) * stepSize;
		const float mieStep		 = expf(-currHeight / heightMie) * stepSize;
		rayleighDist += rayleighStep;
		mieDist += mieStep;
		const V3 e = -(mieDist * v3(kMie) + rayleighDist * kRayleigh);
		const V3 directAttenuation = V3{expf(e.x), expf(e.y), expf(e.z)};
		const float relativeHeight	 = (length3(currPos) - groundRadius) / (topRadius - groundRadius);
		const float relativeCosAngle = -0.5f * sunDir.y + 0.5f;
		float uvx = (511.0f / 512.0f) * relativeHeight + 0.5f / 512.0f;
		float uvy = (511.0f / 512.0f) * relativeCosAngle + 0.5f / 512.0f;
		uvx = glm_clamp(uvx, 0.5f / 512.0f, 511.0f / 512.0f);
		uvy = glm_clamp(uvy, 0.5f / 512.0f, 511.0f / 512.0f);
		const float4 lut = rb_image_rgbal_p(texels, textures, skyLut, uvx, uvy);
		const V3 secondaryAttenuation = V3{lut.x, lut.y, lut.z};
		const V3 attenuation		  = directAttenuation * secondaryAttenuation;
		rayleighScatt += rayleighStep * attenuation;
		mieScatt += mieStep * attenuation;
		transmittance += directAttenuation;
	}
	// MaterialSky::rayleighPhase / miePhase (MaterialSky.cpp:89-98)
	const float k			  = 1.0f / (4.0f * RB_PI);
	const float rayleighPhase = k * 3.0f / 4.0f * (1.0f + cosViewSun * cosViewSun);
	const float g2			  = gMie * gMie;
	const float miePhase	  = k * 3.0f * (1.0f - g2) / (2.0f * (2.0f + g2)) * (1.0f + cosViewSun * cosViewSun) / powf(1.0f + g2 - 2.0f * gMie * cosViewSun, 3.0f / 2.0f);
	const V3 rayleighParticipation = rayleighPhase * kRayleigh * rayleighScatt;
	const V3 mieParticipation	   = (kMie * miePhase) * mieScatt;
	V3 sunRadiance = v3(0.0f);
	const bool didHitGroundForward = didHitGround && gr1 > 0.0f;
	if(!didHitGroundForward && dot3(rayDir, sunDir) > sunRadiusCos) {
		sunRadiance = sunColor / (RB_PI * sunRadius * sunRadius);
	}
	return sunIntensity * (rayleighParticipation + mieParticipation) + transmittance * sunRadiance;
}

// PathTracer::evalBackground (PathTracer.cpp:23-55)
RB_D V3 rb_eval_background(const RbSceneDev & sc, V3 rayDir, V3 rayPos, float ndcx, float ndcy, bool directHit) {
	const uint32_t mode = sc.bgMode;
	if(directHit) {
		if(mode == RB_BG_IMAGE) return rb_image_rgbl(sc, sc.bgTex[0], ndcx, ndcy);
		if(mode == RB_BG_SKYBOX) return rb_sample_cubemap(sc, sc.bgTex, normalize3(rayDir));
		if(mode == RB_BG_ATMOSPHERE) return rb_sky_eval(sc, rayPos, normalize3(rayDir), V3{sc.sunDir[0], sc.sunDir[1], sc.sunDir[2]});
		return V3{sc.bgColor[0], sc.bgColor[1], sc.bgColor[2]};
	}
	if(mode == RB_BG_SKYBOX) return rb_sample_cubemap(sc, sc.bgTex, normalize3(rayDir));
	if(mode == RB_BG_ATMOSPHERE) return rb_sky_eval(sc, rayPos, normalize3(rayDir), V3{sc.sunDir[0], sc.sunDir[1], sc.sunDir[2]});
	return v3(0.0f);
}

// ---------------------------------------------------------------------------------------------- surface
struct RbSurface {
	V3 n, t, b; // interpolated, un-normalised
	float u, v; // texture coordinates
};

// Raycaster::interpolateAttribute (Raycaster.hpp:90-97): w*a[i0] + u*a[i1] + v*a[i2], for normals, tangents,
// bitangents and texcoords at once, from the triangle's pre-gathered record (one dependent load level: hit -> record,
// instead of hit -> vertex ids -> three vertices).
RB_D void rb_fetch_surface(const RbSceneDev & sc, int tri, float hu, float hv, RbSurface & s, uint32_t & meshId) {
	const float hw = 1.0f - hu - hv;
	const float4 * p = sc.triShade + 9 * size_t(tri);
	const float4 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
	const float4 b0 = __ldg(p + 3), b1 = __ldg(p + 4), b2 = __ldg(p + 5);
	const float4 c0 = __ldg(p + 6), c1 = __ldg(p + 7), c2 = __ldg(p + 8);
	meshId = __float_as_uint(a2.w);
#define RB_INTERP(fa, fb, fc) (hw * (fa) + hu * (fb) + hv * (fc))
	s.n = V3{RB_INTERP(a0.x, b0.x, c0.x), RB_INTERP(a0.y, b0.y, c0.y), RB_INTERP(a0.z, b0.z, c0.z)};
	s.t = V3{RB_INTERP(a0.w, b0.w, c0.w), RB_INTERP(a1.x, b1.x, c1.x), RB_INTERP(a1.y, b1.y, c1.y)};
	s.b = V3{RB_INTERP(a1.z, b1.z, c1.z), RB_INTERP(a1.w, b1.w, c1.w), RB_INTERP(a2.x, b2.x, c2.x)};
	s.u = RB_INTERP(a2.y, b2.y, c2.y);
	s.v = RB_INTERP(a2.z, b2.z, c2.z);
#undef RB_INTERP
}
RB_D void rb_fetch_uv(const RbSceneDev & sc, int tri, float hu, float hv, float & u, float & v, uint32_t & meshId) {
	const float hw	 = 1.0f - hu - hv;
	const float4 * p = sc.triShade + 9 * size_t(tri);
	const float4 a2 = __ldg(p + 2), b2 = __ldg(p + 5), c2 = __ldg(p + 8);
	meshId = __float_as_uint(a2.w);
	u = hw * a2.y + hu * b2.y + hv * c2.y;
	v = hw * a2.z + hu * b2.z + hv * c2.z;
}

// PathTracer::buildLocalFrame (PathTracer.cpp:86-122).
RB_D M3 rb_build_local_frame(const RbSceneDev & sc, const RbObject & obj, const RbSurface & s, V3 rayDir) {
	const V3 n = normalize3(s.n);
	const V3 t = normalize3(s.t);
	const V3 b = normalize3(s.b);
	M3 invtp;
	invtp.c[0] = V3{obj.invtp[0], obj.invtp[1], obj.invtp[2]};
	invtp.c[1] = V3{obj.invtp[3], obj.invtp[4], obj.invtp[5]};
	invtp.c[2] = V3{obj.invtp[6], obj.invtp[7], obj.invtp[8]};
	M3 tbn;
	tbn.c[0] = mul(invtp, t);
	tbn.c[1] = mul(invtp, b);
	tbn.c[2] = mul(invtp, n);
	const bool frontFacing = dot3(tbn.c[2], rayDir) < 0.0f;
	if(!frontFacing) {
		tbn.c[2] *= -1.0f;
	}
	if((obj.flags & RB_OBJ_USE_UV) && !(obj.flags & RB_OBJ_EMISSIVE)) {
		const float4 img	 = rb_image_rgbal(sc, obj.texNormal, s.u, s.v);
		const V3 localNormal = normalize3(2.0f * V3{img.x, img.y, img.z} - v3(1.0f));
		tbn.c[2]			 = mul(tbn, localNormal);
	}
	tbn.c[2] = normalize3(tbn.c[2]);
	tbn.c[1] = normalize3(cross3(tbn.c[2], normalize3(tbn.c[0])));
	tbn.c[0] = normalize3(cross3(tbn.c[1], tbn.c[2]));
	return tbn;
}
