/* TEST INFRASTRUCTURE — NOT PRODUCT CODE. See rendu_oracle.h for scope and pinning.
 * Path-tracer part of the oracle: a plain-C restatement of PathTracer::render and everything it calls
 * (src/apps/pathtracer/PathTracer.cpp, MaterialGGX.cpp, MaterialSky.cpp; src/engine/resources/Image.cpp, Texture.cpp;
 * src/engine/scene/lights/{Point,Spot,Directional}Light.cpp; src/engine/generation/Random.cpp), each function citing the lines it follows.
 * Random numbers come through a callback table so that the same code runs either on the reference's mt19937 stream
 * (rendu_oracle_std.cpp; bit-identical images to the reference in single-thread mode — that is the pin) or on the
 * counter-based generator the CUDA path uses (orc_rng_counter below; same image as the GPU up to libm/rounding).
 */
#include "rendu_oracle.h"
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;
typedef struct { v3 c[3]; } m3; /* column-major, like glm::mat3 */

static v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static v3 scale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }   /* vec * scalar */
static v3 scaleL(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }  /* scalar * vec */
static v3 divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }            /* func_geometric.inl:48-56 */
static v3 cross(v3 x, v3 y) { return V(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); } /* :70-80 */
static v3 normalize(v3 v) { return scale(v, 1.0f / sqrtf(dot(v, v))); }                /* :83-92 */
static float length(v3 v) { return sqrtf(dot(v, v)); }
static float gmin(float x, float y) { return (y < x) ? y : x; }   /* glm::min */
static float gmax(float x, float y) { return (x < y) ? y : x; }   /* glm::max */
static float smax(float a, float b) { return (a < b) ? b : a; }   /* std::max */
static float smin(float a, float b) { return (b < a) ? b : a; }   /* std::min */
static float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
static float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
static v3 mix3(v3 x, v3 y, float a) { return add(scale(x, 1.0f - a), scale(y, a)); }
static v3 m3mul(const m3 * m, v3 v) { /* type_mat3x3.inl:468-474 */
	return V(m->c[0].x * v.x + m->c[1].x * v.y + m->c[2].x * v.z, m->c[0].y * v.x + m->c[1].y * v.y + m->c[2].y * v.z, m->c[0].z * v.x + m->c[1].z * v.y + m->c[2].z * v.z);
}
static v3 m3mulT(const m3 * m, v3 v) { /* transpose(m) * v */
	return V(m->c[0].x * v.x + m->c[0].y * v.y + m->c[0].z * v.z, m->c[1].x * v.x + m->c[1].y * v.y + m->c[1].z * v.z, m->c[2].x * v.x + m->c[2].y * v.y + m->c[2].z * v.z);
}

#define ONE_OVER_PI 0.318309886183790671537767526745028724f
#define TWO_PI 6.28318530717958647692528676655900576f
#define PI_F 3.14159265358979323846264338327950288f
#define QUARTER_PI 0.785398163397448309615660845819875721f
#define HALF_PI 1.57079632679489661923132169163975144f

/* ------------------------------------------------------------------ counter-based generator (mirrors rb_shade.cuh) */
static uint64_t mix64(uint64_t z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
uint64_t orc_rng_key(uint64_t seed, uint32_t pixel, uint32_t sid) { return mix64(mix64(seed) ^ (((uint64_t)pixel << 32) | (uint64_t)sid)); }
float orc_rng_float(uint64_t key, uint32_t dim) {
	const uint64_t z = mix64(key + (uint64_t)dim * 0xD1B54A32D192ED03ull);
	return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
}

/* Per-path random source. In counter mode `dim` selects the variate; in stream mode (mt19937) draws are sequential. */
typedef struct {
	const orc_rng_stream * stream; /* NULL => counter mode */
	void * streamState;
	uint64_t key;
	uint32_t dimBase;
} rng_t;
static float rnd(rng_t * r, uint32_t dim) { return r->stream ? r->stream->next_float(r->streamState) : orc_rng_float(r->key, r->dimBase + dim); }

/* ------------------------------------------------------------------ images */
static int modPos(int x, int w) { return ((x % w) + w) % w; } /* Image.hpp:189-191 */

/* Image::rgbal (Image.cpp:86-106) / rgbl (:64-84): `comps` floats per texel, rgb or rgba read at that stride. */
static v4 image_sample(const orc_image * im, float x, float y, int wantAlpha) {
	const float xi = x * (float)im->width, yi = y * (float)im->height;
	const float xb = floorf(xi), yb = floorf(yi);
	const float dx = xi - xb, dy = yi - yb;
	const int x0 = modPos((int)xb, (int)im->width), y0 = modPos((int)yb, (int)im->height);
	const int x1 = modPos((int)xb + 1, (int)im->width), y1 = modPos((int)yb + 1, (int)im->height);
	const float * p00 = im->pixels + ((size_t)y0 * im->width + x0) * im->components;
	const float * p01 = im->pixels + ((size_t)y1 * im->width + x0) * im->components;
	const float * p10 = im->pixels + ((size_t)y0 * im->width + x1) * im->components;
	const float * p11 = im->pixels + ((size_t)y1 * im->width + x1) * im->components;
	float o[4] = {0, 0, 0, 0};
	const int n = wantAlpha ? 4 : 3;
	for(int c = 0; c < n; ++c) {
		o[c] = (1.0f - dx) * ((1.0f - dy) * p00[c] + dy * p01[c]) + dx * ((1.0f - dy) * p10[c] + dy * p11[c]);
	}
	v4 r = {o[0], o[1], o[2], o[3]};
	return r;
}
void orc_image_rgbal(const orc_image * im, const float * uv, size_t n, float * out4) {
	for(size_t i = 0; i < n; ++i) {
		const v4 r = image_sample(im, uv[2 * i], uv[2 * i + 1], 1);
		out4[4 * i] = r.x; out4[4 * i + 1] = r.y; out4[4 * i + 2] = r.z; out4[4 * i + 3] = r.w;
	}
}

/* Texture::sampleCubemap (Texture.cpp:141-191) */
static v3 sample_cubemap(const orc_image * faces, v3 dir) {
	const float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
	int side = 0;
	float x = 0.0f, y = 0.0f, denom = 1.0f;
	if(ax >= ay && ax >= az) {
		denom = ax; y = dir.y;
		if(dir.x >= 0.0f) { side = 0; x = -dir.z; } else { side = 1; x = dir.z; }
	} else if(ay >= ax && ay >= az) {
		denom = ay; x = dir.x;
		if(dir.y >= 0.0f) { side = 2; y = -dir.z; } else { side = 3; y = dir.z; }
	} else if(az >= ax && az >= ay) {
		denom = az; y = dir.y;
		if(dir.z >= 0.0f) { side = 4; x = dir.x; } else { side = 5; x = -dir.x; }
	}
	x = 0.5f * (x / denom) + 0.5f;
	y = 0.5f * (-y / denom) + 0.5f;
	const uint32_t md = faces[side].width < faces[side].height ? faces[side].width : faces[side].height;
	const float eps	  = 1.0f / (float)md;
	x = gclamp(x, 0.0f + eps, 1.0f - eps);
	y = gclamp(y, 0.0f + eps, 1.0f - eps);
	const v4 r = image_sample(&faces[side], x, y, 0);
	return V(r.x, r.y, r.z);
}
void orc_cubemap(const orc_image * faces6, const float * dirs, size_t n, float * out3) {
	for(size_t i = 0; i < n; ++i) {
		const v3 c = sample_cubemap(faces6, V(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
		out3[3 * i] = c.x; out3[3 * i + 1] = c.y; out3[3 * i + 2] = c.z;
	}
}

/* ------------------------------------------------------------------ lights */
/* PointLight.cpp:68-87, SpotLight.cpp:52-81 (note cos(cos(outer)) at :68), DirectionalLight.cpp:40-44 */
static v3 light_sample(const orc_light * l, v3 position, float * dist, float * attenuation) {
	if(l->kind == 2) {
		*attenuation = 1.0f;
		*dist		 = FLT_MAX;
		return V(-l->direction[0], -l->direction[1], -l->direction[2]);
	}
	v3 direction = sub(V(l->position[0], l->position[1], l->position[2]), position);
	*dist		 = length(direction);
	*attenuation = 0.0f;
	if(*dist > l->radius) return V(0, 0, 0);
	if(*dist > 0.0f) direction = divs(direction, *dist);
	float angleAttenuation = 1.0f;
	if(l->kind == 1) {
		const float currentCos = dot(neg(direction), V(l->direction[0], l->direction[1], l->direction[2]));
		const float outerCos   = cosf(l->outer);
		if(currentCos < cosf(outerCos)) return V(0, 0, 0);
		const float innerCos = cosf(l->inner);
		angleAttenuation	 = gclamp((currentCos - outerCos) / (innerCos - outerCos), 0.0f, 1.0f);
	}
	const float radiusRatio	 = *dist / l->radius;
	const float radiusRatio2 = radiusRatio * radiusRatio;
	const float attenNum	 = gclamp(1.0f - radiusRatio2, 0.0f, 1.0f);
	*attenuation = (l->kind == 1) ? (angleAttenuation * attenNum * attenNum) : (attenNum * attenNum);
	return direction;
}
void orc_light_sample(const orc_light * l, const float * pos, size_t n, float * out5) {
	for(size_t i = 0; i < n; ++i) {
		float dist = 0.0f, att = 0.0f;
		const v3 d = light_sample(l, V(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), &dist, &att);
		out5[5 * i] = d.x; out5[5 * i + 1] = d.y; out5[5 * i + 2] = d.z; out5[5 * i + 3] = dist; out5[5 * i + 4] = att;
	}
}

/* ------------------------------------------------------------------ MaterialGGX */
static float alphaFromRoughness(float roughness) { /* MaterialGGX.cpp:24-28 */
	const float roughClamp = smax(0.045f, roughness);
	return smax(0.0001f, roughClamp * roughClamp);
}
static float ggxD(float NdotH, float alpha) { /* :9-13 */
	const float halfDenum = NdotH * NdotH * (alpha * alpha - 1.0f) + 1.0f;
	const float halfTerm  = alpha / smax(0.0001f, halfDenum);
	return halfTerm * halfTerm * ONE_OVER_PI;
}
static float ggxV(float NdotL, float NdotV, float alpha) { /* :15-21 */
	const float alpha2 = alpha * alpha;
	const float visL   = NdotV * sqrtf((-NdotL * alpha2 + NdotL) * NdotL + alpha2);
	const float visV   = NdotL * sqrtf((-NdotV * alpha2 + NdotV) * NdotV + alpha2);
	return 0.5f / smax(0.0001f, visV + visL);
}
static v3 ggx(v3 wo, v3 baseColor, float alpha, float metallic, v3 wi, float * pdf) { /* :30-52 */
	const v3 h		  = normalize(add(wi, wo));
	const float NdotH = smax(h.z, 0.0f);
	const float VdotH = smax(dot(wi, h), 0.0f);
	const float NdotL = smax(wo.z, 0.0f);
	const float NdotV = smax(wi.z, 0.0f);
	const float Dh	  = ggxD(NdotH, alpha);
	if(pdf) {
		const float hPdf = Dh * NdotH;
		*pdf			 = hPdf / (4.0f * smax(0.0001f, VdotH));
	}
	const v3 F0		  = mix3(V(0.04f, 0.04f, 0.04f), baseColor, metallic);
	const v3 F		  = add(F0, scaleL(powf(1.0f - VdotH, 5.0f), sub(V(1, 1, 1), F0))); /* :5-7 */
	const v3 specular = scaleL(Dh * ggxV(NdotL, NdotV, alpha), F);
	const v3 diffuse  = mulv(scaleL((1.0f - metallic) * ONE_OVER_PI, baseColor), sub(V(1, 1, 1), F0));
	const v3 multiAdj = add(V(1, 1, 1), scaleL(2.0f * alpha * alpha * NdotV, F0));
	return scale(add(diffuse, mulv(specular, multiAdj)), NdotV);
}
static v3 ggx_eval(v3 wo, v3 baseColor, float roughness, float metallic, v3 wi) { /* :94-101 */
	if(wi.z < 0.0f) return V(0, 0, 0);
	return ggx(wo, baseColor, alphaFromRoughness(roughness), metallic, wi, NULL);
}
static v3 cosine_hemisphere(float u0, float u1) { /* Random.cpp:41-70 */
	const float x = 2.0f * u0 - 1.0f;
	const float y = 2.0f * u1 - 1.0f;
	float dx = 0.0f, dy = 0.0f;
	if(!(x == 0.0f && y == 0.0f)) {
		float angle, radius;
		if(fabsf(x) > fabsf(y)) {
			radius = x;
			angle  = QUARTER_PI * y / x;
		} else {
			radius = y;
			angle  = HALF_PI - QUARTER_PI * x / y;
		}
		dx = radius * cosf(angle);
		dy = radius * sinf(angle);
	}
	return V(dx, dy, sqrtf(smax(0.0f, 1.0f - dx * dx - dy * dy)));
}
static v3 ggx_sample_and_eval(v3 wo, v3 baseColor, float roughness, float metallic, rng_t * r, v3 * wi) { /* :54-92 */
	const float probaSpecular = mixf(1.0f / (dot(baseColor, V(1, 1, 1)) / 3.0f + 1.0f), 1.0f, metallic);
	const float alpha		  = alphaFromRoughness(roughness);
	if(rnd(r, 1) < probaSpecular) {
		const float a2			 = alpha * alpha;
		const float x			 = rnd(r, 2);
		const float phiH		 = rnd(r, 3) * TWO_PI;
		const float cosThetaHSqr = smin((1.0f - x) / ((a2 - 1.0f) * x + 1.0f), 1.0f);
		const float cosThetaH	 = sqrtf(cosThetaHSqr);
		const float sinThetaH	 = sqrtf(1.0f - cosThetaHSqr);
		/* The reference calls the unqualified C `cos`/`sin` here (MaterialGGX.cpp:69): double precision, and the product
		 * with sinThetaH is formed in double before glm::vec3 narrows it to float. */
		const v3 lh = V((float)((double)sinThetaH * cos((double)phiH)), (float)((double)sinThetaH * sin((double)phiH)), cosThetaH);
		*wi = sub(scaleL(2.0f * dot(wo, lh), lh), wo);
	} else {
		const float u0 = rnd(r, 2);
		const float u1 = rnd(r, 3);
		*wi = cosine_hemisphere(u0, u1);
		if(wo.z < 0.0f) wi->z *= -1.0f;
	}
	if(wi->z < 0.0f) return V(0, 0, 0);
	float pdfSpec = 0.0f;
	const v3 brdf = ggx(wo, baseColor, alpha, metallic, *wi, &pdfSpec);
	const float pdf = mixf(ONE_OVER_PI * smax(wi->z, 0.0f), pdfSpec, probaSpecular);
	if(pdf == 0.0f) return V(0, 0, 0);
	return divs(brdf, pdf);
}
void orc_ggx_eval(const float * wo, const float * base, const float * rough, const float * metal, const float * wi, size_t n, float * out3) {
	for(size_t i = 0; i < n; ++i) {
		const v3 r = ggx_eval(V(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), V(base[3 * i], base[3 * i + 1], base[3 * i + 2]), rough[i], metal[i], V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]));
		out3[3 * i] = r.x; out3[3 * i + 1] = r.y; out3[3 * i + 2] = r.z;
	}
}
typedef struct { const float * u; int k; } fixed_u;
static float fixed_next(void * s) { fixed_u * f = (fixed_u *)s; return f->u[f->k++]; }
/* MaterialGGX::sampleAndEval with the three uniforms it consumes given explicitly (u3: lobe pick, then two lobe variates). */
void orc_ggx_sample(const float * wo, const float * base, const float * rough, const float * metal, const float * u3, size_t n, float * outWi, float * outEval) {
	orc_rng_stream st = {fixed_next, NULL};
	for(size_t i = 0; i < n; ++i) {
		fixed_u f = {u3 + 3 * i, 0};
		rng_t r	  = {&st, &f, 0, 0};
		v3 wi	  = V(0, 0, 0);
		const v3 e = ggx_sample_and_eval(V(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), V(base[3 * i], base[3 * i + 1], base[3 * i + 2]), rough[i], metal[i], &r, &wi);
		outWi[3 * i] = wi.x; outWi[3 * i + 1] = wi.y; outWi[3 * i + 2] = wi.z;
		outEval[3 * i] = e.x; outEval[3 * i + 1] = e.y; outEval[3 * i + 2] = e.z;
	}
}

/* ------------------------------------------------------------------ MaterialSky */
static int sphere(v3 rayOrigin, v3 rayDir, float radius, float * r0, float * r1) { /* Intersection.cpp:7-20 */
	const float a = dot(rayDir, rayDir), b = dot(rayOrigin, rayDir), c = dot(rayOrigin, rayOrigin) - radius * radius;
	const float delta = b * b - a * c;
	if(delta < 0.0f) return 0;
	const float dsqrt = sqrtf(delta);
	*r0 = (-b - dsqrt) / a;
	*r1 = (-b + dsqrt) / a;
	return 1;
}
static v3 sky_eval(const orc_image * lut, v3 rayOrigin, v3 rayDir, v3 sunDir) { /* MaterialSky.cpp:6-87, Sky.hpp:53-65 */
	const v3 sunColor = V(1.474f, 1.8504f, 1.91198f), kRayleigh = V(5.5e-6f, 13.0e-6f, 22.4e-6f);
	const float groundRadius = 6371e3f, topRadius = 6471e3f, sunIntensity = 20.0f, kMie = 21e-6f;
	const float heightRayleigh = 8000.0f, heightMie = 1200.0f, gMie = 0.758f, sunRadius = 0.04675f, sunRadiusCos = 0.998f;
	const unsigned samplesCount = 16;
	const v3 planetPos = add(add(rayOrigin, V(0.0f, groundRadius, 0.0f)), V(0.0f, 1.0f, 0.0f));
	float top0 = 0.0f, top1 = 0.0f, gr0 = 0.0f, gr1 = 0.0f;
	if(!sphere(planetPos, rayDir, topRadius, &top0, &top1)) return V(0, 0, 0);
	const int didHitGround		= sphere(planetPos, rayDir, groundRadius, &gr0, &gr1);
	const float distanceToInter = smin(top1, didHitGround ? gr0 : 0.0f);
	const float stepSize		= (distanceToInter - top0) / (float)samplesCount;
	const float cosViewSun		= dot(rayDir, sunDir);
	float rayleighDist = 0.0f, mieDist = 0.0f;
	v3 rayleighScatt = V(0, 0, 0), mieScatt = V(0, 0, 0), transmittance = V(0, 0, 0);
	for(unsigned i = 0; i < samplesCount; ++i) {
		const v3 currPos = add(planetPos, scaleL(((float)i + 0.5f) * stepSize, rayDir));
		float currHeight = length(currPos) - groundRadius;
		if(i == (samplesCount - 1) && currHeight < 0.0f) currHeight = 0.0f;
		const float rayleighStep = expf(-currHeight / heightRayleigh) * stepSize;
		const float mieStep		 = expf(-currHeight / heightMie) * stepSize;
		rayleighDist += rayleighStep;
		mieDist += mieStep;
		const v3 e = neg(add(scaleL(mieDist, V(kMie, kMie, kMie)), scaleL(rayleighDist, kRayleigh)));
		const v3 directAttenuation = V(expf(e.x), expf(e.y), expf(e.z));
		const float relativeHeight	 = (length(currPos) - groundRadius) / (topRadius - groundRadius);
		const float relativeCosAngle = -0.5f * sunDir.y + 0.5f;
		float uvx = (511.0f / 512.0f) * relativeHeight + 0.5f / 512.0f;
		float uvy = (511.0f / 512.0f) * relativeCosAngle + 0.5f / 512.0f;
		uvx = gclamp(uvx, 0.5f / 512.0f, 511.0f / 512.0f);
		uvy = gclamp(uvy, 0.5f / 512.0f, 511.0f / 512.0f);
		const v4 s4 = image_sample(lut, uvx, uvy, 0);
		const v3 attenuation = mulv(directAttenuation, V(s4.x, s4.y, s4.z));
		rayleighScatt = add(rayleighScatt, scaleL(rayleighStep, attenuation));
		mieScatt	  = add(mieScatt, scaleL(mieStep, attenuation));
		transmittance = add(transmittance, directAttenuation);
	}
	const float k			  = 1.0f / (4.0f * PI_F);
	const float rayleighPhase = k * 3.0f / 4.0f * (1.0f + cosViewSun * cosViewSun); /* :89-92 */
	const float g2			  = gMie * gMie;
	const float miePhase	  = k * 3.0f * (1.0f - g2) / (2.0f * (2.0f + g2)) * (1.0f + cosViewSun * cosViewSun) / powf(1.0f + g2 - 2.0f * gMie * cosViewSun, 3.0f / 2.0f); /* :94-98 */
	const v3 rayleighParticipation = mulv(scaleL(rayleighPhase, kRayleigh), rayleighScatt);
	const v3 mieParticipation	   = scaleL(kMie * miePhase, mieScatt);
	v3 sunRadiance = V(0, 0, 0);
	const int didHitGroundForward = didHitGround && gr1 > 0.0f;
	if(!didHitGroundForward && dot(rayDir, sunDir) > sunRadiusCos) sunRadiance = divs(sunColor, PI_F * sunRadius * sunRadius);
	return add(scaleL(sunIntensity, add(rayleighParticipation, mieParticipation)), mulv(transmittance, sunRadiance));
}
void orc_sky_eval(const orc_image * lut, const float * org, const float * dirs, const float * sun, size_t n, float * out3) {
	for(size_t i = 0; i < n; ++i) {
		const v3 c = sky_eval(lut, V(org[3 * i], org[3 * i + 1], org[3 * i + 2]), V(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), V(sun[0], sun[1], sun[2]));
		out3[3 * i] = c.x; out3[3 * i + 1] = c.y; out3[3 * i + 2] = c.z;
	}
}

/* ------------------------------------------------------------------ PathTracer */
static v3 eval_background(const orc_scene * s, v3 rayDir, v3 rayPos, float ndcx, float ndcy, int directHit) { /* PathTracer.cpp:23-55 */
	const uint32_t mode = s->bgMode;
	if(directHit) {
		if(mode == 1) { const v4 c = image_sample(&s->images[s->bgTex[0]], ndcx, ndcy, 0); return V(c.x, c.y, c.z); }
		if(mode == 2) { orc_image f[6]; for(int k = 0; k < 6; ++k) f[k] = s->images[s->bgTex[k]]; return sample_cubemap(f, normalize(rayDir)); }
		if(mode == 3) return sky_eval(&s->images[s->skyLut], rayPos, normalize(rayDir), V(s->sunDir[0], s->sunDir[1], s->sunDir[2]));
		return V(s->bgColor[0], s->bgColor[1], s->bgColor[2]);
	}
	if(mode == 2) { orc_image f[6]; for(int k = 0; k < 6; ++k) f[k] = s->images[s->bgTex[k]]; return sample_cubemap(f, normalize(rayDir)); }
	if(mode == 3) return sky_eval(&s->images[s->skyLut], rayPos, normalize(rayDir), V(s->sunDir[0], s->sunDir[1], s->sunDir[2]));
	return V(0, 0, 0);
}

/* Raycaster::interpolateAttribute (Raycaster.hpp:90-97) */
static v3 interp3(const orc_hit * h, const orc_object * o, const float * attr) {
	const uint32_t i0 = o->indices[h->localId], i1 = o->indices[h->localId + 1], i2 = o->indices[h->localId + 2];
	v3 r;
	r.x = h->w * attr[3 * i0] + h->u * attr[3 * i1] + h->v * attr[3 * i2];
	r.y = h->w * attr[3 * i0 + 1] + h->u * attr[3 * i1 + 1] + h->v * attr[3 * i2 + 1];
	r.z = h->w * attr[3 * i0 + 2] + h->u * attr[3 * i1 + 2] + h->v * attr[3 * i2 + 2];
	return r;
}
static void interp2(const orc_hit * h, const orc_object * o, const float * attr, float * u, float * v) {
	const uint32_t i0 = o->indices[h->localId], i1 = o->indices[h->localId + 1], i2 = o->indices[h->localId + 2];
	*u = h->w * attr[2 * i0] + h->u * attr[2 * i1] + h->v * attr[2 * i2];
	*v = h->w * attr[2 * i0 + 1] + h->u * attr[2 * i1 + 1] + h->v * attr[2 * i2 + 1];
}

/* glm::inverse(glm::transpose(glm::mat3(model))) (PathTracer.cpp:94; func_matrix.inl:87-105, :269-291) */
static m3 inverse_transpose(const float * model) {
	float m[3][3], inv[3][3];
	for(int c = 0; c < 3; ++c) for(int r = 0; r < 3; ++r) m[c][r] = model[r * 4 + c];
	const float d = 1.0f / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
	inv[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * d;
	inv[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * d;
	inv[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * d;
	inv[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * d;
	inv[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * d;
	inv[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * d;
	inv[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * d;
	inv[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * d;
	inv[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * d;
	m3 r;
	for(int c = 0; c < 3; ++c) r.c[c] = V(inv[c][0], inv[c][1], inv[c][2]);
	return r;
}

static m3 build_local_frame(const orc_scene * s, const orc_object * obj, const orc_hit * hit, v3 rayDir, float u, float v) { /* PathTracer.cpp:86-122 */
	const v3 n = normalize(interp3(hit, obj, obj->normals));
	const v3 t = normalize(interp3(hit, obj, obj->tangents));
	const v3 b = normalize(interp3(hit, obj, obj->bitangents));
	const m3 invtp = inverse_transpose(obj->model);
	m3 tbn;
	tbn.c[0] = m3mul(&invtp, t);
	tbn.c[1] = m3mul(&invtp, b);
	tbn.c[2] = m3mul(&invtp, n);
	const int frontFacing = dot(tbn.c[2], rayDir) < 0.0f;
	if(!frontFacing) tbn.c[2] = scale(tbn.c[2], -1.0f);
	if(obj->useTexCoords && !obj->emissive) {
		const v4 img = image_sample(&s->images[obj->tex[1]], u, v, 1);
		const v3 localNormal = normalize(sub(scaleL(2.0f, V(img.x, img.y, img.z)), V(1, 1, 1)));
		tbn.c[2] = m3mul(&tbn, localNormal);
	}
	tbn.c[2] = normalize(tbn.c[2]);
	tbn.c[1] = normalize(cross(tbn.c[2], normalize(tbn.c[0])));
	tbn.c[0] = normalize(cross(tbn.c[1], tbn.c[2]));
	return tbn;
}

static int check_visibility(const orc_scene * s, v3 startPos, v3 rayDir, float maxDist, uint64_t * rays) { /* PathTracer.cpp:124-157 */
	v3 lpos = startPos;
	while(maxDist > 0.0f) {
		const float o[3] = {lpos.x, lpos.y, lpos.z}, d[3] = {rayDir.x, rayDir.y, rayDir.z};
		const orc_hit lhit = orc_intersect(&s->bvh, o, d, 0.001f, maxDist);
		++*rays;
		if(!lhit.hit) return 1;
		const orc_object * lobj = &s->objects[lhit.meshId];
		if(!lobj->masked || !lobj->useTexCoords) return 0;
		float lu, lv;
		interp2(&lhit, lobj, lobj->texcoords, &lu, &lv);
		const float alpha = image_sample(&s->images[lobj->tex[0]], lu, lv, 1).w;
		if(alpha < 0.01f) {
			maxDist = maxDist - lhit.dist;
			lpos	= add(lpos, scaleL(lhit.dist, rayDir));
		} else {
			return 0;
		}
	}
	return 1;
}

typedef struct {
	const orc_scene * scene; const orc_render_args * a; float * pixels; uint32_t cellsX, cellsY; float cellSizeX, cellSizeY;
	uint32_t y0, y1; const orc_rng_stream * stream; void * streamState; uint64_t rays;
} render_job;

/* One pixel sample: the body of the `sid` loop of PathTracer::render (PathTracer.cpp:188-288). */
static v3 trace_sample(render_job * j, uint32_t x, uint32_t y, uint32_t sid) {
	const orc_scene * s		  = j->scene;
	const orc_render_args * a = j->a;
	rng_t r = {j->stream, j->streamState, 0, 0};
	if(!j->stream) r.key = orc_rng_key(a->seed, y * a->width + x, sid);
	/* getSamplePosition (:71-84) */
	const int sidy = (int)sid / (int)j->cellsX, sidx = (int)sid % (int)j->cellsX;
	const float jx = rnd(&r, 0);
	const float jy = rnd(&r, 1);
	const float lx = ((float)sidx + jx) * j->cellSizeX, ly = ((float)sidy + jy) * j->cellSizeY;
	const float sx = (float)x + lx, sy = (float)y + ly;
	const float ndcx = sx / (float)a->width, ndcy = sy / (float)a->height;
	const v3 corner = V(a->corner[0], a->corner[1], a->corner[2]), dx = V(a->dx[0], a->dx[1], a->dx[2]), dy = V(a->dy[0], a->dy[1], a->dy[2]);
	const v3 eye	= V(a->eye[0], a->eye[1], a->eye[2]);
	const v3 worldPos = add(add(corner, scaleL(ndcx, dx)), scaleL(ndcy, dy));
	v3 rayPos = eye;
	v3 rayDir = normalize(sub(worldPos, eye));
	v3 sampleColor = V(0, 0, 0), attenuation = V(1, 1, 1);
	for(uint32_t did = 0; did < a->depth; ++did) {
		r.dimBase = 8u * (did + 1u);
		const float o[3] = {rayPos.x, rayPos.y, rayPos.z}, d[3] = {rayDir.x, rayDir.y, rayDir.z};
		const orc_hit hit = orc_intersect(&s->bvh, o, d, 0.0001f, 1e8f);
		++j->rays;
		if(!hit.hit) {
			sampleColor = add(sampleColor, mulv(attenuation, eval_background(s, rayDir, rayPos, ndcx, ndcy, did == 0)));
			break;
		}
		const orc_object * obj = &s->objects[hit.meshId];
		const v3 p = add(rayPos, scaleL(hit.dist, rayDir));
		float u = 0.5f, v = 0.5f;
		if(obj->useTexCoords) interp2(&hit, obj, obj->texcoords, &u, &v);
		const v4 bCol = image_sample(&s->images[obj->tex[0]], u, v, 1);
		if(obj->masked && bCol.w < 0.01f) {
			rayPos = p;
			continue;
		}
		if(obj->emissive) {
			sampleColor = add(sampleColor, mulv(attenuation, V(bCol.x, bCol.y, bCol.z)));
			break;
		}
		const m3 tbn	   = build_local_frame(s, obj, &hit, rayDir, u, v);
		const v3 wo		   = normalize(m3mulT(&tbn, neg(rayDir)));
		const v3 baseColor = V(powf(bCol.x, 2.2f), powf(bCol.y, 2.2f), powf(bCol.z, 2.2f));
		const v4 rmao	   = image_sample(&s->images[obj->tex[2]], u, v, 1);
		if(s->nlights > 0) {
			uint32_t lid;
			if(j->stream) lid = (uint32_t)j->stream->next_int(j->streamState, 0, (int)s->nlights - 1);
			else {
				int l = (int)(rnd(&r, 0) * (float)s->nlights);
				lid	  = (uint32_t)(l < (int)s->nlights ? l : (int)s->nlights - 1);
			}
			const orc_light * light = &s->lights[lid];
			const v3 pShift = add(p, scaleL(0.001f, tbn.c[2]));
			float maxDist, falloff;
			const v3 direction = light_sample(light, pShift, &maxDist, &falloff);
			int visible = falloff > 0.0f;
			if(visible && light->castsShadow) visible = check_visibility(s, pShift, direction, maxDist, &j->rays);
			if(visible) {
				const v3 lwi		= normalize(m3mulT(&tbn, direction));
				const v3 evalLight	= ggx_eval(wo, baseColor, rmao.x, rmao.y, lwi);
				const float lightPdf = 1.0f / (float)s->nlights;
				const v3 illumination = divs(mulv(scaleL(falloff, evalLight), V(light->intensity[0], light->intensity[1], light->intensity[2])), lightPdf);
				sampleColor = add(sampleColor, mulv(attenuation, illumination));
			}
		}
		v3 wi = V(0, 0, 0);
		const v3 eval		= ggx_sample_and_eval(wo, baseColor, rmao.x, rmao.y, &r, &wi);
		const v3 nextRayDir = normalize(m3mul(&tbn, wi));
		attenuation = mulv(attenuation, eval);
		if(did < a->depth - 1) {
			rayPos = p;
			rayDir = normalize(nextRayDir);
		}
	}
	return V(gmin(sampleColor.x, 5.0f), gmin(sampleColor.y, 5.0f), gmin(sampleColor.z, 5.0f)); /* :288 */
}

static void * render_rows(void * arg) {
	render_job * j = (render_job *)arg;
	const orc_render_args * a = j->a;
	const int windowed = a->window[2] > a->window[0] && a->window[3] > a->window[1];
	for(uint32_t y = j->y0; y < j->y1; ++y) {
		if(windowed && (y < a->window[1] || y >= a->window[3])) continue;
		for(uint32_t x = 0; x < a->width; ++x) {
			if(windowed && (x < a->window[0] || x >= a->window[2])) continue;
			float * px = j->pixels + ((size_t)y * a->width + x) * a->components;
			for(uint32_t sid = 0; sid < a->samples; ++sid) {
				if(a->sampleCount > 1 && sid % a->sampleCount != a->sampleIndex) continue;
				const v3 c = trace_sample(j, x, y, sid);
				px[0] += c.x; px[1] += c.y; px[2] += c.z;
			}
			px[0] = px[0] / (float)a->samples; /* :291-292 */
			px[1] = px[1] / (float)a->samples;
			px[2] = px[2] / (float)a->samples;
		}
	}
	return NULL;
}

/* PathTracer::render (PathTracer.cpp:159-299). `samples` must already be the power of two the reference rounds to.
 * stream == NULL: counter-based RNG, any thread count. stream != NULL: sequential generator (mt19937), one thread. */
uint64_t orc_render(const orc_scene * scene, const orc_render_args * a, const orc_rng_stream * stream, void * streamState, int threads, float * pixels) {
	const int k = (int)floor(log2((double)a->samples)); /* getSampleGrid (:57-69) */
	uint32_t cx = (uint32_t)pow(2, k / 2), cy = cx;
	if(k % 2 == 1) {
		cx = (uint32_t)pow(2, (k + 1) / 2);
		cy = (uint32_t)pow(2, (k - 1) / 2);
	}
	if(stream || threads < 1) threads = 1;
	if((uint32_t)threads > a->height) threads = (int)a->height;
	render_job * jobs = (render_job *)calloc((size_t)threads, sizeof(render_job));
	pthread_t * tid	  = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
	const uint32_t span = a->height / (uint32_t)threads;
	for(int t = 0; t < threads; ++t) {
		render_job j = {scene, a, pixels, cx, cy, 1.0f / (float)(int)cx, 1.0f / (float)(int)cy, (uint32_t)t * span, (t == threads - 1) ? a->height : (uint32_t)(t + 1) * span, stream, streamState, 0};
		jobs[t] = j;
		if(threads == 1) render_rows(&jobs[t]);
		else pthread_create(&tid[t], NULL, render_rows, &jobs[t]);
	}
	uint64_t rays = 0;
	for(int t = 0; t < threads; ++t) {
		if(threads > 1) pthread_join(tid[t], NULL);
		rays += jobs[t].rays;
	}
	free(jobs);
	free(tid);
	return rays;
}

/* Output stage of the headless render: tonemap of renderOneShot (src/apps/pathtracer/main.cpp:109-116) followed by
 * Image::saveAsLDR's conversion (src/engine/resources/Image.cpp:138-152) with SRGB_LDR | IGNORE_ALPHA (main.cpp:119). */
void orc_output_ldr(const float * pixels, uint32_t width, uint32_t height, uint32_t components, int tonemap, int srgb, uint8_t * out) {
	const size_t npx = (size_t)width * height;
	for(size_t pid = 0; pid < npx; ++pid) {
		for(uint32_t cid = 0; cid < components; ++cid) {
			const size_t cur = components * pid + cid;
			float value = pixels[cur];
			if(tonemap && cid < 3) value = 1.0f - expf(-value * 1.f);
			if(srgb && cid != 3) value = powf(value, 1.0f / 2.2f);
			const float newValue = smin(255.0f, smax(0.0f, 255.0f * value));
			out[cur] = (unsigned char)newValue;
			if(cid == 3) out[cur] = 255;
		}
	}
}

/* float -> half exactly as the EXR writer the reference links does it (src/libs/tinyexr/tinyexr.h:971-1008,
 * float_to_half_full): zero/denormal floats flush to signed zero, Inf/NaN keep their class (NaN -> 0x200 payload),
 * overflow saturates to infinity, and rounding adds one when the first dropped mantissa bit is set (round half up, not
 * round-to-nearest-even); the increment may carry into the exponent. */
uint16_t orc_float_to_half(float value) {
	uint32_t bits;
	memcpy(&bits, &value, sizeof(bits));
	const uint32_t sign = bits >> 31, exponent = (bits >> 23) & 0xffu, mantissa = bits & 0x7fffffu;
	uint32_t h = 0;
	if(exponent == 0) {
		h = 0;
	} else if(exponent == 255) {
		h = 0x7c00u | (mantissa ? 0x200u : 0u);
	} else {
		const int newexp = (int)exponent - 127 + 15;
		if(newexp >= 31) {
			h = 0x7c00u;
		} else if(newexp <= 0) {
			if((14 - newexp) <= 24) {
				const uint32_t mant = mantissa | 0x800000u;
				h = (mant >> (14 - newexp)) & 0x3ffu;
				if((mant >> (13 - newexp)) & 1u) h++;
			}
		} else {
			h = ((uint32_t)newexp << 10) | (mantissa >> 13);
			if(mantissa & 0x1000u) h++;
		}
	}
	return (uint16_t)((h & 0x7fffu) | (sign << 15));
}

/* half -> float (tinyexr.h:946-969), used by the tests to compare with what the reference's loader reads back. */
float orc_half_to_float(uint16_t h) {
	const uint32_t sign = (uint32_t)(h >> 15) << 31, exponent = (h >> 10) & 0x1fu, mantissa = h & 0x3ffu;
	uint32_t bits;
	float out;
	if(exponent == 31) {
		bits = sign | 0x7f800000u | (mantissa << 13);
	} else if(exponent == 0) {
		/* zero / denormal: mantissa * 2^-24 is exact in fp32 */
		out = (float)mantissa * (1.0f / 16777216.0f);
		memcpy(&bits, &out, sizeof(bits));
		bits |= sign;
	} else {
		bits = sign | ((exponent + 112u) << 23) | (mantissa << 13);
	}
	memcpy(&out, &bits, sizeof(out));
	return out;
}

/* Image::saveAsHDR's layer split (src/engine/resources/Image.cpp:178-232) and the half conversion of the writer:
 * channels = components, except 2 -> 3 (third layer zero); planes are written in the file's channel order — (A)BGR for
 * 4/3 channels, the single layer for 1 — each width*height halves; IGNORE_ALPHA stores 1.0 in A; flip reverses rows. */
void orc_output_hdr(const float * pixels, uint32_t width, uint32_t height, uint32_t components, int ignore_alpha, int flip, uint16_t * planes) {
	const uint32_t channels = components == 2 ? 3 : components;
	const size_t npx = (size_t)width * height;
	for(uint32_t y = 0; y < height; ++y) {
		for(uint32_t x = 0; x < width; ++x) {
			const size_t dst = (size_t)y * width + x;
			const size_t src = flip ? ((size_t)(height - 1 - y) * width + x) : dst;
			for(uint32_t j = 0; j < channels; ++j) {
				float value = j < components ? pixels[(size_t)components * src + j] : 0.0f;
				if(j == 3 && ignore_alpha) value = 1.0f;
				const uint32_t plane = channels == 1 ? 0 : channels - 1 - j; /* R is the last plane */
				planes[plane * npx + dst] = orc_float_to_half(value);
			}
		}
	}
}
